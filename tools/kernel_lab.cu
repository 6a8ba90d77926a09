// Developer tool (not part of librlic_b200.so): times the shipped pass kernel and
// candidate formulations on the headline workload, checks candidates bit for bit
// against the shipped kernel, and checks the short division used by the fast
// path against the library's IEEE division.
// Build: tools/build_lab.sh    Run on a B200: tools/kernel_lab [n] [taps] [name-filter]
//
// History of what was tried here (numbers in profiles/r1_lab*.txt):
//   lab1  tile shapes 32x2 ... 8x32, wall handling as select / inline branch / out-of-line call
//   lab2  (at, j) walker state, merged rare path, unrolling
//   lab3  precomputed refined reciprocals in the packed field  -> adopted
//   lab5-10  launch bounds, sign arithmetic, zero components, admission tests, tap addressing
//   lab11 wall sentinels in padded buffers (no column counter, no wall compares)  -> adopted
//   lab18 (written without a GPU, to be run first thing in round 2) the grouped walk: WALK
//         bits x sign flavours 0-3; static SASS counts in profiles/r1_sass_steps_grouped_walk.txt
#include "../rlic_b200/csrc/lic_walk.cuh"

#include <cuda.h>
#include <cudaTypedefs.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
    fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

using rlic::Fp;
using rlic::PackedField;
using rlic::PassGeom;

static int g_split_field = 0;   // 1: BASELINE config 3's field (u = -1 | +1 split at mid-width, v = 0)

template <typename T>
__global__ void fill_inputs(T *tex, T *u, T *v, int n, int split)
{
    long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (long long)n * n) return;
    int i = (int)(p / n), j = (int)(p % n);
    unsigned h = (unsigned)p * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    tex[p] = (T)((h >> 8) * (1.0 / 16777216.0));
    if (split) {
        u[p] = j < n / 2 ? (T)-1 : (T)1;
        v[p] = (T)0;
    } else {
        u[p] = (T)(-(-1.0 + 2.0 * i / (n - 1)));
        v[p] = (T)(-1.0 + 2.0 * j / (n - 1));
    }
}

// div_tail against IEEE division on the ranges the fast path admits:
// |b| in [2^-40, 2^40], |a| in [2^-60, 4).
template <typename T> struct Bits;
template <> struct Bits<float> {
    static __device__ float make(int e, unsigned long long m, bool neg) {
        return __int_as_float((neg ? 0x80000000u : 0u) | ((unsigned)(e + 127) << 23) | (unsigned)(m & 0x7fffff));
    }
    static __device__ bool same(float a, float b) { return __float_as_int(a) == __float_as_int(b); }
    static constexpr unsigned long long ones = 0x7fffff;
};
template <> struct Bits<double> {
    static __device__ double make(int e, unsigned long long m, bool neg) {
        return __longlong_as_double((long long)((neg ? 0x8000000000000000ull : 0ull) |
                                                ((unsigned long long)(e + 1023) << 52) | (m & 0xfffffffffffffull)));
    }
    static __device__ bool same(double a, double b) { return __double_as_longlong(a) == __double_as_longlong(b); }
    static constexpr unsigned long long ones = 0xfffffffffffffull;
};

template <typename T>
__global__ void divcheck_kernel(unsigned long long seed, unsigned long long *bad, unsigned long long n_per_thread)
{
    unsigned long long x = seed + 0x9E3779B97F4A7C15ull * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 1);
    unsigned long long local_bad = 0;
    for (unsigned long long it = 0; it < n_per_thread; ++it) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        unsigned long long ma = x, y = x * 0xD6E8FEB86659FD93ull;
        y ^= y >> 32;
        unsigned long long mb = y;
        const unsigned sel = (unsigned)(x >> 40) ^ (unsigned)(y >> 11);
        const int eb = (int)(sel % 81) - 40, ea = (int)((sel / 81) % 62) - 60;
        const unsigned mode = (sel >> 14) & 7;      // bias towards nasty mantissas
        if (mode == 0) mb = Bits<T>::ones; else if (mode == 1) mb = 0; else if (mode == 2) ma = Bits<T>::ones;
        else if (mode == 3) ma = 0; else if (mode == 4) { mb = Bits<T>::ones; ma &= 0xff; }
        else if (mode == 5) { mb = Bits<T>::ones - (mb & 0xf); }
        const T a = Bits<T>::make(ea, ma, sel & (1u << 16));
        const T b = Bits<T>::make(eb, mb, sel & (1u << 17));
        const T want = Fp<T>::div(a, b);
        const T got = rlic::div_tail<T>(a, b, Fp<T>::refined_rcp(b));
        if (!Bits<T>::same(want, got)) ++local_bad;
    }
    if (local_bad) atomicAdd(bad, local_bad);
}

template <typename T> void divcheck(const char *name)
{
    unsigned long long *bad; CK(cudaMalloc(&bad, 8)); CK(cudaMemset(bad, 0, 8));
    const unsigned long long per_thread = sizeof(T) == 4 ? 40000 : 8000;
    divcheck_kernel<T><<<148 * 16, 256>>>(0xC0FFEEull, bad, per_thread);
    CK(cudaDeviceSynchronize());
    unsigned long long h_bad; CK(cudaMemcpy(&h_bad, bad, 8, cudaMemcpyDeviceToHost));
    printf("divcheck %s: %llu mismatches in %.3g samples\n", name, h_bad, (double)per_thread * 148 * 16 * 256);
    CK(cudaFree(bad));
}

// Candidate: persistent CTAs that pull tiles from per-SM lists, so that the CTAs
// resident on one SM work on adjacent tiles (a 4 x 2 block of 16 x 16 tiles) and
// share L1 lines, instead of whatever tiles the hardware scheduler hands out.
template <typename T, bool POL, typename Taps, int TW, int TH, int UNROLL, int MINB, int FLAVOR, int ADMIT, int SW, int SH>
__global__ void __launch_bounds__(TW *TH, MINB)
persistent_pass_kernel(const T *__restrict__ tex, const PackedField<T> *__restrict__ field,
                       T *__restrict__ out, const __grid_constant__ PassGeom g,
                       const __grid_constant__ Taps taps, const int ntaps, unsigned *counters, int nlists)
{
    using F = Fp<T>;
    __shared__ unsigned s_item;
    unsigned smid;
    asm("mov.u32 %0, %%smid;" : "=r"(smid));
    const int tiles_y = g.tiles_per_field / g.tiles_x;
    const int sup_x = (g.tiles_x + SW - 1) / SW, sup_y = (tiles_y + SH - 1) / SH;
    const int nsup = sup_x * sup_y;
    const long long base = g.pitch;
    tex += base;
    out += base;
    typename rlic::FieldAccess<T>::Ptr fcell = rlic::FieldAccess<T>::block(field, 0, g.field_stride) + g.pitch;
    asm volatile("" : "+l"(tex), "+l"(fcell));
    const int pitch = g.pitch;
    const int kmid = ntaps >> 1;
    int owner = (int)(smid % (unsigned)nlists);
    for (int tries = 0; tries < nlists; ++tries, owner = (owner + 1 == nlists ? 0 : owner + 1)) {
        for (;;) {
            if (threadIdx.x == 0) s_item = atomicAdd(&counters[owner], 1u);
            __syncthreads();
            const unsigned k = s_item;
            __syncthreads();
            const int sup = owner + nlists * (int)(k / (SW * SH));
            if (sup >= nsup) break;
            const int slot = (int)(k % (SW * SH));
            const int tile_x = (sup % sup_x) * SW + slot % SW;
            const int tile_y = (sup / sup_x) * SH + slot / SW;
            const int j = tile_x * TW + (int)(threadIdx.x % TW);
            const int r = tile_y * TH + (int)(threadIdx.x / TW);
            if (tile_x >= g.tiles_x || tile_y >= tiles_y || j >= g.nx || r >= g.out_rows) continue;
            const int row = g.first_row + r;
            const int at = row * pitch + j;
            T acc = F::fma(taps.get(kmid), __ldg(tex + at), T(0));
            acc = rlic::half_walk<T, POL, +1, Taps, int, UNROLL, FLAVOR, ADMIT>(acc, at, tex, fcell, taps, kmid + 1, ntaps, pitch, (int)g.field_stride);
            acc = rlic::half_walk<T, POL, -1, Taps, int, UNROLL, FLAVOR, ADMIT>(acc, at, tex, fcell, taps, kmid - 1, -1, pitch, (int)g.field_stride);
            out[at] = acc;
            if (j == g.j_above_to) out[row * pitch + g.nx] = acc;
            if (j == g.j_below_to) out[row * pitch - 1] = acc;
            if (g.lo_wall && row == g.i_below_to) out[-pitch + j] = acc;
            if (g.hi_wall && row == g.i_above_to) out[g.rows * pitch + j] = acc;
        }
    }
}

// (the gather-ceiling kernel lives in lic_walk.cuh: the library exposes it as a measurement entry point)
using rlic::gather_ceiling_kernel;

// lab20 (round 2): the north star's "texture tiles staged in shared memory (TMA) with a halo of
// half the kernel length", built to be measured rather than argued about.  f32 velocity.
// One CTA per TW x TH tile of output pixels.  A CTA whose tile plus a halo of H = ntaps / 2 cells
// lies inside the image stages that (TW + 2H) x (TH + 2H) window of the texture with ONE TMA
// box copy (cp.async.bulk.tensor.2d, completion on an mbarrier) and its walkers read the
// texture from shared memory: a walker takes at most H steps per direction, one cell each, so
// it cannot leave the window, and it cannot meet a wall cell.  CTAs near the image border
// (3 % of them at 4096^2) run the global-memory walk.  The field is gathered from global
// memory in both cases.  TMA wants row strides that are multiples of 16 bytes; the padded
// layout's pitch of nx + 2 cells is not one for nx = 4096, so the window is read from a copy of
// the padded texture with its pitch rounded up to a multiple of four cells (a product version
// would round the padded layout's pitch instead).
// The step is walk_step's flavour 2 with one more piece of state: the walker's byte offset in
// the shared window, moved by the same unit step (+-4 bytes or +-row bytes).
template <int TW, int TH, int H> struct StagedShape {
    static constexpr int SW = TW + 2 * H, SH = TH + 2 * H;
    static constexpr unsigned bytes = SW * SH * sizeof(float);
};

template <int DIR, typename Taps>
__device__ __forceinline__ float staged_half_walk(float acc, int at, unsigned sat, const float *__restrict__ tex,
                                                  const float4 *__restrict__ field, const char *smem_tile,
                                                  const Taps &taps, int k, const int k_end, const int pitch,
                                                  const int srow_bytes, const bool staged)
{
    using F = Fp<float>;
    using S = rlic::SignWord<float>;
    constexpr bool kNeg = DIR < 0;
    float fx = 0.5f, fy = 0.5f;
    const int kb_end = k_end * 4;
#pragma unroll 4
    for (int kb = k * 4; kb != kb_end; kb += DIR * 4) {
        PackedField<float> p = rlic::FieldAccess<float>::load(field, at, 0);
        const float sgx = F::with_sign_of(1.0f, p.u), sgy = F::with_sign_of(1.0f, p.v);
        const float ax = kNeg ? F::sub(1.0f, sgx) : F::add(1.0f, sgx);
        const float ay = kNeg ? F::sub(1.0f, sgy) : F::add(1.0f, sgy);
        const float remx = F::fma(ax, F::sub(0.5f, fx), fx), remy = F::fma(ay, F::sub(0.5f, fy), fy);
        const float tx = F::abs(rlic::div_tail(remx, p.u, p.ru)), ty = F::abs(rlic::div_tail(remy, p.v, p.rv));
        const bool x_first = tx < ty;
        const float fy_if_x = F::fma(tx, kNeg ? -p.v : p.v, fy), fx_if_y = F::fma(ty, kNeg ? -p.u : p.u, fx);
        const int unit = x_first ? S::unit_step_of_two(ax) : S::unit_step_of_two(ay);
        int at2 = at + unit * (x_first ? 1 : pitch);
        unsigned sat2 = sat + (unsigned)(unit * (x_first ? 4 : srow_bytes));
        float fx2 = x_first ? F::fma(ax, -0.5f, 1.0f) : fx_if_y;
        float fy2 = x_first ? fy_if_x : F::fma(ay, -0.5f, 1.0f);
        if (!rlic::fast_path_admits<float, 3>(remx, remy, p.ru)) {
            float pu = p.u, pv = p.v;
            if (rlic::is_sentinel(p)) {               // border CTAs only
                at += rlic::Sentinel<float>::decode<int>(p);
                p = rlic::FieldAccess<float>::load(field, at, 0);
                pu = p.u; pv = p.v;
            }
            if (kNeg) { pu = -pu; pv = -pv; }
            if (pu != pu || pv != pv)
                break;
            const rlic::Moved<float, int> m = rlic::generic_step<float, int, true>(pu, pv, at, fx, fy, pitch);
            const int d = m.at - at;                  // 0, +-1 or +-pitch
            sat2 = sat + (unsigned)(d == 0 ? 0 : (d == 1 ? 4 : (d == -1 ? -4 : (d > 0 ? srow_bytes : -srow_bytes))));
            at2 = m.at; fx2 = m.fx; fy2 = m.fy;
        }
        at = at2; sat = sat2; fx = fx2; fy = fy2;
        const float t = staged ? *reinterpret_cast<const float *>(smem_tile + sat) : __ldg(tex + at);
        acc = F::fma(taps.at_byte(kb), t, acc);
    }
    return acc;
}

template <int TW, int TH, int H, typename Taps>
__global__ void __launch_bounds__(TW *TH, (TW * TH >= 1024 ? 2 : (TW * TH >= 512 ? 3 : 8)))
staged_pass_kernel(const float *__restrict__ tex, const PackedField<float> *__restrict__ field,
                   float *__restrict__ out, const __grid_constant__ PassGeom g,
                   const __grid_constant__ Taps taps, const int ntaps,
                   const __grid_constant__ CUtensorMap tmap)
{
    using Shape = StagedShape<TW, TH, H>;
    extern __shared__ __align__(128) char smem_tile[];
    __shared__ __align__(8) unsigned long long bar;
    const unsigned tile_y = blockIdx.x / (unsigned)g.tiles_x, tile_x = blockIdx.x - tile_y * (unsigned)g.tiles_x;
    const int x0 = (int)tile_x * TW, y0 = (int)tile_y * TH;
    // the window [x0 - H, x0 + TW + H) x [y0 - H, y0 + TH + H) must lie inside the image
    const bool staged = ntaps / 2 <= H && x0 >= H && y0 >= H && x0 + TW + H <= g.nx && y0 + TH + H <= g.rows;
    const unsigned bar_addr = (unsigned)__cvta_generic_to_shared(&bar);
    if (staged) {
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_tile);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr), "r"(Shape::bytes)
                         : "memory");
            // coordinates: {column, buffer row}; buffer row = image row + 1 (one guard row)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
                         "[%0], [%1, {%2, %3}], [%4];"
                         ::"r"(dst), "l"(&tmap), "r"(x0 - H), "r"(y0 - H + 1), "r"(bar_addr)
                         : "memory");
        }
    }
    const int j = x0 + (int)(threadIdx.x % TW), r = y0 + (int)(threadIdx.x / TW);
    const bool live = j < g.nx && r < g.out_rows;
    tex += g.pitch;
    out += g.pitch;
    const float4 *fcell = reinterpret_cast<const float4 *>(field) + g.pitch;
    asm volatile("" : "+l"(tex), "+l"(fcell));
    const int pitch = g.pitch, kmid = ntaps >> 1;
    const int at = r * pitch + j;
    if (staged) {
        unsigned done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(bar_addr) : "memory");
    }
    if (!live)
        return;
    const unsigned sat = (unsigned)((((int)(threadIdx.x / TW) + H) * Shape::SW + (int)(threadIdx.x % TW) + H) * 4);
    const float centre = staged ? *reinterpret_cast<const float *>(smem_tile + sat) : __ldg(tex + at);
    float acc = Fp<float>::fma(taps.get(kmid), centre, 0.0f);
    acc = staged_half_walk<+1>(acc, at, sat, tex, fcell, smem_tile, taps, kmid + 1, ntaps, pitch, Shape::SW * 4, staged);
    acc = staged_half_walk<-1>(acc, at, sat, tex, fcell, smem_tile, taps, kmid - 1, -1, pitch, Shape::SW * 4, staged);
    out[at] = acc;
    if (j == g.j_above_to) out[r * pitch + g.nx] = acc;
    if (j == g.j_below_to) out[r * pitch - 1] = acc;
    if (g.lo_wall && r == g.i_below_to) out[-pitch + j] = acc;
    if (g.hi_wall && r == g.i_above_to) out[g.rows * pitch + j] = acc;
}

struct Result { std::string name; float ms; bool same; int regs; };

template <typename T>
void run_type(const char *tname, int n, int L, const char *only)
{
    // LAB_REPS=1 for a run under ncu
    const int reps = getenv("LAB_REPS") ? atoi(getenv("LAB_REPS")) : 5;
    const size_t count = (size_t)n * n;
    PassGeom g{};
    g.nx = n; g.pitch = n + 2; g.rows = n; g.field_stride = rlic::padded_cells(n, n);
    g.j_below_to = 0; g.j_above_to = n - 1; g.i_below_to = 0; g.i_above_to = n - 1;   // closed walls
    g.lo_wall = g.hi_wall = 1;
    g.first_row = 0; g.out_rows = n;
    const size_t cells = (size_t)g.field_stride;

    T *tex, *u, *v, *ptex, *ref, *out; PackedField<T> *field;
    CK(cudaMalloc(&tex, count * sizeof(T))); CK(cudaMalloc(&u, count * sizeof(T)));
    CK(cudaMalloc(&v, count * sizeof(T)));
    CK(cudaMalloc(&ptex, cells * sizeof(T))); CK(cudaMalloc(&ref, cells * sizeof(T)));
    CK(cudaMalloc(&out, cells * sizeof(T)));
    CK(cudaMalloc(&field, cells * sizeof(PackedField<T>)));
    fill_inputs<T><<<(unsigned)((count + 255) / 256), 256>>>(tex, u, v, n, g_split_field);
    rlic::pack_field_kernel<T><<<148 * 16, 256>>>(u, v, field, g, 0, n, 1);
    rlic::pad_texture_kernel<T><<<148 * 16, 256>>>(tex, ptex, g, 0, n, 1, nullptr);
    CK(cudaDeviceSynchronize());

    using PT = rlic::ParamTaps<T, rlic::kParamTapBytes / (int)sizeof(T)>;
    PT taps{};
    for (int k = 0; k < L; ++k) taps.w[k] = (T)(1.0 - fabs(-1.0 + 2.0 * k / (L - 1)));

    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    std::vector<Result> results;
    std::vector<T> h_ref(cells), h_out(cells);

    g.tiles_x = (n + rlic::kTileW - 1) / rlic::kTileW;
    g.tiles_per_field = g.tiles_x * ((n + rlic::kTileH - 1) / rlic::kTileH);
    {
        auto k = rlic::lic_pass_kernel<T, false, PT, int>;
        float best = 1e9;
        CK(cudaMemset(ref, 0, cells * sizeof(T)));
        for (int r = 0; r < reps + 1; ++r) {
            CK(cudaEventRecord(e0));
            k<<<g.tiles_per_field, rlic::kThreads>>>(ptex, field, ref, g, taps, L, rlic::PathPlanes{});
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (r) best = fminf(best, ms);
        }
        CK(cudaGetLastError());
        cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k));
        results.push_back({"shipped", best, true, fa.numRegs});   // always listed: the yardstick of the run
        CK(cudaMemcpy(h_ref.data(), ref, cells * sizeof(T), cudaMemcpyDeviceToHost));
    }
#define CAND(NAME, TW, TH, UNROLL, MINB, FLAVOR, ADMIT) do { \
        if (only && !strstr(NAME, only)) break; \
        auto k = rlic::lic_pass_kernel<T, false, PT, int, TW, TH, UNROLL, MINB, FLAVOR, ADMIT>; \
        PassGeom gc = g; \
        gc.tiles_x = (n + TW - 1) / TW; \
        gc.tiles_per_field = gc.tiles_x * ((n + TH - 1) / TH); \
        float best = 1e9; \
        CK(cudaMemset(out, 0, cells * sizeof(T))); \
        for (int r = 0; r < reps + 1; ++r) { \
            CK(cudaEventRecord(e0)); \
            k<<<gc.tiles_per_field, TW * TH>>>(ptex, field, out, gc, taps, L, rlic::PathPlanes{}); \
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); \
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (r) best = fminf(best, ms); \
        } \
        CK(cudaGetLastError()); \
        CK(cudaMemcpy(h_out.data(), out, cells * sizeof(T), cudaMemcpyDeviceToHost)); \
        bool same = memcmp(h_out.data(), h_ref.data(), cells * sizeof(T)) == 0; \
        cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k)); \
        results.push_back({NAME, best, same, fa.numRegs}); \
    } while (0)
#define CANDP(NAME, TW, TH, UNROLL, MINB, FLAVOR, ADMIT) do { \
        if (only && !strstr(NAME, only)) break; \
        auto k = rlic::lic_pass_kernel<T, true, PT, int, TW, TH, UNROLL, MINB, FLAVOR, ADMIT>; \
        float best = 1e9; \
        for (int r = 0; r < reps + 1; ++r) { \
            CK(cudaEventRecord(e0)); \
            k<<<g.tiles_per_field, TW * TH>>>(ptex, field, out, g, taps, L, rlic::PathPlanes{}); \
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); \
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (r) best = fminf(best, ms); \
        } \
        CK(cudaGetLastError()); \
        cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k)); \
        results.push_back({NAME, best, true, fa.numRegs}); \
    } while (0)
    CANDP("pol u2 b6 f0 a2", 16, 16, 2, 6, 0, 2);
    CANDP("pol u2 b5 f0 a2", 16, 16, 2, 5, 0, 2);
    CANDP("pol u2 b4 f0 a2", 16, 16, 2, 4, 0, 2);
    CANDP("pol u2 b3 f0 a2", 16, 16, 2, 3, 0, 2);
    CANDP("pol u4 b8 f1 a3", 16, 16, 4, 8, 1, 3);
    CANDP("pol u2 b8 f0 a2", 16, 16, 2, 8, 0, 2);
    CANDP("pol u2 b6 f1 a3", 16, 16, 2, 6, 1, 3);
    CANDP("pol u2 b8 f1 a3", 16, 16, 2, 8, 1, 3);
    {
        unsigned *counters; CK(cudaMalloc(&counters, 1024 * sizeof(unsigned)));
        int nsm = 0; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
#define CANDPERS(NAME, MINB, UNROLL, FLAVOR, ADMIT, SW, SH) do { \
        if (only && !strstr(NAME, only)) break; \
        auto k = persistent_pass_kernel<T, false, PT, 16, 16, UNROLL, MINB, FLAVOR, ADMIT, SW, SH>; \
        float best = 1e9; \
        CK(cudaMemset(out, 0, cells * sizeof(T))); \
        for (int r = 0; r < reps + 1; ++r) { \
            CK(cudaMemset(counters, 0, 1024 * sizeof(unsigned))); \
            CK(cudaEventRecord(e0)); \
            k<<<nsm * MINB, 256>>>(ptex, field, out, g, taps, L, counters, nsm); \
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); \
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (r) best = fminf(best, ms); \
        } \
        CK(cudaGetLastError()); \
        CK(cudaMemcpy(h_out.data(), out, cells * sizeof(T), cudaMemcpyDeviceToHost)); \
        bool same = memcmp(h_out.data(), h_ref.data(), cells * sizeof(T)) == 0; \
        cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k)); \
        results.push_back({NAME, best, same, fa.numRegs}); \
    } while (0)
        if (sizeof(T) == 4) {
            CANDPERS("persist b8 4x2", 8, 4, 1, 3, 4, 2);
            CANDPERS("persist b8 2x4", 8, 4, 1, 3, 2, 4);
            CANDPERS("persist b8 8x1", 8, 4, 1, 3, 8, 1);
            CANDPERS("persist b8 1x8", 8, 4, 1, 3, 1, 8);
            CANDPERS("persist b8 4x4", 8, 4, 1, 3, 4, 4);
        } else {
            CANDPERS("persist b6 3x2", 6, 2, 0, 2, 3, 2);
            CANDPERS("persist b6 2x3", 6, 2, 0, 2, 2, 3);
            CANDPERS("persist b6 6x1", 6, 2, 0, 2, 6, 1);
        }
        CK(cudaFree(counters));
    }
#define CANDCEIL(NAME, DEP) do { \
        if (only && !strstr(NAME, only)) break; \
        auto k = gather_ceiling_kernel<T, DEP>; \
        float best = 1e9; \
        for (int r = 0; r < reps + 1; ++r) { \
            CK(cudaEventRecord(e0)); \
            k<<<g.tiles_per_field, 256>>>(ptex, field, out, g, taps, L); \
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); \
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (r) best = fminf(best, ms); \
        } \
        CK(cudaGetLastError()); \
        cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k)); \
        results.push_back({NAME, best, true, fa.numRegs}); \
    } while (0)
    CANDCEIL("ceiling: gathers only, free addresses", false);
    CANDCEIL("ceiling: gathers only, dependent addresses", true);
    // grouped walk (lic_walk.cuh: walk_step / half_walk_grouped).  WALK bits: 1 grouped,
    // 2 pitch pinned in a uniform register, 4 the constant 1.0 pinned, 8 loop control on the
    // byte offset.  "tuned" is what rlic_b200_set_walk(RLIC_B200_WALK_GROUPED) launches.
#define CANDW(NAME, POL, UNROLL, MINB, FLAVOR, ADMIT, WALK) do { \
        if (only && !strstr(NAME, only)) break; \
        auto kref = rlic::lic_pass_kernel<T, POL, PT, int>; \
        auto k = rlic::lic_pass_kernel<T, POL, PT, int, 16, 16, UNROLL, MINB, FLAVOR, ADMIT, true, WALK>; \
        CK(cudaMemset(ref, 0, cells * sizeof(T))); \
        kref<<<g.tiles_per_field, 256>>>(ptex, field, ref, g, taps, L, rlic::PathPlanes{}); \
        CK(cudaMemcpy(h_ref.data(), ref, cells * sizeof(T), cudaMemcpyDeviceToHost)); \
        float best = 1e9; \
        CK(cudaMemset(out, 0, cells * sizeof(T))); \
        for (int r = 0; r < reps + 1; ++r) { \
            CK(cudaEventRecord(e0)); \
            k<<<g.tiles_per_field, 256>>>(ptex, field, out, g, taps, L, rlic::PathPlanes{}); \
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); \
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (r) best = fminf(best, ms); \
        } \
        CK(cudaGetLastError()); \
        CK(cudaMemcpy(h_out.data(), out, cells * sizeof(T), cudaMemcpyDeviceToHost)); \
        bool same = memcmp(h_out.data(), h_ref.data(), cells * sizeof(T)) == 0; \
        cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k)); \
        results.push_back({NAME, best, same, fa.numRegs}); \
    } while (0)
    {
        using TV = rlic::Tune<T, false>;
        using TP = rlic::Tune<T, true>;
        CANDW("grouped tuned vel", false, TV::walk_unroll, TV::walk_min_blocks, TV::walk_flavor, TV::walk_admit, TV::walk);
        CANDW("grouped tuned pol", true, TP::walk_unroll, TP::walk_min_blocks, TP::walk_flavor, TP::walk_admit, TP::walk);
        if constexpr (sizeof(T) == 4) {
            CANDW("grouped vel w1 f1", false, 4, 8, 1, 3, 1);
            CANDW("grouped vel w1 f2", false, 4, 8, 2, 3, 1);
            CANDW("grouped vel w3 f2", false, 4, 8, 2, 3, 3);
            CANDW("grouped vel w5 f2", false, 4, 8, 2, 3, 5);
            CANDW("grouped vel w7 f2", false, 4, 8, 2, 3, 7);
            CANDW("grouped vel w15 f2", false, 4, 8, 2, 3, 15);
            CANDW("grouped vel w7 f2 u2", false, 2, 8, 2, 3, 7);
            CANDW("grouped vel w7 f2 u8", false, 8, 8, 2, 3, 7);
            CANDW("grouped vel w7 f3", false, 4, 8, 3, 3, 7);
            CANDW("grouped vel w1 f0", false, 4, 8, 0, 3, 1);
            CANDW("grouped vel w7 f2 a2", false, 4, 8, 2, 2, 7);
            // lab19 (round 2): the packed-pair step (flavour 4: FADD2 / FMUL2 / FFMA2) and the
            // NaN-propagating three-input minimum in the admission test (admit 4)
            CANDW("packed vel w1 f4 a3", false, 4, 8, 4, 3, 1);
            CANDW("packed vel w1 f4 a4", false, 4, 8, 4, 4, 1);
            CANDW("packed vel w1 f4 a4 u8", false, 8, 8, 4, 4, 1);
            CANDW("packed vel w1 f4 a4 u2", false, 2, 8, 4, 4, 1);
            CANDW("packed vel w3 f4 a4", false, 4, 8, 4, 4, 3);
            CANDW("packed vel w5 f4 a4", false, 4, 8, 4, 4, 5);
            CANDW("packed vel w9 f4 a4", false, 4, 8, 4, 4, 9);
            CANDW("packed vel w7 f2 a4 (scalar, admit 4)", false, 4, 8, 2, 4, 7);
            CANDW("packed vel w7 f2 a4 u8 (scalar)", false, 8, 8, 2, 4, 7);
            CANDW("packed vel w5 f2 a4 (scalar)", false, 4, 8, 2, 4, 5);
            CANDW("packed vel w3 f2 a4 (scalar)", false, 4, 8, 2, 4, 3);
            CANDW("packed vel w7 f2 a4 b7 (scalar)", false, 4, 7, 2, 4, 7);
            CANDW("packed vel w7 f2 a4 b6 (scalar, 40 regs)", false, 4, 6, 2, 4, 7);
            CANDW("packed vel w7 f1 a4 (scalar)", false, 4, 8, 1, 4, 7);
            CANDW("packed vel w7 f0 a4 (scalar)", false, 4, 8, 0, 4, 7);
            CANDW("packed pol w1 f4 a4", true, 4, 8, 4, 4, 1);
            CANDW("packed pol w1 f0 a4 (scalar, admit 4)", true, 4, 8, 0, 4, 1);
            CANDW("grouped pol w1 f0", true, 4, 8, 0, 3, 1);
            CANDW("grouped pol w9 f0", true, 4, 8, 0, 3, 9);
            CANDW("grouped pol w1 f2", true, 4, 8, 2, 3, 1);
            CANDW("grouped pol w7 f2", true, 4, 8, 2, 3, 7);
        } else {
            CANDW("grouped vel w1 f0", false, 2, 6, 0, 2, 1);
            CANDW("grouped vel w9 f0", false, 2, 6, 0, 2, 9);
            CANDW("grouped vel w11 f0", false, 2, 6, 0, 2, 11);
            CANDW("grouped vel w1 f2", false, 2, 6, 2, 2, 1);
            CANDW("grouped vel w5 f2", false, 2, 6, 2, 2, 5);
            CANDW("grouped vel w11 f0 u4", false, 4, 6, 0, 2, 11);
            CANDW("grouped pol w1 f0", true, 2, 5, 0, 2, 1);
            CANDW("grouped pol w9 f0", true, 2, 5, 0, 2, 9);
            CANDW("grouped pol w11 f0", true, 2, 5, 0, 2, 11);
            CANDW("grouped pol w1 f2", true, 2, 5, 2, 2, 1);
            CANDW("grouped pol w9 f0 b4", true, 2, 4, 0, 2, 9);
            // 48 registers (5 CTAs): the lowest static counts of the f64 sweeps
            CANDW("grouped vel w5 f2 b5", false, 2, 5, 2, 2, 5);
            CANDW("grouped vel w1 f0 b5", false, 2, 5, 0, 2, 1);
            CANDW("grouped vel w9 f0 u4 b5", false, 4, 5, 0, 2, 9);
            CANDW("grouped pol w1 f0 b4", true, 2, 4, 0, 2, 1);
            CANDW("grouped pol w9 f0 u4 b4", true, 4, 4, 0, 2, 9);
        }
        // the later CAND()s compare with the velocity result of the shipped kernel
        auto kref = rlic::lic_pass_kernel<T, false, PT, int>;
        kref<<<g.tiles_per_field, 256>>>(ptex, field, ref, g, taps, L, rlic::PathPlanes{});
        CK(cudaMemcpy(h_ref.data(), ref, cells * sizeof(T), cudaMemcpyDeviceToHost));
    }
    // lab20: TMA-staged texture window (f32, this image only: closed walls, 65 taps -> H = 32)
    if constexpr (sizeof(T) == 4) if (L / 2 <= 32 && (!only || strstr("staged", only))) {
        const int pitch4 = (n + 2 + 3) / 4 * 4;                  // row stride a multiple of 16 bytes
        float *ttex;
        CK(cudaMalloc(&ttex, (size_t)(n + 2) * pitch4 * sizeof(float)));
        CK(cudaMemset(ttex, 0, (size_t)(n + 2) * pitch4 * sizeof(float)));
        CK(cudaMemcpy2D(ttex, (size_t)pitch4 * 4, ptex, (size_t)(n + 2) * 4, (size_t)(n + 2) * 4, n + 2,
                        cudaMemcpyDeviceToDevice));
        PFN_cuTensorMapEncodeTiled encode = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qres));
#define CANDSTAGED(NAME, TW, TH) do { \
        using Shape = StagedShape<TW, TH, 32>; \
        CUtensorMap tmap; \
        const cuuint64_t gdim[2] = {(cuuint64_t)pitch4, (cuuint64_t)(n + 2)}; \
        const cuuint64_t gstr[1] = {(cuuint64_t)pitch4 * 4}; \
        const cuuint32_t box[2] = {Shape::SW, Shape::SH}; \
        const cuuint32_t estr[2] = {1, 1}; \
        CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ttex, gdim, gstr, box, estr, \
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, \
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); \
        if (cr != CUDA_SUCCESS) { fprintf(stderr, "cuTensorMapEncodeTiled failed: %d\n", (int)cr); break; } \
        auto k = staged_pass_kernel<TW, TH, 32, PT>; \
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Shape::bytes)); \
        PassGeom gc = g; \
        gc.tiles_x = (n + TW - 1) / TW; \
        gc.tiles_per_field = gc.tiles_x * ((n + TH - 1) / TH); \
        float best = 1e9; \
        CK(cudaMemset(out, 0, cells * sizeof(T))); \
        for (int r = 0; r < reps + 1; ++r) { \
            CK(cudaEventRecord(e0)); \
            k<<<gc.tiles_per_field, TW * TH, Shape::bytes>>>((const float *)ptex, (const PackedField<float> *)field, \
                                                             (float *)out, gc, *(const PT *)&taps, L, tmap); \
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); \
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (r) best = fminf(best, ms); \
        } \
        CK(cudaGetLastError()); \
        CK(cudaMemcpy(h_out.data(), out, cells * sizeof(T), cudaMemcpyDeviceToHost)); \
        bool same = memcmp(h_out.data(), h_ref.data(), cells * sizeof(T)) == 0; \
        cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k)); \
        results.push_back({NAME, best, same, fa.numRegs}); \
    } while (0)
        CANDSTAGED("staged TMA window 16x16 (80x80 f32 = 25.6 KB)", 16, 16);
        CANDSTAGED("staged TMA window 32x16 (96x80 = 30.7 KB)", 32, 16);
        CANDSTAGED("staged TMA window 32x32 (96x96 = 36.9 KB)", 32, 32);
        CK(cudaFree(ttex));
    }
    //    name              TW  TH  unroll minblocks flavor admit
    CAND("u2 b8 f1 a3", 16, 16, 2, 8, 1, 3);
    CAND("u4 b8 f1 a3", 16, 16, 4, 8, 1, 3);
    CAND("u2 b6 f1 a3", 16, 16, 2, 6, 1, 3);
    CAND("u2 b8 f1 a2", 16, 16, 2, 8, 1, 2);
    CAND("u2 b8 f0 a2", 16, 16, 2, 8, 0, 2);
    CAND("u2 b6 f0 a2", 16, 16, 2, 6, 0, 2);
    CAND("u2 b6 f0 a0", 16, 16, 2, 6, 0, 0);
    CAND("u2 b5 f0 a2", 16, 16, 2, 5, 0, 2);
    CAND("u4 b6 f0 a2", 16, 16, 4, 6, 0, 2);
    CAND("u1 b6 f0 a2", 16, 16, 1, 6, 0, 2);
    CAND("32x8 u2 b8 f1 a3", 32, 8, 2, 8, 1, 3);

    const double steps = (double)count * (L - 1);
    const double bytes = (double)count * (3 * (L - 1) + 2) * sizeof(T);
    printf("%s %dx%d, %d taps\n%-44s %8s %10s %8s %6s %5s\n", tname, n, n, L, "variant", "ms", "Gsteps/s",
           "GB/s", "same", "regs");
    for (auto &r : results)
        printf("%-44s %8.3f %10.1f %8.0f %6s %5d\n", r.name.c_str(), r.ms, steps / r.ms / 1e6,
               bytes / r.ms / 1e6, r.same ? "yes" : "NO", r.regs);
    CK(cudaFree(tex)); CK(cudaFree(u)); CK(cudaFree(v)); CK(cudaFree(ptex)); CK(cudaFree(ref));
    CK(cudaFree(out)); CK(cudaFree(field));
}

int main(int argc, char **argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 4096;
    const int L = argc > 2 ? atoi(argv[2]) : 65;
    const char *only = argc > 3 ? argv[3] : nullptr;
    g_split_field = argc > 4 ? atoi(argv[4]) : 0;
    if (!only) {
        divcheck<float>("f32");
        divcheck<double>("f64");
    }
    run_type<float>("f32", n, L, only);
    run_type<double>("f64", n / 2, 2 * L - 1, only);
    return 0;
}
