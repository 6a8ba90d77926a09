// Developer tool (not part of librlic_b200.so): the recording walk and the replay kernel
// (PathPlanes in lic_walk.cuh) on the headline workload -- launch-bound / tile-shape variants
// timed against the plain walk, every replay compared bit for bit with the walk's output.
// Build: tools/build_lab.sh replay    Run on a B200: tools/replay_lab [n] [taps] [name-filter]
//
//   record   lic_pass_kernel<..., REC = true> at 8 / 7 / 6 CTAs per SM (the plain walk sits at
//            the 32-register cap of 8; the record keeps three more words)
//   replay   lic_replay_kernel over tile shapes and register budgets: the replay has no
//            dependent chain, so what it wants is loads in flight (registers) and few L1
//            wavefronts per gather (a warp's 32 pixels on as few rows as possible)
//   ceiling  the replay's loads alone (gather_ceiling_kernel<T, false, false>)
#include "../rlic_b200/csrc/lic_walk.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
    fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

using rlic::PackedField;
using rlic::PassGeom;

template <typename T>
__global__ void fill_inputs(T *tex, T *u, T *v, int n, int second)
{
    long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (long long)n * n) return;
    int i = (int)(p / n), j = (int)(p % n);
    unsigned h = (unsigned)p * 2654435761u + (unsigned)second * 40503u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    tex[p] = (T)((h >> 8) * (1.0 / 16777216.0));
    u[p] = (T)(-(-1.0 + 2.0 * i / (n - 1)));
    v[p] = (T)(-1.0 + 2.0 * j / (n - 1));
}

struct Result { std::string name; float ms; bool same; int regs; };

template <typename T>
void run_type(const char *tname, int n, int L, const char *only)
{
    const int reps = getenv("LAB_REPS") ? atoi(getenv("LAB_REPS")) : 5;
    const size_t count = (size_t)n * n;
    PassGeom g{};
    g.nx = n; g.pitch = n + 2; g.rows = n; g.field_stride = rlic::padded_cells(n, n);
    g.j_below_to = 0; g.j_above_to = n - 1; g.i_below_to = 0; g.i_above_to = n - 1;   // closed walls
    g.lo_wall = g.hi_wall = 1;
    g.first_row = 0; g.out_rows = n;
    const size_t cells = (size_t)g.field_stride;
    const int groups = rlic::path_groups_fwd(L) + rlic::path_groups_bwd(L);
    const size_t rec_words = (size_t)groups * rlic::kPlanesPerGroup * cells;

    T *tex, *u, *v, *ptex, *ptex2, *ref, *ref2, *out; PackedField<T> *field; uint4 *rec;
    CK(cudaMalloc(&tex, count * sizeof(T))); CK(cudaMalloc(&u, count * sizeof(T)));
    CK(cudaMalloc(&v, count * sizeof(T)));
    CK(cudaMalloc(&ptex, cells * sizeof(T))); CK(cudaMalloc(&ptex2, cells * sizeof(T)));
    CK(cudaMalloc(&ref, cells * sizeof(T))); CK(cudaMalloc(&ref2, cells * sizeof(T)));
    CK(cudaMalloc(&out, cells * sizeof(T)));
    CK(cudaMalloc(&field, cells * sizeof(PackedField<T>)));
    CK(cudaMalloc(&rec, rec_words * sizeof(unsigned)));
    fill_inputs<T><<<(unsigned)((count + 255) / 256), 256>>>(tex, u, v, n, 0);
    rlic::pack_field_kernel<T><<<148 * 16, 256>>>(u, v, field, g, 0, n, 1);
    rlic::pad_texture_kernel<T><<<148 * 16, 256>>>(tex, ptex, g, 0, n, 1, nullptr);
    // a second texture: the record made on the first must replay on any other
    fill_inputs<T><<<(unsigned)((count + 255) / 256), 256>>>(tex, u, v, n, 1);
    rlic::pad_texture_kernel<T><<<148 * 16, 256>>>(tex, ptex2, g, 0, n, 1, nullptr);
    CK(cudaDeviceSynchronize());

    using PT = rlic::ParamTaps<T, rlic::kParamTapBytes / (int)sizeof(T)>;
    using ST = rlic::StepTaps<T, rlic::kStepTapsPerHalf<T>>;
    using Tn = rlic::Tune<T, false>;
    PT taps{};
    ST steps{};
    for (int k = 0; k < L; ++k) taps.w[k] = (T)(1.0 - fabs(-1.0 + 2.0 * k / (L - 1)));
    const int kmid = L / 2;
    steps.centre = taps.w[kmid];
    for (int k = kmid + 1; k < L; ++k) steps.fwd[k - kmid - 1] = taps.w[k];
    for (int k = kmid - 1; k >= 0; --k) steps.bwd[kmid - 1 - k] = taps.w[k];

    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    std::vector<Result> results;
    std::vector<T> h_ref(cells), h_ref2(cells), h_out(cells);
    g.tiles_x = (n + rlic::kTileW - 1) / rlic::kTileW;
    g.tiles_per_field = g.tiles_x * ((n + rlic::kTileH - 1) / rlic::kTileH);
    const rlic::PathPlanes none{nullptr, 0, 0};
    const rlic::PathPlanes planes{rec, (long long)cells, rlic::path_groups_fwd(L)};

#define TIME(LAUNCH, BEST) do { \
        BEST = 1e9; \
        for (int r = 0; r < reps + 1; ++r) { \
            CK(cudaEventRecord(e0)); \
            LAUNCH; \
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); \
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (r) BEST = fminf(BEST, ms); \
        } \
        CK(cudaGetLastError()); \
    } while (0)

    // the plain walk (what every pass cost before): the yardstick and the reference output
    {
        auto k = rlic::lic_pass_kernel<T, false, PT, int, rlic::kTileW, rlic::kTileH, Tn::walk_unroll,
                                       Tn::walk_min_blocks, Tn::walk_flavor, Tn::walk_admit, true, Tn::walk>;
        float best;
        CK(cudaMemset(ref, 0, cells * sizeof(T)));
        TIME((k<<<g.tiles_per_field, rlic::kThreads>>>(ptex, field, ref, g, taps, L, none)), best);
        cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k));
        results.push_back({"walk (no record)", best, true, fa.numRegs});
        CK(cudaMemcpy(h_ref.data(), ref, cells * sizeof(T), cudaMemcpyDeviceToHost));
        CK(cudaMemset(ref2, 0, cells * sizeof(T)));
        k<<<g.tiles_per_field, rlic::kThreads>>>(ptex2, field, ref2, g, taps, L, none);
        CK(cudaMemcpy(h_ref2.data(), ref2, cells * sizeof(T), cudaMemcpyDeviceToHost));
    }
#define RECORD(NAME, UNROLL, MINB) do { \
        if (only && !strstr(NAME, only)) break; \
        auto k = rlic::lic_pass_kernel<T, false, PT, int, rlic::kTileW, rlic::kTileH, UNROLL, MINB, \
                                       Tn::walk_flavor, Tn::walk_admit, true, Tn::walk, true>; \
        float best; \
        CK(cudaMemset(out, 0, cells * sizeof(T))); \
        CK(cudaMemset(rec, 0xA5, rec_words * sizeof(unsigned))); \
        TIME((k<<<g.tiles_per_field, rlic::kThreads>>>(ptex, field, out, g, taps, L, planes)), best); \
        CK(cudaMemcpy(h_out.data(), out, cells * sizeof(T), cudaMemcpyDeviceToHost)); \
        bool same = memcmp(h_out.data(), h_ref.data(), cells * sizeof(T)) == 0; \
        cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k)); \
        results.push_back({NAME, best, same, fa.numRegs}); \
    } while (0)
    RECORD("record u8 b8", 8, 8);
    RECORD("record u8 b7", 8, 7);
    RECORD("record u8 b6", 8, 6);
    RECORD("record u4 b8", 4, 8);
    RECORD("record u4 b6", 4, 6);
    RECORD("record u4 b5", 4, 5);
    RECORD("record u4 b4", 4, 4);
    // the record the replays read: made by the library's own choice, on the FIRST texture
    {
        auto k = rlic::lic_pass_kernel<T, false, PT, int, rlic::kTileW, rlic::kTileH, Tn::walk_unroll,
                                       (Tn::walk_min_blocks > 6 ? 6 : Tn::walk_min_blocks), Tn::walk_flavor,
                                       Tn::walk_admit, true, Tn::walk, true>;
        CK(cudaMemset(rec, 0xA5, rec_words * sizeof(unsigned)));
        k<<<g.tiles_per_field, rlic::kThreads>>>(ptex, field, out, g, taps, L, planes);
        CK(cudaDeviceSynchronize());
    }
    // replays run on the SECOND texture and must reproduce the walk over it
#define REPLAY(NAME, GROUPS, TW, TH, MINB) do { \
        if (only && !strstr(NAME, only)) break; \
        if (GROUPS != 0 && (rlic::path_groups_fwd(L) > GROUPS || rlic::path_groups_bwd(L) > GROUPS)) break; \
        auto k = rlic::lic_replay_kernel<T, ST, int, GROUPS, false, TW, TH, MINB>; \
        PassGeom gc = g; \
        gc.tiles_x = (n + TW - 1) / TW; \
        gc.tiles_per_field = gc.tiles_x * ((n + TH - 1) / TH); \
        float best; \
        CK(cudaMemset(out, 0, cells * sizeof(T))); \
        const dim3 grid((unsigned)gc.tiles_x, (unsigned)((n + TH - 1) / TH), 1); \
        TIME((k<<<grid, TW * TH>>>(ptex2, rec, out, gc, steps, L, (long long)cells, nullptr, 0)), best); \
        CK(cudaMemcpy(h_out.data(), out, cells * sizeof(T), cudaMemcpyDeviceToHost)); \
        bool same = memcmp(h_out.data(), h_ref2.data(), cells * sizeof(T)) == 0; \
        cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k)); \
        results.push_back({NAME, best, same, fa.numRegs}); \
    } while (0)
    REPLAY("replay g1 16x16 b8", 1, 16, 16, 8);
    REPLAY("replay g1 16x16 b6", 1, 16, 16, 6);
    REPLAY("replay g1 16x16 b5", 1, 16, 16, 5);
    REPLAY("replay g1 16x16 b4", 1, 16, 16, 4);
    REPLAY("replay g1 16x16 b3", 1, 16, 16, 3);
    REPLAY("replay g1 32x8 b8", 1, 32, 8, 8);
    REPLAY("replay g1 32x8 b6", 1, 32, 8, 6);
    REPLAY("replay g1 32x8 b4", 1, 32, 8, 4);
    REPLAY("replay g1 32x4 b16", 1, 32, 4, 16);
    REPLAY("replay g1 32x4 b12", 1, 32, 4, 12);
    REPLAY("replay g1 32x4 b8", 1, 32, 4, 8);
    REPLAY("replay g1 64x4 b8", 1, 64, 4, 8);
    REPLAY("replay g1 64x4 b6", 1, 64, 4, 6);
    REPLAY("replay g1 8x32 b8", 1, 8, 32, 8);
    REPLAY("replay g1 8x32 b6", 1, 8, 32, 6);
    REPLAY("replay g1 8x16 b12", 1, 8, 16, 12);
    REPLAY("replay g1 16x8 b12", 1, 16, 8, 12);
    REPLAY("replay g2 16x16 b8", 2, 16, 16, 8);
    REPLAY("replay g2 16x16 b6", 2, 16, 16, 6);
    REPLAY("replay g2 16x16 b4", 2, 16, 16, 4);
    REPLAY("replay g2 32x8 b6", 2, 32, 8, 6);
    REPLAY("replay g0 16x16 b8 (loop)", 0, 16, 16, 8);
    REPLAY("replay g0 16x16 b6 (loop)", 0, 16, 16, 6);
    // the texture window of each tile staged in shared memory (lic_replay_staged_kernel): a warp is
    // one row of 32 pixels; TH rows per CTA, PAD extra words per window row (bank spread)
#define STAGED(NAME, TH, PAD, MINB) do { \
        if (only && !strstr(NAME, only)) break; \
        if (rlic::path_groups_fwd(L) > 1 || rlic::path_groups_bwd(L) > 1) break; \
        const int hh = L / 2; \
        const size_t smem = (size_t)(32 + 2 * hh + PAD) * (TH + 2 * hh) * sizeof(T); \
        if (smem > 200 * 1024) break; \
        auto k = rlic::lic_replay_staged_kernel<T, ST, int, false, 32, TH, PAD, MINB>; \
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        PassGeom gc = g; \
        const dim3 grid((unsigned)((n + 31) / 32), (unsigned)((n + TH - 1) / TH), 1); \
        float best; \
        CK(cudaMemset(out, 0, cells * sizeof(T))); \
        TIME((k<<<grid, 32 * TH, smem>>>(ptex2, rec, out, gc, steps, L, (long long)cells, nullptr, 0)), best); \
        CK(cudaMemcpy(h_out.data(), out, cells * sizeof(T), cudaMemcpyDeviceToHost)); \
        bool same = memcmp(h_out.data(), h_ref2.data(), cells * sizeof(T)) == 0; \
        cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k)); \
        results.push_back({NAME, best, same, fa.numRegs}); \
    } while (0)
    STAGED("staged 32x32 pad0 b2", 32, 0, 2);
    STAGED("staged 32x32 pad4 b2", 32, 4, 2);
    STAGED("staged 32x32 pad8 b2", 32, 8, 2);
    STAGED("staged 32x32 pad16 b2", 32, 16, 2);
    STAGED("staged 32x32 pad1 b2", 32, 1, 2);
    STAGED("staged 32x16 pad0 b4", 16, 0, 4);
    STAGED("staged 32x16 pad4 b4", 16, 4, 4);
    STAGED("staged 32x16 pad8 b4", 16, 8, 4);
    STAGED("staged 32x16 pad16 b4", 16, 16, 4);
    STAGED("staged 32x16 pad1 b4", 16, 1, 4);
    STAGED("staged 32x16 pad8 b3", 16, 8, 3);
    STAGED("staged 32x16 pad8 b2", 16, 8, 2);
    STAGED("staged 32x8 pad8 b8", 8, 8, 8);
    STAGED("staged 32x8 pad8 b6", 8, 8, 6);
    STAGED("staged 32x8 pad0 b8", 8, 0, 8);
    STAGED("staged 32x24 pad8 b2", 24, 8, 2);
    if (!only || strstr("ceiling", only)) {
        auto k = rlic::gather_ceiling_kernel<T, false, false>;
        float best;
        TIME((k<<<g.tiles_per_field, 256>>>(ptex2, field, out, g, taps, L)), best);
        cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k));
        results.push_back({"ceiling: one texture gather per step, nothing else", best, true, fa.numRegs});
    }

    const double nsteps = (double)count * (L - 1);
    printf("%s %dx%d, %d taps (%d groups, record %.0f MB)\n%-52s %8s %10s %6s %5s\n", tname, n, n, L, groups,
           rec_words * 4.0 / 1e6, "variant", "ms", "Gsteps/s", "same", "regs");
    for (auto &r : results)
        printf("%-52s %8.3f %10.1f %6s %5d\n", r.name.c_str(), r.ms, nsteps / r.ms / 1e6, r.same ? "yes" : "NO", r.regs);
    CK(cudaFree(tex)); CK(cudaFree(u)); CK(cudaFree(v)); CK(cudaFree(ptex)); CK(cudaFree(ptex2)); CK(cudaFree(ref));
    CK(cudaFree(ref2)); CK(cudaFree(out)); CK(cudaFree(field)); CK(cudaFree(rec));
}

int main(int argc, char **argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 4096;
    const int L = argc > 2 ? atoi(argv[2]) : 65;
    const char *only = argc > 3 ? argv[3] : nullptr;
    run_type<float>("f32", n, L, only);
    run_type<double>("f64", n / 2, 2 * L - 1, only);
    return 0;
}
