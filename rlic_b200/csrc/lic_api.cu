// Host driver and C ABI of librlic_b200.so (see include/rlic_b200.h).
//
// Replaces the iteration driver and FFI shim of the reference,
// /root/reference/src/lib.rs:408-485, with a CUDA host path:
// upload -> `iterations` passes over device-resident ping-pong buffers -> download.
// There is no CPU compute path in this file; without a device every entry
// point fails.
#include "../../include/rlic_b200.h"
#include "lic_walk.cuh"
#include "lic_equalize.cuh"
#include "host_staging.h"

#include <nvtx3/nvToolsExt.h>   // header-only: ranges show up in Nsight Systems / ncu --nvtx, cost nothing otherwise

#include <atomic>
#include <chrono>
#include <climits>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>


namespace {

using rlic::PassGeom;

thread_local std::string tls_error;
thread_local int tls_device = 0;
std::atomic<int64_t> g_launches{0};
std::atomic<int> g_force_wide{0};   // testing hook: use 64-bit element indices for any size
std::atomic<int> g_small_images{1}; // testing hook: 0 keeps small passes on the one-thread-per-pixel kernel
std::atomic<int> g_replay_staging{0}; // testing hook: 0 keeps replayed passes on the kernel that gathers through L1
// The three choices a call can make (include/rlic_b200.h): which reference build to reproduce,
// which formulation of the pass kernels, how the host path orders its launches.  Each has a
// process-wide DEFAULT (the atomics: set once at start-up, e.g. from the environment) and a
// per-thread override (rlic_b200_set_thread_options; -1 = no override) that the calling thread
// sets around its own calls, so that two threads wanting different arithmetic never race on
// shared state.  A call reads its choices through the effective_*() functions on the thread
// that entered the library; the batch entry hands them to its worker threads.
std::atomic<int> g_arithmetic{RLIC_B200_ARITH_FMA_BRANCHLESS};
std::atomic<int> g_walk{RLIC_B200_WALK_GROUPED};
std::atomic<int> g_schedule{RLIC_B200_SCHEDULE_WAVEFRONT};
std::atomic<int> g_paths{RLIC_B200_PATHS_REPLAY};
struct ThreadChoices { int arithmetic = -1, schedule = -1, walk = -1, paths = -1; };
thread_local ThreadChoices tls_choices;
int effective_arithmetic()
{
    return tls_choices.arithmetic >= 0 ? tls_choices.arithmetic : g_arithmetic.load(std::memory_order_relaxed);
}
int effective_schedule()
{
    return tls_choices.schedule >= 0 ? tls_choices.schedule : g_schedule.load(std::memory_order_relaxed);
}
int effective_walk()
{
    return tls_choices.walk >= 0 ? tls_choices.walk : g_walk.load(std::memory_order_relaxed);
}
int effective_paths()
{
    return tls_choices.paths >= 0 ? tls_choices.paths : g_paths.load(std::memory_order_relaxed);
}

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    tls_error = buf;
    return code;
}

#define CUDA_TRY(expr)                                                              \
    do {                                                                            \
        cudaError_t err__ = (expr);                                                 \
        if (err__ != cudaSuccess)                                                   \
            return fail(err__ == cudaErrorNoDevice || err__ == cudaErrorInsufficientDriver \
                            ? RLIC_B200_ENODEVICE                                   \
                            : RLIC_B200_ECUDA,                                      \
                        "%s: %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, __LINE__); \
    } while (0)

// Stream and stream-ordered allocations that clean up on every exit path.
struct Stream {
    cudaStream_t s = nullptr;
    bool borrowed = false;   // from the calling thread's cache (host_stream): not destroyed here
    ~Stream() { if (s && !borrowed) cudaStreamDestroy(s); }
};

// The streams of the host entry points, kept per calling thread and device: creating and
// destroying four streams costs a small image's call several tens of microseconds each way, and
// a thread runs one host call at a time (concurrent calls come from different threads and so
// have their own).  `role` 0..3: copies, conversions, passes, downloads.
cudaError_t host_stream(int device, int role, Stream &out)
{
    struct Cache {
        std::vector<cudaStream_t> streams;   // [device * 4 + role]
        ~Cache()
        {
            for (cudaStream_t st : streams)
                if (st)
                    cudaStreamDestroy(st);   // (fails harmlessly once the runtime is gone)
        }
    };
    thread_local Cache cache;
    const size_t slot = (size_t)device * 4 + (size_t)role;
    if (slot >= cache.streams.size())
        cache.streams.resize(slot + 1, nullptr);
    if (!cache.streams[slot]) {
        cudaError_t e = cudaStreamCreateWithFlags(&cache.streams[slot], cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            cache.streams[slot] = nullptr;
            return e;
        }
    }
    out.s = cache.streams[slot];
    out.borrowed = true;
    return cudaSuccess;
}
// The library's scratch comes from PRIVATE stream-ordered pools, one per device, never from the
// device's default pool: the device entry points run next to a host framework's own allocator
// (torch), and a process-wide release threshold on the shared default pool would keep our
// scratch away from it for good.  A private pool keeps freed blocks cached up to a bounded
// threshold -- half the device's memory, or RLIC_B200_POOL_KEEP_BYTES -- so that repeated calls
// do not pay cudaMalloc every time, and returns the rest to the driver at the next
// synchronisation.
cudaError_t device_pool(int device, cudaMemPool_t *pool)
{
    static std::mutex mu;
    static std::vector<cudaMemPool_t> pools;
    std::lock_guard<std::mutex> lock(mu);
    if (device < 0)
        return cudaErrorInvalidDevice;
    if ((size_t)device >= pools.size())
        pools.resize((size_t)device + 1, nullptr);
    if (!pools[(size_t)device]) {
        cudaMemPoolProps props{};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        cudaMemPool_t created;
        cudaError_t e = cudaMemPoolCreate(&created, &props);
        if (e != cudaSuccess)
            return e;
        uint64_t keep = 0;
        if (const char *env = getenv("RLIC_B200_POOL_KEEP_BYTES")) {
            keep = strtoull(env, nullptr, 10);
        } else {
            cudaDeviceProp dp;
            if (cudaGetDeviceProperties(&dp, device) == cudaSuccess)
                keep = (uint64_t)dp.totalGlobalMem / 2;
        }
        e = cudaMemPoolSetAttribute(created, cudaMemPoolAttrReleaseThreshold, &keep);
        if (e != cudaSuccess) {
            cudaMemPoolDestroy(created);
            return e;
        }
        pools[(size_t)device] = created;
    }
    *pool = pools[(size_t)device];
    return cudaSuccess;
}

struct DeviceBuf {
    void *p = nullptr;
    cudaStream_t s = nullptr;
    ~DeviceBuf() { if (p) cudaFreeAsync(p, s); }
    // from the current device's private pool, ordered on `stream`
    cudaError_t alloc(size_t bytes, cudaStream_t stream)
    {
        s = stream;
        int device = 0;
        cudaError_t e = cudaGetDevice(&device);
        cudaMemPool_t pool = nullptr;
        if (e == cudaSuccess)
            e = device_pool(device, &pool);
        return e != cudaSuccess ? e : cudaMallocFromPoolAsync(&p, bytes ? bytes : 1, pool, stream);
    }
};

// Waits for the streams of a host call before its buffers go: declared AFTER the DeviceBufs, so
// that on every exit path -- early error returns included -- no kernel is still using a buffer
// when cudaFreeAsync hands it back on another stream.
struct StreamDrain {
    const Stream *streams[4];
    ~StreamDrain()
    {
        for (const Stream *st : streams)
            if (st && st->s)
                cudaStreamSynchronize(st->s);
    }
};

struct Walls { int x_left, x_right, y_left, y_right; };

// NVTX range over a scope (ranges around uploads / passes / halo / downloads)
struct Range {
    explicit Range(const char *name) { nvtxRangePushA(name); }
    ~Range() { nvtxRangePop(); }
};

cudaError_t use_device(int device) { return cudaSetDevice(device); }

// device entry points run on whatever device is current for the caller
cudaError_t use_current_device()
{
    int device = 0;
    return cudaGetDevice(&device);
}

int check_common(int64_t ny, int64_t nx, int64_t klen, int uv_mode, const Walls &w)
{
    if (ny < 0 || nx < 0)
        return fail(RLIC_B200_EINVAL, "negative image size %lld x %lld", (long long)ny, (long long)nx);
    if (ny > INT_MAX - 2 || nx > INT_MAX - 2)
        return fail(RLIC_B200_EINVAL, "image side exceeds 2^31-3 pixels");
    if (klen <= 0)
        return fail(RLIC_B200_EINVAL,
                    "empty convolution kernel (the reference aborts the process here, lib.rs:371)");
    if (klen > INT_MAX / 2)
        return fail(RLIC_B200_EINVAL, "kernel too long");
    if (uv_mode != RLIC_B200_VELOCITY && uv_mode != RLIC_B200_POLARIZATION)
        return fail(RLIC_B200_EINVAL, "unknown uv_mode %d", uv_mode);
    const int sides[4] = {w.x_left, w.x_right, w.y_left, w.y_right};
    for (int s : sides)
        if (s != RLIC_B200_CLOSED && s != RLIC_B200_PERIODIC)
            return fail(RLIC_B200_EINVAL, "unknown boundary %d", s);
    return RLIC_B200_OK;
}

// Which rows of the global ny x nx image a device buffer holds: rows
// [row0 - halo_lo, row0 + nrows + halo_hi), of which [row0, row0 + nrows) are
// the ones a launch may compute.  A whole image is {0, ny, 0, 0}.
struct Slab { int64_t row0, nrows, halo_lo, halo_hi; };

int check_slab(int64_t ny, int64_t klen, const Slab &sl, const Walls &w)
{
    if (sl.row0 < 0 || sl.nrows < 0 || sl.row0 + sl.nrows > ny || sl.halo_lo < 0 || sl.halo_hi < 0)
        return fail(RLIC_B200_ESHARD, "slab rows [%lld,%lld) outside image of %lld rows",
                    (long long)sl.row0, (long long)(sl.row0 + sl.nrows), (long long)ny);
    const int64_t reach = klen / 2;   // a walker moves at most one row per tap
    const bool whole = sl.row0 == 0 && sl.nrows == ny && sl.halo_lo == 0 && sl.halo_hi == 0;
    const bool periodic_y = w.y_left == RLIC_B200_PERIODIC || w.y_right == RLIC_B200_PERIODIC;
    // A side needs `reach` halo rows unless a closed wall stops the walker there.
    const bool lo_closed = sl.row0 == 0 && !periodic_y;
    const bool hi_closed = sl.row0 + sl.nrows == ny && !periodic_y;
    if (!whole && ((!lo_closed && sl.halo_lo < reach) || (!hi_closed && sl.halo_hi < reach)))
        return fail(RLIC_B200_ESHARD, "slab halo (%lld,%lld) shorter than the kernel half-width %lld",
                    (long long)sl.halo_lo, (long long)sl.halo_hi, (long long)reach);
    return RLIC_B200_OK;
}

// Geometry of the padded buffers of a slab (wall rules: lib.rs:83-95).
PassGeom make_geometry(int64_t ny, int64_t nx, const Slab &sl, const Walls &w)
{
    PassGeom g{};
    g.nx = (int)nx;
    g.pitch = (int)nx + 2;
    g.rows = (int)(sl.halo_lo + sl.nrows + sl.halo_hi);
    g.field_stride = rlic::padded_cells(g.rows, nx);
    const int64_t shift = sl.row0 - sl.halo_lo;   // global row of buffer row 0
    const bool whole = sl.row0 == 0 && sl.nrows == ny && sl.halo_lo == 0 && sl.halo_hi == 0;
    const bool periodic_y = w.y_left == RLIC_B200_PERIODIC || w.y_right == RLIC_B200_PERIODIC;
    // a whole image wraps onto itself; a slab of a y-periodic image wraps into
    // halos its owner filled (ring order), so it has no reachable row walls
    g.lo_wall = whole || (sl.row0 == 0 && !periodic_y);
    g.hi_wall = whole || (sl.row0 + sl.nrows == ny && !periodic_y);
    g.j_below_to = w.x_left == RLIC_B200_PERIODIC ? (int)nx - 1 : 0;
    g.j_above_to = w.x_right == RLIC_B200_PERIODIC ? 0 : (int)nx - 1;
    g.i_below_to = (int)((w.y_left == RLIC_B200_PERIODIC ? ny - 1 : 0) - shift);
    g.i_above_to = (int)((w.y_right == RLIC_B200_PERIODIC ? 0 : ny - 1) - shift);
    return g;
}

template <typename T> using Field = rlic::PackedField<T>;

unsigned stream_blocks(long long items)
{
    return (unsigned)std::max<long long>(1, std::min<long long>((items + 255) / 256, 148 * 16));
}

template <typename T>
cudaError_t launch_pack(const T *u, const T *v, Field<T> *field, const PassGeom &g, int64_t rb,
                        int64_t re, int64_t nfields, cudaStream_t stream)
{
    if (re <= rb || nfields <= 0)
        return cudaSuccess;
    rlic::pack_field_kernel<T><<<stream_blocks((re - rb + 2) * g.pitch * nfields), 256, 0, stream>>>(
        u, v, field, g, (int)rb, (int)re, (long long)nfields);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_pad(const T *dense, T *padded, const PassGeom &g, int64_t rb, int64_t re,
                       int64_t nfields, int *negative, cudaStream_t stream)
{
    if (re <= rb || nfields <= 0)
        return cudaSuccess;
    rlic::pad_texture_kernel<T><<<stream_blocks((re - rb + 2) * g.pitch * nfields), 256, 0, stream>>>(
        dense, padded, g, (int)rb, (int)re, (long long)nfields, negative);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_unpad(const T *padded, T *dense, const PassGeom &g, int64_t rb, int64_t re,
                         int64_t nfields, cudaStream_t stream)
{
    if (re <= rb || nfields <= 0)
        return cudaSuccess;
    rlic::unpad_texture_kernel<T><<<stream_blocks((re - rb) * g.nx * nfields), 256, 0, stream>>>(
        padded, dense, g, (int)rb, (int)re, (long long)nfields);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

// Taps prepared once per call: either a parameter block or a device copy.
template <typename T> struct TapSet {
    static constexpr int kMaxParam = rlic::kParamTapBytes / (int)sizeof(T);
    rlic::ParamTaps<T, kMaxParam> param;
    rlic::StepTaps<T, rlic::kStepTapsPerHalf<T>> steps;   // the same taps in walking order, for the replay
    DeviceBuf global;
    int ntaps = 0;
    bool in_param = true;

    cudaError_t prepare(const T *host_taps, int64_t klen, cudaStream_t stream)
    {
        ntaps = (int)klen;
        in_param = klen <= kMaxParam;
        if (in_param) {
            std::memset(param.w, 0, sizeof param.w);
            std::memcpy(param.w, host_taps, sizeof(T) * (size_t)klen);
            std::memset(&steps, 0, sizeof steps);
            const int64_t kmid = klen / 2;
            steps.centre = host_taps[kmid];
            for (int64_t k = kmid + 1; k < klen; ++k)
                steps.fwd[k - kmid - 1] = host_taps[k];
            for (int64_t k = kmid - 1; k >= 0; --k)
                steps.bwd[kmid - 1 - k] = host_taps[k];
            return cudaSuccess;
        }
        cudaError_t e = global.alloc(sizeof(T) * (size_t)klen, stream);
        if (e != cudaSuccess)
            return e;
        // the source may be pageable: the copy is staged before this returns
        return cudaMemcpyAsync(global.p, host_taps, sizeof(T) * (size_t)klen,
                               cudaMemcpyHostToDevice, stream);
    }
};

// Where the results of a pass over an edge strip also go when the halo exchange is fused
// into it (rlic_b200_pass_slab_peer_*): the neighbour's padded buffer and the cell offset
// between a row here and the same row there.
template <typename T> struct PeerTarget {
    T *out = nullptr;
    long long delta = 0;
};

// What a pass does with the recorded streamline paths (rlic::PathPlanes in lic_walk.cuh): the
// first pass of a call with more iterations to come records them, the others replay them.
enum class Paths { none, record, replay };
struct PathUse {
    Paths mode = Paths::none;
    uint4 *rec = nullptr;          // groups * group_cells entries (one per pixel and group of 32 steps)
    long long group_cells = 0;     // cells of the whole padded buffer (every field)
};

// words of the record of one padded buffer of `cells` cells, for a kernel of `klen` taps
size_t path_record_words(int64_t cells, int64_t klen)
{
    return (size_t)(rlic::path_groups_fwd(klen) + rlic::path_groups_bwd(klen)) * rlic::kPlanesPerGroup * (size_t)cells;
}

// Registers: the recording walk keeps the planes of its group (three words) next to the walk's
// own state; the f32 kernels sit at the 32-register cap of 8 CTAs per SM, so the recording
// instantiation gets one CTA fewer instead of spills.
template <typename T, bool POL> constexpr int record_min_blocks()
{
    return rlic::Tune<T, POL>::walk_min_blocks > 6 ? 6 : rlic::Tune<T, POL>::walk_min_blocks;
}

template <typename T, bool POL, typename Taps, typename Idx>
cudaError_t launch_one(const T *tex, const Field<T> *field, T *out, const PassGeom &g,
                       const Taps &taps, int ntaps, unsigned blocks, bool branchless, bool grouped,
                       const PeerTarget<T> &peer, const PathUse &paths, cudaStream_t stream)
{
    using Tn = rlic::Tune<T, POL>;
    const rlic::PathPlanes none{nullptr, 0, 0};
    if (paths.mode == Paths::record) {   // grouped walk, default arithmetic (checked by the caller)
        const rlic::PathPlanes planes{paths.rec, paths.group_cells, rlic::path_groups_fwd(ntaps)};
        if (peer.out)
            rlic::lic_pass_peer_kernel<T, POL, Taps, Idx, rlic::kTileW, rlic::kTileH, Tn::walk_unroll,
                                       record_min_blocks<T, POL>(), Tn::walk_flavor, Tn::walk_admit, true, Tn::walk, true>
                <<<blocks, rlic::kThreads, 0, stream>>>(tex, field, out, g, taps, ntaps, peer.out, peer.delta, planes);
        else
            rlic::lic_pass_kernel<T, POL, Taps, Idx, rlic::kTileW, rlic::kTileH, Tn::walk_unroll,
                                  record_min_blocks<T, POL>(), Tn::walk_flavor, Tn::walk_admit, true, Tn::walk, true>
                <<<blocks, rlic::kThreads, 0, stream>>>(tex, field, out, g, taps, ntaps, planes);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        return cudaGetLastError();
    }
    if (peer.out) {   // the default arithmetic only (checked by the caller)
        if (grouped)
            rlic::lic_pass_peer_kernel<T, POL, Taps, Idx, rlic::kTileW, rlic::kTileH, Tn::walk_unroll,
                                       Tn::walk_min_blocks, Tn::walk_flavor, Tn::walk_admit, true, Tn::walk>
                <<<blocks, rlic::kThreads, 0, stream>>>(tex, field, out, g, taps, ntaps, peer.out, peer.delta, none);
        else
            rlic::lic_pass_peer_kernel<T, POL, Taps, Idx>
                <<<blocks, rlic::kThreads, 0, stream>>>(tex, field, out, g, taps, ntaps, peer.out, peer.delta, none);
    } else if (branchless && grouped)
        rlic::lic_pass_kernel<T, POL, Taps, Idx, rlic::kTileW, rlic::kTileH, Tn::walk_unroll,
                              Tn::walk_min_blocks, Tn::walk_flavor, Tn::walk_admit, true, Tn::walk>
            <<<blocks, rlic::kThreads, 0, stream>>>(tex, field, out, g, taps, ntaps, none);
    else if (branchless)
        rlic::lic_pass_kernel<T, POL, Taps, Idx>
            <<<blocks, rlic::kThreads, 0, stream>>>(tex, field, out, g, taps, ntaps, none);
    else
        rlic::lic_pass_kernel<T, POL, Taps, Idx, rlic::kTileW, rlic::kTileH, Tn::unroll, Tn::min_blocks,
                              Tn::flavor, Tn::admit, false>
            <<<blocks, rlic::kThreads, 0, stream>>>(tex, field, out, g, taps, ntaps, none);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

// One pass over buffer rows [first_row, first_row + out_rows) of `nfields`
// fields.  tex, field, out: padded buffers of geometry g.
// Whether a pass is small enough for the two-warps-per-pixel kernel (lic_pass_pair_kernel):
// too few pixels to fill the GPU, parked samples that fit in shared memory, taps in the
// parameter block, 32-bit cell indices, the default arithmetic.
template <typename T>
bool pair_kernel_fits(const PassGeom &g, int64_t nfields, int64_t out_rows, const TapSet<T> &taps)
{
    const int64_t kmid = taps.ntaps / 2;
    return g_small_images.load(std::memory_order_relaxed) != 0 && taps.in_param && kmid >= 1 &&
           out_rows * g.nx * nfields <= rlic::kPairMaxPixels &&
           kmid * rlic::kPairPixels * (int64_t)sizeof(T) <= rlic::kPairSmemBytes &&
           g.field_stride < (int64_t)INT_MAX && g_force_wide.load(std::memory_order_relaxed) == 0 &&
           effective_arithmetic() == RLIC_B200_ARITH_FMA_BRANCHLESS;
}

template <typename T, bool POL>
cudaError_t launch_pair(const T *tex, const Field<T> *field, T *out, PassGeom g, int64_t nfields,
                        int64_t out_rows, const TapSet<T> &taps, T *dense_out, cudaStream_t stream)
{
    using PT = rlic::ParamTaps<T, TapSet<T>::kMaxParam>;
    auto kernel = rlic::lic_pass_pair_kernel<T, POL, PT, int>;
    static std::once_flag once;
    static cudaError_t attr = cudaSuccess;
    std::call_once(once, [&] {
        attr = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, rlic::kPairSmemBytes);
    });
    if (attr != cudaSuccess)
        return attr;
    g.tiles_x = (g.nx + rlic::kPairTileW - 1) / rlic::kPairTileW;
    g.tiles_per_field = g.tiles_x * (int)((out_rows + rlic::kPairTileH - 1) / rlic::kPairTileH);
    const size_t smem = (size_t)(taps.ntaps / 2) * rlic::kPairPixels * sizeof(T);
    kernel<<<(unsigned)(g.tiles_per_field * nfields), rlic::kPairThreads, smem, stream>>>(
        tex, field, out, g, taps.param, taps.ntaps, dense_out);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

// `dense_out`: when not null and the pass runs on the small-image kernel, the results go there
// (dense rows, as rlic_b200_slab_unpad_texture would leave them) instead of the padded `out`,
// and *wrote_dense is set; the caller then skips its un-padding launch.
// Whether a call of `iterations` passes records the paths of its first pass and replays them
// in the others: the option, more than one pass, and the walk that knows how to record.
bool paths_are_replayed(int64_t iterations)
{
    return iterations >= 2 && effective_paths() == RLIC_B200_PATHS_REPLAY &&
           effective_arithmetic() == RLIC_B200_ARITH_FMA_BRANCHLESS && effective_walk() == RLIC_B200_WALK_GROUPED;
}

// One pass by replay of the recorded paths (lic_replay_kernel); g carries the launch's rows.
// The grid is (tile column, tile row, field).
template <typename T, typename Idx>
cudaError_t launch_replay(const T *tex, T *out, const PassGeom &g, int64_t nfields, const TapSet<T> &taps,
                          const PeerTarget<T> &peer, const PathUse &paths, cudaStream_t stream)
{
    using ST = rlic::StepTaps<T, rlic::kStepTapsPerHalf<T>>;
    using GT = rlic::GlobalStepTaps<T>;
    const int groups = std::max(rlic::path_groups_fwd(taps.ntaps), rlic::path_groups_bwd(taps.ntaps));
    // kernels of up to 65 taps whose window fits: the texture window of each tile staged in
    // shared memory (lic_replay_staged_kernel)
    if (taps.in_param && groups <= 1 && sizeof(Idx) == 4 && g_replay_staging.load(std::memory_order_relaxed) != 0 &&
        rlic::staged_window_bytes(taps.ntaps, sizeof(T)) <= rlic::kStagedMaxBytes) {
        const dim3 sgrid((unsigned)((g.nx + rlic::kStagedTW - 1) / rlic::kStagedTW),
                         (unsigned)((g.out_rows + rlic::kStagedTH - 1) / rlic::kStagedTH), (unsigned)nfields);
        if (sgrid.y > 65535u || sgrid.z > 65535u)
            return cudaErrorInvalidConfiguration;
        const size_t smem = (size_t)rlic::staged_window_bytes(taps.ntaps, sizeof(T));
        if (peer.out)
            rlic::lic_replay_staged_kernel<T, ST, int, true, rlic::kStagedTW, rlic::kStagedTH, rlic::kStagedPad,
                                           rlic::kStagedMinBlocks>
                <<<sgrid, rlic::kStagedTW * rlic::kStagedTH, smem, stream>>>(
                    tex, paths.rec, out, g, taps.steps, taps.ntaps, paths.group_cells, peer.out, peer.delta);
        else
            rlic::lic_replay_staged_kernel<T, ST, int, false, rlic::kStagedTW, rlic::kStagedTH, rlic::kStagedPad,
                                           rlic::kStagedMinBlocks>
                <<<sgrid, rlic::kStagedTW * rlic::kStagedTH, smem, stream>>>(
                    tex, paths.rec, out, g, taps.steps, taps.ntaps, paths.group_cells, nullptr, 0);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        return cudaGetLastError();
    }
    const dim3 grid((unsigned)g.tiles_x, (unsigned)((g.out_rows + rlic::kTileH - 1) / rlic::kTileH), (unsigned)nfields);
    if (grid.y > 65535u || grid.z > 65535u)
        return cudaErrorInvalidConfiguration;
#define RLIC_REPLAY(TAPS, TAPV, GROUPS)                                                                    \
    do {                                                                                                   \
        if (peer.out)                                                                                      \
            rlic::lic_replay_kernel<T, TAPS, Idx, GROUPS, true><<<grid, rlic::kThreads, 0, stream>>>(      \
                tex, paths.rec, out, g, TAPV, taps.ntaps, paths.group_cells, peer.out, peer.delta);        \
        else                                                                                               \
            rlic::lic_replay_kernel<T, TAPS, Idx, GROUPS, false><<<grid, rlic::kThreads, 0, stream>>>(     \
                tex, paths.rec, out, g, TAPV, taps.ntaps, paths.group_cells, nullptr, 0);                  \
    } while (0)
    if (!taps.in_param) {
        const GT gt{static_cast<const T *>(taps.global.p), taps.ntaps / 2};
        RLIC_REPLAY(GT, gt, 0);
    } else if (groups <= 1)
        RLIC_REPLAY(ST, taps.steps, 1);
    else if (groups == 2)
        RLIC_REPLAY(ST, taps.steps, 2);
    else
        RLIC_REPLAY(ST, taps.steps, 0);
#undef RLIC_REPLAY
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaGetLastError();
}

template <typename T>
int launch_pass(const T *tex, const Field<T> *field, T *out, PassGeom g, int64_t nfields,
                int64_t first_row, int64_t out_rows, int uv_mode, const TapSet<T> &taps,
                cudaStream_t stream, const PeerTarget<T> &peer = PeerTarget<T>{}, T *dense_out = nullptr,
                bool *wrote_dense = nullptr, const PathUse &paths = PathUse{})
{
    if (wrote_dense)
        *wrote_dense = false;
    if (out_rows <= 0 || g.nx <= 0 || nfields <= 0)
        return RLIC_B200_OK;
    if (paths.mode != Paths::none && !paths.rec)
        return fail(RLIC_B200_EINVAL, "null path record");
    if (!peer.out && paths.mode == Paths::none && pair_kernel_fits<T>(g, nfields, out_rows, taps)) {
        g.first_row = (int)first_row;
        g.out_rows = (int)out_rows;
        T *dense = (dense_out && first_row == 0 && out_rows == g.rows) ? dense_out : nullptr;
        CUDA_TRY(uv_mode == RLIC_B200_POLARIZATION
                     ? (launch_pair<T, true>(tex, field, out, g, nfields, out_rows, taps, dense, stream))
                     : (launch_pair<T, false>(tex, field, out, g, nfields, out_rows, taps, dense, stream)));
        if (wrote_dense)
            *wrote_dense = dense != nullptr;
        return RLIC_B200_OK;
    }
    g.first_row = (int)first_row;
    g.out_rows = (int)out_rows;
    g.tiles_x = (g.nx + rlic::kTileW - 1) / rlic::kTileW;
    const int64_t tiles_y = (out_rows + rlic::kTileH - 1) / rlic::kTileH;
    const int64_t per_field = tiles_y * g.tiles_x;
    const int64_t blocks = per_field * nfields;
    if (per_field > INT_MAX || blocks > INT_MAX)
        return fail(RLIC_B200_EINVAL, "too many tiles for one launch (%lld)", (long long)blocks);
    g.tiles_per_field = (int)per_field;
    // 32-bit cell indices whenever one field's buffer allows it
    const bool wide = g.field_stride >= (int64_t)INT_MAX ||
                      g_force_wide.load(std::memory_order_relaxed) != 0;
    const bool pol = uv_mode == RLIC_B200_POLARIZATION;
    const bool branchless = effective_arithmetic() == RLIC_B200_ARITH_FMA_BRANCHLESS;
    const bool grouped = effective_walk() == RLIC_B200_WALK_GROUPED;
    if (peer.out && (!branchless || nfields != 1))
        return fail(RLIC_B200_EINVAL, "the fused halo exchange needs the default arithmetic and one field");
    if (paths.mode != Paths::none && !(branchless && grouped))
        return fail(RLIC_B200_EINVAL, "paths are recorded and replayed with the default arithmetic and the grouped walk only");

    cudaError_t e;
    if (paths.mode == Paths::replay) {
        e = wide ? launch_replay<T, long long>(tex, out, g, nfields, taps, peer, paths, stream)
                 : launch_replay<T, int>(tex, out, g, nfields, taps, peer, paths, stream);
        CUDA_TRY(e);
        return RLIC_B200_OK;
    }
#define RLIC_LAUNCH(POL, TAPS, TAPV, IDX) \
    e = launch_one<T, POL, TAPS, IDX>(tex, field, out, g, TAPV, taps.ntaps, (unsigned)blocks, branchless, grouped, peer, paths, stream)
    using PT = rlic::ParamTaps<T, TapSet<T>::kMaxParam>;
    using GT = rlic::GlobalTaps<T>;
    const GT gt{static_cast<const T *>(taps.global.p)};
    if (taps.in_param) {
        if (pol) { if (wide) RLIC_LAUNCH(true, PT, taps.param, long long); else RLIC_LAUNCH(true, PT, taps.param, int); }
        else     { if (wide) RLIC_LAUNCH(false, PT, taps.param, long long); else RLIC_LAUNCH(false, PT, taps.param, int); }
    } else {
        if (pol) { if (wide) RLIC_LAUNCH(true, GT, gt, long long); else RLIC_LAUNCH(true, GT, gt, int); }
        else     { if (wide) RLIC_LAUNCH(false, GT, gt, long long); else RLIC_LAUNCH(false, GT, gt, int); }
    }
#undef RLIC_LAUNCH
    CUDA_TRY(e);
    return RLIC_B200_OK;
}

struct Events {
    std::vector<cudaEvent_t> ev;
    ~Events() { for (cudaEvent_t e : ev) cudaEventDestroy(e); }
    cudaError_t make(size_t n)
    {
        for (size_t i = 0; i < n; ++i) {
            cudaEvent_t e;
            cudaError_t rc = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
            if (rc != cudaSuccess)
                return rc;
            ev.push_back(e);
        }
        return cudaSuccess;
    }
};

// Order in which the (pass, band) launches of the wavefront schedule are issued on the one
// compute stream.  Pass p of band b reads pass p-1 of bands b-1, b, b+1 (a band is at
// least two kernel half-widths tall) and overwrites what pass p-1 of those same bands
// read (two ping-pong buffers), so it may run once those three are done.  On an in-order
// stream the tightest order that guarantees it is the diagonal one: slots b + (p - 1), and within
// a slot ascending passes -- (1, s), (2, s - 1), (3, s - 2), ...: each launch finds the last of
// its three predecessors, pass p-1 of band b+1, right in front of it.  The last pass of band b
// therefore runs `iterations - 1` bands behind pass 1, which in turn needs the upload of band
// b + 1: once the last band has arrived, about iterations^2 / 2 band-passes remain, all on the
// small bands at the end of the image (band_plan).  (Round 1 used slots b + 2 (p - 1), twice
// the skew, for the sake of independent launches within a slot -- which a single stream cannot
// exploit.)
struct BandPass { int pass, band; };   // pass is 1-based

std::vector<BandPass> wavefront_order(int64_t nbands, int64_t iterations)
{
    std::vector<BandPass> order;
    order.reserve((size_t)(nbands * iterations));
    const int64_t slots = (nbands - 1) + (iterations - 1) + 1;
    for (int64_t slot = 0; slot < slots; ++slot)
        for (int64_t p = 1; p <= iterations; ++p) {
            const int64_t b = slot - (p - 1);
            if (b >= 0 && b < nbands)
                order.push_back({(int)p, (int)b});
        }
    return order;
}

// Row bands of a host call over one image: edges[b] .. edges[b + 1] are the rows of band b.
// The call is a pipeline -- uploads, passes, downloads -- and since the passes became cheaper
// than the uploads (recorded paths) its length is the uploads plus whatever cannot start before
// the last band has arrived: the last pass of a band runs `iterations` bands behind the uploads
// (wavefront_order), so small bands would make a short tail -- but a pass over a small band is
// bound by the latency of one wave of walkers, not by throughput (measured with ncu, 4096
// columns: replaying 64 rows takes 19 us, 256 rows 39 us, 512 rows 66 us; walking them 40 /
// 116 / 218 us), so small bands make the passes themselves slow.  Sixteen bands is where the
// two meet for the headline image (a model fed with those launch times and the copy rate puts
// 16 x 256 rows at 5.4 ms, 8 x 512 at 5.9, 35 bands of 64-128 rows at 6.2; the last was also
// measured: 6.6).  A band is at least two kernel half-widths tall (a pass reaches one band up
// and down, and overwrites what the previous pass of the neighbouring bands read), at least 64
// rows and 256 Kpix, and a multiple of the tile height; the top band takes the remainder.
std::vector<int64_t> band_plan(int64_t ny, int64_t nx, int64_t reach, int64_t iterations)
{
    (void)iterations;
    const int64_t tile = rlic::kTileH;
    auto round_up = [&](int64_t x) { return (x + tile - 1) / tile * tile; };
    int64_t smallest = round_up(std::max<int64_t>(2 * reach, 64));
    if (nx > 0)
        smallest = std::max(smallest, round_up(((int64_t)1 << 18) / nx));      // >= 256 Kpix per band
    std::vector<int64_t> edges{0, ny};
    if (ny * nx < ((int64_t)1 << 21) || ny < 2 * smallest)
        return edges;
    const int64_t size = std::max(smallest, round_up(ny / 16));
    std::vector<int64_t> sizes;                  // from the bottom of the image up
    int64_t left = ny;
    while (left >= size + smallest && sizes.size() < 47) {
        sizes.push_back(size);
        left -= size;
    }
    sizes.push_back(left);                       // the top band: whatever remains (>= smallest)
    edges.assign(1, 0);
    for (size_t k = sizes.size(); k-- > 0;)
        edges.push_back(edges.back() + sizes[k]);
    return edges;
}

// The record is planes * 4 bytes per cell (32 bytes for a 65-tap kernel): small next to 180 GB,
// but a caller may have filled the device.  Large records are checked against the free memory;
// when one does not fit the call simply walks every pass.
bool path_record_fits(size_t bytes)
{
    if (bytes <= ((size_t)1 << 30))
        return true;
    size_t free_bytes = 0, total = 0;
    if (cudaMemGetInfo(&free_bytes, &total) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return bytes < free_bytes / 2;
}

// The record of a call: allocated when its passes replay, and how pass `p` (1-based) uses it.
struct CallPaths {
    DeviceBuf buf;
    long long group_cells = 0;
    bool on = false;
    cudaError_t prepare(int64_t iterations, int64_t cells, int64_t klen, cudaStream_t stream)
    {
        const size_t bytes = path_record_words(cells, klen) * sizeof(unsigned);
        on = paths_are_replayed(iterations) && bytes > 0 && path_record_fits(bytes);
        group_cells = cells;
        return on ? buf.alloc(bytes, stream) : cudaSuccess;
    }
    PathUse use(int64_t pass) const
    {
        if (!on)
            return PathUse{};
        return PathUse{pass == 1 ? Paths::record : Paths::replay, static_cast<uint4 *>(buf.p), group_cells};
    }
};

// RLIC_B200_TRACE=1: the host entry point reports, on stderr, when (host clock, ms since entry) it
// finished allocating, enqueueing the uploads and the passes, and returning, and when (device
// clock, ms since the first upload began) the last upload, the last pass and the last download
// ended.  A measurement aid (tools/e2e_probe.py); costs nothing when off.
struct HostTrace {
    bool on = false;
    double t0 = 0;
    std::vector<std::pair<const char *, double>> marks;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // begin, uploads, passes, downloads
    static double now()
    {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
    }
    HostTrace()
    {
        static const bool wanted = getenv("RLIC_B200_TRACE") && atoi(getenv("RLIC_B200_TRACE")) != 0;
        on = wanted;
        t0 = on ? now() : 0;
    }
    void mark(const char *what)
    {
        if (on)
            marks.push_back({what, now() - t0});
    }
    void device(int which, cudaStream_t s)
    {
        if (!on)
            return;
        if (!ev[which])
            cudaEventCreate(&ev[which]);
        cudaEventRecord(ev[which], s);
    }
    ~HostTrace()
    {
        if (!on)
            return;
        mark("return");
        std::string line = "rlic_b200 trace: host ms";
        char buf[96];
        for (auto &m : marks) {
            snprintf(buf, sizeof buf, " | %s %.3f", m.first, m.second);
            line += buf;
        }
        line += " || device ms since first upload";
        const char *names[4] = {"", "uploads done", "passes done", "downloads done"};
        for (int k = 1; k < 4; ++k)
            if (ev[0] && ev[k]) {
                float ms = 0;
                if (cudaEventElapsedTime(&ms, ev[0], ev[k]) == cudaSuccess) {
                    snprintf(buf, sizeof buf, " | %s %.3f", names[k], ms);
                    line += buf;
                }
            }
        fprintf(stderr, "%s\n", line.c_str());
        for (cudaEvent_t e : ev)
            if (e)
                cudaEventDestroy(e);
    }
};

// Host entry: upload -> passes -> download, pipelined over row bands.
//   stream `io`  : uploads (band by band: u, v -> packed field; texture -> padded
//                  buffer) and downloads (padded -> dense -> host)
//   stream `run` : the passes; pass 1 of a band waits only for the bands it can
//                  reach (kernel half-width), the last pass releases each band
//                  to the download as soon as it is done.
template <typename T>
int convolve_host(const T *tex, const T *u, const T *v, int64_t nfields, int64_t ny, int64_t nx,
                  const T *kernel, int64_t klen, int uv_mode, const Walls &w,
                  int64_t iterations, T *out, int device, int *texture_has_negative = nullptr)
{
    Range whole("rlic_b200 convolve (host)");
    HostTrace trace;
    if (texture_has_negative)
        *texture_has_negative = 0;
    if (int rc = check_common(ny, nx, klen, uv_mode, w))
        return rc;
    const size_t count = (size_t)nfields * (size_t)ny * (size_t)nx;
    if (count == 0)
        return RLIC_B200_OK;
    if (!tex || !u || !v || !kernel || !out)
        return fail(RLIC_B200_EINVAL, "null pointer argument");
    if (iterations <= 0) {   // lib.rs:423-432: the loop never runs, output stays zero
        std::memset(out, 0, count * sizeof(T));
        return RLIC_B200_OK;
    }
    const size_t bytes = count * sizeof(T);
    const PassGeom g = make_geometry(ny, nx, Slab{0, ny, 0, 0}, w);
    const size_t padded_bytes = (size_t)g.field_stride * (size_t)nfields * sizeof(T);

    // Row bands (single images only; a batch chunk is pipelined by its caller).
    const int64_t reach = klen / 2;
    const std::vector<int64_t> edges =
        nfields == 1 ? band_plan(ny, nx, reach, iterations) : std::vector<int64_t>{0, ny};
    const int64_t nbands = (int64_t)edges.size() - 1;
    const bool periodic_y = w.y_left == RLIC_B200_PERIODIC || w.y_right == RLIC_B200_PERIODIC;
    auto band_begin = [&](int64_t b) { return edges[(size_t)b]; };
    auto band_of_row = [&](int64_t row) {
        return (int64_t)(std::upper_bound(edges.begin(), edges.end(), row) - edges.begin()) - 1;
    };
    // the last band a pass-1 walker of band b can reach: it must be on the device first
    auto upload_needed = [&](int64_t b) {
        return periodic_y ? nbands - 1 : band_of_row(std::min(ny - 1, band_begin(b + 1) - 1 + reach));
    };

    CUDA_TRY(use_device(device));
    Stream io, prep, run, back;   // `back` is created by the wavefront schedule only
    CUDA_TRY(host_stream(device, 0, io));
    CUDA_TRY(host_stream(device, 2, run));
    if (nbands > 1)   // the layout conversions of a band run beside the next band's copies
        CUDA_TRY(host_stream(device, 1, prep));
    // Two padded texture buffers (the uploaded texture's doubles as the second
    // work buffer), the packed field, and dense staging for the three uploads
    // (the texture's staging is reused for the download).
    DeviceBuf d_tex, d_work, d_field, d_su, d_sv, d_st, d_flag;
    CallPaths paths;
    TapSet<T> taps;                // (owns a device copy of kernels beyond the parameter block)
    StreamDrain drain{{&io, &run, &back, &prep}};
    CUDA_TRY(d_tex.alloc(padded_bytes, io.s));
    CUDA_TRY(d_work.alloc(padded_bytes, io.s));
    CUDA_TRY(d_field.alloc(4 * padded_bytes, io.s));
    CUDA_TRY(d_su.alloc(bytes, io.s));
    CUDA_TRY(d_sv.alloc(bytes, io.s));
    CUDA_TRY(d_st.alloc(bytes, io.s));
    if (texture_has_negative) {
        CUDA_TRY(d_flag.alloc(sizeof(int), io.s));
        CUDA_TRY(cudaMemsetAsync(d_flag.p, 0, sizeof(int), io.s));
    }
    CUDA_TRY(taps.prepare(kernel, klen, io.s));
    // (allocated on `io`, first used on `run` behind an `uploaded` event recorded later on `io`)
    CUDA_TRY(paths.prepare(iterations, g.field_stride * nfields, klen, io.s));
    Events uploaded, done, copied;
    CUDA_TRY(uploaded.make((size_t)nbands));
    CUDA_TRY(done.make((size_t)nbands));
    if (prep.s)
        CUDA_TRY(copied.make((size_t)nbands));

    trace.mark("allocated");
    trace.device(0, io.s);
    T *const t_tex = static_cast<T *>(d_tex.p);
    T *const t_work = static_cast<T *>(d_work.p);
    T *const s_u = static_cast<T *>(d_su.p);
    T *const s_v = static_cast<T *>(d_sv.p);
    T *const s_t = static_cast<T *>(d_st.p);
    Field<T> *const t_field = static_cast<Field<T> *>(d_field.p);
    int *const flag = static_cast<int *>(d_flag.p);
    // pass n writes work[(n-1) % 2]: d_work, then back over the texture copy, ...
    // exactly two texture-sized work buffers (README.md:158-164 of the reference)
    T *const bufs[2] = {t_work, t_tex};
    const bool single = iterations == 1;

    auto first_pass_band = [&](int64_t b) -> int {
        // rows this band's walkers can reach must be on the device
        CUDA_TRY(cudaStreamWaitEvent(run.s, uploaded.ev[(size_t)upload_needed(b)], 0));
        int rc = launch_pass<T>(t_tex, t_field, bufs[0], g, nfields, band_begin(b),
                                band_begin(b + 1) - band_begin(b), uv_mode, taps, run.s, PeerTarget<T>{},
                                nullptr, nullptr, paths.use(1));
        if (rc)
            return rc;
        if (single)
            CUDA_TRY(cudaEventRecord(done.ev[(size_t)b], run.s));
        return RLIC_B200_OK;
    };

    // One band of the inputs: u, v -> packed field, texture -> padded buffer, all on `io`.
    auto enqueue_band_upload = [&](int64_t b) -> int {
        Range r("upload band: u, v -> packed field; texture -> padded");
        const int64_t rb = band_begin(b), re = band_begin(b + 1);
        const size_t off = (size_t)rb * (size_t)nx * (size_t)nfields;
        const size_t n = (size_t)(re - rb) * (size_t)nx * (size_t)nfields;
        // one sweep for the three arrays: pageable sources are staged by all the helper threads at once
        const HostToDevice jobs[3] = {{s_u + off, u + off, n * sizeof(T)},
                                      {s_v + off, v + off, n * sizeof(T)},
                                      {s_t + off, tex + off, n * sizeof(T)}};
        CUDA_TRY(upload(jobs, 3, io.s));
        // the conversions: on their own stream when there are bands, so that the copy engine goes
        // straight on with the next band (the first allocation-ordered use of the buffers is `io`'s)
        cudaStream_t conv = prep.s ? prep.s : io.s;
        if (prep.s) {
            CUDA_TRY(cudaEventRecord(copied.ev[(size_t)b], io.s));
            CUDA_TRY(cudaStreamWaitEvent(prep.s, copied.ev[(size_t)b], 0));
        }
        CUDA_TRY(launch_pack<T>(s_u + off, s_v + off, t_field, g, rb, re, nfields, conv));
        CUDA_TRY(launch_pad<T>(s_t + off, t_tex, g, rb, re, nfields, flag, conv));
        CUDA_TRY(cudaEventRecord(uploaded.ev[(size_t)b], conv));
        return RLIC_B200_OK;
    };

    // ---- wavefront schedule (opt-in): every pass trails the uploads, band by band ----
    // Early bands run through all their passes while later bands are still on the bus,
    // and their results go back while later bands still compute.  Needs a top and a
    // bottom that do not depend on each other (no wrap in y) and something to skew.
    if (effective_schedule() == RLIC_B200_SCHEDULE_WAVEFRONT && !periodic_y &&
        iterations >= 2 && nbands >= 2) {
        const std::vector<BandPass> order = wavefront_order(nbands, iterations);
        auto band_pass = [&](const BandPass &bp) -> int {
            const int64_t b = bp.band;
            // pass p reads what pass p-1 wrote (pass 1: the uploaded texture) and writes bufs[(p-1) % 2]
            const T *from = bp.pass == 1 ? t_tex : bufs[(bp.pass - 2) & 1];
            if (bp.pass == 1)
                CUDA_TRY(cudaStreamWaitEvent(run.s, uploaded.ev[(size_t)upload_needed(b)], 0));
            if (int rc = launch_pass<T>(from, t_field, bufs[(bp.pass - 1) & 1], g, nfields, band_begin(b),
                                        band_begin(b + 1) - band_begin(b), uv_mode, taps, run.s, PeerTarget<T>{},
                                        nullptr, nullptr, paths.use(bp.pass)))
                return rc;
            if (bp.pass == iterations)
                CUDA_TRY(cudaEventRecord(done.ev[(size_t)b], run.s));
            return RLIC_B200_OK;
        };
        size_t next = 0;
        for (int64_t b = 0; b < nbands; ++b) {
            if (int rc = enqueue_band_upload(b))
                return rc;
            // issue everything that no longer waits for a band still to be enqueued
            for (; next < order.size(); ++next) {
                if (order[next].pass == 1 && upload_needed(order[next].band) > b)
                    break;
                if (int rc = band_pass(order[next]))
                    return rc;
            }
        }
        trace.mark("uploads enqueued");
        trace.device(1, io.s);
        for (; next < order.size(); ++next)
            if (int rc = band_pass(order[next]))
                return rc;
        trace.mark("passes enqueued");
        trace.device(2, run.s);
        T *const result = bufs[(iterations - 1) & 1];
        if (is_pageable(out))
            prefault_for_write(out, bytes);
        // results leave on their own stream: the bus is full duplex, and `io` may still be uploading
        CUDA_TRY(host_stream(device, 3, back));
        for (int64_t b = 0; b < nbands; ++b) {
            const int64_t rb = band_begin(b), re = band_begin(b + 1);
            const size_t off = (size_t)rb * (size_t)nx;
            const size_t n = (size_t)(re - rb) * (size_t)nx;
            CUDA_TRY(cudaStreamWaitEvent(back.s, done.ev[(size_t)b], 0));
            // the texture's staging band was consumed by the padding that `uploaded` covers
            CUDA_TRY(cudaStreamWaitEvent(back.s, uploaded.ev[(size_t)b], 0));
            CUDA_TRY(launch_unpad<T>(result, s_t + off, g, rb, re, 1, back.s));
            CUDA_TRY(cudaMemcpyAsync(out + off, s_t + off, n * sizeof(T), cudaMemcpyDeviceToHost, back.s));
        }
        trace.mark("downloads enqueued");
        trace.device(3, back.s);
        CUDA_TRY(cudaStreamSynchronize(back.s));
        if (prep.s)
            CUDA_TRY(cudaStreamSynchronize(prep.s));   // the sign flag is raised by the padding kernels
        if (texture_has_negative)
            CUDA_TRY(cudaMemcpyAsync(texture_has_negative, flag, sizeof(int), cudaMemcpyDeviceToHost, io.s));
        CUDA_TRY(cudaStreamSynchronize(io.s));
        CUDA_TRY(cudaStreamSynchronize(run.s));
        trace.mark("synchronised");
        return RLIC_B200_OK;
    }

    // ---- uploads, with pass 1 trailing behind them ----
    int64_t next_band = 0;   // next band of pass 1 to launch
    for (int64_t b = 0; b < nbands; ++b) {
        if (int rc = enqueue_band_upload(b))
            return rc;
        // launch every band of pass 1 whose reach is now covered
        while (next_band < nbands && !periodic_y) {
            if (upload_needed(next_band) > b)
                break;
            if (int rc = first_pass_band(next_band))
                return rc;
            ++next_band;
        }
    }
    for (; next_band < nbands; ++next_band)
        if (int rc = first_pass_band(next_band))
            return rc;

    // ---- middle passes: whole image; last pass: band by band ----
    const T *src = bufs[0];
    T *result = bufs[0];
    for (int64_t it = 1; it < iterations; ++it) {
        T *dst = bufs[it & 1];
        if (it < iterations - 1) {
            if (int rc = launch_pass<T>(src, t_field, dst, g, nfields, 0, ny, uv_mode, taps, run.s, PeerTarget<T>{},
                                        nullptr, nullptr, paths.use(it + 1)))
                return rc;
        } else {
            for (int64_t b = 0; b < nbands; ++b) {
                if (int rc = launch_pass<T>(src, t_field, dst, g, nfields, band_begin(b),
                                            band_begin(b + 1) - band_begin(b), uv_mode, taps, run.s, PeerTarget<T>{},
                                            nullptr, nullptr, paths.use(it + 1)))
                    return rc;
                CUDA_TRY(cudaEventRecord(done.ev[(size_t)b], run.s));
            }
        }
        src = dst;
        result = dst;
    }

    Range downloads("downloads behind the last pass");
    // ---- downloads, trailing behind the last pass ----
    // the passes are running: get a pageable destination's pages ready meanwhile
    if (is_pageable(out))
        prefault_for_write(out, bytes);
    for (int64_t b = 0; b < nbands; ++b) {
        const int64_t rb = band_begin(b), re = band_begin(b + 1);
        const size_t off = (size_t)rb * (size_t)nx * (size_t)nfields;
        const size_t n = (size_t)(re - rb) * (size_t)nx * (size_t)nfields;
        CUDA_TRY(cudaStreamWaitEvent(io.s, done.ev[(size_t)b], 0));
        CUDA_TRY(launch_unpad<T>(result, s_t + off, g, rb, re, nfields, io.s));
        CUDA_TRY(cudaMemcpyAsync(out + off, s_t + off, n * sizeof(T), cudaMemcpyDeviceToHost, io.s));
    }
    if (prep.s)
        CUDA_TRY(cudaStreamSynchronize(prep.s));       // the sign flag is raised by the padding kernels
    if (texture_has_negative)
        CUDA_TRY(cudaMemcpyAsync(texture_has_negative, flag, sizeof(int), cudaMemcpyDeviceToHost, io.s));
    CUDA_TRY(cudaStreamSynchronize(io.s));
    CUDA_TRY(cudaStreamSynchronize(run.s));
    return RLIC_B200_OK;
}

// `iterations` passes from a dense device texture to a dense device result,
// field already packed (lib.rs:432-440 without the copy-back: two padded work
// buffers swap roles).
template <typename T>
int run_device(const T *d_tex, const Field<T> *d_field, int64_t ny, int64_t nx, const TapSet<T> &taps,
               int uv_mode, const Walls &w, int64_t iterations, T *d_out, cudaStream_t s,
               int64_t nfields = 1)
{
    Range whole("rlic_b200 passes (device)");
    const PassGeom g = make_geometry(ny, nx, Slab{0, ny, 0, 0}, w);
    const size_t padded_bytes = (size_t)g.field_stride * (size_t)nfields * sizeof(T);
    CUDA_TRY(use_current_device());
    DeviceBuf a, b;   // stream-ordered scratch, returned to the pool when the work is enqueued
    CUDA_TRY(a.alloc(padded_bytes, s));
    if (iterations > 1)
        CUDA_TRY(b.alloc(padded_bytes, s));
    DeviceBuf in;
    CUDA_TRY(in.alloc(padded_bytes, s));
    CallPaths paths;
    CUDA_TRY(paths.prepare(iterations, g.field_stride * nfields, taps.ntaps, s));
    CUDA_TRY(launch_pad<T>(d_tex, static_cast<T *>(in.p), g, 0, ny, nfields, nullptr, s));
    const T *src = static_cast<const T *>(in.p);
    T *dst = static_cast<T *>(a.p);
    bool wrote_dense = false;
    for (int64_t it = 0; it < iterations; ++it) {
        // pass 1: in -> a; pass 2: a -> b; pass 3: b -> a; ...
        dst = static_cast<T *>(it == 0 ? a.p : ((it & 1) ? b.p : a.p));
        // a small image's last pass writes the dense result itself (no un-padding launch)
        if (int rc = launch_pass<T>(src, d_field, dst, g, nfields, 0, ny, uv_mode, taps, s, PeerTarget<T>{},
                                    it == iterations - 1 ? d_out : nullptr, &wrote_dense, paths.use(it + 1)))
            return rc;
        src = dst;
    }
    if (!wrote_dense)
        CUDA_TRY(launch_unpad<T>(dst, d_out, g, 0, ny, nfields, s));
    return RLIC_B200_OK;
}

template <typename T>
int check_device_call(const void *a, const void *b, const void *c, const void *kernel, const void *out,
                      int64_t ny, int64_t nx, int64_t klen, int uv_mode, const Walls &w)
{
    if (int rc = check_common(ny, nx, klen, uv_mode, w))
        return rc;
    if (ny == 0 || nx == 0)
        return RLIC_B200_OK;
    if (!a || !b || !c || !kernel || !out)
        return fail(RLIC_B200_EINVAL, "null pointer argument");
    return RLIC_B200_OK;
}

template <typename T>
int convolve_device(const T *d_tex, const T *d_u, const T *d_v, int64_t ny, int64_t nx,
                    const T *kernel, int64_t klen, int uv_mode, const Walls &w,
                    int64_t iterations, T *d_out, void *stream, int64_t nfields = 1)
{
    if (int rc = check_device_call<T>(d_tex, d_u, d_v, kernel, d_out, ny, nx, klen, uv_mode, w))
        return rc;
    if (nfields < 0)
        return fail(RLIC_B200_EINVAL, "negative field count");
    if (ny == 0 || nx == 0 || nfields == 0)
        return RLIC_B200_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t count = (size_t)ny * (size_t)nx * (size_t)nfields;
    if (iterations <= 0) {
        CUDA_TRY(cudaMemsetAsync(d_out, 0, sizeof(T) * count, s));
        return RLIC_B200_OK;
    }
    CUDA_TRY(use_current_device());
    TapSet<T> taps;
    CUDA_TRY(taps.prepare(kernel, klen, s));
    const PassGeom g = make_geometry(ny, nx, Slab{0, ny, 0, 0}, w);
    DeviceBuf d_field;
    CUDA_TRY(d_field.alloc(4 * sizeof(T) * (size_t)g.field_stride * (size_t)nfields, s));
    CUDA_TRY(launch_pack<T>(d_u, d_v, static_cast<Field<T> *>(d_field.p), g, 0, ny, nfields, s));
    return run_device<T>(d_tex, static_cast<Field<T> *>(d_field.p), ny, nx, taps, uv_mode, w,
                         iterations, d_out, s, nfields);
}

template <typename T>
int pack_field(const T *d_u, const T *d_v, int64_t ny, int64_t nx, const Walls &w, T *d_field,
               void *stream)
{
    if (int rc = check_device_call<T>(d_u, d_v, d_field, d_field, d_field, ny, nx, 1, 0, w))
        return rc;
    if (ny == 0 || nx == 0)
        return RLIC_B200_OK;
    const PassGeom g = make_geometry(ny, nx, Slab{0, ny, 0, 0}, w);
    CUDA_TRY(launch_pack<T>(d_u, d_v, reinterpret_cast<Field<T> *>(d_field), g, 0, ny, 1,
                            static_cast<cudaStream_t>(stream)));
    return RLIC_B200_OK;
}

template <typename T>
int convolve_packed(const T *d_tex, const T *d_field, int64_t ny, int64_t nx, const T *kernel,
                    int64_t klen, int uv_mode, const Walls &w, int64_t iterations, T *d_out,
                    void *stream)
{
    if (int rc = check_device_call<T>(d_tex, d_field, d_field, kernel, d_out, ny, nx, klen, uv_mode, w))
        return rc;
    if (ny == 0 || nx == 0)
        return RLIC_B200_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (iterations <= 0) {
        CUDA_TRY(cudaMemsetAsync(d_out, 0, sizeof(T) * (size_t)ny * (size_t)nx, s));
        return RLIC_B200_OK;
    }
    TapSet<T> taps;
    CUDA_TRY(taps.prepare(kernel, klen, s));
    return run_device<T>(d_tex, reinterpret_cast<const Field<T> *>(d_field), ny, nx, taps, uv_mode, w,
                         iterations, d_out, s);
}

// ---- slab building blocks (padded buffers owned by the caller) ----
// Owned rows [sub0, sub0 + subn) of a slab (the whole slab: 0, sl.nrows).
int check_sub_rows(const Slab &sl, int64_t sub0, int64_t subn)
{
    if (sub0 < 0 || subn < 0 || sub0 + subn > sl.nrows)
        return fail(RLIC_B200_ESHARD, "rows [%lld,%lld) outside the slab's %lld rows",
                    (long long)sub0, (long long)(sub0 + subn), (long long)sl.nrows);
    return RLIC_B200_OK;
}

template <typename T>
int slab_pack_field(const T *d_u, const T *d_v, int64_t ny, int64_t nx, const Slab &sl, const Walls &w,
                    T *d_field, void *stream, int64_t sub0 = 0, int64_t subn = -1)
{
    if (subn < 0)
        subn = sl.nrows;
    if (int rc = check_common(ny, nx, 1, 0, w))
        return rc;
    if (int rc = check_slab(ny, 1, sl, w))
        return rc;
    if (int rc = check_sub_rows(sl, sub0, subn))
        return rc;
    if (subn == 0 || nx == 0)
        return RLIC_B200_OK;
    if (!d_u || !d_v || !d_field)
        return fail(RLIC_B200_EINVAL, "null pointer argument");
    const PassGeom g = make_geometry(ny, nx, sl, w);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    Field<T> *field = reinterpret_cast<Field<T> *>(d_field);
    // owned rows from the planar components, with the sentinels of any image
    // wall they touch; halo rows arrive from the neighbours, already packed
    CUDA_TRY(launch_pack<T>(d_u, d_v, field, g, sl.halo_lo + sub0, sl.halo_lo + sub0 + subn, 1, s));
    return RLIC_B200_OK;
}

template <typename T>
int slab_pad_texture(const T *d_tex, int64_t ny, int64_t nx, const Slab &sl, const Walls &w,
                     T *d_padded, void *stream, int64_t sub0 = 0, int64_t subn = -1)
{
    if (subn < 0)
        subn = sl.nrows;
    if (int rc = check_common(ny, nx, 1, 0, w))
        return rc;
    if (int rc = check_slab(ny, 1, sl, w))
        return rc;
    if (int rc = check_sub_rows(sl, sub0, subn))
        return rc;
    if (subn == 0 || nx == 0)
        return RLIC_B200_OK;
    if (!d_tex || !d_padded)
        return fail(RLIC_B200_EINVAL, "null pointer argument");
    const PassGeom g = make_geometry(ny, nx, sl, w);
    CUDA_TRY(launch_pad<T>(d_tex, d_padded, g, sl.halo_lo + sub0, sl.halo_lo + sub0 + subn, 1, nullptr,
                           static_cast<cudaStream_t>(stream)));
    return RLIC_B200_OK;
}

template <typename T>
int slab_unpad_texture(const T *d_padded, int64_t ny, int64_t nx, const Slab &sl, const Walls &w,
                       T *d_tex, void *stream, int64_t sub0 = 0, int64_t subn = -1)
{
    if (subn < 0)
        subn = sl.nrows;
    if (int rc = check_common(ny, nx, 1, 0, w))
        return rc;
    if (int rc = check_slab(ny, 1, sl, w))
        return rc;
    if (int rc = check_sub_rows(sl, sub0, subn))
        return rc;
    if (subn == 0 || nx == 0)
        return RLIC_B200_OK;
    if (!d_tex || !d_padded)
        return fail(RLIC_B200_EINVAL, "null pointer argument");
    const PassGeom g = make_geometry(ny, nx, sl, w);
    CUDA_TRY(launch_unpad<T>(d_padded, d_tex, g, sl.halo_lo + sub0, sl.halo_lo + sub0 + subn, 1,
                             static_cast<cudaStream_t>(stream)));
    return RLIC_B200_OK;
}

template <typename T>
int pass_slab(const T *d_tex, const T *d_field, T *d_out, int64_t ny, int64_t nx, const Slab &sl,
              int64_t sub0, int64_t subn, const T *kernel, int64_t klen, int uv_mode,
              const Walls &w, void *stream, T *peer_out = nullptr, int64_t peer_row_delta = 0,
              int paths_mode = RLIC_B200_PASS_WALK, uint32_t *d_paths = nullptr)
{
    if (paths_mode != RLIC_B200_PASS_WALK && paths_mode != RLIC_B200_PASS_RECORD &&
        paths_mode != RLIC_B200_PASS_REPLAY)
        return fail(RLIC_B200_EINVAL, "unknown paths mode %d", paths_mode);
    if (int rc = check_common(ny, nx, klen, uv_mode, w))
        return rc;
    if (int rc = check_slab(ny, klen, sl, w))
        return rc;
    if (int rc = check_sub_rows(sl, sub0, subn))
        return rc;
    if (subn == 0 || nx == 0)
        return RLIC_B200_OK;
    // (a replayed pass does not look at the field)
    if (!d_tex || (!d_field && paths_mode != RLIC_B200_PASS_REPLAY) || !d_out || !kernel ||
        (paths_mode != RLIC_B200_PASS_WALK && !d_paths))
        return fail(RLIC_B200_EINVAL, "null pointer argument");
    const PassGeom g = make_geometry(ny, nx, sl, w);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    TapSet<T> taps;
    CUDA_TRY(taps.prepare(kernel, klen, s));
    PeerTarget<T> peer;
    peer.out = peer_out;
    peer.delta = (long long)peer_row_delta * g.pitch;
    PathUse paths;
    if (paths_mode != RLIC_B200_PASS_WALK)
        paths = PathUse{paths_mode == RLIC_B200_PASS_RECORD ? Paths::record : Paths::replay,
                        reinterpret_cast<uint4 *>(d_paths), (long long)g.field_stride};
    return launch_pass<T>(d_tex, reinterpret_cast<const Field<T> *>(d_field), d_out, g, 1,
                          sl.halo_lo + sub0, subn, uv_mode, taps, s, peer, nullptr, nullptr, paths);
}

// Whole fields split over devices.  Two host threads per device take chunks
// alternately: while one chunk is being computed or downloaded, the next one is
// already uploading.
template <typename T>
int convolve_batch(const T *tex, const T *u, const T *v, int64_t nfields, int64_t ny, int64_t nx,
                   const T *kernel, int64_t klen, int uv_mode, const Walls &w,
                   int64_t iterations, const int *devices, int ndev, T *out,
                   int *texture_has_negative = nullptr)
{
    if (texture_has_negative)
        *texture_has_negative = 0;
    if (int rc = check_common(ny, nx, klen, uv_mode, w))
        return rc;
    if (nfields < 0)
        return fail(RLIC_B200_EINVAL, "negative field count");
    if (nfields == 0 || ny == 0 || nx == 0)
        return RLIC_B200_OK;
    int visible = 0;
    CUDA_TRY(cudaGetDeviceCount(&visible));
    std::vector<int> devs;
    if (devices && ndev > 0) {
        for (int d = 0; d < ndev; ++d) {
            if (devices[d] < 0 || devices[d] >= visible)
                return fail(RLIC_B200_ESHARD, "device %d not visible (%d devices)", devices[d], visible);
            devs.push_back(devices[d]);
        }
    } else {
        for (int d = 0; d < visible; ++d)
            devs.push_back(d);
    }
    if (devs.empty())
        return fail(RLIC_B200_ENODEVICE, "no CUDA device");
    const int64_t nd = (int64_t)devs.size();
    const size_t field_elems = (size_t)ny * (size_t)nx;
    // chunk: enough fields to fill the GPU (~16 Mpix) but at least 1
    const int64_t chunk = std::max<int64_t>(1, (int64_t)((size_t)(16u << 20) / field_elems));

    const int lanes = 2;
    const ThreadChoices mine{effective_arithmetic(), effective_schedule(), effective_walk(), effective_paths()};
    std::atomic<int> any_negative{0};
    std::vector<int> rcs(devs.size() * lanes, 0);
    std::vector<std::string> msgs(devs.size() * lanes);
    std::vector<std::atomic<int64_t>> cursor(devs.size());
    std::vector<std::thread> threads;
    for (int64_t d = 0; d < nd; ++d) {
        const int64_t f0 = nfields * d / nd, f1 = nfields * (d + 1) / nd;
        cursor[(size_t)d].store(f0);
        for (int lane = 0; lane < lanes; ++lane) {
            threads.emplace_back([&, d, f1, lane]() {
                tls_choices = mine;   // the caller's choices, not this fresh thread's defaults
                int rc = 0;
                while (!rc) {
                    const int64_t f = cursor[(size_t)d].fetch_add(chunk);
                    if (f >= f1)
                        break;
                    const int64_t n = std::min(chunk, f1 - f);
                    const size_t off = (size_t)f * field_elems;
                    int negative = 0;
                    rc = convolve_host<T>(tex + off, u + off, v + off, n, ny, nx, kernel, klen,
                                          uv_mode, w, iterations, out + off, devs[(size_t)d],
                                          texture_has_negative ? &negative : nullptr);
                    if (negative)
                        any_negative.store(1, std::memory_order_relaxed);
                }
                rcs[(size_t)d * lanes + lane] = rc;
                if (rc)
                    msgs[(size_t)d * lanes + lane] = tls_error;
            });
        }
    }
    for (auto &t : threads)
        t.join();
    for (size_t i = 0; i < rcs.size(); ++i)
        if (rcs[i]) {
            tls_error = msgs[i];
            return rcs[i];
        }
    if (texture_has_negative)
        *texture_has_negative = any_negative.load();
    return RLIC_B200_OK;
}

template <typename T>
int equalize_device(const T *d_image, int64_t ny, int64_t nx, int64_t nbins, T *d_out, cudaStream_t s)
{
    if (ny < 0 || nx < 0)
        return fail(RLIC_B200_EINVAL, "negative image size %lld x %lld", (long long)ny, (long long)nx);
    if (nbins < 1 || nbins > ((int64_t)1 << 24))
        return fail(RLIC_B200_EINVAL, "nbins must be between 1 and 2^24, got %lld", (long long)nbins);
    const long long n = (long long)ny * nx;
    if (n == 0)
        return RLIC_B200_OK;
    if (!d_image || !d_out)
        return fail(RLIC_B200_EINVAL, "null pointer argument");
    Range whole("rlic_b200 equalize_histogram");
    CUDA_TRY(use_current_device());
    // scratch: extrema and count, 64-bit counts per bin, the cumulative distribution
    DeviceBuf scratch;
    const size_t hist_off = 64, cdf_off = hist_off + sizeof(unsigned long long) * (size_t)nbins;
    CUDA_TRY(scratch.alloc(cdf_off + sizeof(T) * (size_t)nbins, s));
    char *base = static_cast<char *>(scratch.p);
    auto *st = reinterpret_cast<rlic::EqualizeScratch *>(base);
    auto *hist = reinterpret_cast<unsigned long long *>(base + hist_off);
    T *cdf = reinterpret_cast<T *>(base + cdf_off);
    const unsigned blocks = stream_blocks(n);
    rlic::equalize_init_kernel<<<stream_blocks(nbins), 256, 0, s>>>(st, hist, nbins);
    rlic::equalize_extrema_kernel<T><<<blocks, 256, 0, s>>>(d_image, n, st);
    rlic::equalize_histogram_kernel<T><<<blocks, 256, 0, s>>>(d_image, n, st, hist, nbins);
    rlic::equalize_cdf_kernel<T><<<1, 1024, 0, s>>>(hist, nbins, st, cdf);
    rlic::equalize_map_kernel<T><<<blocks, 256, 0, s>>>(d_image, n, st, cdf, nbins, d_out);
    g_launches.fetch_add(5, std::memory_order_relaxed);
    CUDA_TRY(cudaGetLastError());
    return RLIC_B200_OK;
}

template <typename T>
int equalize_host(const T *image, int64_t ny, int64_t nx, int64_t nbins, T *out, int device)
{
    if (ny < 0 || nx < 0)
        return fail(RLIC_B200_EINVAL, "negative image size %lld x %lld", (long long)ny, (long long)nx);
    const size_t bytes = (size_t)ny * (size_t)nx * sizeof(T);
    if (bytes == 0)
        return RLIC_B200_OK;
    if (!image || !out)
        return fail(RLIC_B200_EINVAL, "null pointer argument");
    CUDA_TRY(use_device(device));
    Stream st;
    CUDA_TRY(cudaStreamCreateWithFlags(&st.s, cudaStreamNonBlocking));
    DeviceBuf d_in, d_out;
    StreamDrain drain{{&st, nullptr, nullptr, nullptr}};
    CUDA_TRY(d_in.alloc(bytes, st.s));
    CUDA_TRY(d_out.alloc(bytes, st.s));
    const HostToDevice job{d_in.p, image, bytes};
    CUDA_TRY(upload(&job, 1, st.s));
    if (int rc = equalize_device<T>(static_cast<const T *>(d_in.p), ny, nx, nbins, static_cast<T *>(d_out.p), st.s))
        return rc;
    if (is_pageable(out))
        prefault_for_write(out, bytes);
    CUDA_TRY(cudaMemcpyAsync(out, d_out.p, bytes, cudaMemcpyDeviceToHost, st.s));
    CUDA_TRY(cudaStreamSynchronize(st.s));
    return RLIC_B200_OK;
}
}  // namespace

extern "C" {

int rlic_b200_abi_version(void) { return RLIC_B200_ABI_VERSION; }

const char *rlic_b200_last_error(void) { return tls_error.c_str(); }

int rlic_b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int64_t rlic_b200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int64_t rlic_b200_padded_cells(int64_t rows, int64_t nx) { return rlic::padded_cells(rows, nx); }

int rlic_b200_debug_wall_cell(int64_t ny, int64_t nx, int64_t row0, int64_t nrows, int64_t halo_lo,
                              int64_t halo_hi, int x_left, int x_right, int y_left, int y_right,
                              int64_t cell, int64_t *out)
{
    tls_error.clear();
    const Walls w{x_left, x_right, y_left, y_right};
    if (int rc = check_common(ny, nx, 1, 0, w))
        return rc;
    const Slab sl{row0, nrows, halo_lo, halo_hi};
    if (int rc = check_slab(ny, 1, sl, w))
        return rc;
    if (!out)
        return fail(RLIC_B200_EINVAL, "null pointer argument");
    const PassGeom g = make_geometry(ny, nx, sl, w);
    if (cell < 0 || cell >= g.field_stride)
        return fail(RLIC_B200_EINVAL, "cell %lld outside the buffer of %lld cells", (long long)cell,
                    (long long)g.field_stride);
    const rlic::CellSource s = rlic::cell_source(cell, g);
    out[0] = s.pixel;
    out[1] = s.reachable;
    out[2] = s.row;
    out[3] = s.col;
    out[4] = s.shift;
    return RLIC_B200_OK;
}

int rlic_b200_set_arithmetic(int which)
{
    tls_error.clear();
    if (which != RLIC_B200_ARITH_FMA_BRANCHLESS && which != RLIC_B200_ARITH_FMA)
        return fail(RLIC_B200_EINVAL, "unknown arithmetic %d", which);
    g_arithmetic.store(which, std::memory_order_relaxed);
    return RLIC_B200_OK;
}

int rlic_b200_get_arithmetic(void) { return g_arithmetic.load(std::memory_order_relaxed); }

int rlic_b200_set_walk(int which)
{
    tls_error.clear();
    if (which != RLIC_B200_WALK_PER_STEP && which != RLIC_B200_WALK_GROUPED)
        return fail(RLIC_B200_EINVAL, "unknown walk %d", which);
    g_walk.store(which, std::memory_order_relaxed);
    return RLIC_B200_OK;
}

int rlic_b200_get_walk(void) { return g_walk.load(std::memory_order_relaxed); }

int rlic_b200_set_schedule(int which)
{
    tls_error.clear();
    if (which != RLIC_B200_SCHEDULE_TRAILING && which != RLIC_B200_SCHEDULE_WAVEFRONT)
        return fail(RLIC_B200_EINVAL, "unknown schedule %d", which);
    g_schedule.store(which, std::memory_order_relaxed);
    return RLIC_B200_OK;
}

int rlic_b200_get_schedule(void) { return g_schedule.load(std::memory_order_relaxed); }

int rlic_b200_set_thread_options(int arithmetic, int schedule, int walk)
{
    tls_error.clear();
    if (arithmetic != -1 && arithmetic != RLIC_B200_ARITH_FMA_BRANCHLESS && arithmetic != RLIC_B200_ARITH_FMA)
        return fail(RLIC_B200_EINVAL, "unknown arithmetic %d", arithmetic);
    if (schedule != -1 && schedule != RLIC_B200_SCHEDULE_TRAILING && schedule != RLIC_B200_SCHEDULE_WAVEFRONT)
        return fail(RLIC_B200_EINVAL, "unknown schedule %d", schedule);
    if (walk != -1 && walk != RLIC_B200_WALK_PER_STEP && walk != RLIC_B200_WALK_GROUPED)
        return fail(RLIC_B200_EINVAL, "unknown walk %d", walk);
    tls_choices.arithmetic = arithmetic;
    tls_choices.schedule = schedule;
    tls_choices.walk = walk;
    return RLIC_B200_OK;
}

int rlic_b200_set_paths(int which)
{
    tls_error.clear();
    if (which != RLIC_B200_PATHS_RECOMPUTE && which != RLIC_B200_PATHS_REPLAY)
        return fail(RLIC_B200_EINVAL, "unknown paths choice %d", which);
    g_paths.store(which, std::memory_order_relaxed);
    return RLIC_B200_OK;
}

int rlic_b200_get_paths(void) { return g_paths.load(std::memory_order_relaxed); }

int rlic_b200_set_thread_paths(int which)
{
    tls_error.clear();
    if (which != -1 && which != RLIC_B200_PATHS_RECOMPUTE && which != RLIC_B200_PATHS_REPLAY)
        return fail(RLIC_B200_EINVAL, "unknown paths choice %d", which);
    tls_choices.paths = which;
    return RLIC_B200_OK;
}

int rlic_b200_get_thread_paths(void) { return tls_choices.paths; }

int rlic_b200_get_effective_paths(void) { return effective_paths(); }

int64_t rlic_b200_path_record_bytes(int64_t rows, int64_t nx, int64_t klen)
{
    if (rows < 0 || nx < 0 || klen <= 0)
        return 0;
    return (int64_t)(path_record_words(rlic::padded_cells(rows, nx), klen) * sizeof(unsigned));
}

void rlic_b200_get_thread_options(int *arithmetic, int *schedule, int *walk)
{
    if (arithmetic) *arithmetic = tls_choices.arithmetic;
    if (schedule) *schedule = tls_choices.schedule;
    if (walk) *walk = tls_choices.walk;
}

void rlic_b200_get_effective_options(int *arithmetic, int *schedule, int *walk)
{
    if (arithmetic) *arithmetic = effective_arithmetic();
    if (schedule) *schedule = effective_schedule();
    if (walk) *walk = effective_walk();
}

int64_t rlic_b200_debug_band_plan(int64_t ny, int64_t nx, int64_t klen, int64_t iterations, int64_t *edges,
                                  int64_t capacity)
{
    if (ny <= 0 || nx <= 0 || klen <= 0)
        return 0;
    const std::vector<int64_t> plan = band_plan(ny, nx, klen / 2, iterations);
    for (size_t k = 0; k < plan.size() && (int64_t)k < capacity && edges; ++k)
        edges[k] = plan[k];
    return (int64_t)plan.size();
}

int64_t rlic_b200_debug_wavefront_order(int64_t nbands, int64_t iterations, int32_t *pass_band,
                                        int64_t capacity)
{
    if (nbands <= 0 || iterations <= 0 || nbands * iterations > (int64_t)1 << 24)
        return 0;
    const std::vector<BandPass> order = wavefront_order(nbands, iterations);
    for (size_t k = 0; k < order.size() && (int64_t)k < capacity && pass_band; ++k) {
        pass_band[2 * k] = order[k].pass;
        pass_band[2 * k + 1] = order[k].band;
    }
    return (int64_t)order.size();
}

int rlic_b200_debug_geometry(int64_t ny, int64_t nx, int64_t row0, int64_t nrows, int64_t halo_lo,
                             int64_t halo_hi, int x_left, int x_right, int y_left, int y_right,
                             int64_t klen, int64_t *out)
{
    tls_error.clear();
    const Walls w{x_left, x_right, y_left, y_right};
    if (int rc = check_common(ny, nx, klen, 0, w))
        return rc;
    const Slab sl{row0, nrows, halo_lo, halo_hi};
    if (int rc = check_slab(ny, klen, sl, w))
        return rc;
    if (!out)
        return fail(RLIC_B200_EINVAL, "null pointer argument");
    const PassGeom g = make_geometry(ny, nx, sl, w);
    out[0] = g.nx;
    out[1] = g.pitch;
    out[2] = g.rows;
    out[3] = g.field_stride;
    out[4] = g.j_below_to;
    out[5] = g.j_above_to;
    out[6] = g.i_below_to;
    out[7] = g.i_above_to;
    out[8] = g.lo_wall;
    out[9] = g.hi_wall;
    return RLIC_B200_OK;
}

void *rlic_b200_result_alloc(int64_t bytes)
{
    if (bytes <= 0 || rlic_b200_device_count() == 0)
        return nullptr;
    return result_blocks().acquire((size_t)bytes);
}

void rlic_b200_result_free(void *block)
{
    if (block)
        result_blocks().release(block);
}

void rlic_b200_debug_force_wide_index(int on) { g_force_wide.store(on ? 1 : 0); }

void rlic_b200_debug_small_image_kernel(int on) { g_small_images.store(on ? 1 : 0); }

void rlic_b200_debug_replay_staging(int on) { g_replay_staging.store(on ? 1 : 0); }

int rlic_b200_set_device(int device)
{
    int n = rlic_b200_device_count();
    if (n == 0)
        return fail(RLIC_B200_ENODEVICE, "no CUDA device is visible");
    if (device < 0 || device >= n)
        return fail(RLIC_B200_EINVAL, "device %d out of range (%d visible)", device, n);
    tls_device = device;
    return RLIC_B200_OK;
}

#define RLIC_DEFINE(T, sfx)                                                                      \
    int rlic_b200_convolve_##sfx(const T *texture, const T *u, const T *v, int64_t ny,           \
                                 int64_t nx, const T *kernel, int64_t klen, int uv_mode,         \
                                 int x_left, int x_right, int y_left, int y_right,               \
                                 int64_t iterations, T *out)                                     \
    {                                                                                            \
        tls_error.clear();                                                                       \
        return convolve_host<T>(texture, u, v, 1, ny, nx, kernel, klen, uv_mode,                 \
                                Walls{x_left, x_right, y_left, y_right}, iterations, out,        \
                                tls_device);                                                     \
    }                                                                                            \
    int rlic_b200_convolve_checked_##sfx(const T *texture, const T *u, const T *v, int64_t ny,   \
                                         int64_t nx, const T *kernel, int64_t klen, int uv_mode, \
                                         int x_left, int x_right, int y_left, int y_right,       \
                                         int64_t iterations, T *out, int *texture_has_negative)  \
    {                                                                                            \
        tls_error.clear();                                                                       \
        if (!texture_has_negative)                                                               \
            return fail(RLIC_B200_EINVAL, "texture_has_negative is null");                       \
        return convolve_host<T>(texture, u, v, 1, ny, nx, kernel, klen, uv_mode,                 \
                                Walls{x_left, x_right, y_left, y_right}, iterations, out,        \
                                tls_device, texture_has_negative);                               \
    }                                                                                            \
    int rlic_b200_convolve_batch_##sfx(const T *texture, const T *u, const T *v,                 \
                                       int64_t nfields, int64_t ny, int64_t nx, const T *kernel, \
                                       int64_t klen, int uv_mode, int x_left, int x_right,       \
                                       int y_left, int y_right, int64_t iterations,              \
                                       const int *devices, int ndev, T *out)                     \
    {                                                                                            \
        tls_error.clear();                                                                       \
        return convolve_batch<T>(texture, u, v, nfields, ny, nx, kernel, klen, uv_mode,          \
                                 Walls{x_left, x_right, y_left, y_right}, iterations, devices,   \
                                 ndev, out);                                                     \
    }                                                                                            \
    int rlic_b200_convolve_batch_checked_##sfx(const T *texture, const T *u, const T *v,         \
                                               int64_t nfields, int64_t ny, int64_t nx,          \
                                               const T *kernel, int64_t klen, int uv_mode,       \
                                               int x_left, int x_right, int y_left, int y_right, \
                                               int64_t iterations, const int *devices, int ndev, \
                                               T *out, int *texture_has_negative)                \
    {                                                                                            \
        tls_error.clear();                                                                       \
        if (!texture_has_negative)                                                               \
            return fail(RLIC_B200_EINVAL, "texture_has_negative is null");                       \
        return convolve_batch<T>(texture, u, v, nfields, ny, nx, kernel, klen, uv_mode,          \
                                 Walls{x_left, x_right, y_left, y_right}, iterations, devices,   \
                                 ndev, out, texture_has_negative);                               \
    }                                                                                            \
    int rlic_b200_convolve_device_##sfx(const T *d_texture, const T *d_u, const T *d_v,          \
                                        int64_t ny, int64_t nx, const T *kernel, int64_t klen,   \
                                        int uv_mode, int x_left, int x_right, int y_left,        \
                                        int y_right, int64_t iterations, T *d_out, void *stream) \
    {                                                                                            \
        tls_error.clear();                                                                       \
        return convolve_device<T>(d_texture, d_u, d_v, ny, nx, kernel, klen, uv_mode,            \
                                  Walls{x_left, x_right, y_left, y_right}, iterations, d_out,    \
                                  stream);                                                       \
    }                                                                                            \
    int rlic_b200_convolve_device_batch_##sfx(const T *d_texture, const T *d_u, const T *d_v,    \
                                              int64_t nfields, int64_t ny, int64_t nx,           \
                                              const T *kernel, int64_t klen, int uv_mode,        \
                                              int x_left, int x_right, int y_left, int y_right,  \
                                              int64_t iterations, T *d_out, void *stream)        \
    {                                                                                            \
        tls_error.clear();                                                                       \
        return convolve_device<T>(d_texture, d_u, d_v, ny, nx, kernel, klen, uv_mode,            \
                                  Walls{x_left, x_right, y_left, y_right}, iterations, d_out,    \
                                  stream, nfields);                                              \
    }                                                                                            \
    int rlic_b200_pack_field_##sfx(const T *d_u, const T *d_v, int64_t ny, int64_t nx,           \
                                   int x_left, int x_right, int y_left, int y_right,             \
                                   T *d_field, void *stream)                                     \
    {                                                                                            \
        tls_error.clear();                                                                       \
        return pack_field<T>(d_u, d_v, ny, nx, Walls{x_left, x_right, y_left, y_right}, d_field, \
                             stream);                                                            \
    }                                                                                            \
    int rlic_b200_convolve_packed_##sfx(const T *d_texture, const T *d_field, int64_t ny,        \
                                        int64_t nx, const T *kernel, int64_t klen, int uv_mode,  \
                                        int x_left, int x_right, int y_left, int y_right,        \
                                        int64_t iterations, T *d_out, void *stream)              \
    {                                                                                            \
        tls_error.clear();                                                                       \
        return convolve_packed<T>(d_texture, d_field, ny, nx, kernel, klen, uv_mode,             \
                                  Walls{x_left, x_right, y_left, y_right}, iterations, d_out,    \
                                  stream);                                                       \
    }                                                                                            \
    int rlic_b200_slab_pack_field_##sfx(const T *d_u, const T *d_v, int64_t ny, int64_t nx,      \
                                        int64_t row0, int64_t nrows, int64_t halo_lo,            \
                                        int64_t halo_hi, int x_left, int x_right, int y_left,    \
                                        int y_right, T *d_field, void *stream)                   \
    {                                                                                            \
        tls_error.clear();                                                                       \
        return slab_pack_field<T>(d_u, d_v, ny, nx, Slab{row0, nrows, halo_lo, halo_hi},         \
                                  Walls{x_left, x_right, y_left, y_right}, d_field, stream);     \
    }                                                                                            \
    int rlic_b200_slab_pad_texture_##sfx(const T *d_texture, int64_t ny, int64_t nx,             \
                                         int64_t row0, int64_t nrows, int64_t halo_lo,           \
                                         int64_t halo_hi, int x_left, int x_right, int y_left,   \
                                         int y_right, T *d_padded, void *stream)                 \
    {                                                                                            \
        tls_error.clear();                                                                       \
        return slab_pad_texture<T>(d_texture, ny, nx, Slab{row0, nrows, halo_lo, halo_hi},       \
                                   Walls{x_left, x_right, y_left, y_right}, d_padded, stream);   \
    }                                                                                            \
    int rlic_b200_slab_unpad_texture_##sfx(const T *d_padded, int64_t ny, int64_t nx,            \
                                           int64_t row0, int64_t nrows, int64_t halo_lo,         \
                                           int64_t halo_hi, int x_left, int x_right, int y_left, \
                                           int y_right, T *d_texture, void *stream)              \
    {                                                                                            \
        tls_error.clear();                                                                       \
        return slab_unpad_texture<T>(d_padded, ny, nx, Slab{row0, nrows, halo_lo, halo_hi},      \
                                     Walls{x_left, x_right, y_left, y_right}, d_texture,         \
                                     stream);                                                    \
    }                                                                                            \
    int rlic_b200_pass_slab_##sfx(const T *d_texture, const T *d_field, T *d_out, int64_t ny,    \
                                  int64_t nx, int64_t row0, int64_t nrows, int64_t halo_lo,      \
                                  int64_t halo_hi, int64_t sub_row0, int64_t sub_nrows,          \
                                  const T *kernel, int64_t klen, int uv_mode, int x_left,        \
                                  int x_right, int y_left, int y_right, void *stream)            \
    {                                                                                            \
        tls_error.clear();                                                                       \
        return pass_slab<T>(d_texture, d_field, d_out, ny, nx,                                   \
                            Slab{row0, nrows, halo_lo, halo_hi}, sub_row0, sub_nrows, kernel,    \
                            klen, uv_mode, Walls{x_left, x_right, y_left, y_right}, stream);     \
    }

RLIC_DEFINE(float, f32)
RLIC_DEFINE(double, f64)

// Row-range variants of the three layout conversions: owned rows [sub_row0, sub_row0 + sub_nrows)
// only, the dense array holding just those rows.  They let a caller upload, convert, compute
// and download a slab band by band (rlic_b200/sharded.py: convolve_host).
#define RLIC_DEFINE_ROWS(T, sfx)                                                                 \
    int rlic_b200_slab_pack_field_rows_##sfx(const T *d_u, const T *d_v, int64_t ny, int64_t nx, \
                                             int64_t row0, int64_t nrows, int64_t halo_lo,       \
                                             int64_t halo_hi, int64_t sub_row0,                  \
                                             int64_t sub_nrows, int x_left, int x_right,         \
                                             int y_left, int y_right, T *d_field, void *stream)  \
    {                                                                                            \
        tls_error.clear();                                                                       \
        if (sub_nrows < 0)                                                                       \
            return fail(RLIC_B200_ESHARD, "negative row count");                                 \
        return slab_pack_field<T>(d_u, d_v, ny, nx, Slab{row0, nrows, halo_lo, halo_hi},         \
                                  Walls{x_left, x_right, y_left, y_right}, d_field, stream,      \
                                  sub_row0, sub_nrows);                                          \
    }                                                                                            \
    int rlic_b200_slab_pad_texture_rows_##sfx(const T *d_texture, int64_t ny, int64_t nx,        \
                                              int64_t row0, int64_t nrows, int64_t halo_lo,      \
                                              int64_t halo_hi, int64_t sub_row0,                 \
                                              int64_t sub_nrows, int x_left, int x_right,        \
                                              int y_left, int y_right, T *d_padded, void *stream) \
    {                                                                                            \
        tls_error.clear();                                                                       \
        if (sub_nrows < 0)                                                                       \
            return fail(RLIC_B200_ESHARD, "negative row count");                                 \
        return slab_pad_texture<T>(d_texture, ny, nx, Slab{row0, nrows, halo_lo, halo_hi},       \
                                   Walls{x_left, x_right, y_left, y_right}, d_padded, stream,    \
                                   sub_row0, sub_nrows);                                         \
    }                                                                                            \
    int rlic_b200_slab_unpad_texture_rows_##sfx(const T *d_padded, int64_t ny, int64_t nx,       \
                                                int64_t row0, int64_t nrows, int64_t halo_lo,    \
                                                int64_t halo_hi, int64_t sub_row0,               \
                                                int64_t sub_nrows, int x_left, int x_right,      \
                                                int y_left, int y_right, T *d_texture,           \
                                                void *stream)                                    \
    {                                                                                            \
        tls_error.clear();                                                                       \
        if (sub_nrows < 0)                                                                       \
            return fail(RLIC_B200_ESHARD, "negative row count");                                 \
        return slab_unpad_texture<T>(d_padded, ny, nx, Slab{row0, nrows, halo_lo, halo_hi},      \
                                     Walls{x_left, x_right, y_left, y_right}, d_texture, stream, \
                                     sub_row0, sub_nrows);                                       \
    }
RLIC_DEFINE_ROWS(float, f32)
RLIC_DEFINE_ROWS(double, f64)

#define RLIC_DEFINE_PEER(T, sfx)                                                                 \
    int rlic_b200_pass_slab_peer_##sfx(const T *d_texture, const T *d_field, T *d_out,           \
                                       int64_t ny, int64_t nx, int64_t row0, int64_t nrows,      \
                                       int64_t halo_lo, int64_t halo_hi, int64_t sub_row0,       \
                                       int64_t sub_nrows, const T *kernel, int64_t klen,         \
                                       int uv_mode, int x_left, int x_right, int y_left,         \
                                       int y_right, T *d_peer_out, int64_t peer_row_delta,       \
                                       void *stream)                                             \
    {                                                                                            \
        tls_error.clear();                                                                       \
        if (!d_peer_out)                                                                         \
            return fail(RLIC_B200_EINVAL, "null peer buffer");                                   \
        return pass_slab<T>(d_texture, d_field, d_out, ny, nx,                                   \
                            Slab{row0, nrows, halo_lo, halo_hi}, sub_row0, sub_nrows, kernel,    \
                            klen, uv_mode, Walls{x_left, x_right, y_left, y_right}, stream,      \
                            d_peer_out, peer_row_delta);                                         \
    }
RLIC_DEFINE_PEER(float, f32)
RLIC_DEFINE_PEER(double, f64)

// rlic_b200_pass_slab_* / _pass_slab_peer_* with the streamline paths recorded or replayed
// (d_peer_out may be null: no neighbour behind these rows).
#define RLIC_DEFINE_SLAB_PATHS(T, sfx)                                                           \
    int rlic_b200_pass_slab_paths_##sfx(const T *d_texture, const T *d_field, T *d_out,          \
                                        int64_t ny, int64_t nx, int64_t row0, int64_t nrows,     \
                                        int64_t halo_lo, int64_t halo_hi, int64_t sub_row0,      \
                                        int64_t sub_nrows, const T *kernel, int64_t klen,        \
                                        int uv_mode, int x_left, int x_right, int y_left,        \
                                        int y_right, T *d_peer_out, int64_t peer_row_delta,      \
                                        int paths_mode, uint32_t *d_paths, void *stream)         \
    {                                                                                            \
        tls_error.clear();                                                                       \
        return pass_slab<T>(d_texture, d_field, d_out, ny, nx,                                   \
                            Slab{row0, nrows, halo_lo, halo_hi}, sub_row0, sub_nrows, kernel,    \
                            klen, uv_mode, Walls{x_left, x_right, y_left, y_right}, stream,      \
                            d_peer_out, peer_row_delta, paths_mode, d_paths);                    \
    }
RLIC_DEFINE_SLAB_PATHS(float, f32)
RLIC_DEFINE_SLAB_PATHS(double, f64)

// ---- histogram equalisation of a result (semantics: lic_equalize.cuh) ----
#define RLIC_DEFINE_EQUALIZE(T, sfx)                                                             \
    int rlic_b200_equalize_histogram_device_##sfx(const T *d_image, int64_t ny, int64_t nx,      \
                                                  int64_t nbins, T *d_out, void *stream)         \
    {                                                                                            \
        tls_error.clear();                                                                       \
        return equalize_device<T>(d_image, ny, nx, nbins, d_out, static_cast<cudaStream_t>(stream)); \
    }                                                                                            \
    int rlic_b200_equalize_histogram_##sfx(const T *image, int64_t ny, int64_t nx, int64_t nbins, \
                                           T *out)                                               \
    {                                                                                            \
        tls_error.clear();                                                                       \
        return equalize_host<T>(image, ny, nx, nbins, out, tls_device);                          \
    }
RLIC_DEFINE_EQUALIZE(float, f32)
RLIC_DEFINE_EQUALIZE(double, f64)

// ---- measurement: the gather ceiling of the memory system for the walk's access pattern ----
// ("an L2 gather peak measured by the build's own microbenchmark, same
// access count, straight-line walkers").  Launches gather_ceiling_kernel (lic_walk.cuh): the
// loads and the tap FMA of a pass and nothing else, walkers climbing a staircase.  The caller
// times it (bench.py: CUDA events) and uses the result as the denominator of the roofline
// fraction.  Not a convolution: `d_out` receives a checksum-like image nobody should use.
#define RLIC_DEFINE_CEILING(T, sfx)                                                              \
    int rlic_b200_measure_gather_ceiling_##sfx(const T *d_padded_texture, const T *d_field,      \
                                               T *d_padded_out, int64_t ny, int64_t nx,          \
                                               const T *kernel, int64_t klen, int dependent,     \
                                               void *stream)                                     \
    {                                                                                            \
        tls_error.clear();                                                                       \
        const Walls w{RLIC_B200_CLOSED, RLIC_B200_CLOSED, RLIC_B200_CLOSED, RLIC_B200_CLOSED};   \
        if (int rc = check_common(ny, nx, klen, 0, w))                                           \
            return rc;                                                                           \
        if (!d_padded_texture || !d_field || !d_padded_out || !kernel || ny == 0 || nx == 0)     \
            return fail(RLIC_B200_EINVAL, "null pointer or empty image");                        \
        if (klen > TapSet<T>::kMaxParam || rlic::padded_cells(ny, nx) >= (int64_t)INT_MAX)       \
            return fail(RLIC_B200_EINVAL, "kernel or image too large for the ceiling probe");    \
        PassGeom g = make_geometry(ny, nx, Slab{0, ny, 0, 0}, w);                                \
        g.first_row = 0;                                                                         \
        g.out_rows = (int)ny;                                                                    \
        g.tiles_x = (g.nx + 15) / 16;                                                            \
        g.tiles_per_field = g.tiles_x * (int)((ny + 15) / 16);                                   \
        cudaStream_t s = static_cast<cudaStream_t>(stream);                                      \
        TapSet<T> taps;                                                                          \
        CUDA_TRY(taps.prepare(kernel, klen, s));                                                 \
        const Field<T> *field = reinterpret_cast<const Field<T> *>(d_field);                     \
        if (dependent == 2)   /* the replay's loads: one texture value per step, no field */     \
            rlic::gather_ceiling_kernel<T, false, false><<<g.tiles_per_field, 256, 0, s>>>(      \
                d_padded_texture, field, d_padded_out, g, taps.param, (int)klen);                \
        else if (dependent)                                                                      \
            rlic::gather_ceiling_kernel<T, true><<<g.tiles_per_field, 256, 0, s>>>(              \
                d_padded_texture, field, d_padded_out, g, taps.param, (int)klen);                \
        else                                                                                     \
            rlic::gather_ceiling_kernel<T, false><<<g.tiles_per_field, 256, 0, s>>>(             \
                d_padded_texture, field, d_padded_out, g, taps.param, (int)klen);                \
        g_launches.fetch_add(1, std::memory_order_relaxed);                                      \
        CUDA_TRY(cudaGetLastError());                                                            \
        return RLIC_B200_OK;                                                                     \
    }
RLIC_DEFINE_CEILING(float, f32)
RLIC_DEFINE_CEILING(double, f64)

// ---- peer memory and flags of the fused halo exchange (one process per GPU) ----
int rlic_b200_peer_alloc(int64_t bytes, void **ptr, unsigned char *handle)
{
    tls_error.clear();
    if (bytes <= 0 || !ptr || !handle)
        return fail(RLIC_B200_EINVAL, "bad argument to peer_alloc");
    static_assert(sizeof(cudaIpcMemHandle_t) == RLIC_B200_PEER_HANDLE_BYTES, "handle size");
    CUDA_TRY(use_current_device());                   // the caller's device, as for every device entry point
    void *p = nullptr;
    CUDA_TRY(cudaMalloc(&p, (size_t)bytes));          // not the stream-ordered pool: IPC needs a plain allocation
    cudaError_t e = cudaMemset(p, 0, (size_t)bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess)
        e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        CUDA_TRY(e);
    }
    std::memcpy(handle, &h, sizeof h);
    *ptr = p;
    return RLIC_B200_OK;
}

int rlic_b200_peer_open(const unsigned char *handle, void **ptr)
{
    tls_error.clear();
    if (!handle || !ptr)
        return fail(RLIC_B200_EINVAL, "bad argument to peer_open");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof h);
    CUDA_TRY(use_current_device());
    CUDA_TRY(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return RLIC_B200_OK;
}

int rlic_b200_peer_close(void *ptr)
{
    tls_error.clear();
    if (ptr)
        CUDA_TRY(cudaIpcCloseMemHandle(ptr));
    return RLIC_B200_OK;
}

int rlic_b200_peer_free(void *ptr)
{
    tls_error.clear();
    if (ptr)
        CUDA_TRY(cudaFree(ptr));
    return RLIC_B200_OK;
}

int rlic_b200_peer_signal(uint32_t *d_flag, uint32_t value, void *stream)
{
    tls_error.clear();
    if (!d_flag)
        return fail(RLIC_B200_EINVAL, "null flag");
    rlic::peer_signal_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(d_flag, value);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CUDA_TRY(cudaGetLastError());
    return RLIC_B200_OK;
}

int rlic_b200_peer_signal2(uint32_t *d_flag_a, uint32_t *d_flag_b, uint32_t value, void *stream)
{
    tls_error.clear();
    if (!d_flag_a && !d_flag_b)
        return fail(RLIC_B200_EINVAL, "null flags");
    rlic::peer_signal2_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(d_flag_a, d_flag_b, value);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CUDA_TRY(cudaGetLastError());
    return RLIC_B200_OK;
}

int rlic_b200_peer_wait4(const uint32_t *d_flag0, uint32_t value0, const uint32_t *d_flag1, uint32_t value1,
                         const uint32_t *d_flag2, uint32_t value2, const uint32_t *d_flag3, uint32_t value3,
                         int64_t timeout_ms, int *d_timed_out, void *stream)
{
    tls_error.clear();
    if ((!d_flag0 && !d_flag1 && !d_flag2 && !d_flag3) || timeout_ms <= 0)
        return fail(RLIC_B200_EINVAL, "bad argument to peer_wait4");
    const rlic::PeerWaits w{{d_flag0, d_flag1, d_flag2, d_flag3}, {value0, value1, value2, value3}};
    rlic::peer_wait4_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(
        w, (long long)timeout_ms * 1000000ll, d_timed_out);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CUDA_TRY(cudaGetLastError());
    return RLIC_B200_OK;
}

int rlic_b200_peer_wait(const uint32_t *d_flag, uint32_t value, int64_t timeout_ms, int *d_timed_out,
                        void *stream)
{
    tls_error.clear();
    if (!d_flag || timeout_ms <= 0)
        return fail(RLIC_B200_EINVAL, "bad argument to peer_wait");
    rlic::peer_wait_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(
        d_flag, value, (long long)timeout_ms * 1000000ll, d_timed_out);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CUDA_TRY(cudaGetLastError());
    return RLIC_B200_OK;
}

}  // extern "C"
