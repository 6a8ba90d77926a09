// The library's kernels (rlic_b200/csrc/lic_walk.cuh), compiled for the CPU and driven
// block by block.  TEST INFRASTRUCTURE: see cuda_on_cpu.h.  Built by
// tests/kernel_emulation/__init__.py into tests/kernel_emulation/libkernel_emulation.so.
//
// The launch shapes mirror rlic_b200/csrc/lic_api.cu (launch_pack / launch_pad /
// launch_unpad / launch_pass); the geometry is not restated here: the tests obtain it
// from the library itself (rlic_b200_debug_geometry) and pass it in.
#define RLIC_HOST_EMULATION 1
#include "../../rlic_b200/csrc/lic_walk.cuh"

#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

// Built as two translation units (-DEMU_UNIT_F32 / -DEMU_UNIT_F64, compiled in parallel: the
// template instantiations dominate the build); the shared state lives in the f32 unit.
#if !defined(EMU_UNIT_F32) && !defined(EMU_UNIT_F64)
#define EMU_UNIT_F32 1
#define EMU_UNIT_F64 1
#endif
#ifdef EMU_UNIT_F32
thread_local Dim3 threadIdx, blockIdx, blockDim, gridDim;
thread_local emulated::StepCounts emulated::step_counts;
thread_local bool emulated::past_barrier;
alignas(16) thread_local unsigned char emulated::shared_bytes[256 * 1024];
#endif

namespace {

using rlic::PassGeom;

// geometry as rlic_b200_debug_geometry reports it
PassGeom geometry_from(const int64_t *v)
{
    PassGeom g{};
    g.nx = (int)v[0];
    g.pitch = (int)v[1];
    g.rows = (int)v[2];
    g.field_stride = v[3];
    g.j_below_to = (int)v[4];
    g.j_above_to = (int)v[5];
    g.i_below_to = (int)v[6];
    g.i_above_to = (int)v[7];
    g.lo_wall = (int)v[8];
    g.hi_wall = (int)v[9];
    return g;
}

unsigned stream_blocks(long long items)
{
    return (unsigned)std::max<long long>(1, std::min<long long>((items + 255) / 256, 148 * 16));
}

// Runs `body()` once per (block, thread) of a 1-D launch, blocks spread over host threads.
// The kernels involved never synchronise within a block, so threads run to completion
// one after another.
}  // namespace
#ifdef EMU_UNIT_F32
std::atomic<unsigned long long> emu_counts[4];
#else
extern std::atomic<unsigned long long> emu_counts[4];
#endif
namespace {
std::atomic<unsigned long long> (&g_counts)[4] = emu_counts;

template <typename Body> void launch(unsigned blocks, unsigned threads, Body body)
{
    const unsigned workers = std::max(1u, std::min(blocks, std::thread::hardware_concurrency()));
    std::atomic<unsigned> next{0};
    auto run = [&] {
        gridDim = {blocks, 1, 1};
        blockDim = {threads, 1, 1};
        emulated::step_counts = {};
        for (unsigned b = next.fetch_add(1); b < blocks; b = next.fetch_add(1)) {
            blockIdx = {b, 0, 0};
            for (unsigned t = 0; t < threads; ++t) {
                threadIdx = {t, 0, 0};
                body();
            }
        }
        g_counts[0] += emulated::step_counts.step;
        g_counts[1] += emulated::step_counts.declined;
        g_counts[2] += emulated::step_counts.wall;
        g_counts[3] += emulated::step_counts.generic;
    };
    std::vector<std::thread> pool;
    for (unsigned w = 1; w < workers; ++w) pool.emplace_back(run);
    run();
    for (auto &th : pool) th.join();
}

// The same for a three-dimensional grid (the replay kernels: tile column, tile row, field).
// sweeps = 2: a kernel with one barrier (see cuda_on_cpu.h: past_barrier).
template <typename Body> void launch3(unsigned gx, unsigned gy, unsigned gz, unsigned threads, Body body,
                                      int sweeps = 1)
{
    const unsigned blocks = gx * gy * gz;
    const unsigned workers = std::max(1u, std::min(blocks, std::thread::hardware_concurrency()));
    std::atomic<unsigned> next{0};
    auto run = [&] {
        gridDim = {gx, gy, gz};
        blockDim = {threads, 1, 1};
        for (unsigned b = next.fetch_add(1); b < blocks; b = next.fetch_add(1)) {
            blockIdx = {b % gx, (b / gx) % gy, b / (gx * gy)};
            if (sweeps > 1)   // a fresh CTA's shared memory holds nothing useful: NaN patterns, so that a
                std::memset(emulated::shared_bytes, 0xFF, sizeof emulated::shared_bytes);   // read of an unfilled cell shows
            for (int sweep = 0; sweep < sweeps; ++sweep) {
                emulated::past_barrier = sweep == sweeps - 1;
                for (unsigned t = 0; t < threads; ++t) {
                    threadIdx = {t, 0, 0};
                    body();
                }
            }
        }
    };
    std::vector<std::thread> pool;
    for (unsigned w = 1; w < workers; ++w) pool.emplace_back(run);
    run();
    for (auto &th : pool) th.join();
}

template <typename T>
void pack_field(const T *u, const T *v, T *field, const int64_t *geom, int64_t rb, int64_t re, int64_t nfields)
{
    const PassGeom g = geometry_from(geom);
    if (re <= rb || nfields <= 0) return;
    auto *f = reinterpret_cast<rlic::PackedField<T> *>(field);
    launch(stream_blocks((re - rb + 2) * g.pitch * nfields), 256,
           [&] { rlic::pack_field_kernel<T>(u, v, f, g, (int)rb, (int)re, (long long)nfields); });
}

template <typename T>
void pad_texture(const T *dense, T *padded, const int64_t *geom, int64_t rb, int64_t re, int64_t nfields,
                 int *negative)
{
    const PassGeom g = geometry_from(geom);
    if (re <= rb || nfields <= 0) return;
    launch(stream_blocks((re - rb + 2) * g.pitch * nfields), 256,
           [&] { rlic::pad_texture_kernel<T>(dense, padded, g, (int)rb, (int)re, (long long)nfields, negative); });
}

template <typename T>
void unpad_texture(const T *padded, T *dense, const int64_t *geom, int64_t rb, int64_t re, int64_t nfields)
{
    const PassGeom g = geometry_from(geom);
    if (re <= rb || nfields <= 0) return;
    launch(stream_blocks((re - rb) * g.nx * nfields), 256,
           [&] { rlic::unpad_texture_kernel<T>(padded, dense, g, (int)rb, (int)re, (long long)nfields); });
}

constexpr int kDefault = -1;

// One pass, as launch_pass() issues it.  flavor / admit: kDefault selects the library's
// per-type tuning (rlic::Tune), other values pick a formulation explicitly.
template <typename T, bool POL, typename Taps, typename Idx, int FLAVOR, int ADMIT, bool BRANCHLESS = true,
          int WALK = 0>
void run_pass(const T *tex, const T *field, T *out, const PassGeom &g, const Taps &taps, int ntaps,
              unsigned blocks)
{
    using Tn = rlic::Tune<T, POL>;
    auto *f = reinterpret_cast<const rlic::PackedField<T> *>(field);
    launch(blocks, rlic::kThreads, [&] {
        rlic::lic_pass_kernel<T, POL, Taps, Idx, rlic::kTileW, rlic::kTileH, (WALK ? Tn::walk_unroll : Tn::unroll),
                              (WALK ? Tn::walk_min_blocks : Tn::min_blocks), FLAVOR, ADMIT, BRANCHLESS, WALK>(
            tex, f, out, g, taps, ntaps, rlic::PathPlanes{});
    });
}

// An explicitly chosen formulation (32-bit indices, taps in the parameter block only).
// walk: 0 the per-step loop (flavours 0, 1), 1 the grouped walk, 9 the grouped walk with
// its loop control on the byte offset (flavours 0-3); the register-pinning bits of WALK
// mean nothing on a CPU.
template <typename T, bool POL, typename Taps>
int pick_formulation(const T *tex, const T *field, T *out, const PassGeom &g, const Taps &taps, int ntaps,
                     unsigned blocks, int flavor, int admit, int walk)
{
#define RLIC_CASE(F, A) \
    if (flavor == F && admit == A) { \
        if (walk == 1) run_pass<T, POL, Taps, int, F, A, true, 1>(tex, field, out, g, taps, ntaps, blocks); \
        else if (walk == 9) run_pass<T, POL, Taps, int, F, A, true, 9>(tex, field, out, g, taps, ntaps, blocks); \
        else if (walk == 0 && F < 2) run_pass<T, POL, Taps, int, (F < 2 ? F : 0), A>(tex, field, out, g, taps, ntaps, blocks); \
        else return 1; \
        return 0; }
    RLIC_CASE(0, 0) RLIC_CASE(0, 1) RLIC_CASE(0, 2) RLIC_CASE(0, 3)
    RLIC_CASE(1, 0) RLIC_CASE(1, 1) RLIC_CASE(1, 2) RLIC_CASE(1, 3)
    RLIC_CASE(2, 0) RLIC_CASE(2, 1) RLIC_CASE(2, 2) RLIC_CASE(2, 3)
    RLIC_CASE(3, 0) RLIC_CASE(3, 1) RLIC_CASE(3, 2) RLIC_CASE(3, 3)
    RLIC_CASE(0, 4) RLIC_CASE(1, 4) RLIC_CASE(2, 4) RLIC_CASE(3, 4)
#undef RLIC_CASE
    // flavour 4: the packed-pair formulation of the grouped walk, single precision only
    if constexpr (sizeof(T) == 4) {
        if (flavor == 4 && (walk == 1 || walk == 9) && (admit == 3 || admit == 4)) {
            if (walk == 1 && admit == 3) run_pass<T, POL, Taps, int, 4, 3, true, 1>(tex, field, out, g, taps, ntaps, blocks);
            else if (walk == 1) run_pass<T, POL, Taps, int, 4, 4, true, 1>(tex, field, out, g, taps, ntaps, blocks);
            else if (admit == 3) run_pass<T, POL, Taps, int, 4, 3, true, 9>(tex, field, out, g, taps, ntaps, blocks);
            else run_pass<T, POL, Taps, int, 4, 4, true, 9>(tex, field, out, g, taps, ntaps, blocks);
            return 0;
        }
    }
    return 1;
}

// The library's own choice (rlic::Tune), as launch_one() dispatches it: walk != 0 is
// rlic_b200_set_walk(RLIC_B200_WALK_GROUPED).
template <typename T, bool POL, typename Taps>
int tuned(const T *tex, const T *field, T *out, const PassGeom &g, const Taps &taps, int ntaps, unsigned blocks,
          int wide, int branchless, int walk)
{
    using Tn = rlic::Tune<T, POL>;
    if (branchless && walk) {
        if (wide) run_pass<T, POL, Taps, long long, Tn::walk_flavor, Tn::walk_admit, true, Tn::walk>(tex, field, out, g, taps, ntaps, blocks);
        else run_pass<T, POL, Taps, int, Tn::walk_flavor, Tn::walk_admit, true, Tn::walk>(tex, field, out, g, taps, ntaps, blocks);
    } else if (branchless) {
        if (wide) run_pass<T, POL, Taps, long long, Tn::flavor, Tn::admit>(tex, field, out, g, taps, ntaps, blocks);
        else run_pass<T, POL, Taps, int, Tn::flavor, Tn::admit>(tex, field, out, g, taps, ntaps, blocks);
    } else {
        if (wide) run_pass<T, POL, Taps, long long, Tn::flavor, Tn::admit, false>(tex, field, out, g, taps, ntaps, blocks);
        else run_pass<T, POL, Taps, int, Tn::flavor, Tn::admit, false>(tex, field, out, g, taps, ntaps, blocks);
    }
    return 0;
}

template <typename T>
int pass(const T *tex, const T *field, T *out, const int64_t *geom, int64_t nfields, int64_t first_row,
         int64_t out_rows, int uv_mode, const T *host_taps, int64_t klen, int wide, int flavor, int admit,
         int branchless, int walk)
{
    PassGeom g = geometry_from(geom);
    if (out_rows <= 0 || g.nx <= 0 || nfields <= 0) return 0;
    g.first_row = (int)first_row;
    g.out_rows = (int)out_rows;
    g.tiles_x = (g.nx + rlic::kTileW - 1) / rlic::kTileW;
    const int64_t tiles_y = (out_rows + rlic::kTileH - 1) / rlic::kTileH;
    const int64_t per_field = tiles_y * g.tiles_x;
    g.tiles_per_field = (int)per_field;
    const unsigned blocks = (unsigned)(per_field * nfields);
    const bool pol = uv_mode == 1;
    const bool chosen = flavor != kDefault || admit != kDefault;

    constexpr int kMaxParam = rlic::kParamTapBytes / (int)sizeof(T);
    using PT = rlic::ParamTaps<T, kMaxParam>;
    using GT = rlic::GlobalTaps<T>;
    const int ntaps = (int)klen;
    if (klen <= kMaxParam) {
        PT pt;
        std::memset(pt.w, 0, sizeof pt.w);
        std::memcpy(pt.w, host_taps, sizeof(T) * (size_t)klen);
        if (chosen) {
            if (wide || !branchless) return 1;
            return pol ? pick_formulation<T, true, PT>(tex, field, out, g, pt, ntaps, blocks, flavor, admit, walk)
                       : pick_formulation<T, false, PT>(tex, field, out, g, pt, ntaps, blocks, flavor, admit, walk);
        }
        return pol ? tuned<T, true, PT>(tex, field, out, g, pt, ntaps, blocks, wide, branchless, walk)
                   : tuned<T, false, PT>(tex, field, out, g, pt, ntaps, blocks, wide, branchless, walk);
    }
    if (chosen) return 1;
    const GT gt{host_taps};
    return pol ? tuned<T, true, GT>(tex, field, out, g, gt, ntaps, blocks, wide, branchless, walk)
               : tuned<T, false, GT>(tex, field, out, g, gt, ntaps, blocks, wide, branchless, walk);
}

// The fused pass + halo shipment (lic_pass_peer_kernel), dispatched as launch_one() does for
// the default arithmetic: per-step walk, or the tuned grouped walk.
template <typename T, bool POL, typename Taps>
void run_pass_peer(const T *tex, const T *field, T *out, const PassGeom &g, const Taps &taps, int ntaps,
                   unsigned blocks, int walk, T *peer_out, long long peer_delta)
{
    using Tn = rlic::Tune<T, POL>;
    auto *f = reinterpret_cast<const rlic::PackedField<T> *>(field);
    if (walk)
        launch(blocks, rlic::kThreads, [&] {
            rlic::lic_pass_peer_kernel<T, POL, Taps, int, rlic::kTileW, rlic::kTileH, Tn::walk_unroll,
                                       Tn::walk_min_blocks, Tn::walk_flavor, Tn::walk_admit, true, Tn::walk>(tex, f, out, g, taps, ntaps,
                                                                                    peer_out, peer_delta, rlic::PathPlanes{});
        });
    else
        launch(blocks, rlic::kThreads, [&] {
            rlic::lic_pass_peer_kernel<T, POL, Taps, int>(tex, f, out, g, taps, ntaps, peer_out, peer_delta,
                                                          rlic::PathPlanes{});
        });
}

template <typename T>
int pass_peer(const T *tex, const T *field, T *out, const int64_t *geom, int64_t first_row, int64_t out_rows,
              int uv_mode, const T *host_taps, int64_t klen, int walk, T *peer_out, int64_t peer_row_delta)
{
    PassGeom g = geometry_from(geom);
    if (out_rows <= 0 || g.nx <= 0) return 0;
    g.first_row = (int)first_row;
    g.out_rows = (int)out_rows;
    g.tiles_x = (g.nx + rlic::kTileW - 1) / rlic::kTileW;
    g.tiles_per_field = (int)(((out_rows + rlic::kTileH - 1) / rlic::kTileH) * g.tiles_x);
    const unsigned blocks = (unsigned)g.tiles_per_field;
    const long long delta = (long long)peer_row_delta * g.pitch;      // as pass_slab() in lic_api.cu
    constexpr int kMaxParam = rlic::kParamTapBytes / (int)sizeof(T);
    using PT = rlic::ParamTaps<T, kMaxParam>;
    using GT = rlic::GlobalTaps<T>;
    const int ntaps = (int)klen;
    const bool pol = uv_mode == 1;
    if (klen <= kMaxParam) {
        PT pt;
        std::memset(pt.w, 0, sizeof pt.w);
        std::memcpy(pt.w, host_taps, sizeof(T) * (size_t)klen);
        if (pol) run_pass_peer<T, true, PT>(tex, field, out, g, pt, ntaps, blocks, walk, peer_out, delta);
        else run_pass_peer<T, false, PT>(tex, field, out, g, pt, ntaps, blocks, walk, peer_out, delta);
    } else {
        const GT gt{host_taps};
        if (pol) run_pass_peer<T, true, GT>(tex, field, out, g, gt, ntaps, blocks, walk, peer_out, delta);
        else run_pass_peer<T, false, GT>(tex, field, out, g, gt, ntaps, blocks, walk, peer_out, delta);
    }
    return 0;
}

// One pass that records the streamline paths (mode 1: the tuned grouped walk with REC) or
// replays them (mode 2: lic_replay_kernel), as launch_pass() dispatches them in lic_api.cu:
// the record of a 65-tap kernel goes through the unrolled one-group kernel, 129 taps through
// the two-group one, longer ones through the loop; kernels beyond the parameter block read
// their taps from memory.  peer_out: the doubled stores of the fused halo exchange.
template <typename T, bool POL, typename Taps, typename Idx>
void run_record(const T *tex, const T *field, T *out, const PassGeom &g, const Taps &taps, int ntaps,
                unsigned blocks, unsigned *rec, long long plane_cells, T *peer_out, long long peer_delta)
{
    using Tn = rlic::Tune<T, POL>;
    auto *f = reinterpret_cast<const rlic::PackedField<T> *>(field);
    const rlic::PathPlanes planes{reinterpret_cast<uint4 *>(rec), plane_cells, rlic::path_groups_fwd(ntaps)};
    constexpr int kFlavor = Tn::walk_flavor == 4 ? 2 : Tn::walk_flavor;
    if (peer_out)
        launch(blocks, rlic::kThreads, [&] {
            rlic::lic_pass_peer_kernel<T, POL, Taps, Idx, rlic::kTileW, rlic::kTileH, Tn::walk_unroll,
                                       Tn::walk_min_blocks, kFlavor, Tn::walk_admit, true, Tn::walk, true>(
                tex, f, out, g, taps, ntaps, peer_out, peer_delta, planes);
        });
    else
        launch(blocks, rlic::kThreads, [&] {
            rlic::lic_pass_kernel<T, POL, Taps, Idx, rlic::kTileW, rlic::kTileH, Tn::walk_unroll, Tn::walk_min_blocks,
                                  kFlavor, Tn::walk_admit, true, Tn::walk, true>(tex, f, out, g, taps, ntaps, planes);
        });
}

template <typename T, typename Taps, typename Idx, int GROUPS>
void run_replay(const T *tex, T *out, const PassGeom &g, const Taps &taps, int ntaps, unsigned nfields,
                const unsigned *rec, long long group_cells, T *peer_out, long long peer_delta)
{
    const auto *entries = reinterpret_cast<const uint4 *>(rec);
    const unsigned gy = (unsigned)((g.out_rows + rlic::kTileH - 1) / rlic::kTileH);
    if (peer_out)
        launch3((unsigned)g.tiles_x, gy, nfields, rlic::kThreads, [&] {
            rlic::lic_replay_kernel<T, Taps, Idx, GROUPS, true>(tex, entries, out, g, taps, ntaps, group_cells, peer_out,
                                                                peer_delta);
        });
    else
        launch3((unsigned)g.tiles_x, gy, nfields, rlic::kThreads, [&] {
            rlic::lic_replay_kernel<T, Taps, Idx, GROUPS, false>(tex, entries, out, g, taps, ntaps, group_cells, nullptr, 0);
        });
}

template <typename T>
int pass_paths(const T *tex, const T *field, T *out, const int64_t *geom, int64_t nfields, int64_t first_row,
               int64_t out_rows, int uv_mode, const T *host_taps, int64_t klen, int wide, int mode, unsigned *rec,
               T *peer_out, int64_t peer_row_delta)
{
    PassGeom g = geometry_from(geom);
    if (out_rows <= 0 || g.nx <= 0 || nfields <= 0) return 0;
    g.first_row = (int)first_row;
    g.out_rows = (int)out_rows;
    g.tiles_x = (g.nx + rlic::kTileW - 1) / rlic::kTileW;
    const int64_t per_field = ((out_rows + rlic::kTileH - 1) / rlic::kTileH) * g.tiles_x;
    g.tiles_per_field = (int)per_field;
    const unsigned blocks = (unsigned)(per_field * nfields);
    const long long plane_cells = g.field_stride * nfields;
    const long long delta = (long long)peer_row_delta * g.pitch;
    const bool pol = uv_mode == 1;
    const int ntaps = (int)klen;
    constexpr int kMaxParam = rlic::kParamTapBytes / (int)sizeof(T);
    using PT = rlic::ParamTaps<T, kMaxParam>;
    using GT = rlic::GlobalTaps<T>;
    using ST = rlic::StepTaps<T, rlic::kStepTapsPerHalf<T>>;
    using GST = rlic::GlobalStepTaps<T>;
    const bool in_param = klen <= kMaxParam;
    if (mode == 1) {
        PT pt;
        std::memset(pt.w, 0, sizeof pt.w);
        if (in_param) std::memcpy(pt.w, host_taps, sizeof(T) * (size_t)klen);
        const GT gt{host_taps};
#define EMU_RECORD(POL, TAPS, TAPV)                                                                                   \
    do {                                                                                                              \
        if (wide) run_record<T, POL, TAPS, long long>(tex, field, out, g, TAPV, ntaps, blocks, rec, plane_cells, peer_out, delta); \
        else run_record<T, POL, TAPS, int>(tex, field, out, g, TAPV, ntaps, blocks, rec, plane_cells, peer_out, delta); \
    } while (0)
        if (in_param) { if (pol) EMU_RECORD(true, PT, pt); else EMU_RECORD(false, PT, pt); }
        else          { if (pol) EMU_RECORD(true, GT, gt); else EMU_RECORD(false, GT, gt); }
#undef EMU_RECORD
        return 0;
    }
    if (mode != 2 && mode != 3) return 1;
    // the taps in walking order, as TapSet::prepare() lays them out
    ST st;
    std::memset(&st, 0, sizeof st);
    const int64_t kmid = klen / 2;
    if (in_param) {
        st.centre = host_taps[kmid];
        for (int64_t k = kmid + 1; k < klen; ++k) st.fwd[k - kmid - 1] = host_taps[k];
        for (int64_t k = kmid - 1; k >= 0; --k) st.bwd[kmid - 1 - k] = host_taps[k];
    }
    const GST gst{host_taps, (int)kmid};
    const int groups = std::max(rlic::path_groups_fwd(klen), rlic::path_groups_bwd(klen));
    if (mode == 3) {
        // lic_replay_staged_kernel, under the conditions launch_replay() uses it
        if (!in_param || groups > 1 || wide || rlic::staged_window_bytes(klen, sizeof(T)) > rlic::kStagedMaxBytes)
            return 2;
        const auto *entries = reinterpret_cast<const uint4 *>(rec);
        const unsigned gx = (unsigned)((g.nx + rlic::kStagedTW - 1) / rlic::kStagedTW);
        const unsigned gy = (unsigned)((g.out_rows + rlic::kStagedTH - 1) / rlic::kStagedTH);
        if (peer_out)
            launch3(gx, gy, (unsigned)nfields, rlic::kStagedTW * rlic::kStagedTH, [&] {
                rlic::lic_replay_staged_kernel<T, ST, int, true, rlic::kStagedTW, rlic::kStagedTH, rlic::kStagedPad,
                                               rlic::kStagedMinBlocks>(tex, entries, out, g, st, ntaps, plane_cells,
                                                                       peer_out, delta);
            }, 2);
        else
            launch3(gx, gy, (unsigned)nfields, rlic::kStagedTW * rlic::kStagedTH, [&] {
                rlic::lic_replay_staged_kernel<T, ST, int, false, rlic::kStagedTW, rlic::kStagedTH, rlic::kStagedPad,
                                               rlic::kStagedMinBlocks>(tex, entries, out, g, st, ntaps, plane_cells,
                                                                       nullptr, 0);
            }, 2);
        return 0;
    }
#define EMU_REPLAY(TAPS, TAPV, GROUPS)                                                                               \
    do {                                                                                                              \
        if (wide) run_replay<T, TAPS, long long, GROUPS>(tex, out, g, TAPV, ntaps, (unsigned)nfields, rec, plane_cells, peer_out, delta); \
        else run_replay<T, TAPS, int, GROUPS>(tex, out, g, TAPV, ntaps, (unsigned)nfields, rec, plane_cells, peer_out, delta);   \
    } while (0)
    if (!in_param) EMU_REPLAY(GST, gst, 0);
    else if (groups <= 1) EMU_REPLAY(ST, st, 1);
    else if (groups == 2) EMU_REPLAY(ST, st, 2);
    else EMU_REPLAY(ST, st, 0);
#undef EMU_REPLAY
    return 0;
}

}  // namespace

// Step counters accumulated over every pass since the last reset.
#ifdef EMU_UNIT_F32
extern "C" void emu_step_counts(unsigned long long *out, int reset)
{
    for (int i = 0; i < 4; ++i) {
        out[i] = g_counts[i].load();
        if (reset) g_counts[i].store(0);
    }
}
#endif

#define EMU_DEFINE(SFX, T)                                                                                  \
    extern "C" void emu_pack_field_##SFX(const T *u, const T *v, T *field, const int64_t *geom, int64_t rb,  \
                                         int64_t re, int64_t nfields)                                        \
    { pack_field<T>(u, v, field, geom, rb, re, nfields); }                                                   \
    extern "C" void emu_pad_texture_##SFX(const T *dense, T *padded, const int64_t *geom, int64_t rb,        \
                                          int64_t re, int64_t nfields, int *negative)                        \
    { pad_texture<T>(dense, padded, geom, rb, re, nfields, negative); }                                      \
    extern "C" void emu_unpad_texture_##SFX(const T *padded, T *dense, const int64_t *geom, int64_t rb,      \
                                            int64_t re, int64_t nfields)                                     \
    { unpad_texture<T>(padded, dense, geom, rb, re, nfields); }                                              \
    extern "C" int emu_pass_##SFX(const T *tex, const T *field, T *out, const int64_t *geom, int64_t nfields, \
                                  int64_t first_row, int64_t out_rows, int uv_mode, const T *taps,           \
                                  int64_t klen, int wide, int flavor, int admit, int branchless,   \
                                  int walk)                                                                  \
    { return pass<T>(tex, field, out, geom, nfields, first_row, out_rows, uv_mode, taps, klen, wide, flavor, admit, \
                     branchless, walk); }

#ifdef EMU_UNIT_F32
EMU_DEFINE(f32, float)
#endif
#ifdef EMU_UNIT_F64
EMU_DEFINE(f64, double)
#endif

#define EMU_DEFINE_PEER(SFX, T)                                                                              \
    extern "C" int emu_pass_peer_##SFX(const T *tex, const T *field, T *out, const int64_t *geom,            \
                                       int64_t first_row, int64_t out_rows, int uv_mode, const T *taps,      \
                                       int64_t klen, int walk, T *peer_out, int64_t peer_row_delta)          \
    { return pass_peer<T>(tex, field, out, geom, first_row, out_rows, uv_mode, taps, klen, walk, peer_out,   \
                          peer_row_delta); }
#ifdef EMU_UNIT_F32
EMU_DEFINE_PEER(f32, float)
#endif
#ifdef EMU_UNIT_F64
EMU_DEFINE_PEER(f64, double)
#endif

#define EMU_DEFINE_PATHS(SFX, T)                                                                             \
    extern "C" int emu_pass_paths_##SFX(const T *tex, const T *field, T *out, const int64_t *geom,           \
                                        int64_t nfields, int64_t first_row, int64_t out_rows, int uv_mode,   \
                                        const T *taps, int64_t klen, int wide, int mode, unsigned *rec,      \
                                        T *peer_out, int64_t peer_row_delta)                                 \
    { return pass_paths<T>(tex, field, out, geom, nfields, first_row, out_rows, uv_mode, taps, klen, wide,   \
                           mode, rec, peer_out, peer_row_delta); }
#ifdef EMU_UNIT_F32
EMU_DEFINE_PATHS(f32, float)
#endif
#ifdef EMU_UNIT_F64
EMU_DEFINE_PATHS(f64, double)
#endif
