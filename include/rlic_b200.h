/*
 * rlic_b200 — C ABI of the B200-native line integral convolution library
 * (librlic_b200.so, built from rlic_b200/csrc/ for sm_100a).
 *
 * This is the drop-in boundary for the one native seam of the reference:
 * the PyO3 module `rlic._core` with its two functions
 *     convolve_f32(texture, (u, v, uv_mode), kernel, ((xl,xr),(yl,yr)), iterations)
 *     convolve_f64(...)
 * declared at /root/reference/src/lib.rs:451-482 (typed in
 * /root/reference/src/rlic/_core.pyi:8-29) and called from
 * /root/reference/src/rlic/_lib.py:228-235.  A maintainer binds these entry
 * points with ctypes; rlic_b200/_core.py is that binding.
 *
 * Conventions
 *   - plain C types only: pointers, int, int64_t; no torch / CUDA types
 *     (a CUDA stream is passed as void*)
 *   - images are C-contiguous, `ny` rows by `nx` columns; x is the column axis
 *     (parallel to u), y the row axis (parallel to v)   [_lib.py:78-79]
 *   - uv_mode:   0 = "velocity", 1 = "polarization"      [lib.rs:19-26]
 *   - boundary:  0 = "closed",   1 = "periodic"          [lib.rs:60-67]
 *   - every function returns 0 on success or an RLIC_B200_E* code; the message
 *     is available from rlic_b200_last_error() on the calling thread.  Nothing
 *     aborts the process (the reference panics->aborts on bad enums and on an
 *     empty kernel, lib.rs:24,65,371; Cargo.toml:23).
 *   - re-entrant: no mutable global state besides per-device caches guarded by
 *     a mutex; each call runs on its own stream (the reference declares
 *     gil_used = false, lib.rs:448).
 *   - there is NO CPU fallback: without a CUDA device every compute entry
 *     point fails with RLIC_B200_ENODEVICE.
 *
 * Arithmetic contract: the kernels evaluate the reference's default feature
 * set (fma + branchless, Cargo.toml:28) operation by operation in the input
 * dtype, so results are bit-identical to that build on the same inputs;
 * rlic_b200_set_arithmetic() selects the `fma`-only build instead.
 */
#ifndef RLIC_B200_H
#define RLIC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RLIC_B200_OK 0
#define RLIC_B200_EINVAL 1    /* bad enum, negative size, empty kernel, null pointer */
#define RLIC_B200_ENODEVICE 2 /* no usable CUDA device / driver */
#define RLIC_B200_ECUDA 3     /* a CUDA runtime call failed (incl. out of memory) */
#define RLIC_B200_ESHARD 4    /* impossible sharding request */

#define RLIC_B200_VELOCITY 0
#define RLIC_B200_POLARIZATION 1
#define RLIC_B200_CLOSED 0
#define RLIC_B200_PERIODIC 1

/* ABI version of this header; bumped on any signature change. */
#define RLIC_B200_ABI_VERSION 4
int rlic_b200_abi_version(void);

/* Message of the last failure on the calling thread ("" if none). */
const char *rlic_b200_last_error(void);

/* Number of visible CUDA devices (0 when there is no driver/GPU; never fails). */
int rlic_b200_device_count(void);

/* Which build of the reference the kernels reproduce bit for bit.  The crate's features
 * (Cargo.toml:27-30) change one function, the time to the next pixel edge
 * (lib.rs:157-180), and the builds in circulation differ in them:
 *   RLIC_B200_ARITH_FMA_BRANCHLESS  `fma` + `branchless`: the crate default, i.e. a build
 *                                   from source / the sdist, and the aarch64 wheels
 *                                   (.github/workflows/cd.yml:93,143,210).  Default here.
 *   RLIC_B200_ARITH_FMA             `fma` alone: the x86-64 wheels (cd.yml:89,139,210).
 * The two agree on almost every pixel (a last-bit difference in an edge time matters only
 * where it flips a `tx < ty` decision).  This sets the PROCESS-WIDE DEFAULT (set it once, at
 * start-up); a thread that wants something else for its own calls uses
 * rlic_b200_set_thread_options below, which never touches shared state. */
#define RLIC_B200_ARITH_FMA_BRANCHLESS 0
#define RLIC_B200_ARITH_FMA 1
int rlic_b200_set_arithmetic(int which);
int rlic_b200_get_arithmetic(void);

/* How the host entry points (rlic_b200_convolve_*, _convolve_checked_*) order their work
 * for a single large image.  Both upload, compute and download in row bands and give the
 * same bits; they differ in how much of the transfers hides behind the passes.
 *   RLIC_B200_SCHEDULE_TRAILING   pass 1 follows the uploads band by band, the middle
 *                                 passes run over the whole image, the last pass releases
 *                                 bands to the download.
 *   RLIC_B200_SCHEDULE_WAVEFRONT  every pass follows the uploads band by band (pass p of
 *                                 band b runs once pass p-1 of bands b-1..b+1 is done), so
 *                                 early bands finish, and leave, while late bands are still
 *                                 arriving.  Used when y is not periodic and iterations >= 2;
 *                                 otherwise the call falls back to the trailing order.
 *                                 Default since round 2 (measured on a B200: 4096^2 x 5 passes
 *                                 end to end 11.7 -> 11.1 ms, 16384^2 x 20 599 -> 559 ms).
 * A process-wide default, like rlic_b200_set_arithmetic. */
#define RLIC_B200_SCHEDULE_TRAILING 0
#define RLIC_B200_SCHEDULE_WAVEFRONT 1
int rlic_b200_set_schedule(int which);
int rlic_b200_get_schedule(void);

/* How the pass kernels are written.  Same results bit for bit; a process-wide default like the
 * two above.
 *   RLIC_B200_WALK_GROUPED   (default since round 2) the loop-exit test once per group of steps,
 *                            no negation of the record in the backward pass: 6 (f32) / 1-2 (f64)
 *                            fewer instructions per step in the sm_100a SASS; measured on a B200
 *                            at 1.43 ms against 1.59 ms per 4096^2 x 65-tap f32 pass
 *                            (profiles/r2_session_lab_grouped_f32_f64.txt).  Applies to the
 *                            default arithmetic only.
 *   RLIC_B200_WALK_PER_STEP  the kernels every measurement of round 1 was made with; kept as
 *                            the yardstick of tools/kernel_lab and for the `fma` arithmetic. */
#define RLIC_B200_WALK_PER_STEP 0
#define RLIC_B200_WALK_GROUPED 1
int rlic_b200_set_walk(int which);
int rlic_b200_get_walk(void);

/* Per-thread overrides of the three choices above, for the calls the CALLING THREAD makes from
 * now on (every entry point reads its choices once, on the thread that entered the library;
 * the batch entry hands them to its worker threads).  -1 = no override, use the process-wide
 * default.  Nothing shared is written: two threads that want different arithmetic do not
 * race (the reference is re-entrant, src/lib.rs:448 `gil_used = false`).
 * rlic_b200_get_thread_options reads the overrides back, rlic_b200_get_effective_options
 * what a call made now by this thread would use. */
int rlic_b200_set_thread_options(int arithmetic, int schedule, int walk);
void rlic_b200_get_thread_options(int *arithmetic, int *schedule, int *walk);
void rlic_b200_get_effective_options(int *arithmetic, int *schedule, int *walk);

/* What the passes of a call after its first do.  Which pixels a streamline visits depends on
 * the vector field, the mode and the boundaries -- never on the texture (lib.rs:305-362; the
 * texture only enters the accumulation, :353-360) -- and the reference hands the same u, v to
 * every iteration (lib.rs:432-440), so every pass of a call walks the same paths.
 *   RLIC_B200_PATHS_REPLAY     (default) the first pass of a call with iterations >= 2 records,
 *                              per pixel and step, which way its walker went (bit planes in
 *                              device memory, 16 bytes per pixel and 32 steps: 32 bytes per
 *                              pixel for 65 taps); the other passes replay the record:
 *                              per step a move, the texture gather and the reference's fused
 *                              multiply-add with that step's tap, in the reference's order --
 *                              the same bits for about a fifth of the instructions.
 *                              Applies with the default arithmetic and the grouped walk; a
 *                              record that does not fit the device's free memory falls back
 *                              to walking.
 *   RLIC_B200_PATHS_RECOMPUTE  every pass walks, as the reference does.
 * A process-wide default like the three above, with its own per-thread override (-1 = none). */
#define RLIC_B200_PATHS_RECOMPUTE 0
#define RLIC_B200_PATHS_REPLAY 1
int rlic_b200_set_paths(int which);
int rlic_b200_get_paths(void);
int rlic_b200_set_thread_paths(int which);
int rlic_b200_get_thread_paths(void);
int rlic_b200_get_effective_paths(void);

/* Testing hook (host code only): the (pass, band) launch order of the wavefront schedule
 * for `nbands` bands and `iterations` passes, as pairs pass_band[2k] = pass (1-based),
 * pass_band[2k+1] = band; returns the number of pairs (at most `capacity` are written). */
int64_t rlic_b200_debug_wavefront_order(int64_t nbands, int64_t iterations, int32_t *pass_band,
                                        int64_t capacity);

/* Testing hook (host code only): the row bands the host entry points cut a single ny x nx image
 * into for a `klen`-tap kernel and `iterations` passes -- edges[b] .. edges[b + 1] are the rows
 * of band b: about sixteen bands, each at least two kernel half-widths (and 64 rows, 256 Kpix)
 * tall.  Returns the number of edges (bands + 1; at most `capacity` are written). */
int64_t rlic_b200_debug_band_plan(int64_t ny, int64_t nx, int64_t klen, int64_t iterations, int64_t *edges,
                                  int64_t capacity);

/* Number of kernel launches issued by this library since it was loaded
 * (all threads, all devices).  bench.py reads it around the timed region. */
int64_t rlic_b200_launch_count(void);

/* Testing hook: when `on` is non-zero every later launch uses the 64-bit element
 * index instantiation of the kernels, which is otherwise selected only for
 * buffers of 2^31 cells or more (a 46340 x 46340 image). */
void rlic_b200_debug_force_wide_index(int on);

/* Testing hook: 0 keeps passes over small images (at most 148 x 2048 pixels) on the
 * one-thread-per-pixel kernel instead of the two-warps-per-pixel one that halves their latency
 * (lic_pass_pair_kernel, lic_walk.cuh).  Same bits either way; default 1. */
void rlic_b200_debug_small_image_kernel(int on);

/* Testing / measurement hook: 0 keeps replayed passes (RLIC_B200_PATHS_REPLAY) on the kernel that
 * gathers the texture through L1; 1 (default) lets kernels of up to 65 taps use the one that
 * stages each tile's texture window in shared memory.  Same bits either way.  Also settable
 * with the environment variable RLIC_B200_REPLAY_STAGING=0 (read by the Python binding). */
void rlic_b200_debug_replay_staging(int on);

/* Testing hook (host code only, no GPU needed): what the padded layout puts in
 * cell `cell` of the buffer of the slab {row0, nrows, halo_lo, halo_hi} of an
 * ny x nx image: out[0] = 1 for a pixel; otherwise a wall cell with out[1] = 1
 * if a walker can land on it; out[2], out[3] = buffer row and column of the pixel
 * itself or of the pixel the wall rule (lib.rs:83-95) continues from; out[4] =
 * that pixel's cell minus `cell` (the sentinel's shift). */
int rlic_b200_debug_wall_cell(int64_t ny, int64_t nx, int64_t row0, int64_t nrows,
                              int64_t halo_lo, int64_t halo_hi,
                              int x_left, int x_right, int y_left, int y_right,
                              int64_t cell, int64_t *out);

/* Testing hook (host code only): the buffer geometry the kernels are launched with for
 * that slab and a `klen`-tap kernel (the slab is checked as rlic_b200_pass_slab_* checks
 * it).  out[0..9] = image width, pitch, rows held, cells per field, the columns a walker
 * continues from after leaving through the left / right wall, the buffer rows after the
 * top / bottom wall, and whether the top / bottom wall is reachable from this buffer. */
int rlic_b200_debug_geometry(int64_t ny, int64_t nx, int64_t row0, int64_t nrows,
                             int64_t halo_lo, int64_t halo_hi,
                             int x_left, int x_right, int y_left, int y_right,
                             int64_t klen, int64_t *out);

/*
 * HOST entry points — replace rlic._core.convolve_f32 / convolve_f64
 * (lib.rs:451-482 -> convolve_iteratively, lib.rs:408-443).
 *
 * All pointers are HOST pointers (pageable or pinned).  `out` receives
 * ny*nx values and must not alias an input.  The library uploads the fields,
 * runs `iterations` passes with device-resident ping-pong buffers, and
 * downloads the result; it keeps no reference to host memory after return.
 * iterations <= 0 writes zeros to `out` (lib.rs:423-432: the loop never runs);
 * the Python layer handles iterations == 0 itself (_lib.py:208-209).
 * Runs on the current default device `rlic_b200_set_device` selected for the
 * calling thread (device 0 if never called).
 */
int rlic_b200_convolve_f32(const float *texture, const float *u, const float *v,
                           int64_t ny, int64_t nx,
                           const float *kernel, int64_t klen,
                           int uv_mode,
                           int x_left, int x_right, int y_left, int y_right,
                           int64_t iterations, float *out);

int rlic_b200_convolve_f64(const double *texture, const double *u, const double *v,
                           int64_t ny, int64_t nx,
                           const double *kernel, int64_t klen,
                           int uv_mode,
                           int x_left, int x_right, int y_left, int y_right,
                           int64_t iterations, double *out);

/*
 * Same as rlic_b200_convolve_*, with the reference's `np.any(texture < 0)`
 * validation (_lib.py:174-179) fused into the upload: *texture_has_negative is
 * set to 1 when the texture holds a negative element (NaN does not count, as on
 * the host), else 0.  The result in `out` is unspecified in that case; the
 * Python layer raises the reference's ValueError.  Saves a host pass over the
 * texture that costs more than a GPU iteration on large images.
 */
int rlic_b200_convolve_checked_f32(const float *texture, const float *u, const float *v,
                                   int64_t ny, int64_t nx,
                                   const float *kernel, int64_t klen,
                                   int uv_mode,
                                   int x_left, int x_right, int y_left, int y_right,
                                   int64_t iterations, float *out,
                                   int *texture_has_negative);

int rlic_b200_convolve_checked_f64(const double *texture, const double *u, const double *v,
                                   int64_t ny, int64_t nx,
                                   const double *kernel, int64_t klen,
                                   int uv_mode,
                                   int x_left, int x_right, int y_left, int y_right,
                                   int64_t iterations, double *out,
                                   int *texture_has_negative);

/*
 * Page-locked host blocks for results.  `out` of the host entry points may be
 * any host memory; a binding that has to return a freshly allocated array (as
 * rlic.convolve does, src/lib.rs:442) can take the array's storage from here
 * instead of from the heap: the download then needs neither page faults nor the
 * driver's bounce buffer.  Blocks are cached and reused; at most 4 GiB are held
 * (cached + handed out), beyond which rlic_b200_result_alloc returns NULL and
 * the caller uses ordinary memory.  Page-locking costs more than one download
 * saves, so the first request of a size also returns NULL (one-shot calls stay
 * on ordinary memory), as do requests above 256 MiB.  Also NULL without a CUDA
 * device.
 */
void *rlic_b200_result_alloc(int64_t bytes);
void rlic_b200_result_free(void *block);

/* Device used by the host entry points on the calling thread. */
int rlic_b200_set_device(int device);

/*
 * DEVICE entry points — same computation on buffers already resident in HBM
 * (used by bench.py's device-resident `value`).
 *
 *   d_texture, d_u, d_v : device pointers, dense ny x nx each, read only
 *   d_out               : device pointer, dense ny x nx, receives the result
 *                         (must not alias an input)
 *   kernel              : HOST pointer to the klen taps
 *   stream              : cudaStream_t, or NULL for the legacy default stream
 *
 * The call only enqueues work on `stream`; it does not synchronise.  Scratch
 * (two padded texture buffers -- the reference's two work buffers,
 * README.md:158-164 -- and the packed field) comes from the device's
 * stream-ordered memory pool and is returned to it in stream order.
 */
int rlic_b200_convolve_device_f32(const float *d_texture, const float *d_u, const float *d_v,
                                  int64_t ny, int64_t nx,
                                  const float *kernel, int64_t klen,
                                  int uv_mode,
                                  int x_left, int x_right, int y_left, int y_right,
                                  int64_t iterations, float *d_out, void *stream);

int rlic_b200_convolve_device_f64(const double *d_texture, const double *d_u, const double *d_v,
                                  int64_t ny, int64_t nx,
                                  const double *kernel, int64_t klen,
                                  int uv_mode,
                                  int x_left, int x_right, int y_left, int y_right,
                                  int64_t iterations, double *d_out, void *stream);

/*
 * Device buffer layout of the kernels ("padded").  An image region of `rows`
 * rows by nx columns is held in (rows + 2) * (nx + 2) cells: a pitch of nx + 2
 * and one guard row above and below.  The two extra cells of a row and the
 * guard rows are WALL CELLS: in a texture buffer they mirror the pixel the
 * boundary rule (lib.rs:83-95) sends a walker to, in the field buffer they hold
 * a sentinel record with the offset to that pixel, so the walk needs neither a
 * column counter nor wall compares.  rlic_b200_padded_cells(rows, nx) is that
 * cell count.  The packed FIELD holds one record of 4 scalars per cell,
 *     { u, v, ru, rv }        (16 bytes for f32, 32 bytes for f64)
 * interleaved for f32; for f64 split into two planes of 2 scalars per cell,
 * { u, v } for all cells of the buffer followed by { ru, rv } for all cells
 * (a halo exchange of f64 field rows therefore moves two ranges),
 * where ru, rv are the refined reciprocals that an IEEE division by u, v
 * computes as its first stage; they are walker- and iteration-invariant, so they
 * are computed once per pixel instead of twice per step in the walk.  A
 * component that is exactly 0 stores 2^120; pixels with a non-finite or extreme
 * (outside [2^-40, 2^40]) component carry ru = NaN and take the kernels' generic
 * step.  The layout is private to the library version that produced it: build
 * it with the functions below, never by hand.
 */
int64_t rlic_b200_padded_cells(int64_t rows, int64_t nx);

/* Packed field of a whole image, from two dense planar device arrays into a
 * device buffer of 4 * rlic_b200_padded_cells(ny, nx) scalars (aligned to 4
 * scalars).  The wall sentinels depend on the boundary kinds. */
int rlic_b200_pack_field_f32(const float *d_u, const float *d_v, int64_t ny, int64_t nx,
                             int x_left, int x_right, int y_left, int y_right,
                             float *d_field, void *stream);
int rlic_b200_pack_field_f64(const double *d_u, const double *d_v, int64_t ny, int64_t nx,
                             int x_left, int x_right, int y_left, int y_right,
                             double *d_field, void *stream);

/* Same as rlic_b200_convolve_device_* with the field already packed (same ny,
 * nx and boundary kinds as given to rlic_b200_pack_field_*): callers that run
 * many convolutions over one field pack once. */
int rlic_b200_convolve_packed_f32(const float *d_texture, const float *d_field,
                                  int64_t ny, int64_t nx,
                                  const float *kernel, int64_t klen,
                                  int uv_mode,
                                  int x_left, int x_right, int y_left, int y_right,
                                  int64_t iterations, float *d_out, void *stream);
int rlic_b200_convolve_packed_f64(const double *d_texture, const double *d_field,
                                  int64_t ny, int64_t nx,
                                  const double *kernel, int64_t klen,
                                  int uv_mode,
                                  int x_left, int x_right, int y_left, int y_right,
                                  int64_t iterations, double *d_out, void *stream);

/*
 * DEVICE-RESIDENT BATCH: rlic_b200_convolve_device_* for a stack of `nfields` independent
 * images of ny x nx pixels each (dense, one after the other, in device memory; so are d_u,
 * d_v and d_out).  One launch of each kernel covers the whole stack: every field has its own
 * guard rows and wall cells in the padded scratch, so out[f] equals the single-image result
 * of field f bit for bit (BASELINE config 5's device-side form).
 */
int rlic_b200_convolve_device_batch_f32(const float *d_texture, const float *d_u, const float *d_v,
                                        int64_t nfields, int64_t ny, int64_t nx,
                                        const float *kernel, int64_t klen, int uv_mode,
                                        int x_left, int x_right, int y_left, int y_right,
                                        int64_t iterations, float *d_out, void *stream);
int rlic_b200_convolve_device_batch_f64(const double *d_texture, const double *d_u, const double *d_v,
                                        int64_t nfields, int64_t ny, int64_t nx,
                                        const double *kernel, int64_t klen, int uv_mode,
                                        int x_left, int x_right, int y_left, int y_right,
                                        int64_t iterations, double *d_out, void *stream);

/*
 * ROW SLABS — the building blocks of row-slab sharding.
 * The image has ny x nx pixels globally; a device holds global rows
 * [row0 - halo_lo, row0 + nrows + halo_hi) in PADDED buffers of
 * rlic_b200_padded_cells(halo_lo + nrows + halo_hi, nx) cells that the caller
 * owns: two for the texture (ping-pong) and one, 4 scalars per cell, for the
 * field.  Buffer row r (0 = first halo row) together with its two wall cells is
 * the contiguous cell range [(r + 1) * (nx + 2) - 1, (r + 2) * (nx + 2) - 1):
 * that is the unit a halo exchange copies.  halo_lo / halo_hi must be >= klen/2
 * unless a closed wall bounds the slab on that side (else RLIC_B200_ESHARD);
 * for y-periodic images the halos of the first and last slab wrap around (ring).
 *
 *   slab_pack_field    dense u, v of the OWNED rows (nrows x nx) -> packed records
 *                      of the owned rows (+ the sentinels of any image wall this
 *                      slab touches).  Halo rows arrive by exchange.
 *   slab_pad_texture   dense texture of the owned rows -> padded buffer, likewise
 *   pass_slab          one pass over owned rows [sub_row0, sub_row0 + sub_nrows),
 *                      reading d_texture / d_field, writing d_out (all padded,
 *                      same geometry); the wall cells of the written rows are
 *                      kept up to date.  Walkers use the image-level wall rules
 *                      in global row numbers, so the stitched result is
 *                      bit-identical to an unsharded pass.
 *   slab_unpad_texture padded owned rows -> dense nrows x nx
 * A whole image is the slab row0 = 0, nrows = ny, no halos.
 */
int rlic_b200_slab_pack_field_f32(const float *d_u, const float *d_v, int64_t ny, int64_t nx,
                                  int64_t row0, int64_t nrows, int64_t halo_lo, int64_t halo_hi,
                                  int x_left, int x_right, int y_left, int y_right,
                                  float *d_field, void *stream);
int rlic_b200_slab_pack_field_f64(const double *d_u, const double *d_v, int64_t ny, int64_t nx,
                                  int64_t row0, int64_t nrows, int64_t halo_lo, int64_t halo_hi,
                                  int x_left, int x_right, int y_left, int y_right,
                                  double *d_field, void *stream);

int rlic_b200_slab_pad_texture_f32(const float *d_texture, int64_t ny, int64_t nx,
                                   int64_t row0, int64_t nrows, int64_t halo_lo, int64_t halo_hi,
                                   int x_left, int x_right, int y_left, int y_right,
                                   float *d_padded, void *stream);
int rlic_b200_slab_pad_texture_f64(const double *d_texture, int64_t ny, int64_t nx,
                                   int64_t row0, int64_t nrows, int64_t halo_lo, int64_t halo_hi,
                                   int x_left, int x_right, int y_left, int y_right,
                                   double *d_padded, void *stream);

int rlic_b200_slab_unpad_texture_f32(const float *d_padded, int64_t ny, int64_t nx,
                                     int64_t row0, int64_t nrows, int64_t halo_lo, int64_t halo_hi,
                                     int x_left, int x_right, int y_left, int y_right,
                                     float *d_texture, void *stream);
int rlic_b200_slab_unpad_texture_f64(const double *d_padded, int64_t ny, int64_t nx,
                                     int64_t row0, int64_t nrows, int64_t halo_lo, int64_t halo_hi,
                                     int x_left, int x_right, int y_left, int y_right,
                                     double *d_texture, void *stream);

/* Row-range variants: owned rows [sub_row0, sub_row0 + sub_nrows) only; the dense array
 * holds just those rows (sub_nrows x nx).  With them a slab can be uploaded, converted,
 * computed and downloaded band by band (rlic_b200/sharded.py: convolve_host). */
int rlic_b200_slab_pack_field_rows_f32(const float *d_u, const float *d_v, int64_t ny, int64_t nx,
                                       int64_t row0, int64_t nrows, int64_t halo_lo, int64_t halo_hi,
                                       int64_t sub_row0, int64_t sub_nrows,
                                       int x_left, int x_right, int y_left, int y_right,
                                       float *d_field, void *stream);
int rlic_b200_slab_pack_field_rows_f64(const double *d_u, const double *d_v, int64_t ny, int64_t nx,
                                       int64_t row0, int64_t nrows, int64_t halo_lo, int64_t halo_hi,
                                       int64_t sub_row0, int64_t sub_nrows,
                                       int x_left, int x_right, int y_left, int y_right,
                                       double *d_field, void *stream);
int rlic_b200_slab_pad_texture_rows_f32(const float *d_texture, int64_t ny, int64_t nx,
                                        int64_t row0, int64_t nrows, int64_t halo_lo, int64_t halo_hi,
                                        int64_t sub_row0, int64_t sub_nrows,
                                        int x_left, int x_right, int y_left, int y_right,
                                        float *d_padded, void *stream);
int rlic_b200_slab_pad_texture_rows_f64(const double *d_texture, int64_t ny, int64_t nx,
                                        int64_t row0, int64_t nrows, int64_t halo_lo, int64_t halo_hi,
                                        int64_t sub_row0, int64_t sub_nrows,
                                        int x_left, int x_right, int y_left, int y_right,
                                        double *d_padded, void *stream);
int rlic_b200_slab_unpad_texture_rows_f32(const float *d_padded, int64_t ny, int64_t nx,
                                          int64_t row0, int64_t nrows, int64_t halo_lo, int64_t halo_hi,
                                          int64_t sub_row0, int64_t sub_nrows,
                                          int x_left, int x_right, int y_left, int y_right,
                                          float *d_texture, void *stream);
int rlic_b200_slab_unpad_texture_rows_f64(const double *d_padded, int64_t ny, int64_t nx,
                                          int64_t row0, int64_t nrows, int64_t halo_lo, int64_t halo_hi,
                                          int64_t sub_row0, int64_t sub_nrows,
                                          int x_left, int x_right, int y_left, int y_right,
                                          double *d_texture, void *stream);

int rlic_b200_pass_slab_f32(const float *d_texture, const float *d_field, float *d_out,
                            int64_t ny, int64_t nx,
                            int64_t row0, int64_t nrows, int64_t halo_lo, int64_t halo_hi,
                            int64_t sub_row0, int64_t sub_nrows,
                            const float *kernel, int64_t klen,
                            int uv_mode,
                            int x_left, int x_right, int y_left, int y_right,
                            void *stream);
int rlic_b200_pass_slab_f64(const double *d_texture, const double *d_field, double *d_out,
                            int64_t ny, int64_t nx,
                            int64_t row0, int64_t nrows, int64_t halo_lo, int64_t halo_hi,
                            int64_t sub_row0, int64_t sub_nrows,
                            const double *kernel, int64_t klen,
                            int uv_mode,
                            int x_left, int x_right, int y_left, int y_right,
                            void *stream);

/*
 * HISTOGRAM EQUALISATION of a result -- the usual step between a line integral convolution and
 * its rendering.  rLIC only declares it (`equalize_histogram_f32 / _f64(image, nbins)`,
 * /root/reference/src/rlic/_core.pyi:30-37: no implementation in src/lib.rs; upstream moved it
 * to its sister project `ahe`, /root/reference/README.md:19-23), so the semantics are this
 * library's own, every operation one IEEE operation in the image's type:
 *
 *     lo, hi = minimum, maximum over the pixels that are not NaN;  w = hi - lo
 *     bin(x) = min(nbins - 1, (int) floor(((x - lo) / w) * nbins))          (0 when w == 0)
 *     cdf[b] = (number of non-NaN pixels in bins 0..b) / (number of non-NaN pixels)
 *     out(x) = cdf[bin(x)],  NaN where x is NaN
 *
 * Values are expected finite or NaN; 1 <= nbins <= 2^24.  The `_device_` form works on device
 * memory in place of a host round trip (enqueued on `stream`, scratch from the library's
 * pool); the plain form takes host pointers.  The adaptive variants the reference's stub also
 * lists (sliding tile, tile interpolation: _core.pyi:38-57) are not built.
 */
int rlic_b200_equalize_histogram_device_f32(const float *d_image, int64_t ny, int64_t nx,
                                            int64_t nbins, float *d_out, void *stream);
int rlic_b200_equalize_histogram_device_f64(const double *d_image, int64_t ny, int64_t nx,
                                            int64_t nbins, double *d_out, void *stream);
int rlic_b200_equalize_histogram_f32(const float *image, int64_t ny, int64_t nx, int64_t nbins,
                                     float *out);
int rlic_b200_equalize_histogram_f64(const double *image, int64_t ny, int64_t nx, int64_t nbins,
                                     double *out);

/*
 * MEASUREMENT — the gather ceiling ("an L2 gather peak measured by the
 * build's own microbenchmark, same access count, straight-line walkers").  One launch performs
 * the loads of a pass -- per step one field record and one texture value at the walker's
 * cell -- and the tap FMA, and nothing else; the walkers climb a staircase (+1 column, +1
 * row, ...) forward and descend it backward, so neighbouring threads touch neighbouring cells
 * as walkers on a smooth field do.  `dependent` = 1 makes each address wait for the record
 * loaded before it, as in the real walk; `dependent` = 2 performs the loads of a REPLAYED pass
 * instead (one texture value per step, no field record: the ceiling of the replay kernel).
 * Buffers: a padded texture, a packed field and a
 * padded output of the whole ny x nx image (closed walls), as the slab entry points above
 * produce them.  The caller times the launch; d_padded_out is NOT a convolution result.
 */
int rlic_b200_measure_gather_ceiling_f32(const float *d_padded_texture, const float *d_field,
                                         float *d_padded_out, int64_t ny, int64_t nx,
                                         const float *kernel, int64_t klen, int dependent,
                                         void *stream);
int rlic_b200_measure_gather_ceiling_f64(const double *d_padded_texture, const double *d_field,
                                         double *d_padded_out, int64_t ny, int64_t nx,
                                         const double *kernel, int64_t klen, int dependent,
                                         void *stream);

/*
 * FUSED HALO EXCHANGE (row-slab sharding, one process per GPU; rlic_b200/sharded.py with
 * exchange="peer").  The default of bench.py --gpus N since round 2 (8 B200s: weak scaling
 * 0.97, profiles/r2_session8b_summary.txt).
 *
 *   pass_slab_peer   rlic_b200_pass_slab_* that also stores every result of the rows it
 *                    computes -- the pixels and the two wall cells that travel with each
 *                    row -- into a neighbour's padded buffer `d_peer_out` (mapped with
 *                    rlic_b200_peer_open), `peer_row_delta` buffer rows away from where
 *                    they land locally: the pass over an edge strip and the shipping of
 *                    that strip into the neighbour's halo are one kernel, the transfer
 *                    going over NVLink as the tiles finish.  Default arithmetic only.
 *   peer_alloc       device memory that other processes can map (cudaMalloc + an IPC
 *                    handle of RLIC_B200_PEER_HANDLE_BYTES bytes), zero-filled
 *   peer_open/close  map / unmap another process's allocation from its handle
 *   peer_signal      after everything enqueued so far on `stream`: raise the 32-bit counter
 *                    `d_flag` (normally in a neighbour's memory) to `value`, with
 *                    system-scope fences around the store
 *   peer_signal2 / peer_wait4   the same for two / four counters with one launch
 *   peer_wait        hold `stream` until the counter `d_flag` (in local memory) has reached
 *                    `value` (signed distance, so counters may wrap); after `timeout_ms`
 *                    the wait gives up and sets *d_timed_out (device memory, may be NULL)
 *                    instead of wedging the device
 */
#define RLIC_B200_PEER_HANDLE_BYTES 64
int rlic_b200_pass_slab_peer_f32(const float *d_texture, const float *d_field, float *d_out,
                                 int64_t ny, int64_t nx,
                                 int64_t row0, int64_t nrows, int64_t halo_lo, int64_t halo_hi,
                                 int64_t sub_row0, int64_t sub_nrows,
                                 const float *kernel, int64_t klen,
                                 int uv_mode,
                                 int x_left, int x_right, int y_left, int y_right,
                                 float *d_peer_out, int64_t peer_row_delta,
                                 void *stream);
int rlic_b200_pass_slab_peer_f64(const double *d_texture, const double *d_field, double *d_out,
                                 int64_t ny, int64_t nx,
                                 int64_t row0, int64_t nrows, int64_t halo_lo, int64_t halo_hi,
                                 int64_t sub_row0, int64_t sub_nrows,
                                 const double *kernel, int64_t klen,
                                 int uv_mode,
                                 int x_left, int x_right, int y_left, int y_right,
                                 double *d_peer_out, int64_t peer_row_delta,
                                 void *stream);
/*
 * SLAB PASSES WITH RECORDED PATHS (see RLIC_B200_PATHS_REPLAY above; replaces the body of the
 * iteration loop lib.rs:432-440 for callers that drive the passes of a slab themselves).
 *   path_record_bytes  size of the record of a padded buffer that holds `rows` image rows
 *                      (halo rows included) of `nx` pixels, for a `klen`-tap kernel
 *   pass_slab_paths    rlic_b200_pass_slab_* (d_peer_out NULL) or _pass_slab_peer_* with
 *                      paths_mode  RLIC_B200_PASS_WALK    as those, d_paths ignored
 *                                  RLIC_B200_PASS_RECORD  the walk also writes the record of the
 *                                                         rows it computes into d_paths
 *                                  RLIC_B200_PASS_REPLAY  the rows are computed from the record
 *                                                         (d_field is not read, may be NULL)
 *                      d_paths must be 16-byte aligned (cudaMalloc and every framework
 *                      allocator are).
 *                      The record belongs to (field, uv_mode, boundaries, kernel length, slab):
 *                      replaying it with anything else is the caller's error.  Default
 *                      arithmetic and the grouped walk only (EINVAL otherwise).
 */
#define RLIC_B200_PASS_WALK 0
#define RLIC_B200_PASS_RECORD 1
#define RLIC_B200_PASS_REPLAY 2
int64_t rlic_b200_path_record_bytes(int64_t rows, int64_t nx, int64_t klen);
int rlic_b200_pass_slab_paths_f32(const float *d_texture, const float *d_field, float *d_out,
                                  int64_t ny, int64_t nx,
                                  int64_t row0, int64_t nrows, int64_t halo_lo, int64_t halo_hi,
                                  int64_t sub_row0, int64_t sub_nrows,
                                  const float *kernel, int64_t klen,
                                  int uv_mode,
                                  int x_left, int x_right, int y_left, int y_right,
                                  float *d_peer_out, int64_t peer_row_delta,
                                  int paths_mode, uint32_t *d_paths,
                                  void *stream);
int rlic_b200_pass_slab_paths_f64(const double *d_texture, const double *d_field, double *d_out,
                                  int64_t ny, int64_t nx,
                                  int64_t row0, int64_t nrows, int64_t halo_lo, int64_t halo_hi,
                                  int64_t sub_row0, int64_t sub_nrows,
                                  const double *kernel, int64_t klen,
                                  int uv_mode,
                                  int x_left, int x_right, int y_left, int y_right,
                                  double *d_peer_out, int64_t peer_row_delta,
                                  int paths_mode, uint32_t *d_paths,
                                  void *stream);
int rlic_b200_peer_alloc(int64_t bytes, void **ptr, unsigned char *handle);
int rlic_b200_peer_open(const unsigned char *handle, void **ptr);
int rlic_b200_peer_close(void *ptr);
int rlic_b200_peer_free(void *ptr);
int rlic_b200_peer_signal(uint32_t *d_flag, uint32_t value, void *stream);
int rlic_b200_peer_wait(const uint32_t *d_flag, uint32_t value, int64_t timeout_ms, int *d_timed_out,
                        void *stream);
/* The same for several counters with one launch each (a replayed pass lasts a few hundred
 * microseconds; the one-thread launches count): signal2 raises up to two counters to `value`,
 * wait4 holds the stream until each of up to four counters has reached its own value.  Null
 * flags are skipped (at least one must be given). */
int rlic_b200_peer_signal2(uint32_t *d_flag_a, uint32_t *d_flag_b, uint32_t value, void *stream);
int rlic_b200_peer_wait4(const uint32_t *d_flag0, uint32_t value0, const uint32_t *d_flag1, uint32_t value1,
                         const uint32_t *d_flag2, uint32_t value2, const uint32_t *d_flag3, uint32_t value3,
                         int64_t timeout_ms, int *d_timed_out, void *stream);

/*
 * BATCH of independent fields on the host (BASELINE config 5): `nfields`
 * images of ny x nx stored back to back in each of texture/u/v/out, split
 * whole-image over the listed devices (one host thread + stream pair per
 * device, uploads/downloads overlapped with compute).  devices == NULL or
 * ndev <= 0 means "all visible devices".
 */
int rlic_b200_convolve_batch_f32(const float *texture, const float *u, const float *v,
                                 int64_t nfields, int64_t ny, int64_t nx,
                                 const float *kernel, int64_t klen,
                                 int uv_mode,
                                 int x_left, int x_right, int y_left, int y_right,
                                 int64_t iterations,
                                 const int *devices, int ndev, float *out);

int rlic_b200_convolve_batch_f64(const double *texture, const double *u, const double *v,
                                 int64_t nfields, int64_t ny, int64_t nx,
                                 const double *kernel, int64_t klen,
                                 int uv_mode,
                                 int x_left, int x_right, int y_left, int y_right,
                                 int64_t iterations,
                                 const int *devices, int ndev, double *out);

/* The batch entry with the texture sign check fused into the uploads, as
 * rlic_b200_convolve_checked_* does for one image: *texture_has_negative is set to 1 when any
 * element of any field is negative (`np.any(textures < 0)`, /root/reference/src/rlic/_lib.py:174,
 * which costs about a second on the host for BASELINE config 5's 4 GiB stack). */
int rlic_b200_convolve_batch_checked_f32(const float *texture, const float *u, const float *v,
                                         int64_t nfields, int64_t ny, int64_t nx,
                                         const float *kernel, int64_t klen, int uv_mode,
                                         int x_left, int x_right, int y_left, int y_right,
                                         int64_t iterations, const int *devices, int ndev, float *out,
                                         int *texture_has_negative);
int rlic_b200_convolve_batch_checked_f64(const double *texture, const double *u, const double *v,
                                         int64_t nfields, int64_t ny, int64_t nx,
                                         const double *kernel, int64_t klen, int uv_mode,
                                         int x_left, int x_right, int y_left, int y_right,
                                         int64_t iterations, const int *devices, int ndev, double *out,
                                         int *texture_has_negative);

#ifdef __cplusplus
}
#endif
#endif /* RLIC_B200_H */
