/*
 * TEST INFRASTRUCTURE — CPU oracle for the rlic.convolve hot path.
 *
 * This is a plain-C restatement of the algorithm in the reference's Rust core
 * (/root/reference/src/lib.rs, 485 lines) used ONLY as a checker by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 * The product (rlic_b200) never imports, links or executes anything in this
 * directory.
 *
 * Pinning status: the Rust reference cannot be compiled in this image (no
 * rustc/cargo/maturin, no network), and its tests hold no stored arrays, so
 * this oracle is pinned against what the reference's own tests DO hold:
 *   - the six Rust unit known-answer tests (src/lib.rs:139-151,182-206,275-297)
 *   - the exact-equality properties of tests/test_convolution.py
 *     (NaN-field scaling kernel[mid]**n, transpose symmetry, eye(5) polarization
 *     cases, default-argument equalities)
 *   - a second, independent pure-Python restatement (oracle/pyoracle.py)
 *   - the only outputs of the real rLIC available offline: the five LIC
 *     panels of its README figures (the PNG files under static/, seeded inputs), which this
 *     oracle reproduces to within rounding of the 8-bit colours
 *     (tests/test_reference_images.py; about 1/256 of the dynamic range per
 *     pixel, so an algorithmic pin, not a bit-level one)
 *   - the eight inputs of the reference's regression test (its tests/
 *     test_regressions.py, held there to vectorplot's LIC at rtol 1.5e-7), against
 *     an exact rational-arithmetic tracer at the same tolerances
 *     (tests/test_regressions.py; vectorplot itself is not installable here)
 * It is NOT pinned against output ARRAYS produced by the reference itself:
 * at the bit level the status is "parity unpinned" (DESIGN.md section 3 says
 * the same).
 *
 * Arithmetic variants mirror the reference's Cargo features
 * (Cargo.toml:27-30): bit 0 = fma, bit 1 = branchless.  Variant 3
 * (fma+branchless) is the crate default and the one the CUDA path targets.
 *
 * Build: see oracle/Makefile  (-O3 -march=x86-64-v3 -ffp-contract=off, which
 * mirrors .cargo/config.toml:3 and Rust's no-contraction rule).
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <sched.h>
#include <stdatomic.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define LIC_ORACLE_EMPTY_KERNEL 1
#define LIC_ORACLE_NOMEM 2
#define LIC_ORACLE_BAD_VARIANT 3

typedef struct {
    size_t ny, nx;            /* rows (y, parallel to v), columns (x, parallel to u) */
    int x_left, x_right;      /* 0 closed, 1 periodic */
    int y_left, y_right;
    int polarization;         /* 0 velocity, 1 polarization */
} lic_geometry;

/* Runs `body(arg)` on `threads` threads (the caller is one of them).  Helper
 * threads are pinned, one per CPU of the caller's affinity mask in turn, so that
 * the all-cores baseline does not depend on where the scheduler puts them (the
 * caller keeps its own mask: the library never changes its caller's affinity). */
static void lic_run_threads(void *(*body)(void *), void *arg, int threads)
{
    if (threads > 256)
        threads = 256;
    pthread_t tid[256];
    int started = 0;
    cpu_set_t allowed;
    int ncpu = 0, cpus[CPU_SETSIZE];
    if (sched_getaffinity(0, sizeof allowed, &allowed) == 0)
        for (int c = 0; c < CPU_SETSIZE; ++c)
            if (CPU_ISSET(c, &allowed))
                cpus[ncpu++] = c;
    for (int t = 1; t < threads; ++t) {
        pthread_attr_t attr;
        pthread_attr_init(&attr);
        if (ncpu > 1) {
            cpu_set_t one;
            CPU_ZERO(&one);
            CPU_SET(cpus[t % ncpu], &one);
            pthread_attr_setaffinity_np(&attr, sizeof one, &one);
        }
        if (pthread_create(&tid[started], &attr, body, arg) == 0)
            ++started;
        else if (pthread_create(&tid[started], NULL, body, arg) == 0)   /* pinning refused: unpinned */
            ++started;
        pthread_attr_destroy(&attr);
    }
    body(arg);
    for (int t = 0; t < started; ++t)
        pthread_join(tid[t], NULL);
}

#define GLUE_(a, b) a##b
#define GLUE(a, b) GLUE_(a, b)

/* ---- float instantiations ---- */
#define REAL float
#define COPYSIGN copysignf
#define FABS fabsf

#define NAME(x) GLUE(x, _f32_v0)
#define FMA(a, b, c) ((a) * (b) + (c))
#define BRANCHLESS 0
#include "lic_walk.inc"
#undef NAME
#undef FMA
#undef BRANCHLESS

#define NAME(x) GLUE(x, _f32_v1)
#define FMA(a, b, c) fmaf((a), (b), (c))
#define BRANCHLESS 0
#include "lic_walk.inc"
#undef NAME
#undef FMA
#undef BRANCHLESS

#define NAME(x) GLUE(x, _f32_v2)
#define FMA(a, b, c) ((a) * (b) + (c))
#define BRANCHLESS 1
#include "lic_walk.inc"
#undef NAME
#undef FMA
#undef BRANCHLESS

#define NAME(x) GLUE(x, _f32_v3)
#define FMA(a, b, c) fmaf((a), (b), (c))
#define BRANCHLESS 1
#include "lic_walk.inc"
#undef NAME
#undef FMA
#undef BRANCHLESS

#undef REAL
#undef COPYSIGN
#undef FABS

/* ---- double instantiations ---- */
#define REAL double
#define COPYSIGN copysign
#define FABS fabs

#define NAME(x) GLUE(x, _f64_v0)
#define FMA(a, b, c) ((a) * (b) + (c))
#define BRANCHLESS 0
#include "lic_walk.inc"
#undef NAME
#undef FMA
#undef BRANCHLESS

#define NAME(x) GLUE(x, _f64_v1)
#define FMA(a, b, c) fma((a), (b), (c))
#define BRANCHLESS 0
#include "lic_walk.inc"
#undef NAME
#undef FMA
#undef BRANCHLESS

#define NAME(x) GLUE(x, _f64_v2)
#define FMA(a, b, c) ((a) * (b) + (c))
#define BRANCHLESS 1
#include "lic_walk.inc"
#undef NAME
#undef FMA
#undef BRANCHLESS

#define NAME(x) GLUE(x, _f64_v3)
#define FMA(a, b, c) fma((a), (b), (c))
#define BRANCHLESS 1
#include "lic_walk.inc"
#undef NAME
#undef FMA
#undef BRANCHLESS

#undef REAL
#undef COPYSIGN
#undef FABS

static lic_geometry make_geometry(int64_t ny, int64_t nx, int polarization,
                                  int x_left, int x_right, int y_left,
                                  int y_right)
{
    lic_geometry g;
    g.ny = (size_t)ny; g.nx = (size_t)nx;
    g.x_left = x_left; g.x_right = x_right;
    g.y_left = y_left; g.y_right = y_right;
    g.polarization = polarization;
    return g;
}

/*
 * Same argument order as the product's C ABI (include/rlic_b200.h) plus
 * `variant` (bit 0 fma, bit 1 branchless) and `threads` (1 = as the reference
 * runs; >1 = rows split over pthreads for the all-cores baseline).
 */
#define DISPATCH(T, sfx)                                                        \
    switch (variant) {                                                          \
    case 0: return run_##sfx##_v0(tex, u, v, taps, (size_t)ntaps, &g, iterations, threads, out); \
    case 1: return run_##sfx##_v1(tex, u, v, taps, (size_t)ntaps, &g, iterations, threads, out); \
    case 2: return run_##sfx##_v2(tex, u, v, taps, (size_t)ntaps, &g, iterations, threads, out); \
    case 3: return run_##sfx##_v3(tex, u, v, taps, (size_t)ntaps, &g, iterations, threads, out); \
    default: return LIC_ORACLE_BAD_VARIANT;                                     \
    }

int lic_oracle_convolve_f32(const float *tex, const float *u, const float *v,
                            int64_t ny, int64_t nx, const float *taps,
                            int64_t ntaps, int polarization, int x_left,
                            int x_right, int y_left, int y_right,
                            int64_t iterations, int variant, int threads,
                            float *out)
{
    lic_geometry g = make_geometry(ny, nx, polarization, x_left, x_right, y_left, y_right);
    DISPATCH(float, f32)
}

int lic_oracle_convolve_f64(const double *tex, const double *u, const double *v,
                            int64_t ny, int64_t nx, const double *taps,
                            int64_t ntaps, int polarization, int x_left,
                            int x_right, int y_left, int y_right,
                            int64_t iterations, int variant, int threads,
                            double *out)
{
    lic_geometry g = make_geometry(ny, nx, polarization, x_left, x_right, y_left, y_right);
    DISPATCH(double, f64)
}

/* One pass restricted to rows [row_begin,row_end): lets a test check a band of
 * a large image against the CUDA path without walking all of it.  Only those
 * rows of `out` are written, on `threads` threads.  Variant 3 (crate default) only. */
int lic_oracle_pass_rows_f32(const float *tex, const float *u, const float *v,
                             int64_t ny, int64_t nx, const float *taps,
                             int64_t ntaps, int polarization, int x_left,
                             int x_right, int y_left, int y_right,
                             int64_t row_begin, int64_t row_end, int threads, float *out)
{
    lic_geometry g = make_geometry(ny, nx, polarization, x_left, x_right, y_left, y_right);
    if (ntaps <= 0) return LIC_ORACLE_EMPTY_KERNEL;
    pass_band_f32_v3(tex, u, v, taps, (size_t)ntaps, &g, out, (size_t)row_begin, (size_t)row_end, threads);
    return 0;
}

int lic_oracle_pass_rows_f64(const double *tex, const double *u, const double *v,
                             int64_t ny, int64_t nx, const double *taps,
                             int64_t ntaps, int polarization, int x_left,
                             int x_right, int y_left, int y_right,
                             int64_t row_begin, int64_t row_end, int threads, double *out)
{
    lic_geometry g = make_geometry(ny, nx, polarization, x_left, x_right, y_left, y_right);
    if (ntaps <= 0) return LIC_ORACLE_EMPTY_KERNEL;
    pass_band_f64_v3(tex, u, v, taps, (size_t)ntaps, &g, out, (size_t)row_begin, (size_t)row_end, threads);
    return 0;
}

/* ---- hooks for the reference's Rust unit KATs ---- */
#define KAT_TIME(T, sfx)                                                        \
    T lic_oracle_edge_time_##sfx(T vel, T frac, int variant)                    \
    {                                                                           \
        switch (variant) {                                                      \
        case 0: return kat_edge_time_##sfx##_v0(vel, frac);                     \
        case 1: return kat_edge_time_##sfx##_v1(vel, frac);                     \
        case 2: return kat_edge_time_##sfx##_v2(vel, frac);                     \
        default: return kat_edge_time_##sfx##_v3(vel, frac);                    \
        }                                                                       \
    }
KAT_TIME(float, f32)
KAT_TIME(double, f64)

#define KAT_CROSS(T, sfx)                                                       \
    void lic_oracle_cross_##sfx(T mu, T mv, int64_t ny, int64_t nx, int x_left, \
                                int x_right, int y_left, int y_right,           \
                                int variant, long *i, long *j, T *fx, T *fy)    \
    {                                                                           \
        lic_geometry g = make_geometry(ny, nx, 0, x_left, x_right, y_left, y_right); \
        switch (variant) {                                                      \
        case 0: kat_cross_##sfx##_v0(mu, mv, &g, i, j, fx, fy); break;          \
        case 1: kat_cross_##sfx##_v1(mu, mv, &g, i, j, fx, fy); break;          \
        case 2: kat_cross_##sfx##_v2(mu, mv, &g, i, j, fx, fy); break;          \
        default: kat_cross_##sfx##_v3(mu, mv, &g, i, j, fx, fy); break;         \
        }                                                                       \
    }
KAT_CROSS(float, f32)
KAT_CROSS(double, f64)

#include <unistd.h>
int lic_oracle_max_threads(void)
{
    /* the CPUs this process may run on (a container's share), not the machine's */
    cpu_set_t allowed;
    if (sched_getaffinity(0, sizeof allowed, &allowed) == 0) {
        int n = CPU_COUNT(&allowed);
        if (n > 0)
            return n;
    }
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}
