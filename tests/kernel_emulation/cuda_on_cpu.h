// CUDA built-ins used by rlic_b200/csrc/lic_walk.cuh, for compiling that file with g++.
//
// TEST INFRASTRUCTURE.  With RLIC_HOST_EMULATION defined, the kernel source of the library
// is compiled for the CPU so that tests can run the real walk, packing and padding code
// against the oracle on a machine without a GPU (tests/test_kernel_emulation.py).  Nothing
// here is part of, linked into, or reachable from librlic_b200.so.
//
// Every arithmetic intrinsic maps to the IEEE operation it names (single rounding,
// round-to-nearest-even; build with -ffp-contract=off).  The one thing a CPU cannot
// reproduce is the hardware's approximate reciprocal (MUFU.RCP), which seeds the packed
// field's refined reciprocals: rcp_approx() returns the correctly rounded reciprocal
// (f32) or its upper word (f64) instead.  The refinement steps that follow are the
// library's own; whether the short division that consumes them is exact with the *real*
// seed is what tools/kernel_lab.cu measured on the GPU.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __grid_constant__

struct float4 { float x, y, z, w; };
struct uint4 { unsigned x, y, z, w; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return {x, y, z, w}; }
struct double2 { double x, y; };
static inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return {x, y}; }

struct Dim3 { unsigned x, y, z; };
extern thread_local Dim3 threadIdx, blockIdx, blockDim, gridDim;

template <typename A, typename B> static inline A bits_as(B b)
{
    static_assert(sizeof(A) == sizeof(B), "size mismatch");
    A a;
    std::memcpy(&a, &b, sizeof a);
    return a;
}

static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }

static inline int __float_as_int(float x) { return bits_as<int>(x); }
static inline unsigned __float_as_uint(float x) { return bits_as<unsigned>(x); }
static inline float __int_as_float(int x) { return bits_as<float>(x); }
static inline int __double2hiint(double x) { return (int)(bits_as<uint64_t>(x) >> 32); }
static inline int __double2loint(double x) { return (int)(uint32_t)bits_as<uint64_t>(x); }
static inline double __hiloint2double(int hi, int lo)
{
    return bits_as<double>(((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo);
}
static inline double __longlong_as_double(long long x) { return bits_as<double>(x); }
static inline long long __double_as_longlong(double x) { return bits_as<long long>(x); }

template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }

namespace emulated {
// How the steps of a pass were decided (see lic_walk.cuh: half_walk): every step, the ones
// the fast path declined, of those the wall crossings (sentinel cells) and the ones that
// went through the generic step (the rest stopped on a NaN).
struct StepCounts { unsigned long long step, declined, wall, generic; };
extern thread_local StepCounts step_counts;
#define RLIC_EMU_EVENT(which) (++emulated::step_counts.which)

// A CTA's shared memory and its barrier, for the one kernel that has them (the staged replay): the
// emulation runs the threads of a block one after another, twice -- a first sweep in which
// every thread stops at the barrier (the window is then complete), a second in which the
// threads carry on past it.  emulate.cpp drives the sweeps.
extern thread_local bool past_barrier;
extern thread_local unsigned char shared_bytes[256 * 1024];
template <typename T> static inline T *shared_window() { return reinterpret_cast<T *>(shared_bytes); }
static inline bool barrier_then_compute() { return past_barrier; }

static inline float rcp_approx(float b) { return (float)(1.0 / (double)b); }
static inline double rcp_approx(double b) { return __hiloint2double(__double2hiint(1.0 / b), 0); }
}  // namespace emulated
