// Histogram equalisation of a device-resident image: the usual step after a line integral
// convolution when the result is rendered.
//
// Reference: rLIC only DECLARES this operation -- `equalize_histogram_f32 / _f64(image, nbins)`
// in /root/reference/src/rlic/_core.pyi:30-37, with no implementation in src/lib.rs: upstream
// moved it to the sister project `ahe` (/root/reference/README.md:19-23), which is not
// available offline.  The semantics below are therefore this repository's own, fixed by the
// oracle oracle/equalize.py and stated in include/rlic_b200.h; every floating-point operation
// is a single IEEE operation in T, so the CUDA path and the NumPy oracle agree bit for bit:
//
//     lo, hi = minimum, maximum over the pixels that are not NaN;  w = hi - lo
//     bin(x) = min(nbins - 1, (int) floor(((x - lo) / w) * T(nbins)))        (0 when w == 0)
//     hist[b] = number of non-NaN pixels in bin b;  n = their total
//     cdf[b] = T(hist[0] + ... + hist[b]) / T(n)
//     out(x) = cdf[bin(x)],  NaN where x is NaN
//
// Four streaming kernels, HBM-bound: extrema, histogram (shared-memory counters per CTA, then
// 64-bit global atomics: integer sums, so the result does not depend on the order), the scan
// that turns counts into the cumulative distribution (one CTA), and the map.  Three reads and
// one write of the image: 16 bytes per f32 pixel (BASELINE config 2's result: 268 MB, about
// 45 us at the measured 6.5 TB/s).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace rlic {

struct EqualizeScratch {            // device memory, zeroed / initialised by equalize_init_kernel
    long long lo_key, hi_key;       // extrema as order-preserving integer keys
    unsigned long long count;       // pixels that are not NaN
};

template <typename T> struct OrderedKey;
template <> struct OrderedKey<float> {
    using Int = int;
    static __device__ __forceinline__ long long encode(float x)
    {
        const int i = __float_as_int(x);
        return (long long)(i ^ ((i >> 31) & 0x7fffffff));
    }
    static __device__ __forceinline__ float decode(long long k)
    {
        const int i = (int)k;
        return __int_as_float(i ^ ((i >> 31) & 0x7fffffff));
    }
};
template <> struct OrderedKey<double> {
    static __device__ __forceinline__ long long encode(double x)
    {
        const long long i = __double_as_longlong(x);
        return i ^ ((i >> 63) & 0x7fffffffffffffffll);
    }
    static __device__ __forceinline__ double decode(long long k)
    {
        return __longlong_as_double(k ^ ((k >> 63) & 0x7fffffffffffffffll));
    }
};

__global__ void equalize_init_kernel(EqualizeScratch *s, unsigned long long *hist, long long nbins)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) {
        s->lo_key = 0x7fffffffffffffffll;
        s->hi_key = -0x7fffffffffffffffll - 1;
        s->count = 0;
    }
    for (long long b = t; b < nbins; b += (long long)gridDim.x * blockDim.x)
        hist[b] = 0;
}

template <typename T>
__global__ void __launch_bounds__(256)
equalize_extrema_kernel(const T *__restrict__ image, long long n, EqualizeScratch *s)
{
    long long lo = 0x7fffffffffffffffll, hi = -0x7fffffffffffffffll - 1;
    unsigned long long count = 0;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const T x = image[t];
        if (x == x) {
            const long long k = OrderedKey<T>::encode(x);
            lo = k < lo ? k : lo;
            hi = k > hi ? k : hi;
            ++count;
        }
    }
    for (int d = 16; d; d >>= 1) {
        const long long lo2 = __shfl_xor_sync(0xffffffffu, lo, d), hi2 = __shfl_xor_sync(0xffffffffu, hi, d);
        lo = lo2 < lo ? lo2 : lo;
        hi = hi2 > hi ? hi2 : hi;
        count += __shfl_xor_sync(0xffffffffu, count, d);
    }
    if ((threadIdx.x & 31) == 0 && count) {
        atomicMin(&s->lo_key, lo);
        atomicMax(&s->hi_key, hi);
        atomicAdd(&s->count, count);
    }
}

// bin of a pixel that is not NaN; lo, w, nbins as in the header comment
template <typename T>
__device__ __forceinline__ long long equalize_bin(T x, T lo, T w, T nbins_t, long long nbins)
{
    if (w == T(0))
        return 0;
    const T scaled = Fp<T>::mul(Fp<T>::div(Fp<T>::sub(x, lo), w), nbins_t);
    long long b = (long long)floor((double)scaled);     // scaled >= 0: exact in double for either T
    return b > nbins - 1 ? nbins - 1 : (b < 0 ? 0 : b);
}

constexpr int kEqualizeSharedBins = 4096;

template <typename T>
__global__ void __launch_bounds__(256)
equalize_histogram_kernel(const T *__restrict__ image, long long n, const EqualizeScratch *s,
                          unsigned long long *hist, long long nbins)
{
    __shared__ unsigned local[kEqualizeSharedBins];
    const bool use_shared = nbins <= kEqualizeSharedBins;
    if (use_shared) {
        for (int b = threadIdx.x; b < nbins; b += blockDim.x)
            local[b] = 0;
        __syncthreads();
    }
    const T lo = OrderedKey<T>::decode(s->lo_key), hi = OrderedKey<T>::decode(s->hi_key);
    const T w = Fp<T>::sub(hi, lo), nbins_t = (T)nbins;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const T x = image[t];
        if (x == x) {
            const long long b = equalize_bin<T>(x, lo, w, nbins_t, nbins);
            if (use_shared)
                atomicAdd(&local[b], 1u);
            else
                atomicAdd(&hist[b], 1ull);
        }
    }
    if (use_shared) {
        __syncthreads();
        for (int b = threadIdx.x; b < nbins; b += blockDim.x)
            if (local[b])
                atomicAdd(&hist[b], (unsigned long long)local[b]);
    }
}

// counts -> cumulative distribution, one CTA of 1024 threads: each thread sums a contiguous
// chunk, the chunk totals are scanned in shared memory, each thread then writes its chunk.
template <typename T>
__global__ void __launch_bounds__(1024)
equalize_cdf_kernel(const unsigned long long *__restrict__ hist, long long nbins, const EqualizeScratch *s,
                    T *__restrict__ cdf)
{
    __shared__ unsigned long long totals[1024];
    const long long chunk = (nbins + 1023) / 1024;
    const long long b0 = (long long)threadIdx.x * chunk, b1 = b0 + chunk < nbins ? b0 + chunk : nbins;
    unsigned long long sum = 0;
    for (long long b = b0; b < b1; ++b)
        sum += hist[b];
    totals[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (int t = 0; t < 1024; ++t) {
            const unsigned long long mine = totals[t];
            totals[t] = run;                              // exclusive prefix of the chunk totals
            run += mine;
        }
    }
    __syncthreads();
    unsigned long long run = totals[threadIdx.x];
    const unsigned long long n = s->count;
    for (long long b = b0; b < b1; ++b) {
        run += hist[b];
        if (sizeof(T) == 4)
            cdf[b] = (T)__fdiv_rn(__ull2float_rn(run), __ull2float_rn(n));
        else
            cdf[b] = (T)__ddiv_rn(__ull2double_rn(run), __ull2double_rn(n));
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
equalize_map_kernel(const T *__restrict__ image, long long n, const EqualizeScratch *s, const T *__restrict__ cdf,
                    long long nbins, T *__restrict__ out)
{
    const T lo = OrderedKey<T>::decode(s->lo_key), hi = OrderedKey<T>::decode(s->hi_key);
    const T w = Fp<T>::sub(hi, lo), nbins_t = (T)nbins;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const T x = image[t];
        out[t] = x == x ? __ldg(cdf + equalize_bin<T>(x, lo, w, nbins_t, nbins)) : x;
    }
}

}  // namespace rlic
