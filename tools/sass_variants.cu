// Instantiates the pass-kernel formulations that tools/sass_steps.py counts (developer tool;
// nothing here is linked into librlic_b200.so).  Compiled to a cubin only, never run.
#include "../rlic_b200/csrc/lic_walk.cuh"
using namespace rlic;
using PT32 = ParamTaps<float, kParamTapBytes / 4>;
using PT64 = ParamTaps<double, kParamTapBytes / 8>;

template <typename T, bool POL, typename PT, int WALK, int FLAVOR> const void *variant()
{
    using Tn = Tune<T, POL>;
    return (const void *)&lic_pass_kernel<T, POL, PT, int, kTileW, kTileH, Tn::walk_unroll, Tn::walk_min_blocks,
                                          FLAVOR, Tn::walk_admit, true, WALK>;
}
template <typename T, bool POL, typename PT, int WALK, int FLAVOR, int ADMIT, int UNROLL> const void *variant_a()
{
    using Tn = Tune<T, POL>;
    return (const void *)&lic_pass_kernel<T, POL, PT, int, kTileW, kTileH, UNROLL, Tn::walk_min_blocks,
                                          FLAVOR, ADMIT, true, WALK>;
}
template <typename T, bool POL, typename PT> const void *shipped()
{
    return (const void *)&lic_pass_kernel<T, POL, PT, int>;
}
template <typename T, bool POL, typename PT> const void *tuned_grouped()
{
    using Tn = Tune<T, POL>;
    return variant<T, POL, PT, Tn::walk, Tn::walk_flavor>();
}

// the recording walk as launch_one() instantiates it (lic_api.cu: record_min_blocks)
template <typename T, bool POL, typename PT> const void *tuned_recording()
{
    using Tn = Tune<T, POL>;
    constexpr int minb = Tn::walk_min_blocks > 6 ? 6 : Tn::walk_min_blocks;
    return (const void *)&lic_pass_kernel<T, POL, PT, int, kTileW, kTileH, Tn::walk_unroll, minb, Tn::walk_flavor,
                                          Tn::walk_admit, true, Tn::walk, true>;
}

#define SWEEP(T, POL, PT, FLAVOR) \
    variant<T, POL, PT, 1, FLAVOR>(), variant<T, POL, PT, 3, FLAVOR>(), variant<T, POL, PT, 5, FLAVOR>(), \
    variant<T, POL, PT, 7, FLAVOR>(), variant<T, POL, PT, 9, FLAVOR>(), variant<T, POL, PT, 11, FLAVOR>()

const void *table[] = {
    variant_a<float, false, PT32, 1, 4, 4, 4>(), variant_a<float, false, PT32, 7, 2, 4, 4>(),
    variant_a<float, false, PT32, 1, 4, 4, 8>(), variant_a<float, false, PT32, 3, 4, 4, 4>(),
    variant_a<float, true, PT32, 1, 4, 4, 4>(), variant_a<float, true, PT32, 1, 0, 4, 4>(),
    variant_a<double, false, PT64, 9, 0, 4, 4>(), variant_a<double, true, PT64, 9, 0, 4, 4>(),
    variant<float, false, PT32, 7, 4>(), variant<float, false, PT32, 5, 4>(), variant<float, false, PT32, 1, 4>(),
    variant<float, true, PT32, 1, 4>(), variant<float, true, PT32, 7, 4>(),
    shipped<float, false, PT32>(), shipped<float, true, PT32>(),
    shipped<double, false, PT64>(), shipped<double, true, PT64>(),
    tuned_grouped<float, false, PT32>(), tuned_grouped<float, true, PT32>(),
    tuned_grouped<double, false, PT64>(), tuned_grouped<double, true, PT64>(),
    tuned_recording<float, false, PT32>(), tuned_recording<float, true, PT32>(),
    tuned_recording<double, false, PT64>(), tuned_recording<double, true, PT64>(),
    SWEEP(float, false, PT32, 1), SWEEP(float, false, PT32, 2), SWEEP(float, false, PT32, 0),
    SWEEP(float, true, PT32, 0), SWEEP(float, true, PT32, 2),
    SWEEP(double, false, PT64, 0), SWEEP(double, false, PT64, 2),
    SWEEP(double, true, PT64, 0), SWEEP(double, true, PT64, 2),
};
