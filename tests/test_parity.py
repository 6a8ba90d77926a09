"""Bit-for-bit parity of the CUDA path with the CPU oracle (crate-default
arithmetic: fma + branchless), through the public API and the raw C ABI.

Covers the edge cases listed in SURVEY.md section 4: exact zeros and signed
zeros, NaN, kernels longer than the image, even kernels, L in {1, 2}, 1xN / Nx1
images, strided and stride-0 inputs, every boundary combination, both modes and
dtypes, tile-edge sizes, the frozen golden vectors, and the BASELINE configs.
"""

import ctypes
import threading

import numpy as np
import pytest
from numpy.testing import assert_array_equal

import oracle
import rlic_b200 as rlic
from rlic_b200 import _core, workloads

from golden_cases import CASES, as_spec, expected, load

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["small-image kernel", "one thread per pixel"])
def both_pass_kernels(request):
    """Passes over small images run on the two-warps-per-pixel kernel (lic_pass_pair_kernel);
    every test of this file runs a second time with that kernel switched off, so that the
    edge cases -- nearly all of them small images -- reach both kernels."""
    _core.lib.rlic_b200_debug_small_image_kernel(1 if request.param == "small-image kernel" else 0)
    try:
        yield
    finally:
        _core.lib.rlic_b200_debug_small_image_kernel(1)


WALLS = {
    "closed": (("closed", "closed"), ("closed", "closed")),
    "periodic": (("periodic", "periodic"), ("periodic", "periodic")),
    "x-periodic": (("periodic", "periodic"), ("closed", "closed")),
    "y-periodic": (("closed", "closed"), ("periodic", "periodic")),
}


def random_case(shape, dtype, klen, seed, specials=True):
    rng = np.random.default_rng(seed)
    tex = rng.random(shape).astype(dtype)
    u = (rng.random(shape) - 0.5).astype(dtype)
    v = (rng.random(shape) - 0.5).astype(dtype)
    if specials and min(shape) >= 8:
        u[1, 2] = v[1, 2] = 0.0
        u[3, 4] = np.nan
        v[5, 1] = -0.0
        u[2, 5], v[2, 5] = -0.0, 0.0
        u[4, 3] = 0.0
        v[6, 6] = np.nan
    kernel = (rng.random(klen) - 0.2).astype(dtype)
    return tex, u, v, kernel


def check(tex, u, v, kernel, mode="velocity", walls="closed", iterations=1):
    bnd = WALLS[walls]
    got = rlic.convolve(tex, u, v, kernel=kernel, uv_mode=mode, boundaries=as_spec(bnd),
                        iterations=iterations)
    want = oracle.convolve(tex, u, v, kernel=kernel, uv_mode=mode, boundaries=bnd,
                           iterations=iterations)
    assert got.dtype == tex.dtype and got.shape == tex.shape
    assert_array_equal(got, want)   # NaNs compare equal position-wise here
    return got


@pytest.mark.parametrize("name", CASES)
def test_golden_vectors(name):
    mode, bnd, its = CASES[name]
    tex, u, v, kernel = load(name)
    got = rlic.convolve(tex, u, v, kernel=kernel, uv_mode=mode, boundaries=as_spec(bnd), iterations=its)
    assert_array_equal(got, expected(name, 3))


@pytest.mark.parametrize("walls", WALLS)
@pytest.mark.parametrize("mode", ["velocity", "polarization"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_random_fields_with_special_pixels(dtype, mode, walls):
    check(*random_case((45, 70), dtype, 23, seed=11), mode=mode, walls=walls, iterations=2)


@pytest.mark.parametrize("klen", [1, 2, 3, 4, 5, 8, 33, 64, 200])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_kernel_lengths_including_longer_than_image(dtype, klen):
    check(*random_case((19, 21), dtype, klen, seed=klen), mode="polarization", walls="x-periodic")
    check(*random_case((19, 21), dtype, klen, seed=klen + 1), walls="periodic", iterations=2)


def test_taps_beyond_the_parameter_block_use_the_global_path():
    # > 768 f32 taps / > 384 f64 taps do not fit the launch-parameter block
    check(*random_case((9, 12), np.float32, 1001, seed=5), walls="periodic")
    check(*random_case((9, 12), np.float64, 500, seed=6), mode="polarization", walls="closed")


@pytest.mark.parametrize(
    "shape", [(1, 1), (1, 40), (40, 1), (2, 2), (8, 32), (9, 33), (7, 31), (16, 64), (17, 65), (64, 3)]
)
def test_degenerate_and_tile_edge_shapes(shape):
    for walls in ("closed", "periodic"):
        check(*random_case(shape, np.float64, 9, seed=sum(shape), specials=False), walls=walls,
              mode="polarization", iterations=2)


def test_uniform_and_axis_aligned_fields():
    rng = np.random.default_rng(3)
    tex = rng.random((40, 50))
    one, zero = np.ones_like(tex), np.zeros_like(tex)
    k = np.linspace(0.1, 1, 15)
    for u, v in ((one, zero), (zero, one), (-one, zero), (zero, -one), (one, one), (-one, one),
                 (zero, zero), (-zero, zero), (one, -one)):
        for walls in WALLS:
            check(tex, u, v, k, walls=walls)
            check(tex, u, v, k, mode="polarization", walls=walls)


def test_fields_with_exact_grid_zeros():
    # the branchless formula turns +-0 velocities into 0/0 when the walker sits
    # exactly on an edge (last table rows)
    n = 64
    x = np.linspace(0, np.pi, n)
    rng = np.random.default_rng(0)
    tex = rng.random((n, n))
    u = np.broadcast_to(np.cos(2 * x), (n, n))
    v = np.broadcast_to(np.sin(x), (n, n))          # v[:, 0] == +0.0 exactly
    assert v[0, 0] == 0.0
    k = workloads.triangle_kernel(65, np.float64)
    for walls in WALLS:
        check(tex, u, v, k, walls=walls, iterations=2)
    check(tex.astype(np.float32), u.astype(np.float32), v.astype(np.float32),
          k.astype(np.float32), walls="periodic")


def test_infinite_and_huge_velocities():
    tex, u, v, k = random_case((24, 24), np.float32, 13, seed=8)
    u[7, 7] = np.inf
    v[8, 8] = -np.inf
    u[9, 9] = 3e38
    v[9, 9] = -3e38
    u[10, 10] = 1e-45    # denormal
    v[11, 11] = -1e-42
    check(tex, u, v, k, walls="periodic", iterations=2)
    check(tex, u, v, k, mode="polarization")


def test_negative_kernel_values_and_nan_texture():
    tex, u, v, k = random_case((20, 20), np.float64, 9, seed=4)
    tex[5, 5] = np.nan
    k[2] = -3.0
    check(tex, u, v, k, iterations=2)


def test_strided_and_broadcast_inputs():
    rng = np.random.default_rng(2)
    big = rng.random((60, 90))
    tex = big[::2, ::3]                                # non-contiguous view
    u = np.broadcast_to(rng.random(30) - 0.5, (30, 30))  # stride-0 rows
    v = (rng.random((30, 30)) - 0.5).T                 # F-order
    k = np.linspace(0, 1, 12)[::-1]                    # negative stride
    got = check(tex, u, v, k, walls="y-periodic")
    assert got.flags.c_contiguous


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_many_iterations(dtype):
    check(*random_case((33, 47), dtype, 7, seed=9), iterations=11, walls="periodic")


def test_c1_readme_example():
    w = workloads.readme_example()
    got = rlic.convolve(w.texture, w.u, w.v, **w.kwargs())
    p = ("periodic", "periodic")
    assert_array_equal(got, oracle.convolve(w.texture, w.u, w.v, kernel=w.kernel, boundaries=(p, p)))
    got5 = rlic.convolve(w.texture, w.u, w.v, kernel=w.kernel, boundaries="periodic", iterations=5)
    assert_array_equal(
        got5, oracle.convolve(w.texture, w.u, w.v, kernel=w.kernel, boundaries=(p, p), iterations=5)
    )


def test_c2_shape_reduced():
    w = workloads.vortex_noise(512, iterations=5)
    got = rlic.convolve(w.texture, w.u, w.v, **w.kwargs())
    want = oracle.convolve(w.texture, w.u, w.v, kernel=w.kernel, iterations=5,
                           threads=oracle.max_threads())
    assert_array_equal(got, want)


def test_c3_polarization_reduced():
    w = workloads.polarization_split(256, taps=129)
    got = rlic.convolve(w.texture, w.u, w.v, **w.kwargs())
    bnd = (("periodic", "periodic"), ("closed", "closed"))
    want = oracle.convolve(w.texture, w.u, w.v, kernel=w.kernel, uv_mode="polarization", boundaries=bnd)
    assert_array_equal(got, want)


def test_c2_full_size_against_oracle_bands_and_properties():
    """4096^2 f32, 65 taps: pass 1 checked on bands against the oracle, all five
    passes checked through size-independent properties."""
    w = workloads.vortex_noise(4096, iterations=5)
    one = rlic.convolve(w.texture, w.u, w.v, kernel=w.kernel, iterations=1)
    for r0 in (0, 1000, 2040, 4096 - 24):
        band = oracle.pass_rows(w.texture, w.u, w.v, kernel=w.kernel, rows=(r0, r0 + 24))
        assert_array_equal(one[r0:r0 + 24], band)
    five = rlic.convolve(w.texture, w.u, w.v, **w.kwargs())
    # iterations compose: 5 = 1 + 4
    four_more = rlic.convolve(one, w.u, w.v, kernel=w.kernel, iterations=4)
    assert_array_equal(five, four_more)
    # (transpose symmetry is NOT asserted here: the reference breaks ties towards y,
    # and this analytic field produces exact ties on its diagonals)
    assert np.isfinite(five).all()


def test_transpose_symmetry_at_scale():
    # reference tests/test_convolution.py:69-82 on a random 1024^2 f64 field.  The
    # property needs tie-free walks (ties go to y, which a transpose turns into x):
    # in f64 an exact tx == ty has probability ~2^-53 per step; in f32 at this size
    # a handful of ties do occur, so f32 is not asserted.
    rng = np.random.default_rng(21)
    n = 1024
    tex = rng.random((n, n))
    u = rng.random((n, n)) - 0.5
    v = rng.random((n, n)) - 0.5
    k = workloads.triangle_kernel(65, np.float64)
    a = rlic.convolve(tex, u, v, kernel=k, iterations=3, boundaries="periodic")
    b = rlic.convolve(tex.T, v.T, u.T, kernel=k, iterations=3, boundaries="periodic").T
    assert_array_equal(a, b)


def test_c2_full_size_full_oracle():
    """The whole 5-iteration headline configuration against the oracle on all
    host cores (a few seconds per iteration per 8 cores)."""
    w = workloads.vortex_noise(4096, iterations=5)
    got = rlic.convolve(w.texture, w.u, w.v, **w.kwargs())
    want = oracle.convolve(w.texture, w.u, w.v, kernel=w.kernel, iterations=5,
                           threads=oracle.max_threads())
    assert_array_equal(got, want)


def test_mismatch_fraction_against_the_pypi_x86_64_variant(capsys):
    """Informational: the published x86_64 wheels are built fma-only (branching
    edge-time formula, cd.yml:89,139,210), the crate default and this package are
    fma+branchless.  The two differ by one ulp in an edge time now and then, which
    occasionally flips a `tx < ty` decision and sends a walker down another path
    (SURVEY.md section 0.3).  Report how rare that is; it must stay a tiny
    minority and everything else must agree to north_star's 1e-5 of the range."""
    w = workloads.vortex_noise(512, iterations=1)
    got = rlic.convolve(w.texture, w.u, w.v, kernel=w.kernel)
    other = oracle.convolve(w.texture, w.u, w.v, kernel=w.kernel, variant=oracle.VARIANT_FMA)
    err = np.abs(got - other) / np.ptp(other)
    diverged = float(np.mean(err > 1e-5))
    bit_equal = float(np.mean(got == other))
    with capsys.disabled():
        print(f"\n[vs fma-only build] bit-equal {bit_equal:.4%}, beyond 1e-5 of range {diverged:.4%}, "
              f"max {err.max():.3e}")
    assert diverged < 5e-3
    assert bit_equal > 0.5


def test_path_divergence_is_zero_against_the_oracle_and_small_against_the_other_build(capsys):
    """north_star: "the fraction of pixels whose traced path diverges is reported".  On the
    path-signature inputs (workloads.path_probe) a pass returns exact integer sums, so the
    comparison counts different *paths*: none against the oracle of the default build, a
    small minority against the `fma`-only build (whose edge times differ in the last bit)."""
    for n, dtype in ((1024, np.float32), (512, np.float64)):
        w = workloads.vortex_noise(n, dtype=dtype, iterations=1)
        probe, ones = workloads.path_probe(w.texture.shape, dtype, w.kernel.size)
        got = rlic.convolve(probe, w.u, w.v, kernel=ones)
        want = oracle.convolve(probe, w.u, w.v, kernel=ones, threads=oracle.max_threads())
        assert_array_equal(got, want)
        other = oracle.convolve(probe, w.u, w.v, kernel=ones, variant=oracle.VARIANT_FMA,
                                threads=oracle.max_threads())
        diverged = float(np.mean(got != other))
        with capsys.disabled():
            print(f"\n[path divergence, {np.dtype(dtype).name} {n}x{n}] vs default build 0, "
                  f"vs fma-only build {diverged:.4%}")
        assert diverged < 0.02


def test_concurrent_calls_are_independent():
    cases = [random_case((64, 64), np.float32, 9 + 2 * t, seed=40 + t) for t in range(6)]
    want = [oracle.convolve(t, u, v, kernel=k, iterations=3) for t, u, v, k in cases]
    got = [None] * len(cases)

    def work(i):
        t, u, v, k = cases[i]
        for _ in range(5):
            got[i] = rlic.convolve(t, u, v, kernel=k, iterations=3)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(cases))]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    for g, w_ in zip(got, want):
        assert_array_equal(g, w_)


def test_batch_entry_point_matches_per_field_calls():
    rng = np.random.default_rng(12)
    nf, ny, nx = 5, 40, 56
    tex = rng.random((nf, ny, nx), dtype=np.float32)
    u = rng.random((nf, ny, nx), dtype=np.float32) - 0.5
    v = rng.random((nf, ny, nx), dtype=np.float32) - 0.5
    u[2, 3, 3] = np.nan
    k = workloads.triangle_kernel(33, np.float32)
    out = np.empty_like(tex)
    p = ctypes.POINTER(ctypes.c_float)
    rc = _core.lib.rlic_b200_convolve_batch_f32(
        tex.ctypes.data_as(p), u.ctypes.data_as(p), v.ctypes.data_as(p), nf, ny, nx,
        k.ctypes.data_as(p), k.size, 0, 0, 0, 1, 1, 3, None, 0, out.ctypes.data_as(p))
    _core.check(rc)
    bnd = (("closed", "closed"), ("periodic", "periodic"))
    for f in range(nf):
        assert_array_equal(out[f], oracle.convolve(tex[f], u[f], v[f], kernel=k, boundaries=bnd, iterations=3))


def test_negative_texture_is_caught_on_the_device():
    # >= 65536 pixels: the sign check runs on the GPU during the upload
    tex, u, v, k = random_case((300, 300), np.float32, 5, seed=2)
    tex[123, 45] = -1e-30
    with pytest.raises(ValueError, match=r"^Found invalid texture element\(s\)\. Expected only positive values\.$"):
        rlic.convolve(tex, u, v, kernel=k)
    tex[123, 45] = np.nan      # NaN is not negative
    tex[0, 0] = 0.0
    tex[1, 1] = -0.0           # neither is -0.0 (np.any(tex < 0) is False)
    check(tex, u, v, k)


def test_64_bit_index_instantiation():
    """Buffers of >= 2^31 elements use 64-bit element indices; force that code path
    on ordinary sizes and hold it to the same bits."""
    _core.lib.rlic_b200_debug_force_wide_index(1)
    try:
        for dtype, mode, walls in ((np.float32, "velocity", "closed"), (np.float64, "polarization", "periodic"),
                                   (np.float32, "polarization", "x-periodic"), (np.float64, "velocity", "y-periodic")):
            check(*random_case((70, 45), dtype, 19, seed=77), mode=mode, walls=walls, iterations=2)
    finally:
        _core.lib.rlic_b200_debug_force_wide_index(0)


@pytest.mark.parametrize("seed", range(12))
def test_randomised_configurations(seed):
    """Seeded fuzzing over shapes, kernel lengths, modes, walls, dtypes and field styles."""
    rng = np.random.default_rng(1000 + seed)
    dtype = [np.float32, np.float64][seed % 2]
    ny, nx = (int(x) for x in rng.integers(1, 90, size=2))
    klen = int(rng.integers(1, 80))
    style = seed % 4
    tex = rng.random((ny, nx)).astype(dtype)
    if style == 0:      # smooth field with exact zeros on grid lines
        y, x = np.meshgrid(np.linspace(-1, 1, ny), np.linspace(-1, 1, nx), indexing="ij")
        u, v = np.sin(3 * y).astype(dtype), (x * y).astype(dtype)
    elif style == 1:    # piecewise constant with sign flips and zero blocks
        u = rng.choice([-1.0, 0.0, 1.0, 0.5], size=(ny, nx)).astype(dtype)
        v = rng.choice([-2.0, 0.0, 1.0], size=(ny, nx)).astype(dtype)
    elif style == 2:    # wide dynamic range, some non-finite
        u = (rng.standard_normal((ny, nx)) * 10.0 ** rng.integers(-30, 30, size=(ny, nx))).astype(dtype)
        v = (rng.standard_normal((ny, nx)) * 10.0 ** rng.integers(-30, 30, size=(ny, nx))).astype(dtype)
        u[rng.random((ny, nx)) < 0.02] = np.nan
        v[rng.random((ny, nx)) < 0.02] = np.inf
    else:               # plain noise
        u = (rng.random((ny, nx)) - 0.5).astype(dtype)
        v = (rng.random((ny, nx)) - 0.5).astype(dtype)
    kernel = (rng.random(klen) - 0.3).astype(dtype)
    mode = ["velocity", "polarization"][int(rng.integers(2))]
    walls = list(WALLS)[int(rng.integers(len(WALLS)))]
    with np.errstate(all="ignore"):
        check(tex, u, v, kernel, mode=mode, walls=walls, iterations=int(rng.integers(1, 4)))


def test_public_batch_api_matches_per_field_convolve():
    rng = np.random.default_rng(14)
    nf, ny, nx = 7, 33, 65
    tex = rng.random((nf, ny, nx))
    u = rng.random((nf, ny, nx)) - 0.5
    v = rng.random((nf, ny, nx)) - 0.5
    u[3, 5, 5] = 0.0
    v[4, 6, 6] = np.nan
    k = workloads.triangle_kernel(17, np.float64)
    got = rlic.convolve_batch(tex, u, v, kernel=k, uv_mode="polarization",
                              boundaries={"x": "periodic", "y": "closed"}, iterations=2, devices=[0])
    for f in range(nf):
        want = rlic.convolve(tex[f], u[f], v[f], kernel=k, uv_mode="polarization",
                             boundaries={"x": "periodic", "y": "closed"}, iterations=2)
        assert_array_equal(got[f], want)


def test_launch_counter_counts_passes():
    tex, u, v, k = random_case((16, 16), np.float32, 5, seed=1, specials=False)
    before = _core.launch_count()
    rlic.convolve(tex, u, v, kernel=k, iterations=4)
    # 4 passes + field packing + texture padding + un-padding of the result
    assert _core.launch_count() - before == 4 + 3
