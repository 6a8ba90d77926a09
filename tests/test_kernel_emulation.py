"""The library's kernel source, compiled for the CPU, against the oracle (no GPU needed).

``tests/kernel_emulation`` compiles ``rlic_b200/csrc/lic_walk.cuh`` itself -- the streamline
walk with its fast path, admission test, generic step and wall sentinels, the field packing
and the texture padding -- with g++ and runs it block by block.  These tests hold it to the
oracle bit for bit over the same cases as the ``-m gpu`` parity tests, so the kernel *logic*
is regression-tested wherever the CPU suite runs; the GPU tests remain the proof for the
shipped binary (code generation, the hardware's reciprocal seed).
"""
from __future__ import annotations

import numpy as np
import pytest
from numpy.testing import assert_array_equal

import kernel_emulation as ke
import oracle
from golden_cases import CASES as GOLDEN_CASES, expected, load
from rlic_b200 import _core, workloads

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


def check(tex, u, v, kernel, mode="velocity", walls="closed", iterations=1, **how):
    bnd = WALLS[walls]
    got = ke.convolve(tex, u, v, kernel=kernel, uv_mode=mode, boundaries=bnd, iterations=iterations, **how)
    # oracle variants mirror the crate's features: 3 = fma + branchless (default), 1 = fma alone
    variant = 3 if how.get("branchless", True) else 1
    want = oracle.convolve(np.ascontiguousarray(tex), np.ascontiguousarray(u), np.ascontiguousarray(v),
                           kernel=kernel, uv_mode=mode, boundaries=bnd, iterations=iterations, variant=variant)
    assert got.dtype == tex.dtype and got.shape == tex.shape
    assert_array_equal(got, want)
    return got


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_golden_vectors(name):
    mode, bnd, its = GOLDEN_CASES[name]
    tex, u, v, kernel = load(name)
    got = ke.convolve(tex, u, v, kernel=kernel, uv_mode=mode, boundaries=bnd, iterations=its)
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
    n = 64
    x = np.linspace(0, np.pi, n)
    tex = np.random.default_rng(0).random((n, n))
    u = np.broadcast_to(np.cos(2 * x), (n, n))
    v = np.broadcast_to(np.sin(x), (n, n))          # v[:, 0] == +0.0 exactly
    k = workloads.triangle_kernel(65, np.float64)
    for walls in WALLS:
        check(tex, u, v, k, walls=walls, iterations=2)
    check(tex.astype(np.float32), u.astype(np.float32), v.astype(np.float32),
          k.astype(np.float32), walls="periodic")


def test_infinite_huge_and_denormal_velocities():
    tex, u, v, k = random_case((24, 24), np.float32, 13, seed=8)
    u[7, 7] = np.inf
    v[8, 8] = -np.inf
    u[9, 9] = 3e38
    v[9, 9] = -3e38
    u[10, 10] = 1e-45
    v[11, 11] = -1e-42
    check(tex, u, v, k, walls="periodic", iterations=2)
    check(tex, u, v, k, mode="polarization")


def test_negative_kernel_values_and_nan_texture():
    tex, u, v, k = random_case((20, 20), np.float64, 9, seed=4)
    tex[5, 5] = np.nan
    k[2] = -3.0
    check(tex, u, v, k, iterations=2)


def fuzz_case(seed):
    rng = np.random.default_rng(1000 + seed)
    dtype = [np.float32, np.float64][seed % 2]
    ny, nx = (int(x) for x in rng.integers(1, 90, size=2))
    klen = int(rng.integers(1, 80))
    style = seed % 4
    tex = rng.random((ny, nx)).astype(dtype)
    with np.errstate(all="ignore"):
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
    return tex, u, v, kernel, mode, walls, int(rng.integers(1, 4))


@pytest.mark.parametrize("seed", range(48))
def test_randomised_configurations(seed):
    """The GPU suite's fuzzer (seeds 0-11 are the same cases), four times as many seeds."""
    tex, u, v, kernel, mode, walls, its = fuzz_case(seed)
    with np.errstate(all="ignore"):
        check(tex, u, v, kernel, mode=mode, walls=walls, iterations=its)


def test_64_bit_index_instantiation():
    for dtype, mode, walls in ((np.float32, "velocity", "closed"), (np.float64, "polarization", "periodic"),
                               (np.float32, "polarization", "x-periodic"), (np.float64, "velocity", "y-periodic")):
        check(*random_case((70, 45), dtype, 19, seed=77), mode=mode, walls=walls, iterations=2, wide=True)


@pytest.mark.parametrize("admit", [0, 1, 2, 3])
@pytest.mark.parametrize("flavor", [0, 1])
def test_every_formulation_of_the_fast_path(flavor, admit):
    """The library picks one sign-handling flavour and one admission test per dtype
    (rlic::Tune); all eight combinations must give the same bits."""
    for seed in (2, 5, 6, 9):      # one fuzz case per field style, both dtypes
        tex, u, v, kernel, mode, walls, its = fuzz_case(seed)
        with np.errstate(all="ignore"):
            check(tex, u, v, kernel, mode=mode, walls=walls, iterations=its, flavor=flavor, admit=admit)
    check(*random_case((45, 70), np.float32, 23, seed=11), mode="polarization", walls="periodic",
          flavor=flavor, admit=admit)
    check(*random_case((45, 70), np.float64, 23, seed=12), walls="x-periodic", flavor=flavor, admit=admit)


def test_c1_readme_example_in_full():
    w = workloads.readme_example()
    check(w.texture, w.u, w.v, w.kernel, walls="periodic")
    check(w.texture, w.u, w.v, w.kernel, walls="periodic", iterations=5)


def test_c2_shape_reduced():
    w = workloads.vortex_noise(512, iterations=5)
    got = ke.convolve(w.texture, w.u, w.v, kernel=w.kernel, iterations=5)
    want = oracle.convolve(w.texture, w.u, w.v, kernel=w.kernel, iterations=5, threads=oracle.max_threads())
    assert_array_equal(got, want)


def test_c3_polarization_reduced():
    w = workloads.polarization_split(256, taps=129)
    check(w.texture, w.u, w.v, w.kernel, mode="polarization", walls="x-periodic")


def test_batch_of_fields_in_one_launch():
    rng = np.random.default_rng(14)
    nf, ny, nx = 5, 33, 65
    tex = rng.random((nf, ny, nx), dtype=np.float32)
    u = (rng.random((nf, ny, nx), dtype=np.float32) - 0.5)
    v = (rng.random((nf, ny, nx), dtype=np.float32) - 0.5)
    u[2, 4, 4] = np.nan
    kernel = np.linspace(0.2, 1.0, 17, dtype=np.float32)
    bnd = WALLS["x-periodic"]
    b = ke.Buffers(np.float32, ny, nx, _core.wall_codes(bnd), kernel.size, nfields=nf)
    b.pack_field(u, v)
    b.pad_texture(tex, 0)
    b.run_pass(0, 1, kernel, "polarization")
    b.run_pass(1, 0, kernel, "polarization")
    got = b.unpad_texture(0)
    for f in range(nf):
        want = oracle.convolve(tex[f], u[f], v[f], kernel=kernel, uv_mode="polarization", boundaries=bnd,
                               iterations=2)
        assert_array_equal(got[f], want)


def test_padding_reports_negative_texture_values_but_not_nan():
    b = ke.Buffers(np.float64, 12, 9, (0, 0, 0, 0), 5)
    tex = np.random.default_rng(1).random((12, 9))
    tex[3, 3] = np.nan
    assert b.pad_texture(tex, 0) is False
    tex[11, 8] = -1e-300
    assert b.pad_texture(tex, 0) is True
    tex[11, 8] = -0.0
    assert b.pad_texture(tex, 0) is False


def test_step_counters_tell_how_steps_were_decided():
    ke.step_counts()                                     # reset
    tex = np.random.default_rng(0).random((32, 48))
    ones, zeros = np.ones_like(tex), np.zeros_like(tex)
    k = np.ones(9)
    # uniform flow to the right, periodic: every step is a fast-path step except the wall crossings
    ke.convolve(tex, ones, zeros, kernel=k, boundaries=WALLS["periodic"])
    c = ke.step_counts()
    assert c["step"] == 32 * 48 * 8
    # a crossing is resolved at the start of the step after the one that left the image: the
    # forward walkers from columns 45..47 and the backward ones from 0..2 have such a step left
    assert c["wall"] == 32 * (3 + 3) and c["declined"] == c["wall"] == c["generic"]
    # a field of NaN stops every walker at its first step; nothing reaches the generic step
    ke.convolve(tex, ones * np.nan, zeros, kernel=k)
    c = ke.step_counts()
    assert c["step"] == c["declined"] == 32 * 48 * 2 and c["generic"] == 0 and c["wall"] == 0
    # zero field: every step is declined (flagged pixel) and the generic step stays put
    ke.convolve(tex, zeros, zeros, kernel=k)
    c = ke.step_counts()
    assert c["step"] == c["declined"] == c["generic"] == 32 * 48 * 8 and c["wall"] == 0


@pytest.mark.parametrize("walls", WALLS)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_band_by_band_upload_builds_the_same_buffers(dtype, walls):
    """The host path uploads an image in row bands (convolve_host in lic_api.cu): packing and
    padding band by band, in any order, must leave exactly the buffers of a single call --
    including the guard rows, whose content comes from the first / last image row."""
    tex, u, v, kernel = random_case((37, 29), dtype, 9, seed=3)
    codes = _core.wall_codes(WALLS[walls])
    whole = ke.Buffers(dtype, 37, 29, codes, kernel.size)
    whole.pack_field(u, v)
    whole.pad_texture(tex, 0)
    banded = ke.Buffers(dtype, 37, 29, codes, kernel.size)
    for rb, re in ((20, 31), (0, 7), (31, 37), (7, 20)):
        banded.pack_field(u[rb:re], v[rb:re], rows=(rb, re))
        banded.pad_texture(tex[rb:re], 0, rows=(rb, re))
    assert_array_equal(banded.field.view(np.uint8), whole.field.view(np.uint8))
    # texture cells no walker can reach (corners, pad columns of guard rows) are never written
    reach_w, reach_b = ~np.isnan(whole.tex[0]), ~np.isnan(banded.tex[0])
    tex_nan = np.isnan(tex)
    if not tex_nan.any():
        assert_array_equal(reach_w, reach_b)
    assert_array_equal(banded.tex[0].view(np.uint8), whole.tex[0].view(np.uint8))


# ---------------------------------------------------------------------------
# The grouped walk (rlic_b200.set_walk("grouped"), WALK template bits in lic_walk.cuh): the
# loop-exit test once per group of steps, a tail loop for the remainder, and the backward
# pass / polarization flip applied to the travel direction instead of the record.  It must
# produce the same bits as the per-step walk, i.e. the oracle's; `walk=1` dispatches as the
# library does (rlic::Tune), explicit (flavor, admit, walk) triples cover the formulations
# the lab sweeps.

@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_grouped_walk_golden_vectors(name):
    mode, bnd, its = GOLDEN_CASES[name]
    tex, u, v, kernel = load(name)
    got = ke.convolve(tex, u, v, kernel=kernel, uv_mode=mode, boundaries=bnd, iterations=its, walk=1)
    assert_array_equal(got, expected(name, 3))


@pytest.mark.parametrize("seed", range(48))
def test_grouped_walk_randomised_configurations(seed):
    tex, u, v, kernel, mode, walls, its = fuzz_case(seed)
    with np.errstate(all="ignore"):
        check(tex, u, v, kernel, mode=mode, walls=walls, iterations=its, walk=1)


@pytest.mark.parametrize("walls", WALLS)
@pytest.mark.parametrize("mode", ["velocity", "polarization"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_grouped_walk_special_pixels(dtype, mode, walls):
    check(*random_case((45, 70), dtype, 23, seed=11), mode=mode, walls=walls, iterations=2, walk=1)


@pytest.mark.parametrize("klen", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 33, 64, 200])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_grouped_walk_every_remainder_of_the_group_size(dtype, klen):
    """Groups are 4 steps (f32) or 2 (f64); the forward and backward halves of an odd and of
    an even kernel leave every possible remainder to the tail loop."""
    check(*random_case((19, 21), dtype, klen, seed=klen), mode="polarization", walls="x-periodic", walk=1)
    check(*random_case((19, 21), dtype, klen, seed=klen + 1), walls="periodic", iterations=2, walk=1)


@pytest.mark.parametrize("walk", [1, 9])
@pytest.mark.parametrize("flavor", [0, 1, 2, 3])
def test_grouped_walk_every_formulation(flavor, walk):
    for admit, seed in ((0, 2), (1, 5), (2, 6), (3, 9)):
        tex, u, v, kernel, mode, walls, its = fuzz_case(seed)
        with np.errstate(all="ignore"):
            check(tex, u, v, kernel, mode=mode, walls=walls, iterations=its, flavor=flavor, admit=admit,
                  walk=walk)
    check(*random_case((45, 70), np.float32, 23, seed=11), mode="polarization", walls="periodic",
          flavor=flavor, admit=3, walk=walk)
    check(*random_case((45, 70), np.float64, 22, seed=12), walls="x-periodic", flavor=flavor, admit=2,
          walk=walk)
    check(*random_case((45, 70), np.float64, 23, seed=13), mode="polarization", walls="y-periodic",
          flavor=flavor, admit=2, walk=walk)


@pytest.mark.parametrize("walk", [1, 9])
@pytest.mark.parametrize("admit", [3, 4])
def test_packed_pair_formulation_f32(admit, walk):
    """Flavour 4: both axes of a step as one packed pair (FADD2 / FMUL2 / FFMA2 on the GPU, two
    scalars here), and ADMIT 4: the flag test folded into a NaN-propagating three-input
    minimum.  Same bits as the oracle on special pixels, every wall kind, both modes, zero and
    signed-zero components, huge and tiny velocities, every remainder of the group size."""
    for mode in ("velocity", "polarization"):
        for walls in WALLS:
            check(*random_case((45, 70), np.float32, 23, seed=11), mode=mode, walls=walls, iterations=2,
                  flavor=4, admit=admit, walk=walk)
    for seed in range(24):
        tex, u, v, kernel, mode, walls, its = fuzz_case(seed)
        if tex.dtype != np.float32:
            continue
        with np.errstate(all="ignore"):
            check(tex, u, v, kernel, mode=mode, walls=walls, iterations=its, flavor=4, admit=admit, walk=walk)
    for klen in (1, 2, 3, 4, 5, 6, 7, 8, 9, 10):
        check(*random_case((19, 21), np.float32, klen, seed=klen), mode="polarization", walls="x-periodic",
              flavor=4, admit=admit, walk=walk)
    rng = np.random.default_rng(3)
    tex = rng.random((40, 50), dtype=np.float32)
    k = np.linspace(0.1, 1, 15, dtype=np.float32)
    for u0, v0 in ((1.0, 0.0), (0.0, -1.0), (-0.0, 1.0), (1e-30, 1.0), (3e20, -2.0), (-1.0, -0.0)):
        u = np.full((40, 50), u0, dtype=np.float32)
        v = np.full((40, 50), v0, dtype=np.float32)
        with np.errstate(all="ignore"):
            check(tex, u, v, k, walls="periodic", flavor=4, admit=admit, walk=walk)
            check(tex, u, v, k, mode="polarization", walls="closed", flavor=4, admit=admit, walk=walk)


def test_admit_4_with_the_scalar_flavours_both_dtypes():
    for dtype in (np.float32, np.float64):
        for flavor in (0, 2):
            for mode in ("velocity", "polarization"):
                check(*random_case((45, 70), dtype, 23, seed=21), mode=mode, walls="y-periodic", iterations=2,
                      flavor=flavor, admit=4, walk=1)


def test_grouped_walk_axis_aligned_zero_and_signed_zero_fields():
    rng = np.random.default_rng(3)
    tex = rng.random((40, 50))
    one, zero = np.ones_like(tex), np.zeros_like(tex)
    k = np.linspace(0.1, 1, 15)
    for u, v in ((one, zero), (zero, one), (-one, zero), (zero, -one), (one, one), (-one, one),
                 (zero, zero), (-zero, zero), (one, -zero), (-one, -zero)):
        for walls in WALLS:
            check(tex, u, v, k, walls=walls, walk=1)
            check(tex, u, v, k, mode="polarization", walls=walls, walk=1)
            check(tex.astype(np.float32), u.astype(np.float32), v.astype(np.float32), k.astype(np.float32),
                  walls=walls, walk=1)


def test_grouped_walk_infinite_huge_and_denormal_velocities():
    tex, u, v, k = random_case((24, 24), np.float32, 13, seed=8)
    u[7, 7] = np.inf
    v[8, 8] = -np.inf
    u[9, 9] = 3e38
    v[9, 9] = -3e38
    u[10, 10] = 1e-45
    v[11, 11] = -1e-42
    check(tex, u, v, k, walls="periodic", iterations=2, walk=1)
    check(tex, u, v, k, mode="polarization", walk=1)


def test_grouped_walk_workloads():
    w = workloads.readme_example()                                     # C1 in full
    check(w.texture, w.u, w.v, w.kernel, walls="periodic", walk=1)
    w = workloads.vortex_noise(512, iterations=3)                      # C2, reduced
    got = ke.convolve(w.texture, w.u, w.v, kernel=w.kernel, iterations=3, walk=1)
    want = oracle.convolve(w.texture, w.u, w.v, kernel=w.kernel, iterations=3, threads=oracle.max_threads())
    assert_array_equal(got, want)
    w = workloads.polarization_split(256, taps=129)                    # C3, reduced
    check(w.texture, w.u, w.v, w.kernel, mode="polarization", walls="x-periodic", walk=1)


def test_grouped_walk_wide_indices_global_taps_and_batches():
    check(*random_case((70, 45), np.float32, 19, seed=77), mode="polarization", walls="x-periodic",
          iterations=2, wide=True, walk=1)
    check(*random_case((70, 45), np.float64, 19, seed=78), walls="y-periodic", iterations=2, wide=True, walk=1)
    check(*random_case((9, 12), np.float32, 1001, seed=5), walls="periodic", walk=1)
    check(*random_case((9, 12), np.float64, 500, seed=6), mode="polarization", walls="closed", walk=1)
    rng = np.random.default_rng(14)
    nf, ny, nx = 4, 33, 65
    tex = rng.random((nf, ny, nx), dtype=np.float32)
    u = (rng.random((nf, ny, nx), dtype=np.float32) - 0.5)
    v = (rng.random((nf, ny, nx), dtype=np.float32) - 0.5)
    kernel = np.linspace(0.2, 1.0, 18, dtype=np.float32)
    bnd = WALLS["y-periodic"]
    b = ke.Buffers(np.float32, ny, nx, _core.wall_codes(bnd), kernel.size, nfields=nf)
    b.pack_field(u, v)
    b.pad_texture(tex, 0)
    b.run_pass(0, 1, kernel, "velocity", walk=1)
    b.run_pass(1, 0, kernel, "velocity", walk=1)
    got = b.unpad_texture(0)
    for f in range(nf):
        assert_array_equal(got[f], oracle.convolve(tex[f], u[f], v[f], kernel=kernel, boundaries=bnd, iterations=2))


def test_grouped_walk_decides_steps_like_the_per_step_walk():
    """Same admissions, same wall crossings, same generic steps: the counters agree."""
    w = workloads.vortex_noise(128, iterations=1)
    counts = []
    for walk in (0, 1):
        ke.step_counts()
        ke.convolve(w.texture, w.u, w.v, kernel=w.kernel, walk=walk)
        counts.append(ke.step_counts())
    assert counts[0] == counts[1] and counts[0]["step"] == 128 * 128 * 64



# ---- the `fma`-only arithmetic (rlic_b200.set_arithmetic("fma"): the x86-64 wheels' build) ----
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_fma_only_golden_vectors(name):
    mode, bnd, its = GOLDEN_CASES[name]
    tex, u, v, kernel = load(name)
    got = ke.convolve(tex, u, v, kernel=kernel, uv_mode=mode, boundaries=bnd, iterations=its, branchless=False)
    assert_array_equal(got, expected(name, 1))


@pytest.mark.parametrize("seed", range(48))
def test_fma_only_randomised_configurations(seed):
    tex, u, v, kernel, mode, walls, its = fuzz_case(seed)
    with np.errstate(all="ignore"):
        check(tex, u, v, kernel, mode=mode, walls=walls, iterations=its, branchless=False)


@pytest.mark.parametrize("walls", WALLS)
@pytest.mark.parametrize("mode", ["velocity", "polarization"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_fma_only_special_pixels(dtype, mode, walls):
    check(*random_case((45, 70), dtype, 23, seed=11), mode=mode, walls=walls, iterations=2, branchless=False)


def test_fma_only_axis_aligned_zero_and_signed_zero_fields():
    rng = np.random.default_rng(3)
    tex = rng.random((40, 50))
    one, zero = np.ones_like(tex), np.zeros_like(tex)
    k = np.linspace(0.1, 1, 15)
    for u, v in ((one, zero), (zero, one), (-one, zero), (zero, -one), (one, -zero), (-zero, -one),
                 (zero, zero), (-zero, zero), (one, -one)):
        for walls in WALLS:
            check(tex, u, v, k, walls=walls, branchless=False)
            check(tex, u, v, k, mode="polarization", walls=walls, branchless=False)


def test_fma_only_workloads_and_the_builds_really_differ():
    w = workloads.readme_example()
    check(w.texture, w.u, w.v, w.kernel, walls="periodic", iterations=5, branchless=False)
    w = workloads.vortex_noise(512, iterations=2)
    default = check(w.texture, w.u, w.v, w.kernel, iterations=2)
    fma_only = check(w.texture, w.u, w.v, w.kernel, iterations=2, branchless=False, wide=True)
    differing = np.mean(default != fma_only)
    # a last-bit difference in an edge time matters where it flips a `tx < ty` decision:
    # after two passes ~8 % of the pixels of this run differ between the two builds
    assert 0 < differing < 0.5, differing


# ---- row slabs: what rlic_b200.sharded and rlic_b200_pass_slab_* rely on ------------------
def slab_pass_all(tex, u, v, kernel, bnd, mode, cuts, bands=1):
    """tests/test_slab.py on the CPU: cut the image into slabs, fill their halos with plain
    copies of padded rows (ring order when y is periodic), run one pass per slab in two
    sub-ranges, stitch."""
    ny, nx = tex.shape
    pitch, h = nx + 2, kernel.size // 2
    walls = _core.wall_codes(bnd)
    periodic_y = bnd[1][0] == "periodic"
    edges = [0, *cuts, ny]
    slabs = []
    for r0, r1 in zip(edges[:-1], edges[1:]):
        lo = h if (r0 > 0 or periodic_y) else 0
        hi = h if (r1 < ny or periodic_y) else 0
        b = ke.Buffers(tex.dtype, ny, nx, walls, kernel.size, slab=(r0, r1 - r0, lo, hi))
        own = (lo, lo + r1 - r0)
        b.pack_field(u[r0:r1], v[r0:r1], rows=own)
        b.pad_texture(tex[r0:r1], 0, rows=own)
        slabs.append(dict(r0=r0, r1=r1, lo=lo, hi=hi, b=b))

    def rows(s, buf, a, width, plane=0):   # padded row `a` of a slab buffer, with its wall cells
        base = plane * s["b"].cells
        return buf[(base + (a + 1) * pitch - 1) * width:(base + (a + 2) * pitch - 1) * width]

    f_width, f_planes = (4, 1) if tex.dtype == np.float32 else (2, 2)

    def owner(g):
        g %= ny
        for s in slabs:
            if s["r0"] <= g < s["r1"]:
                return s, s["lo"] + g - s["r0"]
        raise AssertionError

    for s in slabs:
        wanted = list(range(s["r0"] - s["lo"], s["r0"])) + list(range(s["r1"], s["r1"] + s["hi"]))
        for k, g in enumerate(wanted):
            brow = k if k < s["lo"] else s["lo"] + (s["r1"] - s["r0"]) + (k - s["lo"])
            src, srow = owner(g)
            rows(s, s["b"].tex[0], brow, 1)[:] = rows(src, src["b"].tex[0], srow, 1)
            for plane in range(f_planes):
                rows(s, s["b"].field, brow, f_width, plane)[:] = rows(src, src["b"].field, srow, f_width, plane)

    out = np.empty_like(tex)
    for s in slabs:
        n = s["r1"] - s["r0"]
        s["b"].tex[1][:] = 0
        for a, c in ((0, n // 3), (n // 3, n)):
            s["b"].run_pass(0, 1, kernel, mode, rows=(s["lo"] + a, c - a))
        out[s["r0"]:s["r1"]] = s["b"].unpad_texture(1, rows=(s["lo"], s["lo"] + n))
    return out


SLAB_CASES = {
    "closed-f32": (np.float32, "velocity", "closed", [40, 90]),
    "y-periodic-f64": (np.float64, "velocity", "y-periodic", [64]),
    "all-periodic-pol-f32": (np.float32, "polarization", "periodic", [33, 66, 99]),
    "x-periodic-pol-f64": (np.float64, "polarization", "x-periodic", [50]),
}


@pytest.mark.parametrize("name", SLAB_CASES)
def test_stitched_slabs_equal_the_whole_image(name):
    dtype, mode, walls, cuts = SLAB_CASES[name]
    rng = np.random.default_rng(31)
    ny, nx = 128, 70
    tex = rng.random((ny, nx)).astype(dtype)
    u = (rng.random((ny, nx)) - 0.5).astype(dtype)
    v = (rng.random((ny, nx)) - 0.5).astype(dtype)
    u[5, 5] = np.nan
    v[60, 3] = u[60, 3] = 0
    kernel = (rng.random(33) + 0.1).astype(dtype)
    got = slab_pass_all(tex, u, v, kernel, WALLS[walls], mode, cuts)
    want = oracle.convolve(tex, u, v, kernel=kernel, uv_mode=mode, boundaries=WALLS[walls])
    assert_array_equal(got, want)


# ---- the wavefront schedule of the host path (rlic_b200_set_schedule) ----------------------
def wavefront_order(nbands, iterations):
    import ctypes

    out = np.zeros(2 * nbands * iterations, dtype=np.int32)
    n = _core.lib.rlic_b200_debug_wavefront_order(nbands, iterations,
                                                  out.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                                  nbands * iterations)
    assert n == nbands * iterations
    return [tuple(int(x) for x in pair) for pair in out.reshape(-1, 2)]


def run_in_order(order, tex, u, v, kernel, walls, mode, band_rows, iterations):
    """convolve_host's buffers and roles, its launches issued one after another in `order`:
    exactly what an in-order stream does.  tex[0] is the uploaded texture and doubles as
    the second work buffer; pass p writes work buffer (p - 1) % 2 of {tex[1], tex[0]}."""
    ny, nx = tex.shape
    b = ke.Buffers(tex.dtype, ny, nx, _core.wall_codes(WALLS[walls]), kernel.size)
    b.pack_field(u, v)
    b.pad_texture(tex, 0)
    b.tex[1][:] = np.nan
    work = (1, 0)
    for p, band in order:
        src = 0 if p == 1 else work[(p - 2) & 1]
        r0, r1 = band * band_rows, min(ny, (band + 1) * band_rows)
        b.run_pass(src, work[(p - 1) & 1], kernel, mode, rows=(r0, r1 - r0))
    return b.unpad_texture(work[(iterations - 1) & 1])


@pytest.mark.parametrize("nbands,iterations", [(2, 2), (5, 2), (5, 5), (3, 7), (8, 5), (7, 1)])
def test_wavefront_order_respects_every_dependency(nbands, iterations):
    order = wavefront_order(nbands, iterations)
    assert sorted(order) == [(p, b) for p in range(1, iterations + 1) for b in range(nbands)]
    at = {pb: k for k, pb in enumerate(order)}
    for (p, b), k in at.items():
        if p > 1:   # reads, and overwrites what was read by, pass p-1 of the neighbouring bands
            for nb in (b - 1, b, b + 1):
                if 0 <= nb < nbands:
                    assert at[(p - 1, nb)] < k
    # the diagonal order: slots b + (p - 1) in turn, ascending passes within a slot, so that the
    # last pass of a band runs only `iterations - 1` bands behind the first
    slots = [b + (p - 1) for p, b in order]
    assert slots == sorted(slots)
    for k in range(1, len(order)):
        (p0, b0), (p1, b1) = order[k - 1], order[k]
        if b0 + (p0 - 1) == b1 + (p1 - 1):
            assert p0 < p1
    for b in range(nbands):
        assert at[(iterations, b)] - at[(1, b)] <= iterations * (iterations - 1) + iterations


@pytest.mark.parametrize("dtype,mode,walls,iterations", [
    (np.float32, "velocity", "closed", 5),
    (np.float64, "polarization", "x-periodic", 4),
    (np.float32, "polarization", "closed", 2),
    (np.float64, "velocity", "closed", 3),
])
def test_wavefront_schedule_computes_the_same_image(dtype, mode, walls, iterations):
    ny, band_rows = 150, 32                       # five bands, the last one short
    tex, u, v, kernel = random_case((ny, 41), dtype, 33, seed=21)     # reach 16 = half a band
    nbands = -(-ny // band_rows)
    got = run_in_order(wavefront_order(nbands, iterations), tex, u, v, kernel, walls, mode, band_rows, iterations)
    want = oracle.convolve(tex, u, v, kernel=kernel, uv_mode=mode, boundaries=WALLS[walls], iterations=iterations)
    assert_array_equal(got, want)


def test_an_order_that_breaks_a_dependency_is_caught():
    """The previous test has teeth: issue pass 2 of band 0 before pass 1 of band 1."""
    ny, band_rows, iterations = 150, 32, 3
    tex, u, v, kernel = random_case((ny, 41), np.float32, 33, seed=21)
    order = wavefront_order(5, iterations)
    i, j = order.index((2, 0)), order.index((1, 1))
    assert j < i
    order[i], order[j] = order[j], order[i]
    got = run_in_order(order, tex, u, v, kernel, "closed", "velocity", band_rows, iterations)
    want = oracle.convolve(tex, u, v, kernel=kernel, boundaries=WALLS["closed"], iterations=iterations)
    assert not np.array_equal(got, want, equal_nan=True)


# ---------------------------------------------------------------------------
# Path signatures (workloads.path_probe): an exact measure of path divergence.

def test_path_probe_counts_diverging_paths_not_rounding():
    """On the probe inputs a pass returns exact integer sums.  The kernel source (either walk)
    visits exactly the oracle's pixels; the reference's other build (`fma` alone) takes a
    different path for a small fraction of the walkers, and the probe counts those."""
    w = workloads.vortex_noise(256, iterations=1)
    probe, ones = workloads.path_probe(w.texture.shape, np.float32, w.kernel.size)
    want = oracle.convolve(probe, w.u, w.v, kernel=ones)
    assert_array_equal(want, np.round(want))                      # integer sums: nothing was rounded
    assert want.max() < 2 ** 24
    for walk in (0, 1):
        assert_array_equal(ke.convolve(probe, w.u, w.v, kernel=ones, walk=walk), want)
    other = oracle.convolve(probe, w.u, w.v, kernel=ones, variant=1)
    diverged = float(np.mean(other != want))
    assert 0.0 < diverged < 0.02
    # the value-based comparison on a noise texture sees (at least) the same walkers
    noisy = oracle.convolve(w.texture, w.u, w.v, kernel=w.kernel)
    noisy_other = oracle.convolve(w.texture, w.u, w.v, kernel=w.kernel, variant=1)
    assert float(np.mean(noisy != noisy_other)) >= diverged * 0.9
    with pytest.raises(ValueError):
        workloads.path_probe((4, 4), np.float32, 200)
