"""The two formulations of the pass kernels on the GPU.  The grouped walk (loop-exit test once
per group of steps; backward pass and polarization flip applied to the travel direction instead
of the record) is the default since round 2 and is what every other GPU test exercises; here
the per-step walk of round 1 -- selected for this thread's calls with
``rlic_b200.options(walk="per-step")`` -- must return the same bits, i.e. the oracle's.  The kernel source is held to the oracle on the CPU by
tests/test_kernel_emulation.py (``test_grouped_walk_*``); what only hardware can show is
nvcc's code for it."""
from __future__ import annotations

import numpy as np
import pytest
from numpy.testing import assert_array_equal

import oracle
import rlic_b200
from golden_cases import CASES, as_spec, expected, load
from rlic_b200 import _core, workloads
from test_kernel_emulation import WALLS, fuzz_case, random_case

pytestmark = pytest.mark.gpu


@pytest.fixture
def per_step():
    assert rlic_b200.get_walk() == "grouped"
    with rlic_b200.options(walk="per-step"):
        assert rlic_b200.effective_options()["walk"] == "per-step"
        yield
    assert rlic_b200.effective_options()["walk"] == "grouped"


def check(tex, u, v, kernel, mode="velocity", walls="closed", iterations=1):
    bnd = WALLS[walls]
    with np.errstate(all="ignore"):
        got = rlic_b200.convolve(tex, u, v, kernel=kernel, uv_mode=mode, boundaries=as_spec(bnd),
                                 iterations=iterations)
        want = oracle.convolve(tex, u, v, kernel=kernel, uv_mode=mode, boundaries=bnd, iterations=iterations)
    assert_array_equal(got, want)
    return got


@pytest.mark.parametrize("name", CASES)
def test_golden_vectors(per_step, name):
    mode, bnd, its = CASES[name]
    tex, u, v, kernel = load(name)
    got = rlic_b200.convolve(tex, u, v, kernel=kernel, uv_mode=mode, boundaries=as_spec(bnd), iterations=its)
    assert_array_equal(got, expected(name, 3))


@pytest.mark.parametrize("seed", range(24))
def test_randomised_configurations(per_step, seed):
    tex, u, v, kernel, mode, walls, its = fuzz_case(seed)
    check(tex, u, v, kernel, mode, walls, its)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_special_pixels_every_wall_and_mode(per_step, dtype):
    for mode in ("velocity", "polarization"):
        for walls in WALLS:
            check(*random_case((45, 70), dtype, 23, seed=11), mode, walls, 2)


@pytest.mark.parametrize("klen", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 33, 64, 200, 1001])
def test_every_remainder_of_the_group_size_and_long_kernels(per_step, klen):
    for dtype in (np.float32, np.float64):
        check(*random_case((19, 21), dtype, klen, seed=klen), "polarization", "x-periodic")
        check(*random_case((19, 21), dtype, klen, seed=klen + 1), "velocity", "periodic", 2)


def test_workloads(per_step):
    w = workloads.readme_example()                                     # C1 in full
    check(w.texture, w.u, w.v, w.kernel, walls="periodic")
    for dtype in (np.float32, np.float64):                             # C2, reduced
        w = workloads.vortex_noise(1024, dtype=dtype, iterations=3)
        got = rlic_b200.convolve(w.texture, w.u, w.v, kernel=w.kernel, iterations=3)
        want = oracle.convolve(w.texture, w.u, w.v, kernel=w.kernel, iterations=3, threads=oracle.max_threads())
        assert_array_equal(got, want)
    w = workloads.polarization_split(512, taps=129)                    # C3, reduced
    check(w.texture, w.u, w.v, w.kernel, "polarization", "x-periodic")


def test_full_size_c2_pass_equals_the_default_walk():
    """4096 x 4096 f32, 65 taps, 5 iterations: both formulations, bit for bit."""
    w = workloads.vortex_noise(4096, iterations=5)
    default = rlic_b200.convolve(w.texture, w.u, w.v, **w.kwargs())
    with rlic_b200.options(walk="per-step"):
        before = _core.launch_count()
        other = rlic_b200.convolve(w.texture, w.u, w.v, **w.kwargs())
        assert _core.launch_count() - before >= 5
    assert_array_equal(default, other)


def test_wide_indices_and_batches(per_step):
    _core.lib.rlic_b200_debug_force_wide_index(1)
    try:
        check(*random_case((70, 45), np.float32, 19, seed=77), "polarization", "x-periodic", 2)
        check(*random_case((70, 45), np.float64, 19, seed=78), "velocity", "y-periodic", 2)
    finally:
        _core.lib.rlic_b200_debug_force_wide_index(0)
    rng = np.random.default_rng(14)
    tex = rng.random((6, 33, 65), dtype=np.float32)
    u = rng.random((6, 33, 65), dtype=np.float32) - 0.5
    v = rng.random((6, 33, 65), dtype=np.float32) - 0.5
    kernel = np.linspace(0.2, 1.0, 18, dtype=np.float32)
    got = rlic_b200.convolve_batch(tex, u, v, kernel=kernel, boundaries="periodic", iterations=2)
    for f in range(6):
        want = oracle.convolve(tex[f], u[f], v[f], kernel=kernel, boundaries=WALLS["periodic"], iterations=2)
        assert_array_equal(got[f], want)
