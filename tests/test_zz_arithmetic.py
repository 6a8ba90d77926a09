"""``rlic_b200.set_arithmetic("fma")`` on the GPU: the kernels then reproduce the `fma`-only
build of the reference (what its x86-64 wheels are compiled with; oracle variant 1) bit for
bit, and switching back restores the default build (variant 3).  The same kernel source is
held to the same oracle on the CPU by tests/test_kernel_emulation.py."""
from __future__ import annotations

import numpy as np
import pytest
from numpy.testing import assert_array_equal

import oracle
import rlic_b200
from golden_cases import CASES, as_spec, expected, load
from rlic_b200 import workloads
from test_kernel_emulation import WALLS, fuzz_case, random_case

pytestmark = pytest.mark.gpu


@pytest.fixture
def fma_only():
    with rlic_b200.options(arithmetic="fma"):
        yield


def check(tex, u, v, kernel, mode, walls, iterations, variant):
    bnd = WALLS[walls]
    with np.errstate(all="ignore"):
        got = rlic_b200.convolve(tex, u, v, kernel=kernel, uv_mode=mode, boundaries=as_spec(bnd),
                                 iterations=iterations)
        want = oracle.convolve(tex, u, v, kernel=kernel, uv_mode=mode, boundaries=bnd,
                               iterations=iterations, variant=variant)
    assert_array_equal(got, want)
    return got


@pytest.mark.parametrize("name", CASES)
def test_golden_vectors_of_the_fma_only_build(fma_only, name):
    mode, bnd, its = CASES[name]
    tex, u, v, kernel = load(name)
    got = rlic_b200.convolve(tex, u, v, kernel=kernel, uv_mode=mode, boundaries=as_spec(bnd), iterations=its)
    assert_array_equal(got, expected(name, 1))


@pytest.mark.parametrize("seed", range(16))
def test_randomised_configurations(fma_only, seed):
    tex, u, v, kernel, mode, walls, its = fuzz_case(seed)
    check(tex, u, v, kernel, mode, walls, its, variant=1)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_special_pixels_long_kernels_and_workloads(fma_only, dtype):
    for mode in ("velocity", "polarization"):
        for walls in WALLS:
            check(*random_case((45, 70), dtype, 23, seed=11), mode, walls, 2, variant=1)
    check(*random_case((9, 12), dtype, 1001, seed=5), "velocity", "periodic", 1, variant=1)
    w = workloads.vortex_noise(512, dtype=dtype, iterations=3)
    got = rlic_b200.convolve(w.texture, w.u, w.v, kernel=w.kernel, iterations=3)
    want = oracle.convolve(w.texture, w.u, w.v, kernel=w.kernel, iterations=3, variant=1,
                           threads=oracle.max_threads())
    assert_array_equal(got, want)


def test_switching_back_restores_the_default_build():
    w = workloads.vortex_noise(256, iterations=2)
    default = rlic_b200.convolve(w.texture, w.u, w.v, kernel=w.kernel, iterations=2)
    rlic_b200.set_arithmetic("fma")
    try:
        other = rlic_b200.convolve(w.texture, w.u, w.v, kernel=w.kernel, iterations=2)
    finally:
        rlic_b200.set_arithmetic("fma+branchless")
    again = rlic_b200.convolve(w.texture, w.u, w.v, kernel=w.kernel, iterations=2)
    assert_array_equal(default, again)
    assert_array_equal(default, oracle.convolve(w.texture, w.u, w.v, kernel=w.kernel, iterations=2, variant=3))
    assert_array_equal(other, oracle.convolve(w.texture, w.u, w.v, kernel=w.kernel, iterations=2, variant=1))
    assert (default != other).any()


def test_two_threads_with_different_arithmetic_do_not_race():
    """The choice is per call (per calling thread), not shared mutable state: two threads
    running concurrently, one per build of the reference, each get exactly their build's
    bits, call after call, and so does a batch call (its worker threads inherit the
    caller's choice)."""
    import threading

    w = workloads.vortex_noise(384, iterations=2)
    want = {name: oracle.convolve(w.texture, w.u, w.v, kernel=w.kernel, iterations=2, variant=variant)
            for name, variant in (("fma", 1), ("fma+branchless", 3))}
    assert (want["fma"] != want["fma+branchless"]).any()
    failures = []
    barrier = threading.Barrier(2)

    def worker(name):
        with rlic_b200.options(arithmetic=name):
            barrier.wait()
            for _ in range(12):
                got = rlic_b200.convolve(w.texture, w.u, w.v, kernel=w.kernel, iterations=2)
                if not np.array_equal(got, want[name]):
                    failures.append(name)
            stack = np.stack([w.texture] * 3)
            got = rlic_b200.convolve_batch(stack, np.stack([w.u] * 3), np.stack([w.v] * 3), kernel=w.kernel,
                                           iterations=2)
            if not all(np.array_equal(g, want[name]) for g in got):
                failures.append(name + " (batch)")

    threads = [threading.Thread(target=worker, args=(name,)) for name in want]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not failures
    assert rlic_b200.get_arithmetic() == "fma+branchless"
