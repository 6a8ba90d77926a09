"""The two launch orders of the host path on the GPU: the wavefront schedule (default since
round 2) and the trailing one give the same bits, on images large enough to be cut into row
bands; the tests below pin the wavefront order explicitly for their own thread
(``rlic_b200.options``) whatever the process-wide default is.  The launch order itself is
checked on the CPU (tests/test_kernel_emulation.py: dependencies, and an in-order replay of
the schedule through the emulated kernels)."""
from __future__ import annotations

import numpy as np
import pytest
from numpy.testing import assert_array_equal

import oracle
import rlic_b200
from rlic_b200 import workloads

pytestmark = pytest.mark.gpu


@pytest.fixture
def wavefront():
    with rlic_b200.options(schedule="wavefront"):
        yield


@pytest.mark.parametrize("n,iterations", [(2048, 2), (3072, 5), (4096, 3)])
def test_wavefront_and_trailing_orders_give_the_same_bits(n, iterations):
    w = workloads.vortex_noise(n, iterations=iterations)
    with rlic_b200.options(schedule="trailing"):
        trailing = rlic_b200.convolve(w.texture, w.u, w.v, kernel=w.kernel, iterations=iterations)
    with rlic_b200.options(schedule="wavefront"):
        skewed = rlic_b200.convolve(w.texture, w.u, w.v, kernel=w.kernel, iterations=iterations)
    assert_array_equal(skewed, trailing)


def test_wavefront_against_the_oracle_f64_polarization(wavefront):
    rng = np.random.default_rng(8)
    ny, nx = 2304, 2048                       # > 4 Mpix: at least two bands
    tex = rng.random((ny, nx))
    u, v = rng.random((ny, nx)) - 0.5, rng.random((ny, nx)) - 0.5
    u[100, 100] = np.nan
    kernel = workloads.triangle_kernel(33, np.float64)
    bnd = (("periodic", "periodic"), ("closed", "closed"))
    got = rlic_b200.convolve(tex, u, v, kernel=kernel, uv_mode="polarization",
                             boundaries={"x": "periodic", "y": "closed"}, iterations=3)
    want = oracle.convolve(tex, u, v, kernel=kernel, uv_mode="polarization", boundaries=bnd, iterations=3,
                           threads=oracle.max_threads())
    assert_array_equal(got, want)


def test_periodic_rows_and_single_iterations_fall_back(wavefront):
    w = workloads.vortex_noise(2048, iterations=1)
    one = rlic_b200.convolve(w.texture, w.u, w.v, kernel=w.kernel, iterations=1)
    assert_array_equal(one[:64], oracle.pass_rows(w.texture, w.u, w.v, kernel=w.kernel, rows=(0, 64)))
    p = ("periodic", "periodic")
    got = rlic_b200.convolve(w.texture, w.u, w.v, kernel=w.kernel, boundaries="periodic", iterations=2)
    want = oracle.convolve(w.texture, w.u, w.v, kernel=w.kernel, boundaries=(p, p), iterations=2,
                           threads=oracle.max_threads())
    assert_array_equal(got, want)


def test_negative_texture_is_still_caught(wavefront):
    w = workloads.vortex_noise(2048, iterations=2)
    tex = w.texture.copy()
    tex[2000, 7] = -1.0
    with pytest.raises(ValueError):
        rlic_b200.convolve(tex, w.u, w.v, kernel=w.kernel, iterations=2)
