"""Properties of ``rlic_b200.convolve`` on the GPU.

Re-expresses what the reference asserts of ``rlic.convolve`` in its
tests/test_convolution.py (same fixture recipe: default_rng(0), 128x128 uniform
noise, kernel linspace(0, 1, 11)), and additionally checks each result bit for
bit against the CPU oracle.
"""

from itertools import combinations

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

import oracle
import rlic_b200 as rlic

pytestmark = pytest.mark.gpu

NX = 128


def make_args(nx=NX, klen=11, dtype="float64", seed=0):
    rng = np.random.default_rng(seed)
    dtype = np.dtype(dtype)
    return dict(
        img=rng.random((nx, nx), dtype=dtype),
        u=rng.random((nx, nx), dtype=dtype),
        v=rng.random((nx, nx), dtype=dtype),
        kernel=np.linspace(0, 1, klen, dtype=dtype),
    )


ARGS = make_args()


def test_no_iterations():
    out = rlic.convolve(ARGS["img"], ARGS["u"], ARGS["v"], kernel=ARGS["kernel"], iterations=0)
    assert_array_equal(out, ARGS["img"])
    assert out is not ARGS["img"]


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_default_is_one_iteration(dtype):
    a = make_args(dtype=dtype)
    implicit = rlic.convolve(a["img"], a["u"], a["v"], kernel=a["kernel"])
    explicit = rlic.convolve(a["img"], a["u"], a["v"], kernel=a["kernel"], iterations=1)
    assert_array_equal(implicit, explicit)
    assert implicit.dtype == np.dtype(dtype)
    assert_array_equal(implicit, oracle.convolve(a["img"], a["u"], a["v"], kernel=a["kernel"]))


def test_iterations_change_every_pixel():
    outs = [
        rlic.convolve(ARGS["img"], ARGS["u"], ARGS["v"], kernel=ARGS["kernel"], iterations=n)
        for n in range(3)
    ]
    for a, b in combinations(outs, 2):
        assert np.all(a != b)
    assert_array_equal(
        outs[2], oracle.convolve(ARGS["img"], ARGS["u"], ARGS["v"], kernel=ARGS["kernel"], iterations=2)
    )


def test_transpose_symmetry():
    a = rlic.convolve(ARGS["img"], ARGS["u"], ARGS["v"], kernel=ARGS["kernel"])
    # F-ordered views go in, as in the reference's test
    b = rlic.convolve(ARGS["img"].T, ARGS["v"].T, ARGS["u"].T, kernel=ARGS["kernel"]).T
    assert_array_equal(a, b)


def test_default_mode_is_velocity():
    a = rlic.convolve(ARGS["img"], ARGS["u"], ARGS["v"], kernel=ARGS["kernel"])
    b = rlic.convolve(ARGS["img"], ARGS["u"], ARGS["v"], kernel=ARGS["kernel"], uv_mode="velocity")
    assert_array_equal(a, b)


def test_modes_differ_on_a_sign_flip():
    kernel = np.ones(5)
    ones = np.ones((NX, NX))
    col = np.broadcast_to(np.arange(NX), (NX, NX))
    u1 = np.where(col < NX / 2, ones, -ones)
    u2 = -u1
    v = np.zeros((NX, NX))
    vel1 = rlic.convolve(ARGS["img"], u1, v, kernel=kernel, uv_mode="velocity")
    vel2 = rlic.convolve(ARGS["img"], u2, v, kernel=kernel, uv_mode="velocity")
    assert_allclose(vel2, vel1, atol=1e-14)
    pol1 = rlic.convolve(ARGS["img"], u1, v, kernel=kernel, uv_mode="polarization")
    pol2 = rlic.convolve(ARGS["img"], u2, v, kernel=kernel, uv_mode="polarization")
    assert_allclose(pol2, pol1, atol=1e-14)
    assert np.ptp(vel2 - pol2) > 1
    assert_array_equal(pol2, oracle.convolve(ARGS["img"], u2, v, kernel=kernel, uv_mode="polarization"))


@pytest.mark.parametrize("klen", [3, 4])
def test_modes_agree_for_short_kernels(klen):
    kernel = np.ones(klen)
    vel = rlic.convolve(ARGS["img"], ARGS["u"], ARGS["v"], kernel=kernel, uv_mode="velocity")
    pol = rlic.convolve(ARGS["img"], ARGS["u"], ARGS["v"], kernel=kernel, uv_mode="polarization")
    assert_array_equal(pol, vel)


def test_polarization_ignores_the_sign_of_the_field():
    n = 5
    kernel = np.ones(5)
    img = np.eye(n)
    zero, one = np.zeros((n, n)), np.ones((n, n))
    fwd = rlic.convolve(img, u=one, v=zero, kernel=kernel, uv_mode="polarization")
    bwd = rlic.convolve(img, u=-one, v=zero, kernel=kernel, uv_mode="polarization")
    assert_allclose(bwd, fwd)
    expected = np.array(
        [[3, 2, 1, 0, 0], [1, 1, 1, 1, 0], [1, 1, 1, 1, 1], [0, 1, 1, 1, 1], [0, 0, 1, 2, 3]], float
    )
    assert_array_equal(fwd, expected)
    down = rlic.convolve(img, u=zero, v=one, kernel=kernel, uv_mode="polarization")
    up = rlic.convolve(img, u=zero, v=-one, kernel=kernel, uv_mode="polarization")
    assert_allclose(up, down)
    assert_array_equal(down, expected.T)


@pytest.mark.parametrize("dtype", ["float32", "float64"])
@pytest.mark.parametrize("n", [0, 1, 5])
def test_nan_field_scales_the_texture(dtype, n):
    a = make_args(dtype=dtype)
    nan = np.full_like(a["img"], np.nan)
    out = rlic.convolve(a["img"], nan, nan, kernel=a["kernel"], iterations=n)
    scale = out / a["img"]
    assert np.ptp(scale) == 0.0
    assert scale[0, 0] == a["kernel"][len(a["kernel"]) // 2] ** n


def test_boundary_kinds_give_different_images():
    a = make_args(nx=64, klen=128)
    n = 64
    col = np.broadcast_to(np.arange(n), (n, n))
    u = np.where(col < n / 2, -1.0, 1.0)
    v = np.broadcast_to(np.sin(np.linspace(0, np.pi, n)), (n, n))   # stride-0 view
    closed = rlic.convolve(a["img"], u, v, kernel=a["kernel"], boundaries="closed")
    period = rlic.convolve(a["img"], u, v, kernel=a["kernel"], boundaries="periodic")
    xc_yp = rlic.convolve(a["img"], u, v, kernel=a["kernel"], boundaries={"x": "closed", "y": "periodic"})
    xp_yc = rlic.convolve(a["img"], u, v, kernel=a["kernel"], boundaries={"y": "closed", "x": "periodic"})
    assert np.all(closed != period)
    assert np.all(xc_yp != period)
    assert np.all(xp_yc != closed)
    assert np.all(xc_yp != xp_yc)
    c, p = ("closed", "closed"), ("periodic", "periodic")
    assert_array_equal(xc_yp, oracle.convolve(a["img"], u, v, kernel=a["kernel"], boundaries=(c, p)))
    assert_array_equal(xp_yc, oracle.convolve(a["img"], u, v, kernel=a["kernel"], boundaries=(p, c)))


def test_inputs_are_not_modified_and_output_is_fresh():
    a = make_args(nx=32)
    before = {k: v.copy() for k, v in a.items()}
    for arr in a.values():
        arr.setflags(write=False)   # read-only inputs must be accepted
    out = rlic.convolve(a["img"], a["u"], a["v"], kernel=a["kernel"], iterations=3)
    for k in a:
        assert_array_equal(a[k], before[k])
        assert not np.shares_memory(out, a[k])
    assert out.flags.c_contiguous and out.flags.writeable
