"""Pins the CPU oracle (oracle/) — the checker every GPU parity test relies on.

Sources of truth, in order of independence:
  1. the reference's six Rust unit known-answer tests (src/lib.rs:139-151,
     182-206, 275-297);
  2. answers derived by hand from the algorithm;
  3. the exact-equality properties the reference's tests/test_convolution.py
     asserts of rlic.convolve;
  4. agreement, bit for bit, with a second restatement (oracle/pyoracle.py) and
     with the frozen vectors in tests/golden/.
No GPU is involved anywhere in this file.
"""

import numpy as np
import pytest

import oracle
from oracle import pyoracle

from golden_cases import CASES, expected, load

VARIANTS = [0, 1, 2, 3]


# ---- 1. Rust unit KATs ------------------------------------------------------
@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_rust_kat_time_to_next_pixel(dtype, variant):
    assert oracle.edge_time(1.0, 0.0, dtype, variant) == 1.0     # lib.rs:186-190
    assert oracle.edge_time(-1.0, 1.0, dtype, variant) == 1.0    # lib.rs:191-195
    assert oracle.edge_time(0.0, 0.5, dtype, variant) == np.inf  # lib.rs:196-205


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_rust_kat_advance_with_zero_velocity(dtype, variant):
    # lib.rs:280-296
    state = oracle.cross(0.0, 0.0, 5, 5, 0.5, 0.5, shape=(10, 10), dtype=dtype, variant=variant)
    assert state == (5, 5, 0.5, 0.5)


def test_rust_kat_pixel_select_is_row_major():
    # lib.rs:144-150: arr[[i, j]] with i the row.  A walker that cannot move
    # (NaN field) returns kernel[mid] * texture[i, j] at [i, j].
    tex = np.array([[1.0, 2.0], [3.0, 4.0]])
    nan = np.full_like(tex, np.nan)
    out = oracle.convolve(tex, nan, nan, kernel=np.array([0.0, 1.0, 0.0]))
    np.testing.assert_array_equal(out, tex)


# ---- 2. hand-derived answers ------------------------------------------------
EYE5_U_PLUS = np.array(
    [[3, 2, 1, 0, 0], [1, 1, 1, 1, 0], [1, 1, 1, 1, 1], [0, 1, 1, 1, 1], [0, 0, 1, 2, 3]],
    dtype=float,
)


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_uniform_x_flow_on_identity(dtype, variant):
    img = np.eye(5, dtype=dtype)
    one, zero = np.ones_like(img), np.zeros_like(img)
    k = np.ones(5, dtype=dtype)
    closed = oracle.convolve(img, one, zero, kernel=k, variant=variant)
    np.testing.assert_array_equal(closed, EYE5_U_PLUS.astype(dtype))
    periodic = oracle.convolve(
        img, one, zero, kernel=k, variant=variant,
        boundaries=(("periodic", "periodic"), ("closed", "closed")),
    )
    np.testing.assert_array_equal(periodic, np.ones_like(img))
    # same result for u = -1 in polarization mode, and for the y axis by symmetry
    pol = oracle.convolve(img, -one, zero, kernel=k, uv_mode="polarization", variant=variant)
    np.testing.assert_array_equal(pol, closed)
    along_y = oracle.convolve(img, zero, one, kernel=k, variant=variant)
    np.testing.assert_array_equal(along_y, EYE5_U_PLUS.T.astype(dtype))


@pytest.mark.parametrize("variant", VARIANTS)
def test_diagonal_flow_takes_y_first_on_ties(variant):
    # u = v = 1: samples are (i,j), fwd (i+1,j),(i+1,j+1), bwd (i-1,j),(i-1,j-1)
    n = 7
    tex = np.arange(n * n, dtype=np.float64).reshape(n, n)
    one = np.ones_like(tex)
    out = oracle.convolve(tex, one, one, kernel=np.ones(5), variant=variant)
    i = j = 3
    want = tex[i, j] + tex[i + 1, j] + tex[i + 1, j + 1] + tex[i - 1, j] + tex[i - 1, j - 1]
    assert out[i, j] == want


@pytest.mark.parametrize("variant", VARIANTS)
def test_tap_order_forward_is_upper_half(variant):
    # kernel = [a, b, c, d, e], u = +1: out[j] = c*t[j] + d*t[j+1] + e*t[j+2] + b*t[j-1] + a*t[j-2]
    tex = np.array([[1.0, 10.0, 100.0, 1000.0, 10000.0, 100000.0, 1000000.0]])
    k = np.array([2.0, 3.0, 5.0, 7.0, 11.0])
    out = oracle.convolve(tex, np.ones_like(tex), np.zeros_like(tex), kernel=k, variant=variant)
    j = 3
    want = 5 * tex[0, j] + 7 * tex[0, j + 1] + 11 * tex[0, j + 2] + 3 * tex[0, j - 1] + 2 * tex[0, j - 2]
    assert out[0, j] == want


@pytest.mark.parametrize("variant", VARIANTS)
def test_even_kernel_has_short_forward_half(variant):
    # L = 4: kmid = 2, forward taps {3}, backward taps {1, 0}
    tex = np.array([[1.0, 10.0, 100.0, 1000.0, 10000.0, 100000.0]])
    k = np.array([2.0, 3.0, 5.0, 7.0])
    out = oracle.convolve(tex, np.ones_like(tex), np.zeros_like(tex), kernel=k, variant=variant)
    j = 3
    assert out[0, j] == 5 * tex[0, j] + 7 * tex[0, j + 1] + 3 * tex[0, j - 1] + 2 * tex[0, j - 2]


@pytest.mark.parametrize("variant", VARIANTS)
def test_closed_wall_resamples_the_edge_pixel(variant):
    tex = np.array([[1.0, 10.0, 100.0]])
    out = oracle.convolve(tex, np.ones_like(tex), np.zeros_like(tex), kernel=np.ones(7), variant=variant)
    # from j=1: centre 10; fwd 100,100,100 (stuck at the wall); bwd 1,1,1
    assert out[0, 1] == 10 + 300 + 3


@pytest.mark.parametrize("variant", VARIANTS)
def test_stagnation_pixel_is_resampled_for_every_tap(variant):
    tex = np.array([[1.0, 10.0, 100.0, 1000.0]])
    u = np.array([[1.0, 1.0, 0.0, 1.0]])
    out = oracle.convolve(tex, u, np.zeros_like(u), kernel=np.ones(7), variant=variant)
    # from j=0 forward: 10, then 100 (u=0 there) twice more; backward: wall at j=0 thrice
    assert out[0, 0] == 1 + (10 + 100 + 100) + 3 * 1


@pytest.mark.parametrize("variant", VARIANTS)
def test_nan_velocity_stops_the_walk_there(variant):
    tex = np.array([[1.0, 10.0, 100.0, 1000.0, 10000.0]])
    u = np.array([[1.0, 1.0, np.nan, 1.0, 1.0]])
    out = oracle.convolve(tex, u, np.zeros_like(u), kernel=np.ones(9), variant=variant)
    # from j=0: fwd 10, 100 then stop (u NaN at j=2 is read before the third step); bwd 1 x4
    assert out[0, 0] == 1 + 10 + 100 + 4


# ---- 3. exact properties asserted by the reference's Python tests ----------
@pytest.fixture(scope="module")
def rnd():
    rng = np.random.default_rng(0)
    n = 48
    return dict(
        img=rng.random((n, n)), u=rng.random((n, n)), v=rng.random((n, n)),
        kernel=np.linspace(0, 1, 11),
    )


@pytest.mark.parametrize("variant", VARIANTS)
def test_transpose_symmetry_is_exact(rnd, variant):
    # tests/test_convolution.py:69-82
    a = oracle.convolve(rnd["img"], rnd["u"], rnd["v"], kernel=rnd["kernel"], variant=variant)
    b = oracle.convolve(rnd["img"].T, rnd["v"].T, rnd["u"].T, kernel=rnd["kernel"], variant=variant).T
    np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n", [1, 5])
def test_nan_field_scales_by_centre_tap(rnd, dtype, n):
    # tests/test_convolution.py:165-177
    img = rnd["img"].astype(dtype)
    nan = np.full_like(img, np.nan)
    kernel = rnd["kernel"].astype(dtype)
    out = oracle.convolve(img, nan, nan, kernel=kernel, iterations=n)
    scale = out / img
    assert np.ptp(scale) == 0.0
    assert scale[0, 0] == kernel[len(kernel) // 2] ** n


@pytest.mark.parametrize("klen", [3, 4])
def test_polarization_equals_velocity_for_short_kernels(rnd, klen):
    # tests/test_convolution.py:113-122
    k = np.ones(klen)
    vel = oracle.convolve(rnd["img"], rnd["u"], rnd["v"], kernel=k)
    pol = oracle.convolve(rnd["img"], rnd["u"], rnd["v"], kernel=k, uv_mode="polarization")
    np.testing.assert_array_equal(vel, pol)


def test_velocity_and_polarization_differ_on_a_sign_flip():
    # tests/test_convolution.py:93-110
    n = 64
    rng = np.random.default_rng(0)
    img = rng.random((n, n))
    col = np.broadcast_to(np.arange(n), (n, n))
    u1 = np.where(col < n / 2, 1.0, -1.0)
    u2 = -u1
    v = np.zeros((n, n))
    k = np.ones(5)
    vel1, vel2 = (oracle.convolve(img, u, v, kernel=k) for u in (u1, u2))
    pol1, pol2 = (oracle.convolve(img, u, v, kernel=k, uv_mode="polarization") for u in (u1, u2))
    np.testing.assert_allclose(vel2, vel1, atol=1e-14)
    np.testing.assert_allclose(pol2, pol1, atol=1e-14)
    assert np.ptp(vel2 - pol2) > 1


def test_iterations_compose(rnd):
    k = rnd["kernel"]
    once = oracle.convolve(rnd["img"], rnd["u"], rnd["v"], kernel=k)
    twice = oracle.convolve(once, rnd["u"], rnd["v"], kernel=k)
    both = oracle.convolve(rnd["img"], rnd["u"], rnd["v"], kernel=k, iterations=2)
    np.testing.assert_array_equal(twice, both)
    assert np.all(once != both) and np.all(once != rnd["img"])


def test_threads_do_not_change_results(rnd):
    a = oracle.convolve(rnd["img"], rnd["u"], rnd["v"], kernel=rnd["kernel"], iterations=3)
    b = oracle.convolve(rnd["img"], rnd["u"], rnd["v"], kernel=rnd["kernel"], iterations=3, threads=4)
    np.testing.assert_array_equal(a, b)


def test_band_pass_matches_full_pass(rnd):
    full = oracle.convolve(rnd["img"], rnd["u"] - 0.5, rnd["v"] - 0.5, kernel=rnd["kernel"])
    band = oracle.pass_rows(rnd["img"], rnd["u"] - 0.5, rnd["v"] - 0.5, kernel=rnd["kernel"], rows=(7, 19))
    np.testing.assert_array_equal(full[7:19], band)


def test_boundary_kinds_give_different_images():
    # tests/test_convolution.py:180-207 (128 taps on 64x64)
    rng = np.random.default_rng(0)
    n = 64
    img = rng.random((n, n))
    kernel = np.linspace(0, 1, 128)
    col = np.broadcast_to(np.arange(n), (n, n))
    u = np.where(col < n / 2, -1.0, 1.0)
    v = np.broadcast_to(np.sin(np.linspace(0, np.pi, n)), (n, n))
    c, p = ("closed", "closed"), ("periodic", "periodic")
    res = {
        name: oracle.convolve(img, u, v, kernel=kernel, boundaries=b)
        for name, b in {"cc": (c, c), "pp": (p, p), "cp": (c, p), "pc": (p, c)}.items()
    }
    assert np.all(res["cc"] != res["pp"])
    assert np.all(res["cp"] != res["pp"])
    assert np.all(res["pc"] != res["cc"])
    assert np.all(res["cp"] != res["pc"])


# ---- 4. two restatements and the frozen vectors agree ----------------------
@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("name", CASES)
def test_c_oracle_reproduces_golden_vectors(name, variant):
    mode, bnd, its = CASES[name]
    tex, u, v, kernel = load(name)
    out = oracle.convolve(tex, u, v, kernel=kernel, uv_mode=mode, boundaries=bnd,
                          iterations=its, variant=variant)
    np.testing.assert_array_equal(out, expected(name, variant))


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("mode", ["velocity", "polarization"])
def test_c_and_python_restatements_agree(dtype, mode, variant):
    rng = np.random.default_rng(100 + variant)
    shape = (7, 9)
    tex = rng.random(shape).astype(dtype)
    u = (rng.random(shape) - 0.5).astype(dtype)
    v = (rng.random(shape) - 0.5).astype(dtype)
    u[0, 0] = v[0, 0] = 0
    v[3, 3] = np.nan
    u[5, 5] = -0.0
    kernel = rng.random(10).astype(dtype)
    bnd = (("periodic", "periodic"), ("closed", "closed"))
    a = oracle.convolve(tex, u, v, kernel=kernel, uv_mode=mode, boundaries=bnd, iterations=2, variant=variant)
    b = pyoracle.convolve(tex, u, v, kernel=kernel, uv_mode=mode, boundaries=bnd, iterations=2,
                          fma=bool(variant & 1), branchless=bool(variant & 2))
    np.testing.assert_array_equal(a, b)


def test_variants_really_differ():
    # the four Cargo feature sets are not bit-identical in general; if they were
    # the variant switch would be dead code
    rng = np.random.default_rng(7)
    shape = (32, 32)
    tex, u, v = rng.random(shape), rng.random(shape) - 0.5, rng.random(shape) - 0.5
    k = np.linspace(0, 1, 31)
    outs = [oracle.convolve(tex, u, v, kernel=k, variant=var) for var in VARIANTS]
    assert not np.array_equal(outs[3], outs[0])
    assert not np.array_equal(outs[3], outs[2])
    for o in outs[:3]:
        np.testing.assert_allclose(o, outs[3], rtol=0, atol=1e-12 * np.ptp(outs[3]))


def test_empty_kernel_is_an_error_not_an_abort():
    img = np.eye(4)
    with pytest.raises(RuntimeError):
        oracle.convolve(img, img, img, kernel=np.ones(0))
