"""Argument validation of ``rlic_b200.convolve`` (raises before any GPU work).

Pins the exact messages, their order and the grouping rule of the reference
(``src/rlic/_lib.py:135-205``; spec in ``tests/test_exceptions.py`` there).
"""

import re

import numpy as np
import pytest

import rlic_b200 as rlic

IMG = np.eye(64)
U = IMG.copy()
V = IMG.copy()
KERNEL = np.linspace(0, 1, 10, dtype="float64")

DTYPE_EXPECTATION = (
    r"Expected texture, u, v and kernel with identical dtype, from "
    r"\[dtype\('float32'\), dtype\('float64'\)\]\. "
)


def test_negative_iterations():
    with pytest.raises(
        ValueError,
        match=r"^Invalid number of iterations: -1\nExpected a strictly positive integer\.$",
    ):
        rlic.convolve(IMG, U, V, kernel=KERNEL, iterations=-1)


def test_unknown_uv_mode():
    with pytest.raises(
        ValueError,
        match=r"^Invalid uv_mode 'astral'\. Expected one of \['velocity', 'polarization'\]$",
    ):
        rlic.convolve(IMG, U, V, kernel=KERNEL, uv_mode="astral")


def test_texture_with_three_dimensions_reports_ndim_then_shape():
    cube = np.ones((16, 16, 16))
    with pytest.RaisesGroup(
        pytest.RaisesExc(
            ValueError,
            match=r"^Expected a texture with exactly two dimensions\. Got texture\.ndim=3$",
        ),
        pytest.RaisesExc(
            ValueError,
            match=r"^Shape mismatch: expected texture, u and v with identical shapes\.",
        ),
        match=r"^Invalid inputs were received\.",
    ):
        rlic.convolve(cube, U, V, kernel=KERNEL)


def test_negative_texture_values():
    with pytest.raises(
        ValueError,
        match=r"^Found invalid texture element\(s\)\. Expected only positive values\.$",
    ):
        rlic.convolve(-np.ones((64, 64)), V, V, kernel=KERNEL)


def test_zero_and_nan_texture_values_pass_validation():
    from rlic_b200._lib import _check_inputs

    tex = np.zeros((4, 4))
    tex[1, 1] = np.nan
    z = np.zeros((4, 4))
    problems, walls, deferred = _check_inputs(tex, z, z, np.ones(3), "velocity", "closed", 1)
    assert problems == [] and walls is not None and deferred is None


def test_negative_values_in_a_large_texture_join_the_group_in_order():
    # large textures defer their sign scan to the GPU, unless something else is
    # wrong: then it must still appear, in the reference's position (after ndim,
    # before the shape mismatch)
    tex = np.ones((300, 300))
    tex[7, 7] = -1.0
    with pytest.RaisesGroup(
        pytest.RaisesExc(ValueError, match=r"^Found invalid texture element"),
        pytest.RaisesExc(ValueError, match=r"^Shape mismatch"),
        match=r"^Invalid inputs were received\.",
    ):
        rlic.convolve(tex, U, V, kernel=KERNEL)


@pytest.mark.parametrize(
    "tshape, ushape, vshape",
    [
        ((64, 64), (65, 64), (64, 64)),
        ((64, 64), (64, 64), (63, 64)),
        ((64, 66), (64, 64), (64, 64)),
    ],
)
def test_shape_mismatch(tshape, ushape, vshape):
    rng = np.random.default_rng(0)
    tex, u, v = rng.random(tshape), rng.random(ushape), rng.random(vshape)
    expected = (
        "Shape mismatch: expected texture, u and v with identical shapes. "
        f"Got texture.shape={tshape}, u.shape={ushape}, v.shape={vshape}"
    )
    with pytest.raises(ValueError, match=f"^{re.escape(expected)}$"):
        rlic.convolve(tex, u, v, kernel=KERNEL)


def test_kernel_with_two_dimensions():
    with pytest.raises(
        ValueError,
        match=r"^Expected a kernel with exactly one dimension\. Got kernel\.ndim=2$",
    ):
        rlic.convolve(IMG, U, V, kernel=np.ones((5, 5)))


@pytest.mark.parametrize("bad", [-np.inf, np.inf, np.nan])
def test_non_finite_kernel(bad):
    kernel = np.ones(11)
    kernel[5] = bad
    with pytest.raises(ValueError, match=r"^Found non-finite value\(s\) in kernel\.$"):
        rlic.convolve(IMG, U, V, kernel=kernel)


def test_unsupported_texture_dtype():
    tex = np.ones((64, 64), dtype="complex128")
    with pytest.RaisesGroup(
        pytest.RaisesExc(
            TypeError,
            match=(
                r"^Found unsupported data type\(s\): \[dtype\('complex128'\)\]\. "
                + DTYPE_EXPECTATION
                + r"Got texture\.dtype=dtype\('complex128'\), u\.dtype=dtype\('float64'\), "
                r"v\.dtype=dtype\('float64'\), kernel\.dtype=dtype\('float64'\)$"
            ),
        ),
        pytest.RaisesExc(TypeError, match=r"^Data types mismatch"),
        match=r"^Invalid inputs were received\.",
    ):
        rlic.convolve(tex, U, V, kernel=KERNEL)


def test_unsupported_kernel_dtype():
    with pytest.RaisesGroup(
        pytest.RaisesExc(
            TypeError,
            match=(
                r"^Found unsupported data type\(s\): \[dtype\('complex128'\)\]\. "
                + DTYPE_EXPECTATION
                + r"Got texture\.dtype=dtype\('float64'\), u\.dtype=dtype\('float64'\), "
                r"v\.dtype=dtype\('float64'\), kernel\.dtype=dtype\('complex128'\)$"
            ),
        ),
        pytest.RaisesExc(TypeError, match=r"^Data types mismatch"),
        match=r"^Invalid inputs were received\.",
    ):
        rlic.convolve(IMG, U, V, kernel=-np.ones(5, dtype="complex128"))


def test_mixed_supported_dtypes():
    tex = np.ones((64, 64), dtype="float32")
    with pytest.raises(
        TypeError,
        match=(
            r"^Data types mismatch\. "
            + DTYPE_EXPECTATION
            + r"Got texture\.dtype=dtype\('float32'\), u\.dtype=dtype\('float64'\), "
            r"v\.dtype=dtype\('float64'\), kernel\.dtype=dtype\('float64'\)$"
        ),
    ):
        rlic.convolve(tex, U, V, kernel=KERNEL)


def test_zero_iterations_still_validates():
    with pytest.raises(ValueError, match=r"^Found non-finite value\(s\) in kernel\.$"):
        rlic.convolve(IMG, U, V, kernel=np.full(11, np.nan), iterations=0)


def test_boundaries_of_wrong_type():
    with pytest.raises(TypeError, match=r"^Invalid boundary specification None$"):
        rlic.convolve(IMG, U, V, kernel=KERNEL, boundaries=None)


def test_boundary_name_errors_join_the_group_last():
    with pytest.RaisesGroup(
        pytest.RaisesExc(ValueError, match=r"^Invalid uv_mode 'x'"),
        pytest.RaisesExc(ValueError, match=r"^Unknown left x boundary 'open'$"),
        pytest.RaisesExc(ValueError, match=r"^Unknown right x boundary 'open'$"),
        match=r"^Invalid inputs were received\.",
    ):
        rlic.convolve(IMG, U, V, kernel=KERNEL, uv_mode="x", boundaries={"x": "open", "y": "closed"})


def test_zero_iterations_returns_a_copy_without_touching_the_gpu():
    out = rlic.convolve(IMG, U, V, kernel=KERNEL, iterations=0)
    assert out is not IMG
    assert not np.shares_memory(out, IMG)
    np.testing.assert_array_equal(out, IMG)


def test_signature_matches_the_reference():
    import inspect

    sig = inspect.signature(rlic.convolve)
    kinds = {name: p.kind for name, p in sig.parameters.items()}
    P = inspect.Parameter
    assert list(kinds) == ["texture", "u", "v", "kernel", "uv_mode", "boundaries", "iterations"]
    assert kinds["texture"] is P.POSITIONAL_ONLY
    assert kinds["u"] is kinds["v"] is P.POSITIONAL_OR_KEYWORD
    assert all(kinds[k] is P.KEYWORD_ONLY for k in ("kernel", "uv_mode", "boundaries", "iterations"))
    assert sig.parameters["uv_mode"].default == "velocity"
    assert sig.parameters["boundaries"].default == "closed"
    assert sig.parameters["iterations"].default == 1


def test_batch_entry_validates_like_convolve():
    tex = np.ones((3, 8, 8))
    with pytest.raises(ValueError, match=r"^Expected textures with exactly three dimensions"):
        rlic.convolve_batch(tex[0], tex[0], tex[0], kernel=KERNEL)
    with pytest.raises(ValueError, match=r"^Shape mismatch: expected textures, u and v"):
        rlic.convolve_batch(tex, tex[:2], tex, kernel=KERNEL)
    with pytest.raises(ValueError, match=r"^Invalid uv_mode 'x'"):
        rlic.convolve_batch(tex, tex, tex, kernel=KERNEL, uv_mode="x")
    with pytest.raises(ValueError, match=r"^Found invalid texture element"):
        rlic.convolve_batch(-tex, tex, tex, kernel=KERNEL)
    with pytest.raises(TypeError, match=r"^Data types mismatch"):
        rlic.convolve_batch(tex.astype("float32"), tex, tex, kernel=KERNEL)
    out = rlic.convolve_batch(tex, tex, tex, kernel=KERNEL, iterations=0)
    assert out is not tex and np.array_equal(out, tex)
    assert rlic.convolve_batch(tex[:0], tex[:0], tex[:0], kernel=KERNEL).shape == (0, 8, 8)
