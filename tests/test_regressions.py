"""The reference's regression cases (tests/test_regressions.py of rLIC), re-expressed.

The reference holds ``rlic.convolve`` to vectorplot's LIC on eight small inputs with sharp
sign flips, at rtol 1.5e-7 / atol 1e-6.  vectorplot (0.2.0.post5) is not installed here and
cannot be fetched, so the same inputs are held to an exact-arithmetic tracer
(``regression_cases.exact_streamline_sum``) at the reference's tolerances: on the CPU the
oracle and its slow Python twin and the kernel source run through the emulation, on the GPU
(``-m gpu``) the CUDA path -- which must also equal the oracle bit for bit."""
from __future__ import annotations

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

import kernel_emulation as ke
import oracle
from oracle import pyoracle
from regression_cases import ATOL, CASES, K0, RTOL, TEXTURE, exact_streamline_sum


@pytest.fixture(scope="module")
def expected():
    return {name: exact_streamline_sum(TEXTURE, u, v, K0, mode) for name, (u, v, mode) in CASES.items()}


@pytest.mark.parametrize("name", CASES)
def test_oracle_and_kernel_source_against_exact_arithmetic(name, expected):
    u, v, mode = CASES[name]
    got = oracle.convolve(TEXTURE, u, v, kernel=K0, uv_mode=mode)
    assert got.dtype == np.float32
    assert_allclose(got, expected[name], rtol=RTOL, atol=ATOL)
    assert_array_equal(pyoracle.convolve(TEXTURE, u, v, kernel=K0, uv_mode=mode), got)
    for walk in (0, 1):
        assert_array_equal(ke.convolve(TEXTURE, u, v, kernel=K0, uv_mode=mode, walk=walk), got)
    # the reference's other build takes the same paths on these fields
    assert_allclose(oracle.convolve(TEXTURE, u, v, kernel=K0, uv_mode=mode, variant=oracle.VARIANT_FMA),
                    expected[name], rtol=RTOL, atol=ATOL)


def test_the_cases_are_not_degenerate(expected):
    """Distinct fields give distinct images, and the sharp flips matter: polarization and
    velocity differ exactly where a flip is crossed."""
    zero = expected["0-0-velocity"]
    assert_allclose(zero, TEXTURE.astype(np.float64) * K0.astype(np.float64).sum(), rtol=2e-6)
    assert_array_equal(expected["0-0-polarization"], zero)
    for a, b in (("U1-0-velocity", "0-V1-velocity"), ("U1-0-velocity", "U1-V1-velocity"),
                 ("U1-0-velocity", "U1-0-polarization")):
        assert np.abs(expected[a] - expected[b]).max() > 1e-2, (a, b)
    # where both components flip the flow leaves the centre along the diagonals and no walker
    # ever crosses a flip: the two modes agree
    assert_array_equal(expected["U1-V1-velocity"], expected["U1-V1-polarization"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_outputs(name, expected):
    import rlic_b200 as rlic

    u, v, mode = CASES[name]
    out = rlic.convolve(TEXTURE, u, v, kernel=K0, uv_mode=mode)
    assert_allclose(out, expected[name], rtol=RTOL, atol=ATOL)
    assert_array_equal(out, oracle.convolve(TEXTURE, u, v, kernel=K0, uv_mode=mode))
