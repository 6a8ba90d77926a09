"""Histogram equalisation.  The reference only declares the operation
(/root/reference/src/rlic/_core.pyi:30-37: no implementation offline), so the oracle --
oracle/equalize.py -- DEFINES the semantics (parity with upstream code: unpinned).  CPU tests pin
the oracle to the properties an equalisation must have; GPU tests hold the CUDA path to the
oracle bit for bit."""
from __future__ import annotations

import numpy as np
import pytest
from numpy.testing import assert_array_equal

from oracle.equalize import equalize_histogram as oracle_equalize


def test_oracle_is_a_monotone_map_onto_the_unit_interval():
    rng = np.random.default_rng(0)
    for dtype in (np.float32, np.float64):
        img = (rng.standard_normal((64, 96)) ** 3).astype(dtype)
        for nbins in (1, 2, 7, 256, 10000):
            out = oracle_equalize(img, nbins)
            assert out.dtype == img.dtype and out.shape == img.shape
            assert out.max() == 1.0 and out.min() > 0.0
            order = np.argsort(img, axis=None, kind="stable")
            assert np.all(np.diff(out.ravel()[order]) >= 0)          # never reverses an ordering
            assert len(np.unique(out)) <= nbins
        # many bins on well-spread values: the rank transform
        flat = rng.random((64, 96)).astype(dtype)
        out = oracle_equalize(flat, 1 << 22).astype(np.float64)
        ranks = (np.argsort(np.argsort(flat, axis=None)) + 1).reshape(flat.shape) / flat.size
        assert np.abs(out - ranks).max() < 2e-3


def test_oracle_equalises_and_keeps_nan_and_handles_flat_images():
    rng = np.random.default_rng(1)
    img = rng.random((128, 128)).astype(np.float32) ** 3                # skewed: half the pixels below 1/8
    assert np.histogram(img, bins=8, range=(0, 1))[0][0] > 0.45 * img.size
    out = oracle_equalize(img, 4096)
    hist = np.histogram(out, bins=8, range=(0, 1 + 1e-9))[0] / out.size
    assert np.abs(hist - 1 / 8).max() < 0.03                            # flat histogram afterwards
    img[3, 4] = img[100, 7] = np.nan
    out = oracle_equalize(img, 64)
    assert np.isnan(out[3, 4]) and np.isnan(out[100, 7]) and np.isnan(out).sum() == 2
    assert np.nanmax(out) == 1.0
    assert_array_equal(oracle_equalize(np.full((5, 6), 2.5), 16), np.ones((5, 6)))
    assert np.isnan(oracle_equalize(np.full((2, 2), np.nan, dtype=np.float32), 4)).all()


def test_public_entry_validates_without_a_gpu():
    import rlic_b200

    img = np.zeros((4, 4), dtype=np.float32)
    with pytest.raises(TypeError):
        rlic_b200.equalize_histogram(img.astype(np.int32))
    with pytest.raises(ValueError, match="two dimensions"):
        rlic_b200.equalize_histogram(img[0])
    for bad in (0, -3, 1.5, True, (1 << 24) + 1):
        with pytest.raises(ValueError, match="Invalid number of bins"):
            rlic_b200.equalize_histogram(img, nbins=bad)
    assert rlic_b200.equalize_histogram(img[:0]).shape == (0, 4)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("nbins", [1, 2, 7, 256, 4096, 5000, 1 << 16])
def test_cuda_path_equals_the_oracle(dtype, nbins):
    import rlic_b200

    rng = np.random.default_rng(nbins)
    for shape in ((1, 1), (3, 5), (257, 130), (1024, 1024)):
        img = (rng.standard_normal(shape) ** 3).astype(dtype)
        assert_array_equal(rlic_b200.equalize_histogram(img, nbins=nbins), oracle_equalize(img, nbins))
    img = rng.random((300, 200)).astype(dtype)
    img[rng.random(img.shape) < 0.01] = np.nan
    img[0, 0] = -0.0
    got = rlic_b200.equalize_histogram(img, nbins=nbins)
    assert_array_equal(got, oracle_equalize(img, nbins))
    assert_array_equal(rlic_b200.equalize_histogram(np.full((9, 9), 3.0, dtype=dtype), nbins=nbins),
                       np.ones((9, 9), dtype=dtype))
    assert np.isnan(rlic_b200.equalize_histogram(np.full((4, 4), np.nan, dtype=dtype), nbins=nbins)).all()


@pytest.mark.gpu
def test_equalising_a_convolution_result_on_the_device():
    """The intended use: convolve on the device, equalise the result there, read it back once."""
    import torch

    import oracle
    from rlic_b200 import workloads
    from rlic_b200.device import convolve_device, equalize_histogram_device

    w = workloads.vortex_noise(1024, iterations=2)
    d = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (w.texture, w.u, w.v)]
    lic = convolve_device(*d, kernel=w.kernel, iterations=2)
    eq = equalize_histogram_device(lic, nbins=256)
    want = oracle_equalize(oracle.convolve(w.texture, w.u, w.v, kernel=w.kernel, iterations=2,
                                           threads=oracle.max_threads()), 256)
    assert_array_equal(eq.cpu().numpy(), want)
    into = torch.empty_like(lic)
    assert equalize_histogram_device(lic, nbins=256, out=into) is into
    assert torch.equal(into, eq)
    # full headline size, f32: counts reach 2^24
    big = torch.rand((4096, 4096), device="cuda") ** 4
    assert_array_equal(equalize_histogram_device(big, nbins=1000).cpu().numpy(),
                       oracle_equalize(big.cpu().numpy(), 1000))
