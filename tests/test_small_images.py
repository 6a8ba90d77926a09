"""The small-image pass kernel (lic_pass_pair_kernel: the two directions of a pixel on two warps,
backward samples parked in shared memory, the last pass of a device call writing the dense
result itself) against the CPU oracle and against the one-thread-per-pixel kernel.  The general
edge cases reach it through tests/test_parity.py, which runs every test with and without it;
here are the cases that are about this kernel's own moving parts."""
from __future__ import annotations

import numpy as np
import pytest
import torch
from numpy.testing import assert_array_equal

import oracle
import rlic_b200 as rlic
from rlic_b200 import _core
from rlic_b200.device import convolve_device, convolve_device_batch, pack_field

pytestmark = pytest.mark.gpu

WALLS = {
    "closed": (("closed", "closed"), ("closed", "closed")),
    "periodic": (("periodic", "periodic"), ("periodic", "periodic")),
    "x-periodic": (("periodic", "periodic"), ("closed", "closed")),
}
SPEC = {"closed": "closed", "periodic": "periodic", "x-periodic": {"x": "periodic", "y": "closed"}}


def case(shape, dtype, klen, seed):
    rng = np.random.default_rng(seed)
    tex = rng.random(shape).astype(dtype)
    u = (rng.random(shape) - 0.5).astype(dtype)
    v = (rng.random(shape) - 0.5).astype(dtype)
    u[shape[0] // 2, 3] = np.nan            # stops forward AND backward walks that reach it
    v[1, 1] = u[1, 1] = 0.0                 # stagnation
    u[2, 5], v[2, 5] = -0.0, 0.0
    return tex, u, v, (rng.random(klen) - 0.2).astype(dtype)


def launches(fn):
    before = _core.launch_count()
    out = fn()
    torch.cuda.synchronize()
    return out, _core.launch_count() - before


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("mode", ["velocity", "polarization"])
@pytest.mark.parametrize("klen", [2, 3, 4, 9, 64, 65, 129])
def test_against_the_oracle_and_the_other_kernel(dtype, mode, klen):
    for walls in WALLS:
        tex, u, v, k = case((70, 90), dtype, klen, seed=klen)
        for its in (1, 3):
            with np.errstate(all="ignore"):
                want = oracle.convolve(tex, u, v, kernel=k, uv_mode=mode, boundaries=WALLS[walls], iterations=its)
                got = rlic.convolve(tex, u, v, kernel=k, uv_mode=mode, boundaries=SPEC[walls], iterations=its)
                _core.lib.rlic_b200_debug_small_image_kernel(0)
                try:
                    other = rlic.convolve(tex, u, v, kernel=k, uv_mode=mode, boundaries=SPEC[walls], iterations=its)
                finally:
                    _core.lib.rlic_b200_debug_small_image_kernel(1)
            assert_array_equal(got, want)
            assert_array_equal(other, want)


def test_device_entry_writes_the_dense_result_from_the_last_pass():
    """convolve_device on a small image: pad + passes, and NO un-padding launch (the last pass
    stores the dense result); with the hook off there is one launch more.  Same bits."""
    tex, u, v, k = case((256, 256), np.float64, 65, seed=3)
    d = [torch.from_numpy(a).cuda() for a in (tex, u, v)]
    field = pack_field(d[1], d[2], boundaries="periodic")
    for its in (1, 2, 5):
        want = oracle.convolve(tex, u, v, kernel=k, boundaries=WALLS["periodic"], iterations=its,
                               threads=oracle.max_threads())
        out, n = launches(lambda: convolve_device(d[0], field=field, kernel=k, boundaries="periodic", iterations=its))
        assert_array_equal(out.cpu().numpy(), want)
        # pad, passes; a call that replays (iterations >= 2) records with the one-thread-per-pixel
        # kernel, replays, and un-pads
        assert n == (1 + its if its == 1 else 2 + its)
        with rlic.options(paths="recompute"):
            out1, n1 = launches(lambda: convolve_device(d[0], field=field, kernel=k, boundaries="periodic",
                                                         iterations=its))
        assert_array_equal(out1.cpu().numpy(), want)
        assert n1 == 1 + its                                 # pad, passes: the last one writes the dense result
        _core.lib.rlic_b200_debug_small_image_kernel(0)
        try:
            with rlic.options(paths="recompute"):
                out2, n2 = launches(lambda: convolve_device(d[0], field=field, kernel=k, boundaries="periodic",
                                                            iterations=its))
        finally:
            _core.lib.rlic_b200_debug_small_image_kernel(1)
        assert_array_equal(out2.cpu().numpy(), want)
        assert n2 == 2 + its                                 # pad, passes, un-pad


def test_small_batches_and_images_just_around_the_size_limit():
    rng = np.random.default_rng(8)
    k = np.linspace(0.2, 1.0, 33, dtype=np.float32)
    # a stack of small fields (every field keeps its own walls)
    stack = [rng.random((5, 60, 70), dtype=np.float32) for _ in range(3)]
    stack[1] -= 0.5
    stack[2] -= 0.5
    got = convolve_device_batch(*(torch.from_numpy(a).cuda() for a in stack), kernel=k, iterations=2)
    for f in range(5):
        want = oracle.convolve(stack[0][f], stack[1][f], stack[2][f], kernel=k, iterations=2)
        assert_array_equal(got[f].cpu().numpy(), want)
    # one image below and one above the pixel limit of the small-image kernel (75 776 pixels)
    for shape in ((275, 275), (276, 276)):
        tex, u, v, _ = case(shape, np.float32, 33, seed=shape[0])
        want = oracle.convolve(tex, u, v, kernel=k, iterations=2, threads=oracle.max_threads())
        with np.errstate(all="ignore"):
            assert_array_equal(rlic.convolve(tex, u, v, kernel=k, iterations=2), want)


def test_kernels_too_long_for_the_parked_samples_fall_back():
    """200 taps of f64 park 100 x 128 x 8 B = 100 KB per CTA: beyond the kernel's shared-memory
    budget, so the call runs on the one-thread-per-pixel kernel (and 1001 taps live in global
    memory, which the small-image kernel does not read)."""
    for dtype, klen in ((np.float64, 200), (np.float32, 500), (np.float32, 1001)):
        tex, u, v, k = case((40, 50), dtype, klen, seed=klen)
        with np.errstate(all="ignore"):
            want = oracle.convolve(tex, u, v, kernel=k, boundaries=WALLS["periodic"])
            assert_array_equal(rlic.convolve(tex, u, v, kernel=k, boundaries="periodic"), want)
