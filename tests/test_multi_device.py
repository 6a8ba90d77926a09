"""Single-process row-slab sharding (rlic_b200/multi.py) on the CPU: the slab geometry, the
halo copies and the strips / exchange / interior order run for real on CPU tensors, with the
library's own kernels compiled for the CPU (tests/kernel_emulation) as the compute steps.
The stitched result must equal the unsharded oracle bit for bit.  What this cannot show is
the ordering of asynchronous work across devices; that is the GPU twin's job."""
from __future__ import annotations

import numpy as np
import pytest
from numpy.testing import assert_array_equal

import kernel_emulation
import oracle
from rlic_b200.multi import MultiDeviceConvolver

CASES = {
    # name: (ndev, ny, nx, klen, boundaries, mode, iterations, dtype)
    "closed-2": (2, 40, 23, 9, "closed", "velocity", 3, np.float64),
    "periodic-ring-2": (2, 48, 17, 11, "periodic", "velocity", 3, np.float32),
    "y-periodic-pol-3": (3, 45, 20, 7, {"x": "closed", "y": "periodic"}, "polarization", 4, np.float64),
    "x-periodic-3-uneven": (3, 77, 19, 13, {"x": "periodic", "y": "closed"}, "velocity", 4, np.float32),
    "thin-slabs-no-overlap-4": (4, 31, 19, 13, "closed", "velocity", 3, np.float32),
    "single-device": (1, 20, 21, 9, "periodic", "polarization", 2, np.float64),
    "single-iteration-3": (3, 45, 20, 7, "periodic", "velocity", 1, np.float32),
}


def _inputs(ny, nx, klen, dtype):
    rng = np.random.default_rng(5)
    tex = rng.random((ny, nx)).astype(dtype)
    u = (rng.random((ny, nx)) - 0.5).astype(dtype)
    v = (rng.random((ny, nx)) - 0.5).astype(dtype)
    u[ny // 2, 3] = np.nan
    v[1, 1] = u[1, 1] = 0.0
    return tex, u, v, (rng.random(klen) + 0.1).astype(dtype)


def _pairs(boundaries):
    x = boundaries["x"] if isinstance(boundaries, dict) else boundaries
    y = boundaries["y"] if isinstance(boundaries, dict) else boundaries
    return ((x, x), (y, y))


@pytest.mark.parametrize("name", CASES)
def test_stitched_slabs_equal_the_whole_image(name):
    ndev, ny, nx, klen, boundaries, mode, iterations, dtype = CASES[name]
    tex, u, v, kernel = _inputs(ny, nx, klen, dtype)
    mc = MultiDeviceConvolver(ny, nx, kernel=kernel, uv_mode=mode, boundaries=boundaries,
                              devices=["cpu"] * ndev, ops=kernel_emulation.SlabOps())
    assert len(mc.plans) == ndev
    mc.set_field(u, v)
    before = tex.copy()
    for _ in range(2):                      # a convolver is reusable
        got = mc.convolve(tex, iterations=iterations)
    assert_array_equal(tex, before)         # inputs are never written
    want = oracle.convolve(tex, u, v, kernel=kernel, uv_mode=mode, boundaries=_pairs(boundaries),
                           iterations=iterations)
    assert got.dtype == tex.dtype
    assert_array_equal(got, want)
    assert_array_equal(mc.convolve(tex, iterations=0), tex)


def test_fewer_devices_are_used_when_the_image_is_short():
    kernel = np.ones(21)
    mc = MultiDeviceConvolver(35, 8, kernel=kernel, devices=["cpu"] * 8, ops=kernel_emulation.SlabOps())
    assert len(mc.plans) == 3                # 35 rows, reach 10: three slabs of >= 10 rows
    with pytest.raises(ValueError, match="shape"):
        mc.set_field(np.zeros((3, 3)), np.zeros((3, 3)))
    with pytest.raises(RuntimeError, match="set_field"):
        mc.convolve(np.zeros((35, 8)))


def test_public_entry_validates_like_convolve():
    import rlic_b200

    tex = np.random.default_rng(0).random((16, 16))
    with pytest.raises(ValueError):
        rlic_b200.convolve_sharded(-tex - 1, tex, tex, kernel=np.ones(5))
    with pytest.raises(TypeError):
        rlic_b200.convolve_sharded(tex, tex.astype(np.float32), tex, kernel=np.ones(5))
    out = rlic_b200.convolve_sharded(tex, tex, tex, kernel=np.ones(5), iterations=0)
    assert_array_equal(out, tex) and out is not tex


@pytest.mark.gpu
@pytest.mark.parametrize("ndev", [1, 2, 3])
def test_convolve_sharded_on_gpus_equals_convolve(ndev):
    """With fewer GPUs than slabs the same device is listed several times: the decomposition,
    the copies and their ordering are the same, only the link is missing."""
    import torch

    import rlic_b200

    have = torch.cuda.device_count()
    devices = [f"cuda:{i % have}" for i in range(ndev)]
    rng = np.random.default_rng(3)
    for dtype, mode, bnd in ((np.float32, "velocity", "closed"), (np.float64, "polarization", "periodic")):
        tex = rng.random((700, 300)).astype(dtype)
        u = (rng.random((700, 300)) - 0.5).astype(dtype)
        v = (rng.random((700, 300)) - 0.5).astype(dtype)
        kernel = np.linspace(0.1, 1, 33).astype(dtype)
        want = rlic_b200.convolve(tex, u, v, kernel=kernel, uv_mode=mode, boundaries=bnd, iterations=4)
        got = rlic_b200.convolve_sharded(tex, u, v, kernel=kernel, uv_mode=mode, boundaries=bnd, iterations=4,
                                         devices=devices)
        assert_array_equal(got, want)
