"""The slab pass of the C ABI (rlic_b200_pass_slab_*) on ONE GPU: an image is cut
into row slabs by hand, halos are filled by plain copies (ring order for
y-periodic images), every slab is computed separately, and the stitched result
must equal the whole-image result and the oracle bit for bit.  This is the
arithmetic the multi-GPU driver (rlic_b200.sharded) relies on; its exchange
logic is covered on CPU by tests/test_sharded_gloo.py and end to end by
test_two_gpu_sharded_run below when two devices are present.
"""

import ctypes
import os
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import oracle
from rlic_b200 import _core

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def slab_pass_all(tex, u, v, kernel, bounds_spec, walls, mode, cuts):
    import torch

    from rlic_b200.device import pack_field

    dev = torch.device("cuda", 0)
    ny, nx = tex.shape
    h = kernel.size // 2
    periodic_y = bounds_spec[1][0] == "periodic"
    t_tex = torch.from_numpy(tex).to(dev)
    field = pack_field(torch.from_numpy(np.ascontiguousarray(u)).to(dev),
                       torch.from_numpy(np.ascontiguousarray(v)).to(dev)).uv
    sfx, real = ("f32", ctypes.c_float) if tex.dtype == np.float32 else ("f64", ctypes.c_double)
    out = torch.empty_like(t_tex)
    edges = [0, *cuts, ny]
    for r0, r1 in zip(edges[:-1], edges[1:]):
        lo = h if (r0 > 0 or periodic_y) else 0
        hi = h if (r1 < ny or periodic_y) else 0
        rows = (torch.arange(r0 - lo, r1 + hi, device=dev)) % ny   # ring order when wrapping
        s_tex = t_tex[rows].contiguous()
        s_field = field[rows].contiguous()
        s_out = torch.empty((r1 - r0, nx), dtype=t_tex.dtype, device=dev)
        rc = getattr(_core.lib, f"rlic_b200_pass_slab_{sfx}")(
            s_tex.data_ptr(), s_field.data_ptr(), s_out.data_ptr(), ny, nx, r0, r1 - r0, lo, hi,
            kernel.ctypes.data_as(ctypes.POINTER(real)), kernel.size, mode, *walls,
            int(torch.cuda.current_stream().cuda_stream))
        _core.check(rc)
        out[r0:r1] = s_out
    torch.cuda.synchronize()
    return out.cpu().numpy()


CASES = {
    "closed-f32": (np.float32, "velocity", (("closed", "closed"), ("closed", "closed")), [40, 90]),
    "y-periodic-f64": (np.float64, "velocity", (("closed", "closed"), ("periodic", "periodic")), [64]),
    "all-periodic-pol-f32": (np.float32, "polarization", (("periodic", "periodic"), ("periodic", "periodic")), [33, 66, 99]),
    "x-periodic-pol-f64": (np.float64, "polarization", (("periodic", "periodic"), ("closed", "closed")), [50]),
}


@pytest.mark.parametrize("name", CASES)
def test_stitched_slabs_equal_the_whole_image(name):
    dtype, mode, bnd, cuts = CASES[name]
    rng = np.random.default_rng(31)
    ny, nx = 128, 70
    tex = rng.random((ny, nx)).astype(dtype)
    u = (rng.random((ny, nx)) - 0.5).astype(dtype)
    v = (rng.random((ny, nx)) - 0.5).astype(dtype)
    u[5, 5] = np.nan
    v[60, 3] = u[60, 3] = 0
    kernel = (rng.random(33) + 0.1).astype(dtype)    # reach 16 rows
    walls = _core.wall_codes(bnd)
    got = slab_pass_all(tex, u, v, kernel, bnd, walls, _core.mode_code(mode), cuts)
    want = oracle.convolve(tex, u, v, kernel=kernel, uv_mode=mode, boundaries=bnd)
    np.testing.assert_array_equal(got, want)


def test_halo_shorter_than_the_reach_is_refused():
    import torch

    dev = torch.device("cuda", 0)
    t = torch.zeros((40, 16), device=dev)
    f = torch.zeros((40, 16, 4), device=dev)
    o = torch.zeros((20, 16), device=dev)
    k = np.ones(33, dtype=np.float32)
    rc = _core.lib.rlic_b200_pass_slab_f32(
        t.data_ptr(), f.data_ptr(), o.data_ptr(), 100, 16, 30, 20, 10, 10,
        k.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), k.size, 0, 0, 0, 0, 0, None)
    assert rc == _core.ESHARD
    with pytest.raises(ValueError, match="halo"):
        _core.check(rc)


def test_device_entry_point_matches_host_entry_point():
    import torch

    import rlic_b200
    from rlic_b200.device import convolve_device

    rng = np.random.default_rng(4)
    tex = rng.random((200, 150), dtype=np.float32)
    u = rng.random((200, 150), dtype=np.float32) - 0.5
    v = rng.random((200, 150), dtype=np.float32) - 0.5
    k = np.linspace(0, 1, 21, dtype=np.float32)
    host = rlic_b200.convolve(tex, u, v, kernel=k, iterations=3, boundaries="periodic")
    dev = convolve_device(*(torch.from_numpy(a).cuda() for a in (tex, u, v)), kernel=k,
                          iterations=3, boundaries="periodic")
    np.testing.assert_array_equal(dev.cpu().numpy(), host)
    assert convolve_device(torch.from_numpy(tex).cuda(), torch.from_numpy(u).cuda(),
                           torch.from_numpy(v).cuda(), kernel=k, iterations=0).data_ptr() != 0


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(_core.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_sharded_run():
    """NCCL halo exchange + overlap on two real GPUs equals the single-GPU result."""
    script = ROOT / "tests" / "sharded_nccl_worker.py"
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    proc = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
         "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(script)],
        capture_output=True, text=True, env=env, timeout=600)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    assert "SHARDED_OK" in proc.stdout
