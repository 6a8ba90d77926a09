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
    """Cut the image into slabs, give every slab its halos by plain copies of padded
    rows (ring order when y is periodic), run one slab pass each, stitch."""
    import torch

    dev = torch.device("cuda", 0)
    ny, nx = tex.shape
    P = nx + 2
    h = kernel.size // 2
    periodic_y = bounds_spec[1][0] == "periodic"
    sfx, real = ("f32", ctypes.c_float) if tex.dtype == np.float32 else ("f64", ctypes.c_double)
    stream = lambda: int(torch.cuda.current_stream().cuda_stream)  # noqa: E731
    lib = _core.lib
    edges = [0, *cuts, ny]
    slabs = []
    for r0, r1 in zip(edges[:-1], edges[1:]):
        lo = h if (r0 > 0 or periodic_y) else 0
        hi = h if (r1 < ny or periodic_y) else 0
        args = (ny, nx, r0, r1 - r0, lo, hi)
        cells = _core.padded_cells(lo + (r1 - r0) + hi, nx)
        t_pad = torch.full((cells,), float("nan"), dtype=torch.from_numpy(tex).dtype, device=dev)
        f_pad = torch.full((4 * cells,), float("nan"), dtype=t_pad.dtype, device=dev)
        d_tex = torch.from_numpy(np.ascontiguousarray(tex[r0:r1])).to(dev)
        d_u = torch.from_numpy(np.ascontiguousarray(u[r0:r1])).to(dev)
        d_v = torch.from_numpy(np.ascontiguousarray(v[r0:r1])).to(dev)
        _core.check(getattr(lib, f"rlic_b200_slab_pad_texture_{sfx}")(
            d_tex.data_ptr(), *args, *walls, t_pad.data_ptr(), stream()))
        _core.check(getattr(lib, f"rlic_b200_slab_pack_field_{sfx}")(
            d_u.data_ptr(), d_v.data_ptr(), *args, *walls, f_pad.data_ptr(), stream()))
        slabs.append(dict(r0=r0, r1=r1, lo=lo, hi=hi, args=args, tex=t_pad, field=f_pad))

    def rows(s, buf, a, b, width, plane=0):   # padded rows [a, b) of a slab buffer, with wall cells
        cells = (s["lo"] + s["r1"] - s["r0"] + s["hi"] + 2) * P
        base = plane * cells
        return buf[(base + (a + 1) * P - 1) * width:(base + (b + 1) * P - 1) * width]

    # packed field: interleaved 4-scalar records for f32, two 2-scalar planes for f64
    f_width, f_planes = (4, 1) if tex.dtype == np.float32 else (2, 2)

    def owner(g):                         # slab and buffer row holding global row g
        g %= ny
        for s in slabs:
            if s["r0"] <= g < s["r1"]:
                return s, s["lo"] + g - s["r0"]
        raise AssertionError

    # halos, one global row at a time (a test, not a fast path)
    for s in slabs:
        wanted = list(range(s["r0"] - s["lo"], s["r0"])) + list(range(s["r1"], s["r1"] + s["hi"]))
        for k, g in enumerate(wanted):
            brow = k if k < s["lo"] else s["lo"] + (s["r1"] - s["r0"]) + (k - s["lo"])
            src, srow = owner(g)
            rows(s, s["tex"], brow, brow + 1, 1).copy_(rows(src, src["tex"], srow, srow + 1, 1))
            for plane in range(f_planes):
                rows(s, s["field"], brow, brow + 1, f_width, plane).copy_(
                    rows(src, src["field"], srow, srow + 1, f_width, plane))

    out = np.empty_like(tex)
    for s in slabs:
        n = s["r1"] - s["r0"]
        o_pad = torch.zeros_like(s["tex"])
        # two sub-ranges, as the overlapped multi-GPU driver issues them
        for a, b in ((0, n // 3), (n // 3, n)):
            _core.check(getattr(lib, f"rlic_b200_pass_slab_{sfx}")(
                s["tex"].data_ptr(), s["field"].data_ptr(), o_pad.data_ptr(), *s["args"], a, b - a,
                kernel.ctypes.data_as(ctypes.POINTER(real)), kernel.size, mode, *walls, stream()))
        d_out = torch.empty((n, nx), dtype=o_pad.dtype, device=dev)
        _core.check(getattr(lib, f"rlic_b200_slab_unpad_texture_{sfx}")(
            o_pad.data_ptr(), *s["args"], *walls, d_out.data_ptr(), stream()))
        out[s["r0"]:s["r1"]] = d_out.cpu().numpy()
    return out


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
    cells = _core.padded_cells(40, 16)
    t = torch.zeros(cells, device=dev)
    f = torch.zeros(4 * cells, device=dev)
    o = torch.zeros(cells, device=dev)
    k = np.ones(33, dtype=np.float32)
    rc = _core.lib.rlic_b200_pass_slab_f32(
        t.data_ptr(), f.data_ptr(), o.data_ptr(), 100, 16, 30, 20, 10, 10, 0, 20,
        k.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), k.size, 0, 0, 0, 0, 0, None)
    assert rc == _core.ESHARD
    with pytest.raises(ValueError, match="halo"):
        _core.check(rc)


def test_packed_field_reuse_and_out_argument():
    import torch

    import rlic_b200
    from rlic_b200.device import convolve_device, pack_field

    rng = np.random.default_rng(6)
    tex = rng.random((90, 130))
    u = rng.random((90, 130)) - 0.5
    v = rng.random((90, 130)) - 0.5
    k = np.linspace(0, 1, 15)
    bnd = {"x": "periodic", "y": "closed"}
    want = rlic_b200.convolve(tex, u, v, kernel=k, iterations=2, boundaries=bnd, uv_mode="polarization")
    d = [torch.from_numpy(a).cuda() for a in (tex, u, v)]
    field = pack_field(d[1], d[2], boundaries=bnd)
    out = torch.empty_like(d[0])
    got = convolve_device(d[0], field=field, kernel=k, iterations=2, boundaries=bnd,
                          uv_mode="polarization", out=out)
    assert got is out
    np.testing.assert_array_equal(out.cpu().numpy(), want)
    with pytest.raises(ValueError, match="packed for another"):
        convolve_device(d[0], field=field, kernel=k, boundaries="closed")


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


@pytest.mark.skipif(_core.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_fused_peer_exchange():
    """exchange="peer": the edge-strip passes store into the neighbour's halo through CUDA IPC
    mappings, counters order the iterations; the result equals the single-GPU one."""
    script = ROOT / "tests" / "sharded_nccl_worker.py"
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    proc = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
         "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(script),
         "--exchange", "peer"],
        capture_output=True, text=True, env=env, timeout=300)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    assert "SHARDED_OK" in proc.stdout


def test_batch_over_every_visible_device_with_pageable_inputs():
    """rlic_b200_convolve_batch_* with devices=NULL: whole fields split over all visible GPUs,
    two host lanes per device, pageable inputs of 1 MiB per field staged through the pinned
    pools (one pool per device: an event of one device cannot be recorded on another's stream).
    Every field equals its own single-image convolve."""
    rng = np.random.default_rng(21)
    nf = 4 * max(1, _core.device_count()) + 1
    tex = rng.random((nf, 512, 512), dtype=np.float32)
    u = rng.random((nf, 512, 512), dtype=np.float32) - 0.5
    v = rng.random((nf, 512, 512), dtype=np.float32) - 0.5
    kernel = np.linspace(0.1, 1.0, 33, dtype=np.float32)
    import rlic_b200

    got = rlic_b200.convolve_batch(tex, u, v, kernel=kernel, boundaries="closed", iterations=3)
    for f in sorted({0, 1, nf // 2, nf - 1}):
        want = oracle.convolve(tex[f], u[f], v[f], kernel=kernel, iterations=3, threads=oracle.max_threads())
        np.testing.assert_array_equal(got[f], want)
    # and field by field on every device in turn through the single-image host entry
    for d in range(_core.device_count()):
        _core.check(_core.lib.rlic_b200_set_device(d))
        try:
            one = rlic_b200.convolve(tex[d % nf], u[d % nf], v[d % nf], kernel=kernel, iterations=3)
        finally:
            _core.check(_core.lib.rlic_b200_set_device(0))
        np.testing.assert_array_equal(one, got[d % nf])
