"""Worker for tests/test_slab.py::test_two_gpu_sharded_run and ::test_two_gpu_fused_peer_exchange
(launched by torchrun).  `--exchange peer` runs the sharded passes with the halo exchange fused
into the edge-strip kernels (peer stores through CUDA IPC mappings) instead of NCCL messages."""

import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main() -> None:
    exchange = sys.argv[sys.argv.index("--exchange") + 1] if "--exchange" in sys.argv else "nccl"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from rlic_b200 import _core
    from rlic_b200.device import convolve_device
    from rlic_b200.sharded import ShardedConvolver

    _core.check(_core.lib.rlic_b200_set_device(local))
    ok = True
    for bnd, mode, ny in (("closed", "velocity", 1024), ("periodic", "polarization", 768),
                          ({"x": "periodic", "y": "closed"}, "velocity", 515)):
        rng = np.random.default_rng(9)
        nx = 640
        tex = rng.random((ny, nx), dtype=np.float32)
        u = rng.random((ny, nx), dtype=np.float32) - 0.5
        v = rng.random((ny, nx), dtype=np.float32) - 0.5
        kernel = np.linspace(0.1, 1, 65, dtype=np.float32)
        sc = ShardedConvolver(ny, nx, kernel=kernel, uv_mode=mode, boundaries=bnd, exchange=exchange)
        sc.PEER_TIMEOUT_MS = 5_000
        mine = slice(sc.plan.row0, sc.plan.row1)
        sc.set_field(torch.from_numpy(u[mine].copy()).to(dev), torch.from_numpy(v[mine].copy()).to(dev))
        for overlap in (True, False):
            got = sc.convolve(torch.from_numpy(tex[mine].copy()).to(dev), iterations=4, overlap=overlap)
            want = convolve_device(*(torch.from_numpy(a).to(dev) for a in (tex, u, v)), kernel=kernel,
                                   uv_mode=mode, boundaries=bnd, iterations=4)
            ok &= bool(torch.equal(got, want[mine]))
        # host slabs in, host slabs out (with exchange="peer": the band pipeline on three
        # streams, halos of texture and field by peer copies); then the same field, new texture
        host = sc.convolve_host(tex[mine], u[mine], v[mine], iterations=4, min_band_pixels=1)
        ok &= bool(np.array_equal(host, want[mine].cpu().numpy()))
        tex2 = rng.random((ny, nx), dtype=np.float32)
        host2 = sc.convolve_host(tex2[mine], iterations=3, min_band_pixels=1)
        want2 = convolve_device(*(torch.from_numpy(a).to(dev) for a in (tex2, u, v)), kernel=kernel,
                                uv_mode=mode, boundaries=bnd, iterations=3)
        ok &= bool(np.array_equal(host2, want2[mine].cpu().numpy()))
        if exchange == "peer":
            ok &= not sc.peer_timed_out()
            sc.close()
    flags = [None] * world
    dist.all_gather_object(flags, ok)
    if rank == 0:
        print("SHARDED_OK" if all(flags) else f"SHARDED_MISMATCH {flags}")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if all(flags) else 1)


if __name__ == "__main__":
    main()
