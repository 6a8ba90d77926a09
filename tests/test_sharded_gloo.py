"""Row-slab sharding logic on CPU: world_size 2 and 3 over gloo.

The halo plan, the message ordering and the strip/interior decomposition of
``rlic_b200.sharded`` run for real; only the per-slab compute step is replaced by
the CPU oracle (injected through ``pass_fn`` — test-only use of the oracle; the
product default is the CUDA slab pass).  Buffers are poisoned outside the rows a
rank legitimately holds, so a missing or misplaced halo shows up as a mismatch.
The sharded result must equal the unsharded oracle bit for bit.
"""

import os
import socket
import sys
import traceback
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

WALL_NAMES = {0: "closed", 1: "periodic"}


def oracle_slab_pass(tex, uv, out, plan, row0, nrows, halo_lo, halo_hi, taps, mode, walls):
    """CPU stand-in for rlic_b200_pass_slab_*: same contract, computed by the oracle."""
    import oracle

    ny, nx = plan.ny, plan.nx
    tex, uv = tex.numpy(), uv.numpy()
    rows_alloc = halo_lo + nrows + halo_hi
    periodic_y = walls[2] == 1
    g_tex = np.full((ny, nx), np.nan, dtype=tex.dtype)        # poison: unread rows stay NaN
    g_u = np.full((ny, nx), 1e30, dtype=tex.dtype)
    g_v = np.full((ny, nx), -1e30, dtype=tex.dtype)
    for k in range(rows_alloc):
        g = row0 - halo_lo + k
        if periodic_y:
            g %= ny
        elif not (0 <= g < ny):
            raise AssertionError("buffer row outside a closed image")
        g_tex[g] = tex[k]
        g_u[g] = uv[k, :, 0]
        g_v[g] = uv[k, :, 1]
    bnd = ((WALL_NAMES[walls[0]], WALL_NAMES[walls[1]]), (WALL_NAMES[walls[2]], WALL_NAMES[walls[3]]))
    band = oracle.pass_rows(g_tex, g_u, g_v, kernel=taps, rows=(row0, row0 + nrows),
                            uv_mode="polarization" if mode else "velocity", boundaries=bnd)
    out[:nrows].copy_(torch.from_numpy(band))


def cpu_pack(u, v, owned):
    # the product's packed record is (u, v, ru, rv); the stand-in only needs u, v
    owned[..., 0] = u
    owned[..., 1] = v
    owned[..., 2:] = 0


def _worker(rank, world, port, case, queue):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import oracle
        from rlic_b200.sharded import ShardedConvolver

        ny, nx, klen, boundaries, mode, iterations, dtype = case
        rng = np.random.default_rng(5)
        tex = rng.random((ny, nx)).astype(dtype)
        u = (rng.random((ny, nx)) - 0.5).astype(dtype)
        v = (rng.random((ny, nx)) - 0.5).astype(dtype)
        u[ny // 2, 3] = np.nan
        v[1, 1] = u[1, 1] = 0.0
        kernel = (rng.random(klen) + 0.1).astype(dtype)

        sc = ShardedConvolver(ny, nx, kernel=kernel, uv_mode=mode, boundaries=boundaries,
                              pass_fn=oracle_slab_pass, pack_fn=cpu_pack)
        p = sc.plan
        assert (p.row0, p.row1) == (ny * rank // world, ny * (rank + 1) // world)
        mine = slice(p.row0, p.row1)
        sc.set_field(torch.from_numpy(u[mine].copy()), torch.from_numpy(v[mine].copy()))
        got = sc.convolve(torch.from_numpy(tex[mine].copy()), iterations=iterations).numpy()

        bs_y = boundaries["y"] if isinstance(boundaries, dict) else boundaries
        bs_x = boundaries["x"] if isinstance(boundaries, dict) else boundaries
        want = oracle.convolve(tex, u, v, kernel=kernel, uv_mode=mode,
                               boundaries=((bs_x, bs_x), (bs_y, bs_y)), iterations=iterations)
        ok = np.array_equal(got, want[mine], equal_nan=True)
        gathered = [None] * world
        dist.all_gather_object(gathered, bool(ok))
        if rank == 0:
            queue.put(("ok", gathered))
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        queue.put(("error", f"rank {rank}:\n{traceback.format_exc()}"))
        raise


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


CASES = {
    "closed-2": (2, (40, 23, 9, "closed", "velocity", 3, np.float64)),
    "periodic-ring-2": (2, (36, 17, 11, "periodic", "velocity", 3, np.float32)),
    "y-periodic-pol-3": (3, (45, 20, 7, {"x": "closed", "y": "periodic"}, "polarization", 2, np.float64)),
    "x-periodic-3-uneven": (3, (31, 19, 13, {"x": "periodic", "y": "closed"}, "velocity", 4, np.float32)),
}


@pytest.mark.parametrize("name", CASES)
def test_sharded_equals_unsharded(name):
    world, case = CASES[name]
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, queue)) for r in range(world)]
    for pr in procs:
        pr.start()
    try:
        status, payload = queue.get(timeout=120)
    finally:
        for pr in procs:
            pr.join(timeout=60)
            if pr.is_alive():
                pr.terminate()
    assert status == "ok", payload
    assert all(payload), f"ranks with mismatching slabs: {payload}"


def test_plan_geometry():
    from rlic_b200.sharded import SlabPlan

    plans = [SlabPlan(ny=100, nx=8, world=4, rank=r, reach=5, periodic_y=False) for r in range(4)]
    assert [(p.row0, p.row1) for p in plans] == [(0, 25), (25, 50), (50, 75), (75, 100)]
    assert [p.up for p in plans] == [None, 0, 1, 2]
    assert [p.down for p in plans] == [1, 2, 3, None]
    assert [(p.halo_lo, p.halo_hi) for p in plans] == [(0, 5), (5, 5), (5, 5), (5, 0)]
    ring = [SlabPlan(ny=100, nx=8, world=4, rank=r, reach=5, periodic_y=True) for r in range(4)]
    assert [p.up for p in ring] == [3, 0, 1, 2]
    assert [p.down for p in ring] == [1, 2, 3, 0]
    assert all(p.rows_alloc == 35 for p in ring)
    single = SlabPlan(ny=10, nx=8, world=1, rank=0, reach=5, periodic_y=True)
    assert single.up is None and single.down is None and single.rows_alloc == 10


def test_slabs_thinner_than_the_kernel_reach_are_refused():
    from rlic_b200.sharded import SlabPlan

    with pytest.raises(ValueError, match="non-adjacent"):
        SlabPlan(ny=64, nx=8, world=8, rank=0, reach=32, periodic_y=False).validate()
    SlabPlan(ny=64, nx=8, world=2, rank=0, reach=32, periodic_y=False).validate()
