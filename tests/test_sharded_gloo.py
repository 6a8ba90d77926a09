"""Row-slab sharding logic on CPU: world_size 2 and 3 over gloo.

The halo plan, the message ordering, the padded-buffer row ranges and the
strip/interior decomposition of ``rlic_b200.sharded`` run for real; only the
per-slab compute steps are replaced (injected through ``ops``; the product
default is the CUDA slab API), in two ways: by the CPU oracle on poisoned
buffers, so that a missing or misplaced halo shows up as a mismatch; and by the
library's own kernels compiled for the CPU (tests/kernel_emulation), so that the
real packed records, sentinels and wall cells are what travels.
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


class OracleSlabOps:
    """CPU stand-in for the slab building blocks of the C ABI (rlic_b200_slab_* and
    rlic_b200_pass_slab_*): same contract on the same padded buffers, computed by the
    oracle.  Buffers are poisoned (NaN / huge) wherever a rank legitimately holds
    nothing, and the wall cells of every row a pass may read are checked, so a
    missing, misplaced or truncated halo message shows up as a failure."""

    @staticmethod
    def _walls(plan, walls):
        ny, nx = plan.ny, plan.nx
        periodic_y = walls[2] == 1
        shift = plan.row0 - plan.halo_lo
        return dict(
            j_below_to=nx - 1 if walls[0] == 1 else 0,
            j_above_to=0 if walls[1] == 1 else nx - 1,
            i_below_to=(ny - 1 if walls[2] == 1 else 0) - shift,
            i_above_to=(0 if walls[3] == 1 else ny - 1) - shift,
            lo_wall=plan.world == 1 or (plan.row0 == 0 and not periodic_y),
            hi_wall=plan.world == 1 or (plan.row1 == ny and not periodic_y),
        )

    @staticmethod
    def _cell(plan, r, c):
        """flat cell index of buffer row r, column c (c = -1 .. nx are the wall cells)"""
        return (r + 1) * plan.pitch + c

    def _write_rows(self, flat, plan, walls, rows, values):
        """values: dict row -> 1-D array of nx pixels; writes pixels and their wall cells"""
        w = self._walls(plan, walls)
        for r, line in values.items():
            base = self._cell(plan, r, 0)
            flat[base:base + plan.nx] = torch.from_numpy(np.ascontiguousarray(line))
            flat[self._cell(plan, r, plan.nx)] = float(line[w["j_above_to"]])
            flat[self._cell(plan, r, -1)] = float(line[w["j_below_to"]])
            if w["lo_wall"] and r == w["i_below_to"]:
                g = self._cell(plan, -1, 0)
                flat[g:g + plan.nx] = torch.from_numpy(np.ascontiguousarray(line))
            if w["hi_wall"] and r == w["i_above_to"]:
                g = self._cell(plan, plan.rows_alloc, 0)
                flat[g:g + plan.nx] = torch.from_numpy(np.ascontiguousarray(line))

    @staticmethod
    def _uv_views(field, plan):
        """(u, v, ru, rv) views over the cells of a packed-field buffer, following the
        product's layout: interleaved records for f32, two planes for f64."""
        if field.dtype == torch.float32:
            f = field.view(-1, 4)
            return f[:, 0], f[:, 1], f[:, 2], f[:, 3]
        a = field[:2 * plan.cells].view(-1, 2)
        b = field[2 * plan.cells:].view(-1, 2)
        return a[:, 0], a[:, 1], b[:, 0], b[:, 1]

    def pack_field(self, u, v, field, plan, walls):
        field[:] = float("nan")
        fu, fv, fru, frv = self._uv_views(field, plan)
        for k in range(plan.nrows):
            r = plan.halo_lo + k
            base = self._cell(plan, r, 0)
            fu[base:base + plan.nx] = u[k]
            fv[base:base + plan.nx] = v[k]
            fru[base:base + plan.nx] = 0
            frv[base:base + plan.nx] = 0
            # stand-in sentinels: tag the wall cells of owned rows so that their transport can be checked
            for part in (fu, fv, fru, frv):
                part[self._cell(plan, r, plan.nx)] = 12345.0
                part[self._cell(plan, r, -1)] = 54321.0

    # Whether pad_texture poisons the halo rows too.  The product's kernel writes the owned rows
    # and their wall cells only, and the fused peer exchange relies on that: a neighbour may
    # deliver the halos of a call's input before this rank has padded its own rows (it waits for
    # this rank's PREVIOUS call to end, not for this call to begin).  With message passing the
    # halos arrive after the padding, so there everything can be poisoned.
    poison_halos = True

    def pad_texture(self, texture, padded, plan, walls):
        if self.poison_halos:
            padded[:] = float("nan")
        else:
            own = plan.row_cells(plan.halo_lo, plan.halo_lo + plan.nrows)
            padded[own] = float("nan")
        tex = texture.numpy()
        self._write_rows(padded, plan, walls, None,
                         {plan.halo_lo + k: tex[k] for k in range(plan.nrows)})

    def unpad_texture(self, padded, texture, plan, walls):
        for k in range(plan.nrows):
            base = self._cell(plan, plan.halo_lo + k, 0)
            texture[k] = padded[base:base + plan.nx]

    def pass_rows(self, src, field, dst, plan, a, b, taps, mode, walls):
        import oracle

        ny, nx = plan.ny, plan.nx
        w = self._walls(plan, walls)
        periodic_y = walls[2] == 1
        fu, fv, fru, frv = (t.numpy() for t in self._uv_views(field, plan))
        s = src.numpy()
        g_tex = np.full((ny, nx), np.nan, dtype=s.dtype)        # poison: unread rows stay NaN
        g_u = np.full((ny, nx), 1e30, dtype=s.dtype)
        g_v = np.full((ny, nx), -1e30, dtype=s.dtype)
        reach = taps.size // 2
        lo = max(0, plan.halo_lo + a - reach)
        hi = min(plan.rows_alloc, plan.halo_lo + b + reach)
        for r in range(lo, hi):                                   # rows this pass may read
            g = plan.row0 - plan.halo_lo + r
            if periodic_y:
                g %= ny
            assert 0 <= g < ny, "buffer row outside a closed image"
            base = self._cell(plan, r, 0)
            g_tex[g] = s[base:base + nx]
            g_u[g] = fu[base:base + nx]
            g_v[g] = fv[base:base + nx]
            # the wall cells must have travelled with their rows
            assert s[self._cell(plan, r, nx)] == g_tex[g][w["j_above_to"]] or np.isnan(g_tex[g][w["j_above_to"]])
            assert s[self._cell(plan, r, -1)] == g_tex[g][w["j_below_to"]] or np.isnan(g_tex[g][w["j_below_to"]])
            for part in (fu, fv, fru, frv):
                assert part[self._cell(plan, r, nx)] == 12345.0 and part[self._cell(plan, r, -1)] == 54321.0
        bnd = ((WALL_NAMES[walls[0]], WALL_NAMES[walls[1]]), (WALL_NAMES[walls[2]], WALL_NAMES[walls[3]]))
        band = oracle.pass_rows(g_tex, g_u, g_v, kernel=taps, rows=(plan.row0 + a, plan.row0 + b),
                                uv_mode="polarization" if mode else "velocity", boundaries=bnd)
        self._write_rows(dst, plan, walls, None,
                         {plan.halo_lo + a + k: band[k] for k in range(b - a)})


class SharedMemoryPeers:
    """CPU stand-in for rlic_b200.sharded.CudaPeerMemory (the CUDA IPC entry points of the C
    ABI): named shared-memory segments play the device allocations other ranks can map."""

    def __init__(self):
        self._segments = {}

    def alloc(self, nbytes):
        from multiprocessing import shared_memory

        seg = shared_memory.SharedMemory(create=True, size=nbytes)
        seg.buf[:nbytes] = bytes(nbytes)
        self._segments[id(seg)] = seg
        return id(seg), seg.name.encode()

    def open(self, handle):
        from multiprocessing import shared_memory

        seg = shared_memory.SharedMemory(name=handle.decode())
        self._segments[id(seg)] = seg
        return id(seg)

    def view(self, ptr, count, dtype, device):
        return torch.frombuffer(self._segments[ptr].buf, dtype=dtype, count=count)

    def close(self, ptr):
        self._segments.pop(ptr).close()

    def free(self, ptr):
        seg = self._segments.pop(ptr)
        seg.close()
        seg.unlink()


class HostFlags:
    """signal / wait of the peer exchange on shared memory: program order is stream order."""

    def signal(self, flags, index, value):
        flags[index] = value

    def wait(self, flags, index, value, timeout_ms, timed_out_index):
        import time

        deadline = time.monotonic() + min(timeout_ms, 30_000) / 1e3
        while int(flags[index]) - value < 0:
            if time.monotonic() > deadline:
                flags[timed_out_index] = 1
                return
            time.sleep(0.0005)

    # the one-launch forms (rlic_b200_peer_signal2 / _peer_wait4)
    def signal_many(self, targets, value):
        for flags, index in targets:
            self.signal(flags, index, value)

    def wait_many(self, flags, waits, timeout_ms, timed_out_index):
        for index, value in waits:
            self.wait(flags, index, value, timeout_ms, timed_out_index)


def _peer_methods(cls):
    """Adds the peer-exchange half of the ops interface to a CPU ops class."""

    class WithPeers(cls, HostFlags):
        def pass_rows_peer(self, src, field, dst, plan, a, b, taps, mode, walls, peer, peer_row_delta):
            if hasattr(super(), "pass_rows_peer"):      # the emulated kernels have the real thing
                return super().pass_rows_peer(src, field, dst, plan, a, b, taps, mode, walls, peer,
                                              peer_row_delta)
            self.pass_rows(src, field, dst, plan, a, b, taps, mode, walls)
            cells = plan.row_cells(plan.halo_lo + a, plan.halo_lo + b)
            shift = peer_row_delta * plan.pitch
            peer[cells.start + shift:cells.stop + shift] = dst[cells]

    return WithPeers()


def _make_ops(kind):
    if kind == "oracle stand-in":
        return _peer_methods(OracleSlabOps)
    # the library's own kernels, compiled for the CPU (tests/kernel_emulation): real packed
    # records, sentinels and wall cells travel through the exchange
    sys.path.insert(0, str(ROOT / "tests"))
    import kernel_emulation

    ops = _peer_methods(kernel_emulation.SlabOps)
    ops.walk = 1 if kind.endswith("grouped walk") else 0
    return ops


def _worker(rank, world, port, case, queue, ops_kind="oracle stand-in", exchange="nccl", calls=1):
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

        ops = _make_ops(ops_kind)
        ops.poison_halos = exchange != "peer"
        sc = ShardedConvolver(ny, nx, kernel=kernel, uv_mode=mode, boundaries=boundaries,
                              ops=ops, exchange=exchange,
                              peers=SharedMemoryPeers() if exchange == "peer" else None)
        p = sc.plan
        assert (p.row0, p.row1) == (ny * rank // world, ny * (rank + 1) // world)
        mine = slice(p.row0, p.row1)
        sc.set_field(torch.from_numpy(u[mine].copy()), torch.from_numpy(v[mine].copy()))
        for _ in range(calls):     # a convolver is reused: the counters of the peer exchange carry on
            got = sc.convolve(torch.from_numpy(tex[mine].copy()), iterations=iterations).numpy()
        if exchange == "peer":
            assert not sc.peer_timed_out(), "a wait for a neighbour gave up"
            sc.close()

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


# ---------------------------------------------------------------------------------------------
# The ranks of these tests are three long-lived processes: starting a process and importing
# torch in it takes four to five seconds, and the thirty-odd cases below would spend four minutes
# doing only that.  Every case still gets a fresh process group (its own port, created and
# destroyed by the worker function) and fresh ops / peer objects; a case that fails or times out
# takes the pool down with it, so that the next one starts from new processes.
def _rank_loop(tasks, results):
    while True:
        job = tasks.get()
        if job is None:
            return
        name, args = job
        rank = args[0]
        try:
            globals()[name](*args[:4], results, *args[4:])
            results.put(("done", rank))
        except BaseException:                       # the worker has reported it; this process is done for
            results.put(("died", rank))
            return


class _RankPool:
    SIZE = 3

    def __init__(self):
        ctx = mp.get_context("spawn")
        self.results = ctx.Queue()
        self.tasks = [ctx.Queue() for _ in range(self.SIZE)]
        self.procs = [ctx.Process(target=_rank_loop, args=(t, self.results), daemon=True) for t in self.tasks]
        for pr in self.procs:
            pr.start()

    def stop(self, kill=False):
        for t, pr in zip(self.tasks, self.procs):
            if not kill:
                t.put(None)
        for pr in self.procs:
            if not kill:
                pr.join(timeout=20)
            if pr.is_alive():
                pr.terminate()
                pr.join(timeout=20)


_pool = None


def _run_ranks(worker, world, args, timeout):
    """Runs `worker(rank, world, port, *args[:1], queue, *args[1:])` on `world` ranks of the pool and
    returns what rank 0 reported: ("ok", payload) or ("error", traceback)."""
    import queue as queue_errors
    import time

    global _pool
    if _pool is None:
        _pool = _RankPool()
    pool, port = _pool, _free_port()
    for r in range(world):
        pool.tasks[r].put((worker.__name__, (r, world, port, *args)))
    report, finished, deadline = None, 0, time.monotonic() + timeout
    try:
        while finished < world:
            kind, payload = pool.results.get(timeout=max(0.1, deadline - time.monotonic()))
            if kind in ("done", "died"):
                finished += 1
                if kind == "died":
                    report = report if report and report[0] == "error" else ("error", f"rank {payload} died")
            elif kind == "error" or report is None:
                report = (kind, payload)
            if report is not None and report[0] == "error":     # the others may be waiting for the failed rank
                deadline = min(deadline, time.monotonic() + 10)
    except queue_errors.Empty:
        if report is None or report[0] != "error":
            report = ("error", f"timed out after {timeout} s; last report: {report}")
    if report is None or report[0] != "ok":
        pool.stop(kill=True)
        _pool = None
    return report or ("error", "no rank reported")


@pytest.fixture(scope="module", autouse=True)
def _rank_pool_lifetime():
    yield
    global _pool
    if _pool is not None:
        _pool.stop()
        _pool = None


CASES = {
    "closed-2": (2, (40, 23, 9, "closed", "velocity", 3, np.float64)),
    "periodic-ring-2": (2, (36, 17, 11, "periodic", "velocity", 3, np.float32)),
    "y-periodic-pol-3": (3, (45, 20, 7, {"x": "closed", "y": "periodic"}, "polarization", 2, np.float64)),
    "x-periodic-3-uneven": (3, (31, 19, 13, {"x": "periodic", "y": "closed"}, "velocity", 4, np.float32)),
}


@pytest.mark.parametrize("ops_kind", ["oracle stand-in", "emulated kernels"])
@pytest.mark.parametrize("name", CASES)
def test_sharded_equals_unsharded(name, ops_kind):
    world, case = CASES[name]
    status, payload = _run_ranks(_worker, world, (case, ops_kind), timeout=120)
    assert status == "ok", payload
    assert all(payload), f"ranks with mismatching slabs: {payload}"


PEER_CASES = {
    # slabs of at least four kernel half-widths, as exchange="peer" demands
    "closed-2": (2, (40, 23, 9, "closed", "velocity", 4, np.float64)),
    "periodic-ring-2": (2, (48, 17, 11, "periodic", "velocity", 3, np.float32)),
    "y-periodic-pol-3": (3, (45, 20, 7, {"x": "closed", "y": "periodic"}, "polarization", 5, np.float64)),
    "x-periodic-3-uneven": (3, (77, 19, 13, {"x": "periodic", "y": "closed"}, "velocity", 4, np.float32)),
    "single-iteration-3": (3, (45, 20, 7, "periodic", "velocity", 1, np.float32)),
}


PEER_RUNS = ([(name, "oracle stand-in") for name in PEER_CASES]
             + [(name, "emulated kernels") for name in ("closed-2", "y-periodic-pol-3", "x-periodic-3-uneven")]
             + [(name, "emulated kernels, grouped walk") for name in ("periodic-ring-2", "y-periodic-pol-3")])


@pytest.mark.parametrize("name,ops_kind", PEER_RUNS)
def test_fused_peer_exchange_equals_unsharded(name, ops_kind):
    """exchange="peer": the strips land in the neighbours' halos through mapped memory (here:
    named shared memory) and the halo / free counters order the passes; two calls in a row,
    so that the counters and the start-of-call handshake are exercised too."""
    world, case = PEER_CASES[name]
    status, payload = _run_ranks(_worker, world, (case, ops_kind, "peer", 2), timeout=180)
    assert status == "ok", payload
    assert all(payload), f"ranks with mismatching slabs: {payload}"


def _host_worker(rank, world, port, case, queue, exchange):
    """ShardedConvolver.convolve_host: host slab in, host slab out.  With exchange="peer" this is
    the band pipeline (uploads, pass 1 trailing them, downloads behind the last pass, halos of
    texture and field by peer copies), run here in program order on the emulated kernels."""
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import oracle
        from rlic_b200.sharded import ShardedConvolver

        ny, nx, klen, boundaries, mode, iterations, dtype = case
        rng = np.random.default_rng(17)
        kernel = (rng.random(klen) + 0.1).astype(dtype)
        sc = ShardedConvolver(ny, nx, kernel=kernel, uv_mode=mode, boundaries=boundaries,
                              ops=_make_ops("emulated kernels"), exchange=exchange,
                              peers=SharedMemoryPeers() if exchange == "peer" else None)
        mine = slice(sc.plan.row0, sc.plan.row1)
        bs_y = boundaries["y"] if isinstance(boundaries, dict) else boundaries
        bs_x = boundaries["x"] if isinstance(boundaries, dict) else boundaries
        ok = True
        # three calls on one convolver: new field + texture, same field + new texture (u, v
        # omitted), and a new field again with another iteration count
        for call, (with_field, its) in enumerate(((True, iterations), (False, iterations), (True, 1))):
            tex = rng.random((ny, nx)).astype(dtype)
            if with_field:
                u = (rng.random((ny, nx)) - 0.5).astype(dtype)
                v = (rng.random((ny, nx)) - 0.5).astype(dtype)
                u[ny // 2, 3] = np.nan
                v[1, 1] = u[1, 1] = 0.0
            got = sc.convolve_host(tex[mine], u[mine] if with_field else None, v[mine] if with_field else None,
                                   iterations=its, min_band_pixels=1, min_band_rows=16)
            want = oracle.convolve(tex, u, v, kernel=kernel, uv_mode=mode,
                                   boundaries=((bs_x, bs_x), (bs_y, bs_y)), iterations=its)
            ok &= bool(np.array_equal(got, want[mine], equal_nan=True))
        if exchange == "peer":
            assert not sc.peer_timed_out(), "a wait for a neighbour gave up"
            sc.close()
        gathered = [None] * world
        dist.all_gather_object(gathered, bool(ok))
        if rank == 0:
            queue.put(("ok", gathered))
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        queue.put(("error", f"rank {rank}:\n{traceback.format_exc()}"))
        raise


HOST_CASES = {
    # slabs of 96 / 112 rows cut into bands of 16 and 32 rows
    "closed-2": (2, (192, 23, 9, "closed", "velocity", 4, np.float32)),
    "periodic-ring-2": (2, (224, 17, 11, "periodic", "polarization", 3, np.float64)),
    "x-periodic-3-uneven": (3, (301, 19, 13, {"x": "periodic", "y": "closed"}, "velocity", 2, np.float32)),
    "y-periodic-3": (3, (288, 20, 7, {"x": "closed", "y": "periodic"}, "velocity", 5, np.float64)),
}


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
@pytest.mark.parametrize("name", HOST_CASES)
def test_host_slabs_in_host_slabs_out(name, exchange):
    if exchange == "nccl" and name not in ("closed-2", "y-periodic-3"):
        pytest.skip("the unpipelined path is covered by two cases")
    world, case = HOST_CASES[name]
    status, payload = _run_ranks(_host_worker, world, (case, exchange), timeout=180)
    assert status == "ok", payload
    assert all(payload), f"ranks with mismatching slabs: {payload}"


def test_peer_exchange_rejects_unknown_modes_and_needs_neighbours():
    import torch.distributed as dist

    from rlic_b200.sharded import ShardedConvolver

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(_free_port())
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        with pytest.raises(ValueError, match="unknown exchange"):
            ShardedConvolver(64, 8, kernel=np.ones(5), exchange="carrier pigeon")
        sc = ShardedConvolver(64, 8, kernel=np.ones(5), exchange="peer")
        assert sc.exchange == "nccl"        # nobody to exchange with
    finally:
        dist.destroy_process_group()


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
    assert ring[0].pitch == 10 and ring[0].cells == 37 * 10
    assert ring[0].row_cells(0, 5) == slice(9, 59)      # rows 0..4 with their wall cells
    single = SlabPlan(ny=10, nx=8, world=1, rank=0, reach=5, periodic_y=True)
    assert single.up is None and single.down is None and single.rows_alloc == 10


def test_slabs_thinner_than_the_kernel_reach_are_refused():
    from rlic_b200.sharded import SlabPlan

    with pytest.raises(ValueError, match="non-adjacent"):
        SlabPlan(ny=64, nx=8, world=8, rank=0, reach=32, periodic_y=False).validate()
    SlabPlan(ny=64, nx=8, world=2, rank=0, reach=32, periodic_y=False).validate()


# ---------------------------------------------------------------------------------------------
# A peer set-up that fails on ONE rank must end on EVERY rank, with everything released, so that
# the caller can fall back to the NCCL exchange consistently (PeerMemoryUnavailable).
class _BrokenPeers(SharedMemoryPeers):
    """Fails on one rank, in `alloc` (after `fail_after` successful calls) or in `open`."""

    live = 0          # segments this process holds (allocated or mapped)

    def __init__(self, rank, broken_rank, where, fail_after=2):
        super().__init__()
        self.broken = rank == broken_rank
        self.where, self.left = where, fail_after

    def _maybe_fail(self, where):
        if self.broken and self.where == where:
            if self.left == 0:
                raise RuntimeError(f"no peer access ({where})")
            self.left -= 1

    def alloc(self, nbytes):
        self._maybe_fail("alloc")
        return super().alloc(nbytes)

    def open(self, handle):
        self._maybe_fail("open")
        return super().open(handle)


def _broken_peer_worker(rank, world, port, where, queue):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import oracle
        from rlic_b200.sharded import PeerMemoryUnavailable, ShardedConvolver

        ny, nx, klen = 96, 19, 9
        rng = np.random.default_rng(3)
        tex = rng.random((ny, nx)).astype(np.float32)
        u = (rng.random((ny, nx)) - 0.5).astype(np.float32)
        v = (rng.random((ny, nx)) - 0.5).astype(np.float32)
        kernel = (rng.random(klen) + 0.1).astype(np.float32)
        peers = _BrokenPeers(rank, world - 1, where)
        sc = ShardedConvolver(ny, nx, kernel=kernel, ops=_make_ops("emulated kernels"), exchange="peer", peers=peers)
        mine = slice(sc.plan.row0, sc.plan.row1)
        message = None
        try:
            sc.set_field(torch.from_numpy(u[mine].copy()), torch.from_numpy(v[mine].copy()))
        except PeerMemoryUnavailable as exc:
            message = str(exc)
        assert message is not None, "the broken set-up went unnoticed on this rank"
        assert f"rank {world - 1}: RuntimeError: no peer access ({where})" in message
        assert not peers._segments, "segments left allocated or mapped after the failed set-up"
        assert sc._peer is None
        # ... and the fallback every rank takes gives the right answer
        ops = _make_ops("emulated kernels")
        ops.poison_halos = True
        sc = ShardedConvolver(ny, nx, kernel=kernel, ops=ops, exchange="nccl")
        sc.set_field(torch.from_numpy(u[mine].copy()), torch.from_numpy(v[mine].copy()))
        got = sc.convolve(torch.from_numpy(tex[mine].copy()), iterations=3).numpy()
        want = oracle.convolve(tex, u, v, kernel=kernel, iterations=3)
        gathered = [None] * world
        dist.all_gather_object(gathered, bool(np.array_equal(got, want[mine])))
        if rank == 0:
            queue.put(("ok", gathered))
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        queue.put(("error", f"rank {rank}:\n{traceback.format_exc()}"))
        raise


@pytest.mark.parametrize("world, where", [(2, "alloc"), (3, "open")])
def test_a_peer_set_up_that_fails_on_one_rank_ends_on_every_rank(world, where):
    status, payload = _run_ranks(_broken_peer_worker, world, (where,), timeout=120)
    assert status == "ok", payload
    assert all(payload), f"ranks with mismatching slabs after the fallback: {payload}"
