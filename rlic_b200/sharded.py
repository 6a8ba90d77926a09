"""Row-slab sharding of one image over the ranks of a ``torch.distributed`` group
(BASELINE config 4).

Rank ``r`` owns a contiguous block of rows.  A walker moves at most one row per
tap, so a pass over the owned rows reads at most ``h = len(kernel) // 2`` rows
beyond them: each rank keeps ``h`` halo rows on either side.  The vector field
is iteration-invariant (its halos are exchanged once); the texture halos are
exchanged every iteration with point-to-point sends (NCCL over NVLink on GPUs,
gloo in the CPU tests).  Each iteration computes the two edge strips first,
ships them to the neighbours on a side stream, and computes the interior rows
while that exchange is in flight.

Buffers are the kernels' padded buffers (``include/rlic_b200.h``): a pitch of
``nx + 2`` cells and a guard row above and below; buffer row ``r`` together
with its two wall cells is the contiguous cell range
``[(r + 1) * pitch - 1, (r + 2) * pitch - 1)``, which is what travels.

The per-pixel arithmetic is the single-GPU kernel's, in global row numbers, so
the gathered result is bit-identical to an unsharded run.

``exchange="peer"`` (what ``bench.py --gpus N`` and ``tools/bench_c4_scaling.py`` use; measured
on 2, 4 and 8 B200s: DESIGN.md section 7) fuses the per-iteration texture exchange into the
passes over the edge strips: the pass kernel stores every
row it computes there into the neighbour's halo as well, through that neighbour's
buffer mapped into this process (CUDA IPC; the stores travel over NVLink while the
strip is still being computed), and two pairs of counters per neighbour -- raised by
one-thread kernels after the launches they announce, awaited by one-thread kernels on
the consumer's stream -- order the iterations: ``halo`` ("the rows for your pass n + 1
are in your buffer") and ``free`` ("my pass n is done, its input buffer may be
overwritten").  No separate exchange step, no host synchronisation, nothing on a side
stream.

The compute steps are injectable (``ops``), and so is the peer memory (``peers``), so the
exchange logic can be exercised without a GPU; the defaults are the CUDA slab API and the
CUDA IPC entry points of the C ABI.  There is no CPU compute path in this package.
"""

from __future__ import annotations

__all__ = ["SlabPlan", "ShardedConvolver", "CudaSlabOps", "CudaPeerMemory", "PeerMemoryUnavailable", "pinned_empty"]

import ctypes
from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist

from rlic_b200._boundaries import BoundarySet


@dataclass(frozen=True)
class SlabPlan:
    """Which rows a rank owns and whom it exchanges halos with."""

    ny: int
    nx: int
    world: int
    rank: int
    reach: int          # kernel half-width: rows a walker can travel
    periodic_y: bool

    @property
    def row0(self) -> int:
        return self.ny * self.rank // self.world

    @property
    def row1(self) -> int:
        return self.ny * (self.rank + 1) // self.world

    @property
    def nrows(self) -> int:
        return self.row1 - self.row0

    @property
    def up(self) -> int | None:
        """Rank owning the rows just above ours (smaller row numbers)."""
        if self.world == 1:
            return None
        if self.rank > 0:
            return self.rank - 1
        return self.world - 1 if self.periodic_y else None

    @property
    def down(self) -> int | None:
        if self.world == 1:
            return None
        if self.rank < self.world - 1:
            return self.rank + 1
        return 0 if self.periodic_y else None

    @property
    def halo_lo(self) -> int:
        return self.reach if self.up is not None else 0

    @property
    def halo_hi(self) -> int:
        return self.reach if self.down is not None else 0

    @property
    def rows_alloc(self) -> int:
        return self.halo_lo + self.nrows + self.halo_hi

    @property
    def pitch(self) -> int:
        return self.nx + 2

    @property
    def cells(self) -> int:
        """Cells of one padded buffer (include/rlic_b200.h: rlic_b200_padded_cells)."""
        return (self.rows_alloc + 2) * self.pitch

    def row_cells(self, a: int, b: int) -> slice:
        """Cell range of buffer rows [a, b) with their wall cells."""
        return slice((a + 1) * self.pitch - 1, (b + 1) * self.pitch - 1)

    @property
    def slab_args(self) -> tuple[int, int, int, int, int, int]:
        return (self.ny, self.nx, self.row0, self.nrows, self.halo_lo, self.halo_hi)

    def validate(self) -> None:
        if self.world < 1 or not (0 <= self.rank < self.world):
            raise ValueError(f"bad rank {self.rank} of {self.world}")
        if self.world > 1:
            smallest = min(self.ny * (r + 1) // self.world - self.ny * r // self.world
                           for r in range(self.world))
            if smallest < max(self.reach, 1):
                raise ValueError(
                    f"cannot shard {self.ny} rows over {self.world} ranks with a kernel half-width "
                    f"of {self.reach}: slabs of {smallest} rows would need halos from non-adjacent ranks"
                )


class CudaSlabOps:
    """The slab building blocks of the C ABI, on torch CUDA tensors."""

    @staticmethod
    def _kind(t):
        return ("f32", ctypes.c_float) if t.dtype == torch.float32 else ("f64", ctypes.c_double)

    @staticmethod
    def _stream():
        return int(torch.cuda.current_stream().cuda_stream)

    @staticmethod
    def _need_cuda(t):
        if not t.is_cuda:
            raise RuntimeError("rlic_b200 has no CPU fallback: pass CUDA tensors")

    # pack_field / pad_texture / unpad_texture take an optional ``rows=(a, b)``: owned rows
    # [a, b) only, the dense tensor holding just those rows (the *_rows entry points)
    def pack_field(self, u, v, field, plan, walls, rows=None):
        from rlic_b200 import _core

        self._need_cuda(u)
        sfx, _ = self._kind(u)
        a, b = rows if rows is not None else (0, plan.nrows)
        _core.check(getattr(_core.lib, f"rlic_b200_slab_pack_field_rows_{sfx}")(
            u.data_ptr(), v.data_ptr(), *plan.slab_args, a, b - a, *walls, field.data_ptr(), self._stream()))

    def pad_texture(self, texture, padded, plan, walls, rows=None):
        from rlic_b200 import _core

        self._need_cuda(texture)
        sfx, _ = self._kind(texture)
        a, b = rows if rows is not None else (0, plan.nrows)
        _core.check(getattr(_core.lib, f"rlic_b200_slab_pad_texture_rows_{sfx}")(
            texture.data_ptr(), *plan.slab_args, a, b - a, *walls, padded.data_ptr(), self._stream()))

    def unpad_texture(self, padded, texture, plan, walls, rows=None):
        from rlic_b200 import _core

        sfx, _ = self._kind(texture)
        a, b = rows if rows is not None else (0, plan.nrows)
        _core.check(getattr(_core.lib, f"rlic_b200_slab_unpad_texture_rows_{sfx}")(
            padded.data_ptr(), *plan.slab_args, a, b - a, *walls, texture.data_ptr(), self._stream()))

    def pass_rows(self, src, field, dst, plan, a, b, taps, mode, walls):
        from rlic_b200 import _core

        sfx, real = self._kind(src)
        _core.check(getattr(_core.lib, f"rlic_b200_pass_slab_{sfx}")(
            src.data_ptr(), field.data_ptr(), dst.data_ptr(), *plan.slab_args, a, b - a,
            taps.ctypes.data_as(ctypes.POINTER(real)), taps.size, mode, *walls, self._stream()))


    def pass_rows_peer(self, src, field, dst, plan, a, b, taps, mode, walls, peer, peer_row_delta):
        """The pass over owned rows [a, b) that also stores them into ``peer`` (a neighbour's
        padded buffer mapped into this process), ``peer_row_delta`` buffer rows away."""
        from rlic_b200 import _core

        sfx, real = self._kind(src)
        _core.check(getattr(_core.lib, f"rlic_b200_pass_slab_peer_{sfx}")(
            src.data_ptr(), field.data_ptr(), dst.data_ptr(), *plan.slab_args, a, b - a,
            taps.ctypes.data_as(ctypes.POINTER(real)), taps.size, mode, *walls,
            peer.data_ptr(), peer_row_delta, self._stream()))

    def pass_rows_paths(self, src, field, dst, plan, a, b, taps, mode, walls, paths_mode, paths,
                        peer=None, peer_row_delta=0):
        """``pass_rows`` / ``pass_rows_peer`` with the streamline paths of rows [a, b) recorded into
        (``_core.PASS_RECORD``) or replayed from (``_core.PASS_REPLAY``) the int32 tensor ``paths``
        (``rlic_b200_pass_slab_paths_*``)."""
        from rlic_b200 import _core

        sfx, real = self._kind(src)
        _core.check(getattr(_core.lib, f"rlic_b200_pass_slab_paths_{sfx}")(
            src.data_ptr(), field.data_ptr(), dst.data_ptr(), *plan.slab_args, a, b - a,
            taps.ctypes.data_as(ctypes.POINTER(real)), taps.size, mode, *walls,
            peer.data_ptr() if peer is not None else None, peer_row_delta, paths_mode, paths.data_ptr(),
            self._stream()))

    def signal(self, flags, index, value):
        """After everything enqueued so far: raise counter ``index`` of ``flags`` (int32
        tensor, usually a neighbour's) to ``value``."""
        from rlic_b200 import _core

        _core.check(_core.lib.rlic_b200_peer_signal(flags.data_ptr() + 4 * index, value & 0xFFFFFFFF,
                                                    self._stream()))

    def wait(self, flags, index, value, timeout_ms, timed_out_index):
        """Hold the stream until counter ``index`` of the local ``flags`` has reached ``value``;
        past ``timeout_ms`` give up and set ``flags[timed_out_index]``."""
        from rlic_b200 import _core

        _core.check(_core.lib.rlic_b200_peer_wait(flags.data_ptr() + 4 * index, value & 0xFFFFFFFF, timeout_ms,
                                                  flags.data_ptr() + 4 * timed_out_index, self._stream()))

    def signal_many(self, targets, value):
        """``signal`` for up to two counters -- ``targets``: ``(flags, index)`` pairs -- in one launch."""
        from rlic_b200 import _core

        ptrs = [f.data_ptr() + 4 * i for f, i in targets]
        for k in range(0, len(ptrs), 2):
            pair = ptrs[k:k + 2] + [None] * (2 - len(ptrs[k:k + 2]))
            _core.check(_core.lib.rlic_b200_peer_signal2(pair[0], pair[1], value & 0xFFFFFFFF, self._stream()))

    def wait_many(self, flags, waits, timeout_ms, timed_out_index):
        """``wait`` for up to four counters of the local ``flags`` -- ``waits``: ``(index, value)``
        pairs -- in one launch."""
        from rlic_b200 import _core

        for k in range(0, len(waits), 4):
            args = []
            for index, value in waits[k:k + 4]:
                args += [flags.data_ptr() + 4 * index, value & 0xFFFFFFFF]
            args += [None, 0] * (4 - len(waits[k:k + 4]))
            _core.check(_core.lib.rlic_b200_peer_wait4(*args, timeout_ms, flags.data_ptr() + 4 * timed_out_index,
                                                       self._stream()))


class _DevicePointer:
    """A raw device allocation as a ``__cuda_array_interface__`` producer (zero-copy into torch)."""

    def __init__(self, ptr: int, count: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False),
                                         "strides": None, "version": 3}


class CudaPeerMemory:
    """Device memory other ranks can map: the CUDA IPC entry points of the C ABI."""

    _TYPESTR = {torch.float32: "<f4", torch.float64: "<f8", torch.int32: "<i4"}

    def alloc(self, nbytes: int) -> tuple[int, bytes]:
        from rlic_b200 import _core

        ptr = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        _core.check(_core.lib.rlic_b200_peer_alloc(nbytes, ctypes.byref(ptr), handle))
        return int(ptr.value), bytes(handle)

    def open(self, handle: bytes) -> int:
        from rlic_b200 import _core

        ptr = ctypes.c_void_p()
        raw = (ctypes.c_ubyte * 64).from_buffer_copy(handle)
        _core.check(_core.lib.rlic_b200_peer_open(raw, ctypes.byref(ptr)))
        return int(ptr.value)

    def close(self, ptr: int) -> None:
        from rlic_b200 import _core

        _core.check(_core.lib.rlic_b200_peer_close(ctypes.c_void_p(ptr)))

    def free(self, ptr: int) -> None:
        from rlic_b200 import _core

        _core.check(_core.lib.rlic_b200_peer_free(ctypes.c_void_p(ptr)))

    def view(self, ptr: int, count: int, dtype, device) -> torch.Tensor:
        return torch.as_tensor(_DevicePointer(ptr, count, self._TYPESTR[dtype]), device=device)


def pinned_empty(shape, dtype) -> np.ndarray:
    """A page-locked NumPy array (what ``convolve_host`` copies asynchronously from and to)."""
    t = torch.empty(tuple(shape), dtype=torch.from_numpy(np.empty(0, dtype=dtype)).dtype, pin_memory=True)
    return t.numpy()


class PeerMemoryUnavailable(RuntimeError):
    """``exchange="peer"`` cannot be set up: some rank could not allocate its shareable buffers or
    map a neighbour's (no peer access between the GPUs, CUDA IPC not permitted in this
    container, ...).  Raised on EVERY rank of the group -- the ranks agree on the outcome before
    anyone proceeds, and whatever was allocated or mapped has been released -- so a caller can
    catch it and build its convolver with ``exchange="nccl"`` instead, consistently."""


# counters in a rank's flag block, raised by its neighbours (int32 each)
_HALO_FROM_UP, _HALO_FROM_DOWN, _FREE_FROM_UP, _FREE_FROM_DOWN, _TIMED_OUT, _NFLAGS = 0, 1, 2, 3, 4, 8


class _PeerExchange:
    """One rank's side of the fused halo exchange: its two padded texture buffers, its packed
    field and its flag block in memory the neighbours can map, and the neighbours' mapped
    into this process."""

    def __init__(self, plan: SlabPlan, group, peers, dtype, device):
        self.plan, self.peers = plan, peers
        itemsize = torch.empty((), dtype=dtype).element_size()
        self._own = []                                   # (ptr, handle) of this rank's allocations
        self._opened = {}                                # rank -> [ptr, ...]
        self.remote = {}                                 # rank -> (bufs, flags, that rank's plan, field)
        self.bufs, self.flags, self.field = [], None, None

        def agree(error, payload=None):
            """Every rank learns whether every rank got this far (a failure must not leave the
            others waiting in a collective); returns the payloads, or releases everything and
            raises PeerMemoryUnavailable on all ranks alike."""
            said = [None] * plan.world
            dist.all_gather_object(said, (error, payload), group=group)
            errors = [e for e, _ in said if e]
            if errors:
                try:
                    self._release(group)
                except Exception as exc:  # noqa: BLE001 -- every rank must leave with the same exception
                    errors.append(f"(while releasing on rank {plan.rank}: {type(exc).__name__}: {exc})")
                raise PeerMemoryUnavailable("; ".join(errors))
            return [x for _, x in said]

        def describe(exc):
            return f"rank {plan.rank}: {type(exc).__name__}: {exc}"

        error = None
        try:
            for nbytes in (plan.cells * itemsize, plan.cells * itemsize, 4 * _NFLAGS, 4 * plan.cells * itemsize):
                self._own.append(peers.alloc(nbytes))
            self.bufs = [peers.view(self._own[i][0], plan.cells, dtype, device) for i in (0, 1)]
            self.flags = peers.view(self._own[2][0], _NFLAGS, torch.int32, device)
            self.field = peers.view(self._own[3][0], 4 * plan.cells, dtype, device)   # 4 scalars per cell
        except Exception as exc:  # noqa: BLE001 -- reported to every rank, then raised on all of them
            error = describe(exc)
        handles = agree(error, [h for _, h in self._own])
        try:
            for rank in {plan.up, plan.down} - {None}:
                ptrs = self._opened.setdefault(rank, [])
                for h in handles[rank]:
                    ptrs.append(peers.open(h))
                theirs = SlabPlan(ny=plan.ny, nx=plan.nx, world=plan.world, rank=rank, reach=plan.reach,
                                  periodic_y=plan.periodic_y)
                self.remote[rank] = ([peers.view(ptrs[i], theirs.cells, dtype, device) for i in (0, 1)],
                                     peers.view(ptrs[2], _NFLAGS, torch.int32, device), theirs,
                                     peers.view(ptrs[3], 4 * theirs.cells, dtype, device))
        except Exception as exc:  # noqa: BLE001
            error = describe(exc)
        agree(error)
        # where my edge strips land: my top rows in the upper neighbour's high halo, my
        # bottom rows in the lower neighbour's low halo (buffer rows, theirs minus mine)
        self.delta_up = self.delta_down = 0
        if plan.up is not None:
            theirs = self.remote[plan.up][2]
            self.delta_up = theirs.halo_lo + theirs.nrows - plan.halo_lo
        if plan.down is not None:
            self.delta_down = -(plan.halo_lo + plan.nrows - plan.reach)
        self.tick = 0                                    # passes announced so far (all ranks alike)
        dist.barrier(group=group)                        # every flag block exists and is zero

    def close(self, group) -> None:
        dist.barrier(group=group)                        # nobody still writes into a neighbour
        self._release(group)

    def _release(self, group) -> None:
        """Unmap the neighbours' buffers, then -- once every rank has -- free this rank's own
        (collective; also the way out of a set-up that failed on some rank)."""
        failed = None
        self.remote = {}
        for ptrs in self._opened.values():
            for ptr in ptrs:
                try:
                    self.peers.close(ptr)
                except Exception as exc:  # noqa: BLE001 -- the barrier below must still be reached
                    failed = failed or exc
        self._opened = {}
        dist.barrier(group=group)                        # every mapping is gone before memory is freed
        self.bufs, self.flags, self.field = [], None, None
        for ptr, _ in self._own:
            try:
                self.peers.free(ptr)
            except Exception as exc:  # noqa: BLE001
                failed = failed or exc
        self._own = []
        if failed is not None:
            raise failed


class ShardedConvolver:
    """Holds one rank's slab of a sharded image and runs ``convolve`` on it.

    ``texture``, ``u``, ``v``: this rank's rows ``[plan.row0, plan.row1)`` (dense,
    no halos), tensors on the rank's device.
    """

    PEER_TIMEOUT_MS = 20_000      # a wait for a neighbour gives up after this long (and says so)
    # What happens to the "a wait for a neighbour gave up" flag of exchange="peer" (a result
    # computed from stale halos must never pass for a good one):
    #   "lazy"  (default) every call ends with an asynchronous 4-byte copy of the flag to
    #           page-locked memory; the NEXT call, `synchronize()` and `close()` look at the
    #           copies that have landed and raise.  No host synchronisation: the host keeps
    #           running ahead of the device, which is what hides its launch latencies
    #           (measured on 8 GPUs: a blocking read per call cost 1-2 ms of a 7.6 ms call,
    #           profiles/r2_peer_diag_*.json).  Host entry points (`convolve_host`) synchronise
    #           anyway and check before they return.
    #   True    read the flag back before the call returns (one synchronisation per call)
    #   False   never look
    check_peer_timeouts = "lazy"

    def __init__(self, ny: int, nx: int, *, kernel, uv_mode: str = "velocity",
                 boundaries="closed", group=None, ops=None, exchange: str = "nccl", peers=None,
                 paths: str | None = None):
        from rlic_b200 import _core   # enum tables only; no computation

        # what the passes of a call after its first do (include/rlic_b200.h, RLIC_B200_PATHS_*):
        # None = the library's choice for this thread at the time of each call
        if paths is not None and paths not in _core.PATHS:
            raise ValueError(f"unknown paths choice {paths!r}: expected one of {sorted(_core.PATHS)}")
        self.paths = paths
        self._record = None          # int32 tensor: the recorded paths of this slab (of self.field)
        self._call_paths = False     # whether the call in progress records / replays

        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        bs = BoundarySet.from_spec(boundaries)
        if bs is None:
            raise TypeError(f"Invalid boundary specification {boundaries}")
        bs.validate()
        self.walls = _core.wall_codes((bs.x, bs.y))
        self.mode = _core.mode_code(uv_mode)
        self.taps = np.ascontiguousarray(kernel)
        if self.taps.ndim != 1 or self.taps.size == 0:
            raise ValueError("kernel must be a non-empty 1-D array")
        self.plan = SlabPlan(ny=ny, nx=nx, world=self.world, rank=self.rank,
                             reach=self.taps.size // 2, periodic_y=bs.y[0] == "periodic")
        self.plan.validate()
        self.ops = ops or CudaSlabOps()
        self.field = None
        self.comm_stream = None
        if exchange not in ("nccl", "peer"):
            raise ValueError(f"unknown exchange {exchange!r}: expected 'nccl' or 'peer'")
        if exchange == "peer" and self.world > 1:
            smallest = min(ny * (r + 1) // self.world - ny * r // self.world for r in range(self.world))
            if smallest < 4 * max(self.plan.reach, 1):
                raise ValueError(
                    f"exchange='peer' needs slabs of at least four kernel half-widths "
                    f"({4 * max(self.plan.reach, 1)} rows); the smallest here has {smallest}")
        self.exchange = exchange if self.world > 1 else "nccl"
        self._peer_memory = peers
        self._peer = None            # _PeerExchange, built on the first convolve (needs dtype and device)
        self._work = None            # ((dtype, device), padded buffer, padded buffer) of the NCCL path
        self.trace = None            # diagnostics: list of (label, CUDA event), see _mark
        self._timeout_probes = []    # (pinned copy of the time-out flag, event) of earlier calls
        self._staging = None         # ((dtype, device), dense texture, u, v) of convolve_host
        self._host_streams = None    # (upload stream, download stream) of convolve_host

    # -- halo plumbing ------------------------------------------------------
    @staticmethod
    def field_layout(dtype) -> tuple[int, int]:
        """(scalars per cell, planes) of a packed field buffer: f32 records are
        interleaved (4 scalars per cell); f64 records are split into a {u, v}
        and a {ru, rv} plane of 2 scalars per cell (include/rlic_b200.h)."""
        return (4, 1) if dtype == torch.float32 else (2, 2)

    def _exchange(self, buf: torch.Tensor, width: int = 1, planes: int = 1) -> None:
        """Fill the halo rows of the flat padded buffer ``buf`` (``planes`` planes of
        ``width`` scalars per cell) from the neighbours.

        Every rank posts: send top rows up, send bottom rows down, receive the
        high halo from below, receive the low halo from above.  That order
        makes the k-th message between any pair match on both sides, including
        the two-rank ring where both neighbours are the same peer.
        """
        p = self.plan
        h, lo, n = p.reach, p.halo_lo, p.nrows

        def rows(plane, a, b):
            s = p.row_cells(a, b)
            base = plane * p.cells
            return buf[(base + s.start) * width:(base + s.stop) * width]

        ops = []
        for plane in range(planes):
            if p.up is not None:
                ops.append(dist.P2POp(dist.isend, rows(plane, lo, lo + h), p.up, self.group))
            if p.down is not None:
                ops.append(dist.P2POp(dist.isend, rows(plane, lo + n - h, lo + n), p.down, self.group))
            if p.down is not None:
                ops.append(dist.P2POp(dist.irecv, rows(plane, lo + n, lo + n + h), p.down, self.group))
            if p.up is not None:
                ops.append(dist.P2POp(dist.irecv, rows(plane, 0, lo), p.up, self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()

    def _alloc(self, like: torch.Tensor, width: int = 1) -> torch.Tensor:
        # zero-filled: wall cells nobody can reach are never written
        return torch.zeros(self.plan.cells * width, dtype=like.dtype, device=like.device)

    def set_field(self, u: torch.Tensor, v: torch.Tensor) -> None:
        """Pack (u, v) with halos; done once, the field does not change between passes."""
        p = self.plan
        if tuple(u.shape) != (p.nrows, p.nx) or tuple(v.shape) != (p.nrows, p.nx):
            raise ValueError(f"expected this rank's slab of shape {(p.nrows, p.nx)}")
        width, planes = self.field_layout(u.dtype)
        if self.exchange == "peer":
            # in memory the neighbours can map: convolve_host pushes its halos there
            field = self._peer_exchange(u.dtype, u.device).field
        elif (self.field is not None and self.field.dtype == u.dtype and self.field.device == u.device
              and self.field.numel() == p.cells * width * planes):
            field = self.field       # every reachable cell is rewritten by the pack and the exchange
        else:
            field = self._alloc(u, width * planes)
        self.ops.pack_field(u.contiguous(), v.contiguous(), field, p, self.walls)
        self._exchange(field, width, planes)
        self.field = field

    def _peer_exchange(self, dtype, device) -> "_PeerExchange":
        if self._peer is None:
            self._peer = _PeerExchange(self.plan, self.group, self._peer_memory or CudaPeerMemory(),
                                       dtype, device)
        if self._peer.bufs[0].dtype != dtype:
            raise TypeError("the peer buffers of this convolver hold another dtype")
        return self._peer

    def _begin_call(self, iterations: int, like: torch.Tensor) -> None:
        """Decide whether this call records the streamline paths in its first pass and replays them
        in the others (they depend on the field, never on the texture: lib.rs:305-362), and make
        room for the record.  Same rule as the single-GPU entry points: more than one pass, the
        option, the default arithmetic and walk."""
        from rlic_b200 import _core

        want = self.paths
        if want is None:
            opts = _core.effective_options() if hasattr(self.ops, "pass_rows_paths") else {}
            want = opts.get("paths") if (opts.get("arithmetic") == "fma+branchless"
                                         and opts.get("walk") == "grouped") else "recompute"
        self._call_paths = bool(iterations >= 2 and want == "replay" and hasattr(self.ops, "pass_rows_paths"))
        if self._call_paths:
            words = _core.path_record_bytes(self.plan.rows_alloc, self.plan.nx, int(self.taps.size)) // 4
            if (self._record is None or self._record.device != like.device or self._record.numel() != words):
                self._record = torch.empty(words, dtype=torch.int32, device=like.device)

    def _pass_rows(self, src, dst, a: int, b: int, k: int = 0, peer=None, peer_row_delta: int = 0) -> None:
        """Compute owned rows [a, b) of ``dst`` from ``src`` in pass ``k`` (1-based; 0: a pass outside
        a call that records), optionally storing them into the neighbour's buffer ``peer`` too."""
        from rlic_b200 import _core

        if b <= a:
            return
        if self._call_paths and k >= 1:
            self.ops.pass_rows_paths(src, self.field, dst, self.plan, a, b, self.taps, self.mode, self.walls,
                                     _core.PASS_RECORD if k == 1 else _core.PASS_REPLAY, self._record,
                                     peer, peer_row_delta)
        elif peer is not None:
            self.ops.pass_rows_peer(src, self.field, dst, self.plan, a, b, self.taps, self.mode, self.walls,
                                    peer, peer_row_delta)
        else:
            self.ops.pass_rows(src, self.field, dst, self.plan, a, b, self.taps, self.mode, self.walls)

    def convolve(self, texture: torch.Tensor, iterations: int = 1, overlap: bool = True) -> torch.Tensor:
        """Run ``iterations`` passes; returns this rank's rows of the result."""
        if self.field is None:
            raise RuntimeError("call set_field(u, v) first")
        p = self.plan
        if tuple(texture.shape) != (p.nrows, p.nx):
            raise ValueError(f"expected this rank's slab of shape {(p.nrows, p.nx)}")
        if iterations <= 0:
            return texture.clone()
        if self.exchange == "peer":
            return self._convolve_peer(texture, iterations)
        # the two padded work buffers live as long as the convolver (zero-filled once: wall
        # cells nobody can reach are never written, every other cell is rewritten by each call)
        key = (texture.dtype, texture.device)
        if self._work is None or self._work[0] != key:
            self._work = (key, self._alloc(texture), self._alloc(texture))
        src, dst = self._work[1], self._work[2]
        self._begin_call(iterations, texture)
        self.ops.pad_texture(texture.contiguous(), src, p, self.walls)
        self._exchange(src)
        h = p.reach
        # strips first, interior while the strips travel: worth it only when
        # there is an interior left and someone to talk to
        split = overlap and p.world > 1 and p.nrows >= 4 * max(h, 1)
        side = texture.is_cuda and split
        if side and self.comm_stream is None:
            self.comm_stream = torch.cuda.Stream(device=texture.device)
        for it in range(iterations):
            last = it == iterations - 1
            k = it + 1
            if not split or last:
                self._pass_rows(src, dst, 0, p.nrows, k)
                if not last:
                    self._exchange(dst)
            elif not side:           # same decomposition without streams (CPU tests)
                self._pass_rows(src, dst, 0, h, k)
                self._pass_rows(src, dst, p.nrows - h, p.nrows, k)
                self._exchange(dst)
                self._pass_rows(src, dst, h, p.nrows - h, k)
            else:
                main = torch.cuda.current_stream()
                # 1. the strips the neighbours are waiting for
                self._pass_rows(src, dst, 0, h, k)
                self._pass_rows(src, dst, p.nrows - h, p.nrows, k)
                strips_done = torch.cuda.Event()
                strips_done.record(main)
                # 2. ship them while 3. the interior is computed
                self.comm_stream.wait_event(strips_done)
                with torch.cuda.stream(self.comm_stream):
                    self._exchange(dst)
                    shipped = torch.cuda.Event()
                    shipped.record(self.comm_stream)
                self._pass_rows(src, dst, h, p.nrows - h, k)
                main.wait_event(shipped)
            src, dst = dst, src
        out = torch.empty_like(texture)
        self.ops.unpad_texture(src, out, p, self.walls)
        return out

    # -- fused halo exchange --------------------------------------------------
    def _convolve_peer(self, texture: torch.Tensor, iterations: int) -> torch.Tensor:
        """``convolve`` with ``exchange="peer"``: see the module docstring.  No message-passing
        library is involved once the buffers are mapped: everything below is kernels and
        device-to-device copies on the current stream, ordered across ranks by two counters
        per neighbour that live in the consumer's memory.

        One number line serves both counters.  A call that starts at count ``c`` and runs ``n``
        passes (pass ``k`` reads X[k-1] in ``bufs[(k-1) % 2]`` and writes X[k] into
        ``bufs[k % 2]``; X[0] is the padded texture) uses ``c + 1 .. c + n + 1``:

            halo = c + k + 1   "the halo rows of X[k] are in your buffer"  (k = 0 .. n-1;
                               k = 0 is pushed by a copy, the others by the fused strips)
            free = c + k       "my pass k is over" (k = 1 .. n), and free = c + n + 1
                               "my un-padding is over: this call no longer reads anything"

        X[k]'s halos land in the neighbour's ``bufs[k % 2]``, which it last read during its
        pass k - 1 (k >= 2) or during its previous call (k = 0, 1): they may be stored once
        ``free >= c + max(k - 1, 0)`` -- the previous call ended at exactly ``c``.  Pass k
        needs the halos of X[k-1]: ``halo >= c + k``.  Strips go first, so in the steady
        state every wait finds its counter already raised."""
        p = self.plan
        px = self._peer_exchange(texture.dtype, texture.device)
        self._poll_timeout_probes()
        self._begin_call(iterations, texture)
        self.ops.pad_texture(texture.contiguous(), px.bufs[0], p, self.walls)
        self._peer_passes(px, iterations)
        out = torch.empty_like(texture)
        self.ops.unpad_texture(px.bufs[iterations % 2], out, p, self.walls)
        self._peer_finish(px, iterations)
        return out

    def _mark(self, label: str) -> None:
        """Diagnostics (tools/peer_diag.py): with ``self.trace`` a list, record a timing event on
        the current stream after the operation called ``label``."""
        if self.trace is not None and torch.cuda.is_available():
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.trace.append((label, ev))

    def _peer_signals(self, px):
        p, flags, limit = self.plan, px.flags, self.PEER_TIMEOUT_MS
        up = px.remote.get(p.up) if p.up is not None else None
        down = px.remote.get(p.down) if p.down is not None else None

        def wait(which_up, which_down, value, and_up=None, and_down=None, and_value=0):
            """hold the stream until my counters `which_up` / `which_down` have reached `value` (and,
            when given, `and_up` / `and_down` have reached `and_value`): one launch where the ops can"""
            waits = []
            if up is not None:
                waits.append((which_up, value))
                if and_up is not None:
                    waits.append((and_up, and_value))
            if down is not None:
                waits.append((which_down, value))
                if and_down is not None:
                    waits.append((and_down, and_value))
            if not waits:
                return
            if hasattr(self.ops, "wait_many"):
                self.ops.wait_many(flags, waits, limit, _TIMED_OUT)
            else:
                for index, v in waits:
                    self.ops.wait(flags, index, v, limit, _TIMED_OUT)

        def signal(which_at_up, which_at_down, value):
            # my upper neighbour sees me below it, my lower neighbour sees me above it
            targets = []
            if up is not None:
                targets.append((up[1], which_at_up))
            if down is not None:
                targets.append((down[1], which_at_down))
            if not targets:
                return
            if hasattr(self.ops, "signal_many"):
                self.ops.signal_many(targets, value)
            else:
                for f, index in targets:
                    self.ops.signal(f, index, value)

        return up, down, wait, signal

    def _peer_passes(self, px, n: int, *, push_field: bool = False, rows_runner=None) -> None:
        """The passes of one call under the counter protocol (see ``_convolve_peer``), on the
        current stream.  On entry the edge rows of X[0] in ``px.bufs[0]`` (and, with
        ``push_field``, of the packed field) are ready in stream order; on return X[n] is in
        ``px.bufs[n % 2]``.  ``rows_runner(k, src, dst, a, b)`` computes owned rows [a, b) of
        pass k (default: one launch); the host path cuts them into bands there."""
        p, h = self.plan, self.plan.reach
        bufs = px.bufs
        up, down, wait, signal = self._peer_signals(px)
        c = px.tick
        lo, rows = p.halo_lo, p.nrows
        if rows_runner is None:
            def rows_runner(k, src, dst, a, b):
                self._pass_rows(src, dst, a, b, k)

        def push(theirs, mine, delta_rows, a, b, width=1, planes=1, their_cells=0):
            """buffer rows [a, b) of `mine`, with their wall cells, into `theirs` delta_rows away"""
            cells = p.row_cells(a, b)
            shift = delta_rows * p.pitch
            for plane in range(planes):
                s0, d0 = plane * p.cells, plane * their_cells
                theirs[(d0 + cells.start + shift) * width:(d0 + cells.stop + shift) * width].copy_(
                    mine[(s0 + cells.start) * width:(s0 + cells.stop) * width])

        # X[0]'s halos: my edge rows copied into the neighbours' bufs[0] (and field)
        self._mark("begin")
        wait(_FREE_FROM_UP, _FREE_FROM_DOWN, c)
        self._mark("wait free (previous call)")
        width, planes = self.field_layout(bufs[0].dtype)
        if up is not None:
            push(up[0][0], bufs[0], px.delta_up, lo, lo + h)
            if push_field:
                push(up[3], px.field, px.delta_up, lo, lo + h, width, planes, up[2].cells)
        if down is not None:
            push(down[0][0], bufs[0], px.delta_down, lo + rows - h, lo + rows)
            if push_field:
                push(down[3], px.field, px.delta_down, lo + rows - h, lo + rows, width, planes, down[2].cells)
        signal(_HALO_FROM_DOWN, _HALO_FROM_UP, c + 1)
        self._mark("push initial halos")
        for k in range(1, n + 1):
            src, dst = bufs[(k - 1) % 2], bufs[k % 2]
            if k == n or k < 2:
                wait(_HALO_FROM_UP, _HALO_FROM_DOWN, c + k)
            else:   # the halos of this pass's input, and room for this pass's strips: one launch
                wait(_HALO_FROM_UP, _HALO_FROM_DOWN, c + k, _FREE_FROM_UP, _FREE_FROM_DOWN, c + k - 1)
            self._mark("wait halo")
            if k == n:
                rows_runner(k, src, dst, 0, rows)
                self._mark("last pass")
                break
            # the edge strips, stored into the neighbours' halos as they are computed
            if up is not None:
                self._pass_rows(src, dst, 0, h, k, up[0][k % 2], px.delta_up)
            else:
                self._pass_rows(src, dst, 0, h, k)
            if down is not None:
                self._pass_rows(src, dst, rows - h, rows, k, down[0][k % 2], px.delta_down)
            else:
                self._pass_rows(src, dst, rows - h, rows, k)
            signal(_HALO_FROM_DOWN, _HALO_FROM_UP, c + k + 1)
            self._mark("strips + signal")
            rows_runner(k, src, dst, h, rows - h)
            signal(_FREE_FROM_DOWN, _FREE_FROM_UP, c + k)
            self._mark("interior + signal")

    def _peer_finish(self, px, n: int, blocking: bool = False) -> None:
        """After the last read of this call's buffers (the un-padding) has been enqueued on the
        current stream: tell the neighbours, advance the count, and see to it that a result
        computed from stale halos is reported (`check_peer_timeouts`); `blocking`: the caller
        synchronises anyway, so the flag can be read on the spot."""
        _, _, _, signal = self._peer_signals(px)
        signal(_FREE_FROM_DOWN, _FREE_FROM_UP, px.tick + n + 1)
        px.tick += n + 1
        if self.check_peer_timeouts is True or (blocking and self.check_peer_timeouts):
            self._raise_if_timed_out(bool(int(px.flags[_TIMED_OUT].item()) != 0))
        elif self.check_peer_timeouts == "lazy":
            if px.flags.is_cuda:
                probe = torch.empty(1, dtype=torch.int32, pin_memory=True)
                probe.copy_(px.flags[_TIMED_OUT:_TIMED_OUT + 1], non_blocking=True)
                landed = torch.cuda.Event()
                landed.record()
                self._timeout_probes.append((probe, landed))
            else:
                self._raise_if_timed_out(bool(int(px.flags[_TIMED_OUT]) != 0))

    def _raise_if_timed_out(self, timed_out: bool) -> None:
        if timed_out:
            self._peer.flags[_TIMED_OUT] = 0
            self._timeout_probes.clear()
            raise RuntimeError(
                f"rank {self.plan.rank}: a neighbour did not answer within {self.PEER_TIMEOUT_MS} ms during "
                "the fused halo exchange; results computed since the last check are invalid")

    def _poll_timeout_probes(self, wait: bool = False) -> None:
        """Look at the flag copies of earlier calls that have landed (all of them with `wait`)."""
        keep, bad = [], False
        for probe, landed in self._timeout_probes:
            if wait:
                landed.synchronize()
            if landed.query():
                bad |= bool(int(probe[0]) != 0)
            else:
                keep.append((probe, landed))
        self._timeout_probes = keep
        self._raise_if_timed_out(bad)

    def synchronize(self) -> None:
        """Wait for everything this convolver enqueued and raise if a wait for a neighbour gave up."""
        if self._peer is not None and self._peer.flags.is_cuda:
            torch.cuda.current_stream(self._peer.flags.device).synchronize()
        self._poll_timeout_probes(wait=True)

    # -- host slabs in, host slabs out ------------------------------------------
    def convolve_host(self, texture, u=None, v=None, *, iterations: int = 1, out=None, device=None,
                      max_bands: int = 8, min_band_pixels: int = 1 << 21, min_band_rows: int = 64):
        """This rank's slab from HOST memory to host memory: ``texture`` (and, when given, ``u``
        and ``v``: the field is re-packed; otherwise the one ``set_field`` left is used) are
        NumPy arrays of this rank's rows, ``out`` a NumPy array to fill (allocated page-locked
        if omitted).  Page-locked inputs (``pinned_empty``) are copied asynchronously.

        With ``exchange="peer"`` the call is a pipeline over row bands, like the single-GPU
        host path (``convolve_host`` in lic_api.cu): an upload stream brings the bands in --
        the two edge bands first, so that the halos can leave for the neighbours at once --
        and converts each to the padded layout; pass 1 trails behind it band by band; the last
        pass hands each band to a download stream as soon as it is done.  Halo rows of the
        texture AND of the re-packed field travel by peer copies under the same counters as
        the device path.  Otherwise (NCCL exchange, one rank) the three steps run one after
        the other."""
        import contextlib

        p, h = self.plan, self.plan.reach
        t_tex = torch.from_numpy(np.ascontiguousarray(texture))
        if tuple(t_tex.shape) != (p.nrows, p.nx):
            raise ValueError(f"expected this rank's slab of shape {(p.nrows, p.nx)}")
        if (u is None) != (v is None):
            raise ValueError("pass both u and v, or neither")
        t_u = torch.from_numpy(np.ascontiguousarray(u)) if u is not None else None
        t_v = torch.from_numpy(np.ascontiguousarray(v)) if v is not None else None
        if t_u is not None and (t_u.shape != t_tex.shape or t_v.shape != t_tex.shape or t_u.dtype != t_tex.dtype
                                or t_v.dtype != t_tex.dtype):
            raise ValueError("u and v must have the texture's shape and dtype")
        if out is None:
            out = pinned_empty(t_tex.shape, texture.dtype) if torch.cuda.is_available() else np.empty_like(texture)
        t_out = torch.from_numpy(out)
        if t_out.shape != t_tex.shape or t_out.dtype != t_tex.dtype or not t_out.is_contiguous():
            raise ValueError("out must be a C-contiguous array of the texture's shape and dtype")
        if iterations <= 0:
            t_out.copy_(t_tex)
            return out
        if device is None:
            device = (torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available()
                      else torch.device("cpu"))
        cuda = device.type == "cuda"
        if t_u is None and self.field is None:
            raise RuntimeError("call set_field(u, v) first, or pass u and v")

        key = (t_tex.dtype, device)
        if self._staging is None or self._staging[0] != key:
            dense = lambda: torch.empty(t_tex.shape, dtype=t_tex.dtype, device=device)  # noqa: E731
            self._staging = (key, dense(), dense(), dense())
        _, s_t, s_u, s_v = self._staging

        if self.exchange != "peer" or p.world == 1:
            s_t.copy_(t_tex, non_blocking=True)
            if t_u is not None:
                s_u.copy_(t_u, non_blocking=True)
                s_v.copy_(t_v, non_blocking=True)
                self.set_field(s_u, s_v)
            res = self.convolve(s_t, iterations=iterations)
            t_out.copy_(res, non_blocking=True)
            if cuda:
                torch.cuda.current_stream(device).synchronize()
            return out

        px = self._peer_exchange(t_tex.dtype, device)
        self._poll_timeout_probes()
        self._begin_call(iterations, s_t)
        if t_u is not None:
            self.field = px.field
        n, rows = iterations, p.nrows
        # row bands: multiples of the tile height, at least two kernel half-widths tall
        # (the last band takes the remainder, so no band is shorter than that)
        nb = max(1, min(max_bands, rows * p.nx // max(min_band_pixels, 1), rows // max(2 * h, min_band_rows, 16)))
        band_rows = -(-(-(-rows // nb)) // 16) * 16
        band_rows = max(band_rows, 2 * h, 16)
        nb = max(1, rows // band_rows)
        edge = [b * band_rows for b in range(nb)] + [rows]
        if cuda:
            main = torch.cuda.current_stream(device)
            if self._host_streams is None:
                self._host_streams = (torch.cuda.Stream(device=device), torch.cuda.Stream(device=device))
            io, back = self._host_streams
            on = torch.cuda.stream
            # the previous call's passes and downloads are done with what the uploads overwrite
            io.wait_stream(main)
            io.wait_stream(back)
        else:
            main = io = back = None
            on = lambda _stream: contextlib.nullcontext()  # noqa: E731

        def record(stream):
            if not cuda:
                return None
            ev = torch.cuda.Event()
            ev.record(stream)
            return ev

        def after(stream, ev):
            if cuda and ev is not None:
                stream.wait_event(ev)

        # ---- uploads: the edge bands first (their rows are the neighbours' halos)
        uploaded = {}
        with on(io):
            for b in dict.fromkeys([0, nb - 1, *range(1, nb - 1)]):
                a, e = edge[b], edge[b + 1]
                if t_u is not None:
                    s_u[a:e].copy_(t_u[a:e], non_blocking=True)
                    s_v[a:e].copy_(t_v[a:e], non_blocking=True)
                    self.ops.pack_field(s_u[a:e], s_v[a:e], px.field, p, self.walls, rows=(a, e))
                s_t[a:e].copy_(t_tex[a:e], non_blocking=True)
                self.ops.pad_texture(s_t[a:e], px.bufs[0], p, self.walls, rows=(a, e))
                uploaded[b] = record(io)

        # ---- passes: pass 1 trails the uploads, the last pass releases bands to the download
        done = {}

        def rows_runner(k, src, dst, a, e):
            if k != 1 and k != n:
                self._pass_rows(src, dst, a, e, k)
                return
            for b in range(nb):
                lo_, hi_ = max(a, edge[b]), min(e, edge[b + 1])
                if k == 1:      # rows a walker of this band can reach must have arrived
                    for nbr in (b - 1, b, b + 1):
                        if 0 <= nbr < nb:
                            after(main, uploaded[nbr])
                if hi_ > lo_:
                    self._pass_rows(src, dst, lo_, hi_, k)
                if k == n:
                    done[b] = record(main)

        after(main, uploaded[0])
        after(main, uploaded[nb - 1])
        self._peer_passes(px, n, push_field=t_u is not None, rows_runner=rows_runner)

        # ---- downloads, band by band behind the last pass
        with on(back):
            for b in range(nb):
                a, e = edge[b], edge[b + 1]
                after(back, done[b])
                self.ops.unpad_texture(px.bufs[n % 2], s_t[a:e], p, self.walls, rows=(a, e))
                t_out[a:e].copy_(s_t[a:e], non_blocking=True)
            finished = record(back)
        # the counters are raised on the stream that made the last read of this call's buffers
        after(main, finished)
        if cuda:
            back.synchronize()
        self._peer_finish(px, n, blocking=True)
        return out

    def peer_timed_out(self) -> bool:
        """Whether any wait for a neighbour gave up (synchronises with the device)."""
        return bool(self._peer is not None and int(self._peer.flags[_TIMED_OUT].item()) != 0)

    def close(self) -> None:
        """Release the peer mappings and buffers (collective; only needed with ``exchange="peer"``).
        Raises if a wait for a neighbour gave up since the last check."""
        if self._peer is not None:
            try:
                if self.check_peer_timeouts:
                    self.synchronize()
            finally:
                self._peer.close(self.group)
                self._peer = None
