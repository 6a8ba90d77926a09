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

``exchange="peer"`` (opt-in; written after round 1's GPU time was spent, so the NCCL
exchange stays the default until it has run on hardware) fuses the per-iteration
texture exchange into the passes over the edge strips: the pass kernel stores every
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

__all__ = ["SlabPlan", "ShardedConvolver", "CudaSlabOps", "CudaPeerMemory"]

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

    def pack_field(self, u, v, field, plan, walls):
        from rlic_b200 import _core

        self._need_cuda(u)
        sfx, _ = self._kind(u)
        _core.check(getattr(_core.lib, f"rlic_b200_slab_pack_field_{sfx}")(
            u.data_ptr(), v.data_ptr(), *plan.slab_args, *walls, field.data_ptr(), self._stream()))

    def pad_texture(self, texture, padded, plan, walls):
        from rlic_b200 import _core

        self._need_cuda(texture)
        sfx, _ = self._kind(texture)
        _core.check(getattr(_core.lib, f"rlic_b200_slab_pad_texture_{sfx}")(
            texture.data_ptr(), *plan.slab_args, *walls, padded.data_ptr(), self._stream()))

    def unpad_texture(self, padded, texture, plan, walls):
        from rlic_b200 import _core

        sfx, _ = self._kind(texture)
        _core.check(getattr(_core.lib, f"rlic_b200_slab_unpad_texture_{sfx}")(
            padded.data_ptr(), *plan.slab_args, *walls, texture.data_ptr(), self._stream()))

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


# counters in a rank's flag block, raised by its neighbours (int32 each)
_HALO_FROM_UP, _HALO_FROM_DOWN, _FREE_FROM_UP, _FREE_FROM_DOWN, _TIMED_OUT, _NFLAGS = 0, 1, 2, 3, 4, 8


class _PeerExchange:
    """One rank's side of the fused halo exchange: its two padded texture buffers and its flag
    block in memory the neighbours can map, and the neighbours' mapped into this process."""

    def __init__(self, plan: SlabPlan, group, peers, dtype, device):
        self.plan, self.peers = plan, peers
        itemsize = torch.empty((), dtype=dtype).element_size()
        self._own = []                                   # (ptr, handle) of this rank's allocations
        for nbytes in (plan.cells * itemsize, plan.cells * itemsize, 4 * _NFLAGS):
            self._own.append(peers.alloc(nbytes))
        self.bufs = [peers.view(self._own[i][0], plan.cells, dtype, device) for i in (0, 1)]
        self.flags = peers.view(self._own[2][0], _NFLAGS, torch.int32, device)
        handles = [None] * plan.world
        dist.all_gather_object(handles, [h for _, h in self._own], group=group)
        self._opened = {}                                # rank -> [ptr, ptr, ptr]
        self.remote = {}                                 # rank -> (bufs, flags, that rank's plan)
        for rank in {plan.up, plan.down} - {None}:
            ptrs = [peers.open(h) for h in handles[rank]]
            self._opened[rank] = ptrs
            theirs = SlabPlan(ny=plan.ny, nx=plan.nx, world=plan.world, rank=rank, reach=plan.reach,
                              periodic_y=plan.periodic_y)
            self.remote[rank] = ([peers.view(ptrs[i], theirs.cells, dtype, device) for i in (0, 1)],
                                 peers.view(ptrs[2], _NFLAGS, torch.int32, device), theirs)
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
        self.remote = {}
        for ptrs in self._opened.values():
            for ptr in ptrs:
                self.peers.close(ptr)
        self._opened = {}
        dist.barrier(group=group)                        # every mapping is gone before memory is freed
        self.bufs, self.flags = [], None
        for ptr, _ in self._own:
            self.peers.free(ptr)
        self._own = []


class ShardedConvolver:
    """Holds one rank's slab of a sharded image and runs ``convolve`` on it.

    ``texture``, ``u``, ``v``: this rank's rows ``[plan.row0, plan.row1)`` (dense,
    no halos), tensors on the rank's device.
    """

    PEER_TIMEOUT_MS = 20_000      # a wait for a neighbour gives up after this long (and says so)

    def __init__(self, ny: int, nx: int, *, kernel, uv_mode: str = "velocity",
                 boundaries="closed", group=None, ops=None, exchange: str = "nccl", peers=None):
        from rlic_b200 import _core   # enum tables only; no computation

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
        field = self._alloc(u, width * planes)
        self.ops.pack_field(u.contiguous(), v.contiguous(), field, p, self.walls)
        self._exchange(field, width, planes)
        self.field = field

    def _pass_rows(self, src, dst, a: int, b: int) -> None:
        """Compute owned rows [a, b) of ``dst`` from ``src``."""
        if b > a:
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
        src = self._alloc(texture)
        dst = self._alloc(texture)
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
            if not split or last:
                self._pass_rows(src, dst, 0, p.nrows)
                if not last:
                    self._exchange(dst)
            elif not side:           # same decomposition without streams (CPU tests)
                self._pass_rows(src, dst, 0, h)
                self._pass_rows(src, dst, p.nrows - h, p.nrows)
                self._exchange(dst)
                self._pass_rows(src, dst, h, p.nrows - h)
            else:
                main = torch.cuda.current_stream()
                # 1. the strips the neighbours are waiting for
                self._pass_rows(src, dst, 0, h)
                self._pass_rows(src, dst, p.nrows - h, p.nrows)
                strips_done = torch.cuda.Event()
                strips_done.record(main)
                # 2. ship them while 3. the interior is computed
                self.comm_stream.wait_event(strips_done)
                with torch.cuda.stream(self.comm_stream):
                    self._exchange(dst)
                    shipped = torch.cuda.Event()
                    shipped.record(self.comm_stream)
                self._pass_rows(src, dst, h, p.nrows - h)
                main.wait_event(shipped)
            src, dst = dst, src
        out = torch.empty_like(texture)
        self.ops.unpad_texture(src, out, p, self.walls)
        return out

    # -- fused halo exchange --------------------------------------------------
    def _convolve_peer(self, texture: torch.Tensor, iterations: int) -> torch.Tensor:
        """``convolve`` with ``exchange="peer"``: see the module docstring.  Per pass n (tick t):

            wait   free >= t - 1 from both neighbours      (their pass n - 1 is over: the buffer
                                                             my strips are about to land in is idle)
            strips pass over the two edge strips, stored here AND in the neighbours' halos
            signal halo = t to both neighbours
            pass   over the interior
            signal free = t to both neighbours              (my pass n is over)
            wait   halo >= t from both neighbours           (their strips are in my halos)

        all on the current stream, none of it involving the host.  The first input's halos
        travel by the ordinary exchange, which also tells each rank that its neighbours have
        finished the previous call."""
        p, h = self.plan, self.plan.reach
        if self._peer is None:
            self._peer = _PeerExchange(p, self.group, self._peer_memory or CudaPeerMemory(),
                                       texture.dtype, texture.device)
        px = self._peer
        if px.bufs[0].dtype != texture.dtype:
            raise TypeError("the peer buffers of this convolver hold another dtype")
        bufs, flags = px.bufs, px.flags
        up = px.remote.get(p.up) if p.up is not None else None
        down = px.remote.get(p.down) if p.down is not None else None
        self.ops.pad_texture(texture.contiguous(), bufs[0], p, self.walls)
        self._exchange(bufs[0])
        base, limit = px.tick, self.PEER_TIMEOUT_MS
        for it in range(iterations):
            t = base + it + 1
            src, dst = bufs[it % 2], bufs[(it + 1) % 2]
            if it == iterations - 1:
                self._pass_rows(src, dst, 0, p.nrows)
                break
            if it > 0:
                if up is not None:
                    self.ops.wait(flags, _FREE_FROM_UP, t - 1, limit, _TIMED_OUT)
                if down is not None:
                    self.ops.wait(flags, _FREE_FROM_DOWN, t - 1, limit, _TIMED_OUT)
            if up is not None:
                self.ops.pass_rows_peer(src, self.field, dst, p, 0, h, self.taps, self.mode, self.walls,
                                        up[0][(it + 1) % 2], px.delta_up)
            else:
                self._pass_rows(src, dst, 0, h)
            if down is not None:
                self.ops.pass_rows_peer(src, self.field, dst, p, p.nrows - h, p.nrows, self.taps, self.mode,
                                        self.walls, down[0][(it + 1) % 2], px.delta_down)
            else:
                self._pass_rows(src, dst, p.nrows - h, p.nrows)
            if up is not None:
                self.ops.signal(up[1], _HALO_FROM_DOWN, t)
            if down is not None:
                self.ops.signal(down[1], _HALO_FROM_UP, t)
            self._pass_rows(src, dst, h, p.nrows - h)
            if up is not None:
                self.ops.signal(up[1], _FREE_FROM_DOWN, t)
                self.ops.wait(flags, _HALO_FROM_UP, t, limit, _TIMED_OUT)
            if down is not None:
                self.ops.signal(down[1], _FREE_FROM_UP, t)
                self.ops.wait(flags, _HALO_FROM_DOWN, t, limit, _TIMED_OUT)
        px.tick = base + iterations
        out = torch.empty_like(texture)
        self.ops.unpad_texture(bufs[iterations % 2], out, p, self.walls)
        return out

    def peer_timed_out(self) -> bool:
        """Whether any wait for a neighbour gave up (synchronises with the device)."""
        return bool(self._peer is not None and int(self._peer.flags[_TIMED_OUT].item()) != 0)

    def close(self) -> None:
        """Release the peer mappings and buffers (collective; only needed with ``exchange="peer"``)."""
        if self._peer is not None:
            self._peer.close(self.group)
            self._peer = None
