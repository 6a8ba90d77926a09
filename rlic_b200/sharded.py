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

The compute steps are injectable (``ops``) so the exchange logic can be
exercised without a GPU; the default is the CUDA slab API of the C ABI.  There
is no CPU compute path in this package.
"""

from __future__ import annotations

__all__ = ["SlabPlan", "ShardedConvolver", "CudaSlabOps"]

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


class ShardedConvolver:
    """Holds one rank's slab of a sharded image and runs ``convolve`` on it.

    ``texture``, ``u``, ``v``: this rank's rows ``[plan.row0, plan.row1)`` (dense,
    no halos), tensors on the rank's device.
    """

    def __init__(self, ny: int, nx: int, *, kernel, uv_mode: str = "velocity",
                 boundaries="closed", group=None, ops=None):
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
