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

The per-pixel arithmetic is the single-GPU kernel's, in global row numbers, so
the gathered result is bit-identical to an unsharded run.

The compute step is injectable (``pass_fn``) so the exchange logic can be
exercised without a GPU; the default is the CUDA slab pass of the C ABI.  There
is no CPU compute path in this package.
"""

from __future__ import annotations

__all__ = ["SlabPlan", "ShardedConvolver"]

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


def _cuda_pass(tex, uv, out, plan: SlabPlan, row0, nrows, halo_lo, halo_hi, taps, mode, walls):
    """One CUDA pass over rows [row0, row0+nrows) of the global image."""
    from rlic_b200 import _core

    sfx, real = ("f32", ctypes.c_float) if tex.dtype == torch.float32 else ("f64", ctypes.c_double)
    rc = getattr(_core.lib, f"rlic_b200_pass_slab_{sfx}")(
        tex.data_ptr(), uv.data_ptr(), out.data_ptr(), plan.ny, plan.nx, row0, nrows, halo_lo,
        halo_hi, taps.ctypes.data_as(ctypes.POINTER(real)), taps.size, mode, *walls,
        int(torch.cuda.current_stream().cuda_stream))
    _core.check(rc)


class ShardedConvolver:
    """Holds one rank's slab of a sharded image and runs ``convolve`` on it.

    ``texture``, ``u``, ``v``: this rank's rows ``[plan.row0, plan.row1)`` (no halos),
    tensors on the rank's device.
    """

    def __init__(self, ny: int, nx: int, *, kernel, uv_mode: str = "velocity",
                 boundaries="closed", group=None, pass_fn=None, pack_fn=None):
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
        self._pass = pass_fn or _cuda_pass
        self._pack = pack_fn
        self.uv = None
        self.comm_stream = None

    # -- halo plumbing ------------------------------------------------------
    def _exchange(self, buf: torch.Tensor, async_op: bool = False):
        """Fill the halo rows of ``buf`` (rows_alloc x ...) from the neighbours.

        Every rank posts: send top rows up, send bottom rows down, receive the
        high halo from below, receive the low halo from above.  That order
        makes the k-th message between any pair match on both sides, including
        the two-rank ring where both neighbours are the same peer.
        """
        p = self.plan
        h, lo, n = p.reach, p.halo_lo, p.nrows
        ops = []
        if p.up is not None:
            ops.append(dist.P2POp(dist.isend, buf[lo:lo + h], p.up, self.group))
        if p.down is not None:
            ops.append(dist.P2POp(dist.isend, buf[lo + n - h:lo + n], p.down, self.group))
        if p.down is not None:
            ops.append(dist.P2POp(dist.irecv, buf[lo + n:lo + n + h], p.down, self.group))
        if p.up is not None:
            ops.append(dist.P2POp(dist.irecv, buf[0:lo], p.up, self.group))
        if not ops:
            return []
        reqs = dist.batch_isend_irecv(ops)
        if not async_op:
            for r in reqs:
                r.wait()
            return []
        return reqs

    def _alloc(self, like: torch.Tensor, trailing=()) -> torch.Tensor:
        p = self.plan
        return torch.empty((p.rows_alloc, p.nx, *trailing), dtype=like.dtype, device=like.device)

    def set_field(self, u: torch.Tensor, v: torch.Tensor) -> None:
        """Pack (u, v) with halos; done once, the field does not change between passes."""
        p = self.plan
        uv = self._alloc(u, (4,))
        owned = uv[p.halo_lo:p.halo_lo + p.nrows]
        if self._pack is not None:
            self._pack(u, v, owned)
        elif u.is_cuda:
            from rlic_b200.device import pack_field

            pack_field(u.contiguous(), v.contiguous(), out=owned)
        else:
            raise RuntimeError("rlic_b200 has no CPU fallback: pass CUDA tensors")
        self._exchange(uv)
        self.uv = uv

    def _pass_rows(self, src, dst, a: int, b: int) -> None:
        """Compute owned-relative rows [a, b) of ``dst`` from ``src`` (both with halos)."""
        if b <= a:
            return
        p = self.plan
        # rows of the buffers available around the strip
        lo_avail = p.halo_lo + a
        hi_avail = p.rows_alloc - (p.halo_lo + b)
        closed_top = p.up is None and not p.periodic_y
        closed_bottom = p.down is None and not p.periodic_y
        halo_lo = lo_avail if not (closed_top and a == 0) else 0
        halo_hi = hi_avail if not (closed_bottom and b == p.nrows) else 0
        # a strip that starts at a closed image edge needs no halo there; any
        # other side gets every row the buffer holds (>= reach by construction)
        first = p.halo_lo + a - halo_lo
        self._pass(src[first:], self.uv[first:], dst[p.halo_lo + a:], p, p.row0 + a, b - a,
                   halo_lo, halo_hi, self.taps, self.mode, self.walls)

    def convolve(self, texture: torch.Tensor, iterations: int = 1, overlap: bool = True) -> torch.Tensor:
        """Run ``iterations`` passes; returns this rank's rows of the result."""
        if self.uv is None:
            raise RuntimeError("call set_field(u, v) first")
        p = self.plan
        if texture.shape != (p.nrows, p.nx):
            raise ValueError(f"expected this rank's slab of shape {(p.nrows, p.nx)}")
        if iterations <= 0:
            return texture.clone()
        src = self._alloc(texture)
        dst = self._alloc(texture)
        src[p.halo_lo:p.halo_lo + p.nrows].copy_(texture)
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
        return src[p.halo_lo:p.halo_lo + p.nrows]
