"""One process, several GPUs: row slabs of one image, halos copied device to device
(BASELINE config 4 without ``torchrun``).

``rlic_b200.sharded`` runs one process per GPU under ``torch.distributed``; this module is
the same decomposition driven from a single process, for callers who just have a big
image and a box full of GPUs:

    out = rlic_b200.convolve_sharded(texture, u, v, kernel=kernel, iterations=20)

Device ``d`` holds the rows ``SlabPlan(world=len(devices), rank=d)`` assigns to it plus
``h = len(kernel) // 2`` halo rows on either side, in the kernels' padded buffers.  The
field's halos are copied once, the texture's after every pass: the two edge strips of every
slab are computed first, copied into the neighbours' halos (device-to-device copies; torch
orders a cross-device copy after the work already enqueued on both devices' streams and
makes the destination's stream wait for it, which is exactly the ordering the two ping-pong
buffers need), and the interiors are computed while those copies are in flight.  The
per-pixel arithmetic is the single-GPU kernel's in global row numbers, so the result is
bit-identical to an unsharded run.

Verified on the CPU through the emulated kernels and on B200s against ``convolve``
(``tests/test_multi_device.py``; a box with fewer GPUs than slabs lists a device several
times); a convenience, not benchmarked -- every pass walks (no recorded paths), and end to end
it is bound by one process's uploads.  The compute steps are injectable (``ops``) for the CPU
tests; there is no CPU compute path in this package.
"""

from __future__ import annotations

__all__ = ["MultiDeviceConvolver"]

import contextlib

import numpy as np
import torch

from rlic_b200._boundaries import BoundarySet
from rlic_b200.sharded import CudaSlabOps, ShardedConvolver, SlabPlan


def _usable_devices(ny: int, reach: int, devices) -> list[torch.device]:
    if devices is None:
        devices = [torch.device("cuda", i) for i in range(torch.cuda.device_count())]
        if not devices:
            raise RuntimeError("rlic_b200 has no CPU fallback: no CUDA device is visible")
    devices = [torch.device(d) if not isinstance(d, torch.device) else d for d in devices]
    # every slab must hold its neighbours' reach: use fewer devices for short images
    most = max(1, ny // max(reach, 1)) if ny else 1
    return devices[:max(1, min(len(devices), most))]


class MultiDeviceConvolver:
    """Row slabs of one ``ny`` x ``nx`` image on several devices of this process."""

    def __init__(self, ny: int, nx: int, *, kernel, uv_mode: str = "velocity", boundaries="closed",
                 devices=None, ops=None):
        from rlic_b200 import _core   # enum tables only; no computation

        bs = BoundarySet.from_spec(boundaries)
        if bs is None:
            raise TypeError(f"Invalid boundary specification {boundaries}")
        bs.validate()
        self.walls = _core.wall_codes((bs.x, bs.y))
        self.mode = _core.mode_code(uv_mode)
        self.taps = np.ascontiguousarray(kernel)
        if self.taps.ndim != 1 or self.taps.size == 0:
            raise ValueError("kernel must be a non-empty 1-D array")
        reach = self.taps.size // 2
        self.devices = _usable_devices(ny, reach, devices)
        n = len(self.devices)
        self.plans = [SlabPlan(ny=ny, nx=nx, world=n, rank=d, reach=reach, periodic_y=bs.y[0] == "periodic")
                      for d in range(n)]
        for plan in self.plans:
            plan.validate()
        self.ops = ops or CudaSlabOps()
        self.fields = None

    # -- helpers ----------------------------------------------------------------
    @staticmethod
    def _on(device):
        """Make ``device`` current (kernels are enqueued on its current stream)."""
        return torch.cuda.device(device) if device.type == "cuda" else contextlib.nullcontext()

    def _rows_to(self, array, d: int) -> torch.Tensor:
        """Device ``d``'s rows of a host image, on that device."""
        plan = self.plans[d]
        rows = np.ascontiguousarray(array[plan.row0:plan.row1])
        return torch.from_numpy(rows).to(self.devices[d], non_blocking=True)

    def _exchange(self, bufs, width: int = 1, planes: int = 1) -> None:
        """Fill every slab's halo rows from its neighbours' edge rows (the message pattern of
        ``ShardedConvolver._exchange``, as device-to-device copies)."""
        h = self.plans[0].reach
        for d, p in enumerate(self.plans):
            lo, n = p.halo_lo, p.nrows

            def rows(buf, plan, plane, a, b):
                s = plan.row_cells(a, b)
                base = plane * plan.cells
                return buf[(base + s.start) * width:(base + s.stop) * width]

            for plane in range(planes):
                if p.up is not None:      # my top rows are the upper neighbour's high halo
                    q = self.plans[p.up]
                    rows(bufs[p.up], q, plane, q.halo_lo + q.nrows, q.halo_lo + q.nrows + h).copy_(
                        rows(bufs[d], p, plane, lo, lo + h), non_blocking=True)
                if p.down is not None:    # my bottom rows are the lower neighbour's low halo
                    q = self.plans[p.down]
                    rows(bufs[p.down], q, plane, 0, h).copy_(
                        rows(bufs[d], p, plane, lo + n - h, lo + n), non_blocking=True)

    def _alloc(self, d: int, dtype, width: int = 1) -> torch.Tensor:
        return torch.zeros(self.plans[d].cells * width, dtype=dtype, device=self.devices[d])

    # -- the path ---------------------------------------------------------------
    def set_field(self, u, v) -> None:
        """Pack the whole-image components ``u``, ``v`` (host arrays) slab by slab; done once."""
        ny, nx = self.plans[0].ny, self.plans[0].nx
        if tuple(u.shape) != (ny, nx) or tuple(v.shape) != (ny, nx):
            raise ValueError(f"expected u and v of shape {(ny, nx)}")
        dtype = torch.from_numpy(np.empty(0, dtype=u.dtype)).dtype
        width, planes = ShardedConvolver.field_layout(dtype)
        fields = []
        for d, plan in enumerate(self.plans):
            with self._on(self.devices[d]):
                field = self._alloc(d, dtype, width * planes)
                self.ops.pack_field(self._rows_to(u, d), self._rows_to(v, d), field, plan, self.walls)
                fields.append(field)
        self._exchange(fields, width, planes)
        self.fields = fields

    def _pass(self, d, src, dst, a, b) -> None:
        if b > a:
            with self._on(self.devices[d]):
                self.ops.pass_rows(src[d], self.fields[d], dst[d], self.plans[d], a, b, self.taps, self.mode,
                                   self.walls)

    def convolve(self, texture, iterations: int = 1) -> np.ndarray:
        """``iterations`` passes over the whole host image ``texture``; returns a new host array."""
        if self.fields is None:
            raise RuntimeError("call set_field(u, v) first")
        ny, nx = self.plans[0].ny, self.plans[0].nx
        if tuple(texture.shape) != (ny, nx):
            raise ValueError(f"expected a texture of shape {(ny, nx)}")
        if iterations <= 0:
            return np.array(texture, copy=True)
        dtype = self.fields[0].dtype
        n, h = len(self.plans), self.plans[0].reach
        src, dst, dense = [], [], []
        for d, plan in enumerate(self.plans):
            with self._on(self.devices[d]):
                dense.append(self._rows_to(texture, d))
                src.append(self._alloc(d, dtype))
                dst.append(self._alloc(d, dtype))
                self.ops.pad_texture(dense[d], src[d], plan, self.walls)
        self._exchange(src)
        # strips first, interiors while the strips travel: needs an interior on every slab
        split = n > 1 and all(p.nrows >= 4 * max(h, 1) for p in self.plans)
        for it in range(iterations):
            last = it == iterations - 1
            if split and not last:
                for d, p in enumerate(self.plans):
                    self._pass(d, src, dst, 0, h)
                    self._pass(d, src, dst, p.nrows - h, p.nrows)
                self._exchange(dst)
                for d, p in enumerate(self.plans):
                    self._pass(d, src, dst, h, p.nrows - h)
            else:
                for d, p in enumerate(self.plans):
                    self._pass(d, src, dst, 0, p.nrows)
                if not last:
                    self._exchange(dst)
            src, dst = dst, src
        out = np.empty((ny, nx), dtype=np.asarray(texture).dtype)
        results = []
        for d, plan in enumerate(self.plans):
            with self._on(self.devices[d]):
                result = torch.empty_like(dense[d])      # never the upload buffer: on a CPU
                self.ops.unpad_texture(src[d], result, plan, self.walls)   # device it aliases the input
                results.append(result.to("cpu", non_blocking=False))
        for d, plan in enumerate(self.plans):
            out[plan.row0:plan.row1] = results[d].numpy()
        return out
