"""The library's CUDA kernels run on the CPU -- TEST INFRASTRUCTURE, not a fallback.

``rlic_b200/csrc/lic_walk.cuh`` (the streamline walk, the field packing with its wall
sentinels, the texture padding) is compiled with g++ behind ``RLIC_HOST_EMULATION`` and a
shim of the CUDA built-ins it uses (``cuda_on_cpu.h``), and driven block by block by
``emulate.cpp``.  The tests compare it with the oracle on machines without a GPU, so a
change to the kernel logic is checked before any GPU time is spent on it.  Buffer
geometry comes from the library's own host code (``rlic_b200_debug_geometry``).

What this cannot show: anything about nvcc's code generation or the hardware -- in
particular the approximate reciprocal that seeds the packed field (see ``cuda_on_cpu.h``).
The ``-m gpu`` parity tests remain the proof for the shipped binary.
"""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

from rlic_b200 import _core

HERE = Path(__file__).resolve().parent
LIB = HERE / "libkernel_emulation.so"
KERNEL_SOURCE = HERE.parents[1] / "rlic_b200" / "csrc" / "lic_walk.cuh"
FLAGS = ["-O2", "-std=c++17", "-march=x86-64-v3", "-ffp-contract=off", "-fPIC", "-shared", "-pthread"]

_i64, _int = ctypes.c_int64, ctypes.c_int
_lib = None


def build(force: bool = False) -> Path:
    deps = [HERE / "emulate.cpp", HERE / "cuda_on_cpu.h", KERNEL_SOURCE]
    if force or not LIB.exists() or any(d.stat().st_mtime > LIB.stat().st_mtime for d in deps):
        # one translation unit per scalar type, compiled side by side (the template
        # instantiations are what takes the time), then linked
        compile_flags = [f for f in FLAGS if f != "-shared"]
        units = []
        for unit in ("F32", "F64"):
            obj = HERE / f"emulate_{unit.lower()}.o"
            units.append((obj, subprocess.Popen(
                ["g++", *compile_flags, f"-DEMU_UNIT_{unit}", f"-I{HERE}", "-c", str(HERE / "emulate.cpp"),
                 "-o", str(obj)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        for obj, proc in units:
            out, _ = proc.communicate()
            if proc.returncode:
                raise RuntimeError(f"g++ failed on {obj.name}:\n{out}")
        subprocess.run(["g++", "-shared", "-pthread", *(str(obj) for obj, _ in units), "-o", str(LIB)],
                       check=True, capture_output=True, text=True)
    return LIB


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        cdll = ctypes.CDLL(str(build()))
        for sfx, real in (("f32", ctypes.c_float), ("f64", ctypes.c_double)):
            p, g = ctypes.POINTER(real), ctypes.POINTER(_i64)
            getattr(cdll, f"emu_pack_field_{sfx}").argtypes = [p, p, p, g, _i64, _i64, _i64]
            getattr(cdll, f"emu_pack_field_{sfx}").restype = None
            getattr(cdll, f"emu_pad_texture_{sfx}").argtypes = [p, p, g, _i64, _i64, _i64, ctypes.POINTER(_int)]
            getattr(cdll, f"emu_pad_texture_{sfx}").restype = None
            getattr(cdll, f"emu_unpad_texture_{sfx}").argtypes = [p, p, g, _i64, _i64, _i64]
            getattr(cdll, f"emu_unpad_texture_{sfx}").restype = None
            getattr(cdll, f"emu_pass_{sfx}").argtypes = [p, p, p, g, _i64, _i64, _i64, _int, p, _i64,
                                                        _int, _int, _int, _int, _int]
            getattr(cdll, f"emu_pass_{sfx}").restype = _int
        for sfx, real in (("f32", ctypes.c_float), ("f64", ctypes.c_double)):
            p, g = ctypes.POINTER(real), ctypes.POINTER(_i64)
            getattr(cdll, f"emu_pass_peer_{sfx}").argtypes = [p, p, p, g, _i64, _i64, _int, p, _i64, _int, p, _i64]
            getattr(cdll, f"emu_pass_peer_{sfx}").restype = _int
        for sfx, real in (("f32", ctypes.c_float), ("f64", ctypes.c_double)):
            p, g = ctypes.POINTER(real), ctypes.POINTER(_i64)
            getattr(cdll, f"emu_pass_paths_{sfx}").argtypes = [p, p, p, g, _i64, _i64, _i64, _int, p, _i64, _int, _int,
                                                              ctypes.POINTER(ctypes.c_uint32), p, _i64]
            getattr(cdll, f"emu_pass_paths_{sfx}").restype = _int
        cdll.emu_step_counts.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), _int]
        cdll.emu_step_counts.restype = None
        _lib = cdll
    return _lib


def step_counts(reset: bool = True) -> dict:
    """How the steps of the passes run so far were decided: ``step`` all of them, ``declined``
    the ones the fast path handed over, of which ``wall`` crossed a wall (sentinel cell) and
    ``generic`` took the generic step (a wall crossing usually continues on the fast path's
    behalf through the generic step too)."""
    out = (ctypes.c_ulonglong * 4)()
    lib().emu_step_counts(out, int(reset))
    return dict(zip(("step", "declined", "wall", "generic"), (int(x) for x in out)))


def _kind(dtype):
    return {"float32": ("f32", ctypes.c_float), "float64": ("f64", ctypes.c_double)}[np.dtype(dtype).name]


def _ptr(a: np.ndarray, real):
    return a.ctypes.data_as(ctypes.POINTER(real))


def geometry(ny, nx, slab, walls, klen) -> np.ndarray:
    """The library's buffer geometry for ``slab = (row0, nrows, halo_lo, halo_hi)``."""
    out = np.zeros(10, dtype=np.int64)
    rc = _core.lib.rlic_b200_debug_geometry(ny, nx, *slab, *walls, klen, _ptr(out, _i64))
    _core.check(rc)
    return out


class SlabOps:
    """The ``ops`` interface of ``rlic_b200.sharded.ShardedConvolver`` (what ``CudaSlabOps``
    does through the C ABI's slab entry points), on CPU tensors through the emulated kernels."""

    @staticmethod
    def _np(t):
        return t.numpy()

    @staticmethod
    def _geom(plan, walls, klen=1):
        return geometry(plan.ny, plan.nx, (plan.row0, plan.nrows, plan.halo_lo, plan.halo_hi), walls, klen)

    # rows=(a, b): owned rows [a, b) only, the dense tensor holding just those rows
    # (the rlic_b200_slab_*_rows_* entry points)
    def pack_field(self, u, v, field, plan, walls, rows=None):
        sfx, real = _kind(self._np(u).dtype)
        g = self._geom(plan, walls)
        a, b = rows if rows is not None else (0, plan.nrows)
        getattr(lib(), f"emu_pack_field_{sfx}")(
            _ptr(self._np(u), real), _ptr(self._np(v), real), _ptr(self._np(field), real), _ptr(g, _i64),
            plan.halo_lo + a, plan.halo_lo + b, 1)

    def pad_texture(self, texture, padded, plan, walls, rows=None):
        sfx, real = _kind(self._np(texture).dtype)
        g = self._geom(plan, walls)
        a, b = rows if rows is not None else (0, plan.nrows)
        getattr(lib(), f"emu_pad_texture_{sfx}")(
            _ptr(self._np(texture), real), _ptr(self._np(padded), real), _ptr(g, _i64),
            plan.halo_lo + a, plan.halo_lo + b, 1, None)

    def unpad_texture(self, padded, texture, plan, walls, rows=None):
        sfx, real = _kind(self._np(texture).dtype)
        g = self._geom(plan, walls)
        a, b = rows if rows is not None else (0, plan.nrows)
        getattr(lib(), f"emu_unpad_texture_{sfx}")(
            _ptr(self._np(padded), real), _ptr(self._np(texture), real), _ptr(g, _i64),
            plan.halo_lo + a, plan.halo_lo + b, 1)

    def pass_rows(self, src, field, dst, plan, a, b, taps, mode, walls):
        sfx, real = _kind(self._np(src).dtype)
        g = self._geom(plan, walls, taps.size)
        rc = getattr(lib(), f"emu_pass_{sfx}")(
            _ptr(self._np(src), real), _ptr(self._np(field), real), _ptr(self._np(dst), real), _ptr(g, _i64), 1,
            plan.halo_lo + a, b - a, mode, _ptr(taps, real), taps.size, 0, -1, -1, 1, 0)
        assert rc == 0


    def pass_rows_peer(self, src, field, dst, plan, a, b, taps, mode, walls, peer, peer_row_delta):
        """rlic_b200_pass_slab_peer_*: the pass over owned rows [a, b) that also stores its rows
        into `peer` (the neighbour's padded buffer), `peer_row_delta` buffer rows away."""
        sfx, real = _kind(self._np(src).dtype)
        g = self._geom(plan, walls, taps.size)
        rc = getattr(lib(), f"emu_pass_peer_{sfx}")(
            _ptr(self._np(src), real), _ptr(self._np(field), real), _ptr(self._np(dst), real), _ptr(g, _i64),
            plan.halo_lo + a, b - a, mode, _ptr(taps, real), taps.size, int(getattr(self, "walk", 0)),
            _ptr(self._np(peer), real), peer_row_delta)
        assert rc == 0


    def pass_rows_paths(self, src, field, dst, plan, a, b, taps, mode, walls, paths_mode, paths,
                        peer=None, peer_row_delta=0):
        """rlic_b200_pass_slab_paths_*: rows [a, b) with their streamline paths recorded into (1) or
        replayed from (2) the int32 tensor `paths`; `peer`: the doubled stores of the fused exchange."""
        sfx, real = _kind(self._np(src).dtype)
        g = self._geom(plan, walls, taps.size)
        rec = self._np(paths).view(np.uint32)
        # a replayed pass: the staged kernel where the library would use it (rc 2: not eligible)
        modes = (3, 2) if int(paths_mode) == 2 and getattr(self, "staging", True) else (int(paths_mode),)
        for m in modes:
            rc = getattr(lib(), f"emu_pass_paths_{sfx}")(
                _ptr(self._np(src), real), _ptr(self._np(field), real), _ptr(self._np(dst), real), _ptr(g, _i64), 1,
                plan.halo_lo + a, b - a, mode, _ptr(taps, real), taps.size, 0, m,
                rec.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)),
                None if peer is None else _ptr(self._np(peer), real), peer_row_delta)
            if rc != 2:
                break
        assert rc == 0


class Buffers:
    """Padded device-style buffers of one slab (or a batch of whole images), on the host."""

    def __init__(self, dtype, ny, nx, walls, klen, slab=None, nfields=1):
        self.sfx, self.real = _kind(dtype)
        self.dtype = np.dtype(dtype)
        self.ny, self.nx, self.walls, self.nfields = ny, nx, walls, nfields
        self.slab = slab or (0, ny, 0, 0)
        self.geom = geometry(ny, nx, self.slab, walls, klen)
        self.rows, self.cells = int(self.geom[2]), int(self.geom[3])
        self.field = np.full(4 * self.cells * nfields, np.nan, dtype=self.dtype)
        self.tex = [np.full(self.cells * nfields, np.nan, dtype=self.dtype) for _ in range(2)]
        self._g = _ptr(self.geom, _i64)

    def pack_field(self, u, v, rows=None):
        rb, re = rows or (0, self.rows)
        u, v = (np.ascontiguousarray(a, dtype=self.dtype) for a in (u, v))
        assert u.size == self.nfields * (re - rb) * self.nx == v.size
        getattr(lib(), f"emu_pack_field_{self.sfx}")(_ptr(u, self.real), _ptr(v, self.real),
                                                     _ptr(self.field, self.real), self._g, rb, re, self.nfields)

    def pad_texture(self, dense, which=0, rows=None) -> bool:
        rb, re = rows or (0, self.rows)
        dense = np.ascontiguousarray(dense, dtype=self.dtype)
        assert dense.size == self.nfields * (re - rb) * self.nx
        negative = _int(0)
        getattr(lib(), f"emu_pad_texture_{self.sfx}")(_ptr(dense, self.real), _ptr(self.tex[which], self.real),
                                                      self._g, rb, re, self.nfields, ctypes.byref(negative))
        return bool(negative.value)

    def unpad_texture(self, which, rows=None) -> np.ndarray:
        rb, re = rows or (0, self.rows)
        dense = np.empty((self.nfields, re - rb, self.nx), dtype=self.dtype)
        getattr(lib(), f"emu_unpad_texture_{self.sfx}")(_ptr(self.tex[which], self.real), _ptr(dense, self.real),
                                                        self._g, rb, re, self.nfields)
        return dense[0] if self.nfields == 1 else dense

    def run_pass(self, src, dst, taps, uv_mode, rows=None, wide=False, flavor=-1, admit=-1, branchless=True,
                 walk=0):
        first, count = rows or (self.slab[2], self.slab[1])
        taps = np.ascontiguousarray(taps, dtype=self.dtype)
        rc = getattr(lib(), f"emu_pass_{self.sfx}")(
            _ptr(self.tex[src], self.real), _ptr(self.field, self.real), _ptr(self.tex[dst], self.real),
            self._g, self.nfields, first, count, _core.mode_code(uv_mode), _ptr(taps, self.real), taps.size,
            int(wide), flavor, admit, int(branchless), int(walk))
        assert rc == 0, "no such formulation"


    def path_record(self, klen) -> np.ndarray:
        """Storage for the recorded paths of these buffers (rlic_b200_path_record_bytes); filled with
        a pattern, so that a replay that reads a word the recording pass did not write shows."""
        words = _core.path_record_bytes(self.rows, self.nx, klen) // 4 * self.nfields
        return np.full(words, 0xDEADBEEF, dtype=np.uint32)

    def run_pass_paths(self, src, dst, taps, uv_mode, mode, rec, rows=None, wide=False, peer=None,
                       peer_row_delta=0):
        """One pass that records (mode 1: the tuned grouped walk) or replays the paths (mode 2: gathers
        through L1; mode 3: the texture window staged in shared memory -- returns False when the
        library would not use that kernel for this case)."""
        first, count = rows or (self.slab[2], self.slab[1])
        taps = np.ascontiguousarray(taps, dtype=self.dtype)
        rc = getattr(lib(), f"emu_pass_paths_{self.sfx}")(
            _ptr(self.tex[src], self.real), _ptr(self.field, self.real), _ptr(self.tex[dst], self.real),
            self._g, self.nfields, first, count, _core.mode_code(uv_mode), _ptr(taps, self.real), taps.size,
            int(wide), int(mode), rec.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)),
            None if peer is None else _ptr(peer, self.real), peer_row_delta)
        assert rc in (0, 2), "no such paths mode"
        return rc == 0


def convolve(texture, u, v, *, kernel, uv_mode="velocity", boundaries=(("closed", "closed"),) * 2,
             iterations=1, wide=False, flavor=-1, admit=-1, branchless=True, walk=0, paths=False) -> np.ndarray:
    """The whole-image flow of ``run_device`` in lic_api.cu: pack, pad, passes, un-pad.
    paths=True: the first pass records the streamline paths, the others replay them (what the
    library does by default for iterations >= 2)."""
    ny, nx = texture.shape
    b = Buffers(texture.dtype, ny, nx, _core.wall_codes(boundaries), len(kernel))
    b.pack_field(u, v)
    b.pad_texture(texture, 0)
    src = 0
    rec = b.path_record(len(kernel)) if paths else None
    for it in range(iterations):
        if paths and it == 0:
            b.run_pass_paths(src, 1 - src, kernel, uv_mode, 1, rec, wide=wide)
        elif paths:
            # paths="staged": the shared-memory kernel where the library would launch it
            if not (paths == "staged" and b.run_pass_paths(src, 1 - src, kernel, uv_mode, 3, rec, wide=wide)):
                b.run_pass_paths(src, 1 - src, kernel, uv_mode, 2, rec, wide=wide)
        else:
            b.run_pass(src, 1 - src, kernel, uv_mode, wide=wide, flavor=flavor, admit=admit,
                       branchless=branchless, walk=walk)
        src = 1 - src
    return b.unpad_texture(src)
