"""TEST INFRASTRUCTURE — ctypes loader for the C oracle (``liblic_oracle.so``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs
may import this package.  ``rlic_b200`` never does: the product path has no CPU
fallback.

The oracle restates ``/root/reference/src/lib.rs`` (see ``lic_oracle.c`` for the
line-by-line citations and the pinning status).
"""

from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "liblic_oracle.so"

# Cargo feature sets of the reference (Cargo.toml:27-30): bit 0 fma, bit 1 branchless.
VARIANT_PLAIN = 0
VARIANT_FMA = 1  # PyPI x86_64 wheels (cd.yml: --no-default-features -F fma)
VARIANT_BRANCHLESS = 2
VARIANT_DEFAULT = 3  # crate default: fma + branchless

_BOUNDARY_CODE = {"closed": 0, "periodic": 1}
_MODE_CODE = {"velocity": 0, "polarization": 1}


def build(force: bool = False) -> Path:
    """Compile the oracle with its Makefile (gcc only)."""
    srcs = [_HERE / "lic_oracle.c", _HERE / "lic_walk.inc", _HERE / "Makefile"]
    stale = not _LIB_PATH.exists() or any(
        s.stat().st_mtime > _LIB_PATH.stat().st_mtime for s in srcs
    )
    if force or stale:
        subprocess.run(["make", "-C", str(_HERE), "-B"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(os.fspath(_LIB_PATH))
        for sfx, ct in (("f32", ctypes.c_float), ("f64", ctypes.c_double)):
            fn = getattr(_lib, f"lic_oracle_convolve_{sfx}")
            p = ctypes.POINTER(ct)
            fn.argtypes = [p, p, p, ctypes.c_int64, ctypes.c_int64, p, ctypes.c_int64]
            fn.argtypes += [ctypes.c_int] * 5 + [ctypes.c_int64, ctypes.c_int, ctypes.c_int, p]
            fn.restype = ctypes.c_int
            pr = getattr(_lib, f"lic_oracle_pass_rows_{sfx}")
            pr.argtypes = [p, p, p, ctypes.c_int64, ctypes.c_int64, p, ctypes.c_int64]
            pr.argtypes += [ctypes.c_int] * 5 + [ctypes.c_int64, ctypes.c_int64, ctypes.c_int, p]
            pr.restype = ctypes.c_int
            et = getattr(_lib, f"lic_oracle_edge_time_{sfx}")
            et.argtypes = [ct, ct, ctypes.c_int]
            et.restype = ct
            cr = getattr(_lib, f"lic_oracle_cross_{sfx}")
            cr.argtypes = [ct, ct, ctypes.c_int64, ctypes.c_int64]
            cr.argtypes += [ctypes.c_int] * 5
            cr.argtypes += [ctypes.POINTER(ctypes.c_long)] * 2 + [p, p]
            cr.restype = None
        _lib.lic_oracle_max_threads.restype = ctypes.c_int
    return _lib


def max_threads() -> int:
    return int(lib().lic_oracle_max_threads())


def _sfx(dtype: np.dtype) -> tuple[str, type]:
    if dtype == np.float32:
        return "f32", ctypes.c_float
    if dtype == np.float64:
        return "f64", ctypes.c_double
    raise TypeError(f"oracle handles float32/float64 only, got {dtype}")


def convolve(
    texture,
    u,
    v,
    *,
    kernel,
    uv_mode: str = "velocity",
    boundaries=(("closed", "closed"), ("closed", "closed")),
    iterations: int = 1,
    variant: int = VARIANT_DEFAULT,
    threads: int = 1,
):
    """Run the C oracle.  ``boundaries`` is ``((x_left, x_right), (y_left, y_right))``
    as passed to ``_core.convolve_*`` by the reference (``_lib.py:235``), or a
    single string for all four sides.  No validation: callers pass valid inputs."""
    if isinstance(boundaries, str):
        boundaries = ((boundaries, boundaries), (boundaries, boundaries))
    texture = np.ascontiguousarray(texture)
    sfx, ct = _sfx(texture.dtype)
    u = np.ascontiguousarray(u, dtype=texture.dtype)
    v = np.ascontiguousarray(v, dtype=texture.dtype)
    kernel = np.ascontiguousarray(kernel, dtype=texture.dtype)
    ny, nx = texture.shape
    out = np.empty_like(texture)
    p = ctypes.POINTER(ct)
    (xl, xr), (yl, yr) = boundaries
    rc = getattr(lib(), f"lic_oracle_convolve_{sfx}")(
        texture.ctypes.data_as(p),
        u.ctypes.data_as(p),
        v.ctypes.data_as(p),
        ny,
        nx,
        kernel.ctypes.data_as(p),
        kernel.size,
        _MODE_CODE[uv_mode],
        _BOUNDARY_CODE[xl],
        _BOUNDARY_CODE[xr],
        _BOUNDARY_CODE[yl],
        _BOUNDARY_CODE[yr],
        iterations,
        variant,
        threads,
        out.ctypes.data_as(p),
    )
    if rc != 0:
        raise RuntimeError(f"lic_oracle_convolve_{sfx} failed with code {rc}")
    return out


def pass_rows(texture, u, v, *, kernel, rows, uv_mode="velocity",
              boundaries=(("closed", "closed"), ("closed", "closed")), threads=1):
    """One pass (crate-default arithmetic) over image rows ``rows=(r0, r1)`` only, on
    ``threads`` threads; returns the ``(r1-r0, nx)`` band."""
    if isinstance(boundaries, str):
        boundaries = ((boundaries, boundaries), (boundaries, boundaries))
    texture = np.ascontiguousarray(texture)
    sfx, ct = _sfx(texture.dtype)
    u = np.ascontiguousarray(u, dtype=texture.dtype)
    v = np.ascontiguousarray(v, dtype=texture.dtype)
    kernel = np.ascontiguousarray(kernel, dtype=texture.dtype)
    ny, nx = texture.shape
    r0, r1 = rows
    out = np.zeros_like(texture)
    p = ctypes.POINTER(ct)
    (xl, xr), (yl, yr) = boundaries
    rc = getattr(lib(), f"lic_oracle_pass_rows_{sfx}")(
        texture.ctypes.data_as(p), u.ctypes.data_as(p), v.ctypes.data_as(p), ny, nx,
        kernel.ctypes.data_as(p), kernel.size, _MODE_CODE[uv_mode],
        _BOUNDARY_CODE[xl], _BOUNDARY_CODE[xr], _BOUNDARY_CODE[yl], _BOUNDARY_CODE[yr],
        r0, r1, threads, out.ctypes.data_as(p),
    )
    if rc != 0:
        raise RuntimeError(f"lic_oracle_pass_rows_{sfx} failed with code {rc}")
    return out[r0:r1].copy()


def edge_time(vel, frac, dtype, variant: int = VARIANT_DEFAULT):
    sfx, ct = _sfx(np.dtype(dtype))
    return getattr(lib(), f"lic_oracle_edge_time_{sfx}")(vel, frac, variant)


def cross(mu, mv, i, j, fx, fy, *, shape, dtype, boundaries="closed", variant=VARIANT_DEFAULT):
    """One pixel crossing; returns (i, j, fx, fy).  For the Rust `advance` KAT."""
    if isinstance(boundaries, str):
        boundaries = ((boundaries, boundaries), (boundaries, boundaries))
    sfx, ct = _sfx(np.dtype(dtype))
    ci, cj = ctypes.c_long(i), ctypes.c_long(j)
    cfx, cfy = ct(fx), ct(fy)
    (xl, xr), (yl, yr) = boundaries
    getattr(lib(), f"lic_oracle_cross_{sfx}")(
        mu, mv, shape[0], shape[1],
        _BOUNDARY_CODE[xl], _BOUNDARY_CODE[xr], _BOUNDARY_CODE[yl], _BOUNDARY_CODE[yr],
        variant, ctypes.byref(ci), ctypes.byref(cj), ctypes.byref(cfx), ctypes.byref(cfy),
    )
    return ci.value, cj.value, cfx.value, cfy.value
