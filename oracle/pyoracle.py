"""TEST INFRASTRUCTURE — second, slow restatement of the reference algorithm.

Pure-Python loops over NumPy scalars (so every operation rounds in the array
dtype), with fused multiply-add taken from libm through ctypes (``math.fma`` does
not exist on Python 3.12).  It exists so that a mistake in the C oracle cannot
hide behind self-agreement: tests compare the two on small inputs, bit for bit.

Written independently of ``lic_oracle.c`` from SURVEY.md section 0.3; cites
``/root/reference/src/lib.rs``.  Only small inputs: ~20 microseconds per pixel-step.
"""

from __future__ import annotations

import ctypes
import ctypes.util

import numpy as np

_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.fma.argtypes = [ctypes.c_double] * 3
_libm.fma.restype = ctypes.c_double
_libm.fmaf.argtypes = [ctypes.c_float] * 3
_libm.fmaf.restype = ctypes.c_float


class _Arith:
    """Scalar arithmetic in one dtype with a chosen multiply-add flavour."""

    def __init__(self, dtype, fused: bool):
        self.t = np.dtype(dtype).type
        self.fused = fused
        self._fma = _libm.fmaf if self.t is np.float32 else _libm.fma
        self.zero, self.half, self.one = self.t(0), self.t(0.5), self.t(1)

    def muladd(self, a, b, c):
        with np.errstate(all="ignore"):
            if self.fused:
                return self.t(self._fma(float(a), float(b), float(c)))
            return self.t(self.t(a * b) + c)


def _signum(ar: _Arith, x):
    # std float signum: sign bit decides, NaN stays NaN (num-traits 0.2.19 forwards to it)
    if np.isnan(x):
        return x
    return ar.t(np.copysign(ar.one, x))


def time_to_edge(ar: _Arith, vel, frac, branchless: bool):
    """ref: src/lib.rs:157-180."""
    with np.errstate(all="ignore"):
        if branchless:
            rem = ar.muladd(ar.t(ar.one + _signum(ar, vel)), ar.t(ar.half - frac), frac)
            return ar.t(abs(ar.t(rem / vel)))
        if vel > 0:
            return ar.t(ar.t(ar.one - frac) / vel)
        if vel < 0:
            return ar.t(-ar.t(frac / vel))
        return ar.t(np.inf)


def _fix(c: int, size: int, rules) -> int:
    """ref: src/lib.rs:83-95.  -1 plays the role of usize::MAX."""
    if c == -1:
        return size - 1 if rules[0] == "periodic" else 0
    if c == size:
        return 0 if rules[1] == "periodic" else size - 1
    return c


def step(ar, pos, frac, vel, shape, bounds, branchless):
    """ref: src/lib.rs:236-273.  pos=[i,j], frac=[fy,fx] are mutated in place;
    vel=(mv, mu) uses the same (row-axis, column-axis) ordering."""
    mv, mu = vel
    if mu == 0 and mv == 0:
        return
    t = [time_to_edge(ar, mv, frac[0], branchless), time_to_edge(ar, mu, frac[1], branchless)]
    # axis 1 (x) is taken only on a strict win; ties and NaN go to axis 0 (y)
    ax = 1 if t[1] < t[0] else 0
    other = 1 - ax
    if vel[ax] >= 0:
        pos[ax] += 1
        frac[ax] = ar.zero
    else:
        pos[ax] -= 1
        frac[ax] = ar.one
    with np.errstate(all="ignore"):
        frac[other] = ar.muladd(t[ax], vel[other], frac[other])
    pos[0] = _fix(pos[0], shape[0], bounds[1])  # y rules
    pos[1] = _fix(pos[1], shape[1], bounds[0])  # x rules


def convolve(
    texture,
    u,
    v,
    *,
    kernel,
    uv_mode="velocity",
    boundaries=(("closed", "closed"), ("closed", "closed")),
    iterations=1,
    fma=True,
    branchless=True,
):
    if isinstance(boundaries, str):
        boundaries = ((boundaries, boundaries), (boundaries, boundaries))
    ar = _Arith(texture.dtype, fma)
    ny, nx = texture.shape
    taps = [ar.t(k) for k in kernel]
    mid = len(taps) // 2
    fwd = list(range(mid + 1, len(taps)))
    bwd = list(range(mid - 1, -1, -1))
    src = np.array(texture, copy=True)
    dst = np.zeros_like(src)
    for n in range(iterations):
        if n:
            src, dst = dst.copy(), np.zeros_like(src)
        for i in range(ny):
            for j in range(nx):
                acc = ar.muladd(taps[mid], src[i, j], ar.zero)
                for sign, order in ((1, fwd), (-1, bwd)):
                    pos, frac = [i, j], [ar.half, ar.half]
                    prev = (ar.zero, ar.zero)
                    for k in order:
                        pu, pv = u[pos[0], pos[1]], v[pos[0], pos[1]]
                        if np.isnan(pu) or np.isnan(pv):
                            break
                        if uv_mode == "polarization":
                            with np.errstate(all="ignore"):
                                if ar.t(ar.t(pu * prev[0]) + ar.t(pv * prev[1])) < 0:
                                    pu, pv = ar.t(-pu), ar.t(-pv)
                            prev = (pu, pv)
                        if sign < 0:
                            pu, pv = ar.t(-pu), ar.t(-pv)
                        step(ar, pos, frac, (pv, pu), (ny, nx), boundaries, branchless)
                        with np.errstate(all="ignore"):
                            acc = ar.muladd(taps[k], src[pos[0], pos[1]], acc)
                dst[i, j] = acc
    return dst
