"""The eight inputs of the reference's regression test (tests/test_regressions.py:19-63 of
rLIC): a 31 x 33 f32 noise texture, a sine kernel as long as the short side, and the
products of {0, a field that flips sign at mid-width} x {0, one that flips at mid-height}
x {velocity, polarization}.  The reference compares rlic.convolve with vectorplot's
line_integral_convolution on them (rtol 1.5e-7, atol 1e-6); vectorplot is not installed in
this image, so the expected values here come from an exact-arithmetic tracer instead."""
from __future__ import annotations

from fractions import Fraction
from itertools import product

import numpy as np

SHAPE = (31, 33)
NX, NY = SHAPE                      # the reference's own (swapped) names, kept for comparison
rng = np.random.default_rng(0)
TEXTURE = rng.random(SHAPE, dtype="float32")
ONE, ZERO = np.ones_like(TEXTURE), np.zeros_like(TEXTURE)
ii = np.broadcast_to(np.arange(NY), SHAPE)
U1 = np.where(ii < NY / 2, -ONE, ONE)
jj = np.broadcast_to(np.atleast_2d(np.arange(NX)).T, SHAPE)
V1 = np.where(jj < NX / 2, -ONE, ONE)
KL = min(NX, NY)
K0 = np.sin(np.arange(KL) * np.pi / KL, dtype="float32")
RTOL, ATOL = 1.5e-7, 1e-6           # the reference's tolerances (test_regressions.py:68)

CASES = {
    f"{'U1' if u is U1 else '0'}-{'V1' if v is V1 else '0'}-{mode}": (u, v, mode)
    for (u, v, mode) in product([ZERO, U1], [ZERO, V1], ["velocity", "polarization"])
}


def exact_streamline_sum(texture, u, v, kernel, uv_mode) -> np.ndarray:
    """The algorithm of SURVEY.md section 0.3 in exact rational arithmetic, closed walls, one
    pass: which pixels a walker visits never depends on rounding here, because every field
    value is 0 or +-1 and every position a dyadic fraction.  The weighted sum is taken the
    way vectorplot's C loop takes it -- float32, product then add, centre tap first, then the
    forward taps, then the backward ones -- so that the reference's tolerances apply.  Not a port of anything: positions are (cell, offset in [0, 1]) pairs of
    Fractions and the step is "advance to the nearest cell edge along the velocity"."""
    ny, nx = texture.shape
    mid = len(kernel) // 2
    out = np.zeros(texture.shape, dtype=np.float32)
    f32 = np.float32
    half = Fraction(1, 2)

    def time_to_edge(vel, off):
        if vel > 0:
            return (1 - off) / vel
        if vel < 0:
            return off / -vel
        return None                                      # never

    for i in range(ny):
        for j in range(nx):
            acc = f32(f32(kernel[mid]) * f32(texture[i, j]))
            for sign, taps in ((1, range(mid + 1, len(kernel))), (-1, range(mid - 1, -1, -1))):
                ci, cj, oy, ox = i, j, half, half
                prev = (Fraction(0), Fraction(0))
                for k in taps:
                    pu, pv = Fraction(float(u[ci, cj])), Fraction(float(v[ci, cj]))
                    if uv_mode == "polarization":
                        if pu * prev[0] + pv * prev[1] < 0:
                            pu, pv = -pu, -pv
                        prev = (pu, pv)
                    pu, pv = sign * pu, sign * pv
                    if pu != 0 or pv != 0:
                        tx, ty = time_to_edge(pu, ox), time_to_edge(pv, oy)
                        if tx is not None and (ty is None or tx < ty):      # ties go to y
                            cj += 1 if pu > 0 else -1
                            ox = Fraction(0) if pu > 0 else Fraction(1)
                            oy += tx * pv
                        else:
                            ci += 1 if pv > 0 else -1
                            oy = Fraction(0) if pv > 0 else Fraction(1)
                            ox += ty * pu
                        ci = min(max(ci, 0), ny - 1)                         # closed walls
                        cj = min(max(cj, 0), nx - 1)
                    acc = f32(acc + f32(f32(kernel[k]) * f32(texture[ci, cj])))
            out[i, j] = acc
    return out
