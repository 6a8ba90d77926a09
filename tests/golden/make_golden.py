"""Regenerates tests/golden/lic_small.npz.

The reference's Python layer cannot run here (its native core needs a Rust
toolchain, absent from this image), so these vectors are NOT reference outputs:
they are produced by the slow pure-Python restatement (oracle/pyoracle.py) and
frozen, so that the C oracle, the CUDA path and any later rewrite of either are
all held to the same bits.  Hand-derived answers live
in tests/test_oracle.py as literal arrays.

    python tests/golden/make_golden.py
"""

from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import pyoracle  # noqa: E402

CASES = {
    # name: (shape, dtype, taps, mode, boundaries, iterations)
    "vel_closed_f64": ((14, 17), "float64", 11, "velocity",
                       (("closed", "closed"), ("closed", "closed")), 1),
    "vel_periodic_f32_even": ((12, 10), "float32", 8, "velocity",
                              (("periodic", "periodic"), ("periodic", "periodic")), 2),
    "pol_mixed_f64": ((11, 13), "float64", 9, "polarization",
                      (("periodic", "periodic"), ("closed", "closed")), 1),
    "pol_mixed_f32_long": ((6, 7), "float32", 21, "polarization",
                           (("closed", "closed"), ("periodic", "periodic")), 3),
}


def inputs(name, shape, dtype):
    rng = np.random.default_rng(sum(map(ord, name)))
    tex = rng.random(shape).astype(dtype)
    u = (rng.random(shape) - 0.5).astype(dtype)
    v = (rng.random(shape) - 0.5).astype(dtype)
    # special pixels: stagnation, NaN, signed zeros, one axis exactly zero
    u[1, 2] = 0.0
    v[1, 2] = 0.0
    u[3, 4] = np.nan
    v[5, 1] = -0.0
    u[2, 5] = -0.0
    v[2, 5] = 0.0
    u[4, 3] = 0.0
    return tex, u, v


def main() -> None:
    out = {}
    for name, (shape, dtype, taps, mode, bnd, its) in CASES.items():
        tex, u, v = inputs(name, shape, dtype)
        rng = np.random.default_rng(taps)
        kernel = (rng.random(taps) - 0.2).astype(dtype)
        out[f"{name}/texture"], out[f"{name}/u"], out[f"{name}/v"] = tex, u, v
        out[f"{name}/kernel"] = kernel
        for variant in range(4):
            res = pyoracle.convolve(tex, u, v, kernel=kernel, uv_mode=mode, boundaries=bnd,
                                    iterations=its, fma=bool(variant & 1),
                                    branchless=bool(variant & 2))
            out[f"{name}/out_v{variant}"] = res
    np.savez_compressed(Path(__file__).with_name("lic_small.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
