"""Metadata of the frozen vectors in tests/golden/lic_small.npz (see make_golden.py)."""

from pathlib import Path

import numpy as np

sys_path_root = Path(__file__).resolve().parent
GOLDEN = np.load(sys_path_root / "golden" / "lic_small.npz")

# name: (uv_mode, ((x_left, x_right), (y_left, y_right)), iterations)
CASES = {
    "vel_closed_f64": ("velocity", (("closed", "closed"), ("closed", "closed")), 1),
    "vel_periodic_f32_even": ("velocity", (("periodic", "periodic"), ("periodic", "periodic")), 2),
    "pol_mixed_f64": ("polarization", (("periodic", "periodic"), ("closed", "closed")), 1),
    "pol_mixed_f32_long": ("polarization", (("closed", "closed"), ("periodic", "periodic")), 3),
}


def load(name):
    g = GOLDEN
    return (g[f"{name}/texture"], g[f"{name}/u"], g[f"{name}/v"], g[f"{name}/kernel"])


def expected(name, variant=3):
    return GOLDEN[f"{name}/out_v{variant}"]


def as_spec(bnd):
    """((xl, xr), (yl, yr)) -> the dict form accepted by the public API."""
    return {"x": tuple(bnd[0]), "y": tuple(bnd[1])}
