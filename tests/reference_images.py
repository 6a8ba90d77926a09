"""Forward model of how the reference's published example figures were rendered.

The only outputs of the real rLIC that exist offline are the three figures of its README
(/root/reference/static/*.png, produced by the code blocks at README.md:44-69, 78-93 and
117-148 with a seeded texture).  Each LIC result there went through matplotlib's
``imshow``: min/max normalisation, a 256-entry viridis table, and -- because the 256-cell
image is drawn on ~342 pixels -- a Hanning-windowed resampling of the RGBA image.  This
module restates that rendering so a convolution result can be compared with the figure
pixel by pixel.  ``tests/golden/make_reference_images.py`` crops the panels, calibrates the
colour table and the pixel alignment on the *input texture* panel (which involves no LIC
at all) and writes ``tests/golden/reference_images.npz``.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

FIXTURE = Path(__file__).parent / "golden" / "reference_images.npz"
CELLS = 256
MARGIN = 4  # panel pixels next to the axes spines are not compared


def hann_weights(n_pixels: int, origin: float, extent: float) -> np.ndarray:
    """(n_pixels, CELLS) resampling matrix: panel pixel p (centre p + 0.5) looks at the
    cell coordinate (p + 0.5 - origin) * CELLS / extent and blends the two nearest cell
    centres with a raised-cosine window of radius one cell."""
    p = np.arange(n_pixels)
    t = (p + 0.5 - origin) * CELLS / extent - 0.5
    j0 = np.floor(t).astype(np.int64)
    w_near = 0.5 + 0.5 * np.cos(np.pi * (t - j0))
    weights = np.zeros((n_pixels, CELLS))
    np.add.at(weights, (p, np.clip(j0, 0, CELLS - 1)), w_near)
    np.add.at(weights, (p, np.clip(j0 + 1, 0, CELLS - 1)), 1.0 - w_near)
    return weights


def table_index(image: np.ndarray) -> np.ndarray:
    """imshow's default normalisation followed by the colormap's table lookup."""
    lo, hi = image.min(), image.max()
    return np.clip(((image - lo) / (hi - lo) * 256).astype(np.int64), 0, 255)


def render(image: np.ndarray, table: np.ndarray, wy: np.ndarray, wx: np.ndarray) -> np.ndarray:
    colours = table[table_index(image)]
    return np.stack([wy @ colours[:, :, c] @ wx.T for c in range(3)], axis=-1)


def _registrations(n_pixels: int):
    """Candidate (origin, extent) pairs per axis: matplotlib places the axes box on a
    fractional pixel position that differs between figures and between x and y, so the
    comparison is taken at the best placement within one pixel of the spines."""
    for origin in (-2.0, -1.5, -1.0, -0.5, 0.0):
        for extent in (n_pixels + 1.0, n_pixels + 1.5, n_pixels + 2.0):
            yield origin, extent


def panel_error(image: np.ndarray, panel: np.ndarray, table: np.ndarray) -> tuple[float, float]:
    """(mean, max) absolute difference, in 8-bit levels, between ``image`` rendered the
    way the figure was and the figure's own pixels, at the best registration."""
    image = np.asarray(image, dtype=np.float64)
    colours = table[table_index(image)]
    n = panel.shape[0]
    inner = slice(MARGIN, n - MARGIN)
    target = np.ascontiguousarray(panel[inner, inner].astype(np.float64).transpose(2, 0, 1))
    planes = np.ascontiguousarray(colours.transpose(2, 0, 1))            # (3, CELLS, CELLS)
    weights = [hann_weights(n, o, e)[inner] for o, e in _registrations(n)]
    best = (np.inf, np.inf)
    for wy in weights:
        rows = wy @ planes                                               # (3, n, CELLS)
        for wx in weights:
            diff = np.abs(rows @ wx.T - target)
            best = min(best, (float(diff.mean()), float(diff.max())))
    return best


# ---- the inputs of the README's code blocks ---------------------------------------------
def readme_texture() -> np.ndarray:
    return np.random.default_rng(0).random((CELLS, CELLS))


def readme_kernel() -> np.ndarray:
    return 1 - np.abs(np.linspace(-1, 1, 65))


def base_example_field() -> tuple[np.ndarray, np.ndarray]:
    """README.md:53-55 -- broadcast (stride-0) views, exactly as the README passes them."""
    x = np.linspace(0, np.pi, CELLS)
    shape = (CELLS, CELLS)
    return np.broadcast_to(np.cos(2 * x), shape), np.broadcast_to(np.sin(x).T, shape)


def polarization_example_field() -> tuple[np.ndarray, np.ndarray]:
    """README.md:126-129."""
    shape = (CELLS, CELLS)
    ones = np.ones(shape)
    column = np.broadcast_to(np.arange(CELLS), shape)
    return np.where(column < CELLS / 2, -ones, ones), np.zeros(shape)


def load() -> dict:
    with np.load(FIXTURE) as z:
        return {k: z[k] for k in z.files}
