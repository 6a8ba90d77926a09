"""Builds tests/golden/reference_images.npz from the reference's published figures.

Run in the development container only (needs /root/reference and Pillow):

    python tests/golden/make_reference_images.py

What it does
* crops the five LIC panels out of static/base_example_out.png (iterations 1, 5, 100) and
  static/polarization_example.png (uv_mode velocity / polarization);
* calibrates the rendering model of tests/reference_images.py on the left panel of
  static/base_example_in.png, which is ``imshow(texture)`` of the seeded noise and involves
  no LIC: the pixel alignment by maximising the correlation of the resampled texture with
  the panel's luminance, then the 256-entry colour table by linear least squares (the panel
  is linear in the table once the alignment is known);
* stores the panels (uint8 RGB), the fitted table and the calibration residual.

The figures are assets of rLIC (MIT, see NOTICE); only the cropped panels are kept.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
from PIL import Image
from scipy.optimize import minimize

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import reference_images as ri  # noqa: E402

STATIC = Path("/root/reference/static")
LUMA = np.array([0.2126, 0.7152, 0.0722])


def rgb(name: str) -> np.ndarray:
    return np.asarray(Image.open(STATIC / name).convert("RGB"))


def panel_spans(img: np.ndarray) -> tuple[list[tuple[int, int]], tuple[int, int]]:
    """Column spans and the row span of the colour-mapped panels (saturated pixels)."""
    wide = img.astype(np.int64)
    colourful = (wide.max(-1) - wide.min(-1)) > 40

    def spans(counts):
        idx = np.where(counts > 100)[0]
        cuts = np.where(np.diff(idx) > 1)[0]
        starts = np.r_[idx[0], idx[cuts + 1]]
        ends = np.r_[idx[cuts], idx[-1]]
        return [(int(a), int(b) + 1) for a, b in zip(starts, ends) if b - a > 200]

    return spans(colourful.sum(0)), spans(colourful.sum(1))[0]


def crop_panels(name: str) -> list[np.ndarray]:
    img = rgb(name)
    cols, (r0, r1) = panel_spans(img)
    return [img[r0:r1, a:b] for a, b in cols]


def calibrate(panel: np.ndarray) -> tuple[np.ndarray, float]:
    """Colour table from the imshow(texture) panel (and the panel's registration)."""
    texture = ri.readme_texture()
    n = panel.shape[0]
    luminance = panel.astype(np.float64) @ LUMA
    m = 6
    target_l = luminance[m:-m, m:-m] - luminance[m:-m, m:-m].mean()

    def negative_correlation(p):
        wy, wx = ri.hann_weights(n, p[0], p[1]), ri.hann_weights(n, p[2], p[3])
        r = (wy @ texture @ wx.T)[m:-m, m:-m]
        a = r - r.mean()
        return -float((a * target_l).sum() / np.sqrt((a * a).sum() * (target_l * target_l).sum()))

    coarse = min((negative_correlation((oy, n + e, ox, n + e)), oy, n + e, ox, n + e)
                 for oy in np.arange(-3, 1.01, 0.5) for ox in np.arange(-3, 1.01, 0.5)
                 for e in (1.0, 1.5, 2.0, 2.5))
    fit = minimize(negative_correlation, coarse[1:], method="Nelder-Mead",
                   options={"xatol": 1e-3, "fatol": 1e-6})
    oy, ey, ox, ex = fit.x
    print(f"calibration panel {n} px: rows origin {oy:.3f} extent {ey:.3f}, "
          f"columns origin {ox:.3f} extent {ex:.3f}, correlation {-fit.fun:.4f}")

    wy = ri.hann_weights(n, oy, ey).astype(np.float32)
    wx = ri.hann_weights(n, ox, ex).astype(np.float32)
    idx = ri.table_index(texture)
    onehot = np.zeros((ri.CELLS, ri.CELLS, 256), dtype=np.float32)
    onehot[np.arange(ri.CELLS)[:, None], np.arange(ri.CELLS)[None, :], idx] = 1
    design = np.einsum("pi,ijk->pjk", wy, onehot)
    design = np.einsum("pjk,qj->pqk", design, wx)[ri.MARGIN:-ri.MARGIN, ri.MARGIN:-ri.MARGIN]
    target = panel[ri.MARGIN:-ri.MARGIN, ri.MARGIN:-ri.MARGIN].astype(np.float64)
    table, *_ = np.linalg.lstsq(design.reshape(-1, 256), target.reshape(-1, 3), rcond=None)
    residual = np.abs(design.reshape(-1, 256) @ table - target.reshape(-1, 3))
    print(f"colour table fit: residual mean {residual.mean():.3f} max {residual.max():.2f} levels")
    return table, float(residual.mean())


def main() -> None:
    in_panel = crop_panels("base_example_in.png")[0]
    table, residual = calibrate(in_panel)
    base = crop_panels("base_example_out.png")
    pol = crop_panels("polarization_example.png")
    assert len(base) == 3 and len(pol) == 2, (len(base), len(pol))
    assert all(p.shape == (341, 341, 3) for p in base + pol), [p.shape for p in base + pol]
    np.savez_compressed(
        ri.FIXTURE,
        table=table.astype(np.float64),
        calibration_residual=np.float64(residual),
        base_iter1=base[0], base_iter5=base[1], base_iter100=base[2],
        pol_velocity=pol[0], pol_polarization=pol[1],
    )
    print("wrote", ri.FIXTURE, ri.FIXTURE.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
