"""TEST INFRASTRUCTURE -- NumPy oracle of `rlic_b200.equalize_histogram`.

The reference only declares this operation (`equalize_histogram_f32 / _f64(image, nbins)`,
/root/reference/src/rlic/_core.pyi:30-37); it has no implementation in the reference tree
(upstream moved it to the project `ahe`, /root/reference/README.md:19-23, which is not available
offline).  PARITY UNPINNED against any upstream code: this file DEFINES the semantics that
include/rlic_b200.h states and the CUDA kernels (rlic_b200/csrc/lic_equalize.cuh) implement,
one IEEE operation in the image's dtype at a time, so that the two can be compared bit for bit.
Only tests may import it.
"""
from __future__ import annotations

import numpy as np


def equalize_histogram(image: np.ndarray, nbins: int) -> np.ndarray:
    image = np.asarray(image)
    T = image.dtype.type
    out = np.array(image, copy=True)
    good = ~np.isnan(image)
    if not good.any():
        return out
    values = image[good]
    lo, hi = values.min(), values.max()
    w = T(hi - lo)
    if w == 0:
        bins = np.zeros(values.shape, dtype=np.int64)
    else:
        with np.errstate(all="ignore"):
            scaled = ((values - lo) / w) * T(nbins)              # three operations, each rounded in T
        bins = np.clip(np.floor(scaled.astype(np.float64)).astype(np.int64), 0, nbins - 1)
    counts = np.bincount(bins, minlength=nbins).astype(np.int64)
    running = np.cumsum(counts)
    cdf = running.astype(image.dtype) / T(values.size)           # integer -> T (nearest), one division
    out[good] = cdf[bins]
    return out
