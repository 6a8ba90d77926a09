"""Line Integral Convolution on NVIDIA B200 (sm_100a), API-compatible with rLIC.

``rlic_b200.convolve`` mirrors ``rlic.convolve`` (reference:
``/root/reference/src/rlic/_lib.py:37-235``): same signature, validation, error
messages, dtype rules and semantics; the computation runs in hand-written CUDA
kernels behind a C ABI (``include/rlic_b200.h``).  There is no CPU fallback.

rLIC (MIT, C.M.T. Robert) and the vectorplot code it descends from (BSD-2,
Anne Archibald) are the behavioural reference only; see NOTICE.
"""

__all__ = [
    "convolve",
    "convolve_batch",
    "convolve_sharded",
    "effective_options",
    "equalize_histogram",
    "get_arithmetic",
    "get_paths",
    "get_schedule",
    "get_walk",
    "options",
    "set_arithmetic",
    "set_paths",
    "set_schedule",
    "set_walk",
]

from rlic_b200._core import (effective_options, get_arithmetic, get_paths, get_schedule, get_walk, options,
                             set_arithmetic, set_paths, set_schedule, set_walk)
from rlic_b200._lib import convolve, convolve_batch, convolve_sharded, equalize_histogram
