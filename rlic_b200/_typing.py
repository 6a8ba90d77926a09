"""Annotation helpers.

The public names mirror the ones user code may import from the reference's
``rlic._typing`` (``/root/reference/src/rlic/_typing.py``): scalar aliases, the
dtype type variable, pair helpers, the mode literal and shaped-array aliases.
Nothing here has a runtime role in ``rlic_b200``.
"""

from typing import Literal, TypeAlias, TypeVar

import numpy as np

__all__ = [
    "D1",
    "D2",
    "F",
    "FArray1D",
    "FArray2D",
    "Pair",
    "PairSpec",
    "UVMode",
    "f32",
    "f64",
]

# --- scalars -----------------------------------------------------------------
f32: TypeAlias = np.float32
f64: TypeAlias = np.float64
#: one of the two supported floating-point types; every array of a call shares it
F = TypeVar("F", np.float32, np.float64)

# --- shapes and arrays -------------------------------------------------------
D1: TypeAlias = tuple[int]
D2: TypeAlias = tuple[int, int]
#: convolution kernel
FArray1D: TypeAlias = np.ndarray[D1, np.dtype[F]]
#: texture and vector-field components
FArray2D: TypeAlias = np.ndarray[D2, np.dtype[F]]

# --- options -----------------------------------------------------------------
_Item = TypeVar("_Item")
#: a (left, right) couple, e.g. the two sides of an axis
Pair: TypeAlias = tuple[_Item, _Item]
#: either one value for both sides, or a couple
PairSpec: TypeAlias = _Item | Pair[_Item]
#: how (u, v) is interpreted: as a direction, or as an orientation only
UVMode: TypeAlias = Literal["velocity", "polarization"]
