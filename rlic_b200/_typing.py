"""Type aliases used in annotations (same public names as ``rlic._typing``,
reference ``/root/reference/src/rlic/_typing.py``)."""

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

from typing import Literal, TypeAlias, TypeVar

import numpy as np

f32 = np.float32
f64 = np.float64

T = TypeVar("T")
F = TypeVar("F", np.float32, np.float64)

Pair: TypeAlias = tuple[T, T]
PairSpec: TypeAlias = T | tuple[T, T]

UVMode = Literal["velocity", "polarization"]

D1 = tuple[int]
D2 = tuple[int, int]
FArray1D = np.ndarray[D1, np.dtype[F]]
FArray2D = np.ndarray[D2, np.dtype[F]]
