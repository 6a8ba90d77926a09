"""Boundary specifications: parsing and validation.

Behaviour mirrors ``/root/reference/src/rlic/_boundaries.py:38-114`` (accepted
spec shapes, normalised form, allowed combinations, exact error messages); the
implementation is independent.  ``x`` is the axis parallel to ``u`` (columns),
``y`` the axis parallel to ``v`` (rows).
"""

from __future__ import annotations

__all__ = [
    "COMBO_ALLOWED_BOUNDS",
    "COMBO_DISALLOWED_BOUNDS",
    "SUPPORTED_BOUNDS",
    "BoundarySet",
]

from collections.abc import Mapping, Sequence
from dataclasses import dataclass
from typing import Literal, TypeAlias, TypedDict

BoundaryStr: TypeAlias = Literal["closed", "periodic"]


class BoundaryDictSpec(TypedDict):
    x: "BoundaryStr | tuple[BoundaryStr, BoundaryStr]"
    y: "BoundaryStr | tuple[BoundaryStr, BoundaryStr]"


BoundarySpec: TypeAlias = "BoundaryStr | BoundaryDictSpec"

# a wall of this kind may face any other kind on the opposite side
COMBO_ALLOWED_BOUNDS: frozenset[str] = frozenset({"closed"})
# a wall of this kind needs the very same kind on the opposite side
COMBO_DISALLOWED_BOUNDS: frozenset[str] = frozenset({"periodic"})
SUPPORTED_BOUNDS: frozenset[str] = COMBO_ALLOWED_BOUNDS | COMBO_DISALLOWED_BOUNDS

_AXES = ("x", "y")
_SIDES = ("left", "right")


def _side_pair(entry: object) -> tuple[str, str] | None:
    """``"a"`` -> ``("a", "a")``; a 2-sequence of str -> tuple; anything else -> None."""
    if isinstance(entry, str):
        return (entry, entry)
    if isinstance(entry, Sequence) and not isinstance(entry, (bytes, bytearray)):
        if len(entry) == 2 and all(isinstance(e, str) for e in entry):
            return (entry[0], entry[1])
    return None


@dataclass(frozen=True, slots=True, kw_only=True)
class BoundarySet:
    """Normalised boundaries: ``x=(left, right)``, ``y=(left, right)``."""

    x: tuple[str, str]
    y: tuple[str, str]

    @staticmethod
    def from_spec(spec: object, /) -> "BoundarySet | None":
        """Expand a user specification, or return None when its *shape* is not
        acceptable (names are checked later, by :meth:`collect_exceptions`)."""
        if isinstance(spec, str):
            return BoundarySet(x=(spec, spec), y=(spec, spec))
        if isinstance(spec, Mapping) and len(spec) == 2 and all(ax in spec for ax in _AXES):
            pairs = [_side_pair(spec[ax]) for ax in _AXES]
            if None not in pairs:
                return BoundarySet(x=pairs[0], y=pairs[1])
        return None

    def collect_exceptions(self) -> list[Exception]:
        """All problems with the boundary names, x axis first, left side first."""
        problems: list[Exception] = []
        for axis in _AXES:
            # unpack by iteration: callers may hand in any 2-item iterable
            names = tuple(getattr(self, axis))
            unknown = [n not in SUPPORTED_BOUNDS for n in names]
            if any(unknown):
                # do not pile a combination error on top of what may be a typo
                problems.extend(
                    ValueError(f"Unknown {side} {axis} boundary {name!r}")
                    for side, name, bad in zip(_SIDES, names, unknown)
                    if bad
                )
                continue
            if names[0] == names[1]:
                continue
            for k, side in enumerate(_SIDES):
                if names[k] in COMBO_DISALLOWED_BOUNDS:
                    problems.append(
                        ValueError(
                            f"{side} {axis} boundary {names[k]!r} cannot be combined with "
                            f"a different boundary ({names[1 - k]!r})"
                        )
                    )
        return problems

    def validate(self) -> None:
        problems = self.collect_exceptions()
        if len(problems) == 1:
            raise problems[0]
        if problems:
            raise ExceptionGroup("Found multiple issues with boundary specifications", problems)
