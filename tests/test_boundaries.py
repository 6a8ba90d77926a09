"""Boundary specification handling (no GPU involved).

Same properties as the reference pins in tests/test_boundaries.py:16-131,
re-expressed for ``rlic_b200._boundaries``.
"""

from dataclasses import replace
from itertools import permutations, product

import pytest

from rlic_b200._boundaries import (
    COMBO_ALLOWED_BOUNDS,
    COMBO_DISALLOWED_BOUNDS,
    SUPPORTED_BOUNDS,
    BoundarySet,
)

EXPANSIONS = {
    "single-name": ("a", BoundarySet(x=("a", "a"), y=("a", "a"))),
    "per-axis-names": ({"x": "a", "y": "b"}, BoundarySet(x=("a", "a"), y=("b", "b"))),
    "per-side-tuples": (
        {"x": ("a", "b"), "y": ("c", "w")},
        BoundarySet(x=("a", "b"), y=("c", "w")),
    ),
    "lists-become-tuples": (
        {"x": ["a", "b"], "y": ["c", "w"]},
        BoundarySet(x=("a", "b"), y=("c", "w")),
    ),
    "mixed": ({"y": ("p", "q"), "x": "r"}, BoundarySet(x=("r", "r"), y=("p", "q"))),
}


@pytest.mark.parametrize("case", EXPANSIONS)
def test_spec_expansion(case):
    spec, expected = EXPANSIONS[case]
    assert BoundarySet.from_spec(spec) == expected


@pytest.mark.parametrize(
    "spec",
    [
        123,
        None,
        ["a", "b"],
        ("a", "b"),
        {"x": "a"},
        {"y": "b"},
        {"x": "a", "y": "b", "z": "c"},
        {"x": "a", "z": "c"},
        {"x": ("a",), "y": "b"},
        {"x": ("a", "b", "c"), "y": "b"},
        {"x": ("a", 1), "y": "b"},
        {"x": 1, "y": "b"},
        {"x": b"ab", "y": "b"},
    ],
)
def test_spec_with_wrong_shape_is_rejected(spec):
    assert BoundarySet.from_spec(spec) is None


@pytest.mark.parametrize("bx, by", list(product(sorted(SUPPORTED_BOUNDS), repeat=2)))
def test_same_kind_on_both_sides_is_fine(bx, by):
    bs = BoundarySet(x=(bx, bx), y=(by, by))
    assert bs.collect_exceptions() == []
    assert bs.validate() is None


def test_allowed_combination_sets_are_consistent():
    assert COMBO_ALLOWED_BOUNDS == {"closed"}
    assert COMBO_DISALLOWED_BOUNDS == {"periodic"}
    assert SUPPORTED_BOUNDS == {"closed", "periodic"}


@pytest.mark.parametrize(
    "picky, other",
    [(a, b) for a, b in product(sorted(COMBO_DISALLOWED_BOUNDS), sorted(SUPPORTED_BOUNDS)) if a != b],
)
def test_periodic_needs_a_periodic_partner(picky, other):
    base = BoundarySet(x=(picky, other), y=(picky, picky))
    with pytest.raises(ValueError, match=rf"^left x boundary '{picky}' cannot be combined"):
        base.validate()
    # any 2-item iterable is accepted for a side pair, as with the reference
    with pytest.raises(ValueError, match=rf"^right x boundary '{picky}' cannot be combined"):
        replace(base, x=reversed(base.x)).validate()
    flipped = BoundarySet(x=base.y, y=base.x)
    with pytest.raises(ValueError, match=rf"^left y boundary '{picky}' cannot be combined"):
        flipped.validate()
    with pytest.raises(ValueError, match=rf"^right y boundary '{picky}' cannot be combined"):
        replace(flipped, y=reversed(flipped.y)).validate()


def test_combination_message_is_complete():
    bs = BoundarySet(x=("periodic", "closed"), y=("closed", "closed"))
    (err,) = bs.collect_exceptions()
    assert str(err) == (
        "left x boundary 'periodic' cannot be combined with a different boundary ('closed')"
    )


@pytest.mark.parametrize(
    "names", sorted(set(permutations(("periodic", "periodic", "periodic", "unknown"))))
)
def test_unknown_name_is_reported_alone(names):
    b1, b2, b3, b4 = names
    with pytest.raises(ValueError, match=r"^Unknown (left|right) (x|y) boundary 'unknown'$"):
        BoundarySet(x=(b1, b2), y=(b3, b4)).validate()


def test_several_problems_are_grouped_in_order():
    bs = BoundarySet(x=("unknown", "periodic"), y=("periodic", "closed"))
    with pytest.RaisesGroup(
        pytest.RaisesExc(ValueError, match=r"^Unknown left x boundary 'unknown'$"),
        pytest.RaisesExc(ValueError, match=r"^left y boundary 'periodic' cannot be combined"),
        match="Found multiple issues with boundary specifications",
    ):
        bs.validate()
    assert [str(e) for e in bs.collect_exceptions()] == [
        "Unknown left x boundary 'unknown'",
        "left y boundary 'periodic' cannot be combined with a different boundary ('closed')",
    ]
