"""The padded device layout (include/rlic_b200.h), checked on
the CPU through the library's host-side testing hook: which cells are pixels,
which are wall cells, and that every wall cell points to exactly the pixel the
reference's boundary rule (src/lib.rs:83-95) sends a walker to.  No GPU involved.
"""

import ctypes
from itertools import product

import pytest

from rlic_b200 import _core

CLOSED, PERIODIC = 0, 1


def wall_cell(ny, nx, slab, walls, cell):
    out = (ctypes.c_int64 * 5)()
    rc = _core.lib.rlic_b200_debug_wall_cell(ny, nx, *slab, *walls, cell, out)
    _core.check(rc)
    return dict(pixel=bool(out[0]), reachable=bool(out[1]), row=out[2], col=out[3], shift=out[4])


def reference_rule(c, size, left, right):
    """src/lib.rs:83-95 for a coordinate that stepped to -1 or to `size`."""
    if c == -1:
        return size - 1 if left == PERIODIC else 0
    if c == size:
        return 0 if right == PERIODIC else size - 1
    return c


WALLS = [(xl, xr, yl, yr) for (xl, xr), (yl, yr) in product(
    [(CLOSED, CLOSED), (PERIODIC, PERIODIC)], repeat=2)]


@pytest.mark.parametrize("walls", WALLS)
@pytest.mark.parametrize("shape", [(1, 1), (1, 5), (5, 1), (4, 6), (7, 3)])
def test_whole_image_wall_cells_follow_the_reference_rule(shape, walls):
    ny, nx = shape
    slab = (0, ny, 0, 0)
    P = nx + 2
    assert _core.padded_cells(ny, nx) == (ny + 2) * P

    def cell_of(r, c):           # buffer row r (-1 .. ny), column c (-1 .. nx)
        return (r + 1) * P + c

    # every pixel is a pixel and points to itself
    for r, c in product(range(ny), range(nx)):
        s = wall_cell(ny, nx, slab, walls, cell_of(r, c))
        assert s["pixel"] and (s["row"], s["col"], s["shift"]) == (r, c, 0)
    # stepping off any side from any edge pixel lands on a reachable wall cell whose
    # shift leads to the pixel the reference continues from
    for r in range(ny):
        for c_off in (-1, nx):
            s = wall_cell(ny, nx, slab, walls, cell_of(r, c_off))
            want = (r, reference_rule(c_off, nx, walls[0], walls[1]))
            assert not s["pixel"] and s["reachable"]
            assert (s["row"], s["col"]) == want
            assert cell_of(r, c_off) + s["shift"] == cell_of(*want)
    for c in range(nx):
        for r_off in (-1, ny):
            s = wall_cell(ny, nx, slab, walls, cell_of(r_off, c))
            want = (reference_rule(r_off, ny, walls[2], walls[3]), c)
            assert not s["pixel"] and s["reachable"]
            assert (s["row"], s["col"]) == want
            assert cell_of(r_off, c) + s["shift"] == cell_of(*want)


def test_right_and_left_wall_cells_are_distinct_memory():
    # cell (r, nx) is row r's right wall, cell (r, nx + 1) is row r + 1's left wall
    ny, nx = 3, 4
    P = nx + 2
    walls = (PERIODIC, PERIODIC, CLOSED, CLOSED)
    right_of_row0 = wall_cell(ny, nx, (0, ny, 0, 0), walls, 1 * P + nx)
    left_of_row1 = wall_cell(ny, nx, (0, ny, 0, 0), walls, 1 * P + nx + 1)
    assert (right_of_row0["row"], right_of_row0["col"]) == (0, 0)        # wraps to column 0
    assert (left_of_row1["row"], left_of_row1["col"]) == (1, nx - 1)     # wraps to the last column


@pytest.mark.parametrize("periodic_y", [False, True])
def test_slab_geometry(periodic_y):
    # image of 40 rows in three slabs with a reach of 4 rows
    ny, nx, h = 40, 5, 4
    P = nx + 2
    yw = PERIODIC if periodic_y else CLOSED
    walls = (CLOSED, CLOSED, yw, yw)
    slabs = [(0, 13), (13, 27), (27, 40)]
    for r0, r1 in slabs:
        lo = h if (r0 > 0 or periodic_y) else 0
        hi = h if (r1 < ny or periodic_y) else 0
        slab = (r0, r1 - r0, lo, hi)
        rows = lo + (r1 - r0) + hi
        top = wall_cell(ny, nx, slab, walls, 0 * P + 2)               # guard row above, column 2
        bottom = wall_cell(ny, nx, slab, walls, (rows + 1) * P + 2)   # guard row below
        # a guard row is a live wall only where the slab touches a closed image edge
        assert top["reachable"] == (r0 == 0 and not periodic_y)
        assert bottom["reachable"] == (r1 == ny and not periodic_y)
        if top["reachable"]:
            assert (top["row"], top["col"]) == (0, 2)                  # closed: stay in the first row
        if bottom["reachable"]:
            assert (bottom["row"], bottom["col"]) == (rows - 1, 2)
        # halo rows have their own x-wall cells like any other row
        s = wall_cell(ny, nx, slab, walls, (0 + 1) * P + nx)
        assert s["reachable"] and (s["row"], s["col"]) == (0, nx - 1)


def test_bad_requests_are_refused():
    out = (ctypes.c_int64 * 5)()
    f = _core.lib.rlic_b200_debug_wall_cell
    assert f(8, 8, 0, 8, 0, 0, 0, 0, 0, 0, 10**9, out) == _core.EINVAL       # cell outside the buffer
    assert f(8, 8, 4, 8, 0, 0, 0, 0, 0, 0, 0, out) == _core.ESHARD           # rows beyond the image
    assert f(8, 8, 0, 8, 0, 0, 0, 7, 0, 0, 0, out) == _core.EINVAL           # unknown boundary code


def test_band_plan_of_the_host_path():
    """The row bands of a host call (rlic_b200_debug_band_plan): they tile the image, every band
    is at least two kernel half-widths tall (a pass reaches one band up and down) and at least
    64 rows, inner edges sit on tile rows counted from the bottom, there are about sixteen of
    them, and small or narrow images are not cut at all."""
    import ctypes

    import numpy as np

    from rlic_b200 import _core

    def plan(ny, nx, klen, iterations):
        e = np.zeros(64, dtype=np.int64)
        n = _core.lib.rlic_b200_debug_band_plan(ny, nx, klen, iterations, e.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), 64)
        assert 2 <= n <= 64
        return e[:n]

    for ny, nx, klen, its in ((4096, 4096, 65, 5), (4096, 4096, 65, 1), (16384, 16384, 65, 20), (2048, 2048, 129, 1),
                              (1000, 3000, 65, 3), (4100, 4096, 65, 5), (3000, 5000, 257, 4), (46341, 46341, 65, 2)):
        e = plan(ny, nx, klen, its)
        sizes = np.diff(e)
        assert e[0] == 0 and e[-1] == ny and (sizes > 0).all()
        assert len(sizes) <= 17
        if len(sizes) > 1:
            assert sizes.min() >= max(2 * (klen // 2), 64)
            assert ((ny - e[1:-1]) % 16 == 0).all()
            assert (sizes[1:] == sizes[-1]).all()            # uniform; the top band takes the remainder
    assert np.diff(plan(4096, 4096, 65, 5)).tolist() == [256] * 16
    assert len(plan(256, 256, 65, 2)) == 2                   # too small to pipeline
    assert len(plan(100, 100000, 9, 3)) == 2                 # too few rows
    assert _core.lib.rlic_b200_debug_band_plan(0, 10, 5, 1, None, 0) == 0
