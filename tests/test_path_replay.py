"""Recorded streamline paths: the first pass of a call records, the others replay.

Which pixels a streamline visits depends on the vector field, the mode and the boundaries,
never on the texture (/root/reference/src/lib.rs:305-362; the texture enters at :353-360
only), and the reference hands the same ``u, v`` to every iteration (lib.rs:432-440).  The
library therefore walks once per call and replays the recorded moves in passes 2..n
(``RLIC_B200_PATHS_REPLAY``, include/rlic_b200.h; ``PathPlanes`` in lic_walk.cuh).

CPU half (``-m "not gpu"``): the kernel source itself -- the recording walk and the replay
kernel -- compiled for the CPU (tests/kernel_emulation) against the oracle, which walks every
pass as the reference does.  GPU half (``-m gpu``): the shipped binary through the public API
and the C ABI, against the oracle and against the library's own ``paths="recompute"``.
"""
from __future__ import annotations

import ctypes
import threading

import numpy as np
import pytest
from numpy.testing import assert_array_equal

import kernel_emulation as ke
import oracle
import rlic_b200
from rlic_b200 import _core, workloads
from test_kernel_emulation import WALLS, fuzz_case, random_case


def emulated(tex, u, v, kernel, mode="velocity", walls="closed", iterations=2, **how):
    """record + replay through the emulated kernels -- both replay kernels: the one that gathers
    through L1 and, where the library would launch it, the one with the texture window staged in
    shared memory -- checked against the oracle"""
    bnd = WALLS[walls]
    want = oracle.convolve(np.ascontiguousarray(tex), np.ascontiguousarray(u), np.ascontiguousarray(v),
                           kernel=kernel, uv_mode=mode, boundaries=bnd, iterations=iterations, variant=3)
    for paths in (True, "staged"):
        got = ke.convolve(tex, u, v, kernel=kernel, uv_mode=mode, boundaries=bnd, iterations=iterations, paths=paths,
                          **how)
        assert got.dtype == tex.dtype and got.shape == tex.shape
        assert_array_equal(got, want)
    return got


# ---------------------------------------------------------------------------------- CPU
@pytest.mark.parametrize("walls", WALLS)
@pytest.mark.parametrize("mode", ["velocity", "polarization"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_replay_with_special_pixels(dtype, mode, walls):
    # zero vectors (the walker stays), NaN components (the walk ends), signed zeros
    emulated(*random_case((45, 70), dtype, 23, seed=11), mode=mode, walls=walls, iterations=3)


@pytest.mark.parametrize("klen", [1, 2, 3, 4, 5, 8, 33, 64, 65, 66, 67, 128, 129, 130, 131, 200, 257])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_replay_kernel_lengths(dtype, klen):
    """One group per half (<= 65 taps: the unrolled kernel), two (<= 129), more (the loop); even
    kernels have halves of different lengths; halves that end inside a group."""
    emulated(*random_case((19, 21), dtype, klen, seed=klen), mode="polarization", walls="x-periodic")
    emulated(*random_case((33, 30), dtype, klen, seed=klen + 1), walls="periodic", iterations=3)


def test_replay_taps_beyond_the_parameter_block():
    emulated(*random_case((9, 12), np.float32, 1001, seed=5), walls="periodic")
    emulated(*random_case((9, 12), np.float64, 500, seed=6), mode="polarization", walls="closed")


@pytest.mark.parametrize(
    "shape", [(1, 1), (1, 40), (40, 1), (2, 2), (8, 32), (9, 33), (7, 31), (16, 64), (17, 65), (64, 3)]
)
def test_replay_degenerate_and_tile_edge_shapes(shape):
    for walls in ("closed", "periodic"):
        emulated(*random_case(shape, np.float64, 9, seed=sum(shape), specials=False), walls=walls,
                 mode="polarization", iterations=3)


def test_replay_uniform_axis_aligned_and_stagnant_fields():
    """Walkers pressed against a closed wall re-sample the edge pixel step after step (every
    step a wall crossing), zero fields never move, -0.0 components take the generic step."""
    rng = np.random.default_rng(3)
    tex = rng.random((40, 50))
    one, zero = np.ones_like(tex), np.zeros_like(tex)
    k = np.linspace(0.1, 1, 15)
    for u, v in ((one, zero), (zero, one), (-one, zero), (zero, -one), (one, one), (-one, one),
                 (zero, zero), (-zero, zero), (one, -one)):
        for walls in WALLS:
            emulated(tex, u, v, k, walls=walls)
            emulated(tex, u, v, k, mode="polarization", walls=walls, iterations=3)


def test_replay_all_nan_field_and_nan_texture():
    tex, u, v, k = random_case((20, 24), np.float64, 9, seed=4)
    emulated(tex, np.full_like(u, np.nan), v, k, iterations=3)      # every walk ends at once
    tex[5, 5] = np.nan
    k[2] = -3.0
    emulated(tex, u, v, k, iterations=3)                             # NaN spreads exactly as in the oracle


def test_replay_infinite_huge_and_denormal_velocities():
    tex, u, v, k = random_case((24, 24), np.float32, 13, seed=8)
    u[7, 7] = np.inf
    v[8, 8] = -np.inf
    u[9, 9] = 3e38
    v[9, 9] = -3e38
    u[10, 10] = 1e-45
    v[11, 11] = -1e-42
    emulated(tex, u, v, k, walls="periodic", iterations=3)
    emulated(tex, u, v, k, mode="polarization")


@pytest.mark.parametrize("seed", range(32))
def test_replay_fuzz(seed):
    tex, u, v, kernel, mode, walls, iterations = fuzz_case(seed)
    emulated(tex, u, v, kernel, mode=mode, walls=walls, iterations=max(2, iterations))


def test_replay_c1_and_reduced_c2_c3():
    for w in (workloads.readme_example(), workloads.vortex_noise(192), workloads.polarization_split(128)):
        got = ke.convolve(w.texture, w.u, w.v, kernel=w.kernel, uv_mode=w.uv_mode, boundaries=_bnd(w.boundaries),
                          iterations=3, paths=True)
        want = oracle.convolve(np.ascontiguousarray(w.texture), np.ascontiguousarray(w.u),
                               np.ascontiguousarray(w.v), kernel=w.kernel, uv_mode=w.uv_mode,
                               boundaries=_bnd(w.boundaries), iterations=3)
        assert_array_equal(got, want)


def _bnd(spec):
    from rlic_b200._boundaries import BoundarySet

    bs = BoundarySet.from_spec(spec)
    return (bs.x, bs.y)


def test_replay_with_64_bit_indices():
    emulated(*random_case((23, 37), np.float32, 21, seed=2), walls="periodic", iterations=3, wide=True)
    emulated(*random_case((23, 37), np.float64, 70, seed=3), mode="polarization", walls="y-periodic", wide=True)


def test_replay_of_a_batch_of_fields():
    """Stacked fields (the batch entry points): each has its own guard rows, and a record."""
    rng = np.random.default_rng(12)
    nf, ny, nx, klen = 3, 20, 27, 17
    tex = rng.random((nf, ny, nx)).astype(np.float32)
    u = (rng.random((nf, ny, nx)) - 0.5).astype(np.float32)
    v = (rng.random((nf, ny, nx)) - 0.5).astype(np.float32)
    u[1, 4, 4] = np.nan
    u[2, 7, 7] = v[2, 7, 7] = 0
    k = (rng.random(klen) + 0.1).astype(np.float32)
    for walls in ("closed", "periodic"):
        b = ke.Buffers(np.float32, ny, nx, _core.wall_codes(WALLS[walls]), klen, nfields=nf)
        b.pack_field(u, v)
        b.pad_texture(tex, 0)
        rec = b.path_record(klen)
        src = 0
        for it, how in enumerate((1, 2, 3)):                 # record, replay through L1, staged replay
            assert b.run_pass_paths(src, 1 - src, k, "velocity", how, rec)
            src = 1 - src
        got = b.unpad_texture(src)
        for f in range(nf):
            want = oracle.convolve(tex[f], u[f], v[f], kernel=k, boundaries=WALLS[walls], iterations=3)
            assert_array_equal(got[f], want)


def test_record_in_bands_replay_in_other_bands():
    """The host path records band by band (pass 1 trails the uploads) and replays in whatever
    row ranges its schedule likes: a pixel's record does not depend on the launch that wrote it."""
    tex, u, v, k = random_case((50, 41), np.float32, 19, seed=21)
    b = ke.Buffers(np.float32, 50, 41, _core.wall_codes(WALLS["closed"]), 19)
    b.pack_field(u, v)
    b.pad_texture(tex, 0)
    rec = b.path_record(19)
    for rows in ((0, 16), (16, 16), (32, 18)):
        b.run_pass_paths(0, 1, k, "velocity", 1, rec, rows=rows)
    for rows in ((0, 7), (7, 40), (47, 3)):
        b.run_pass_paths(1, 0, k, "velocity", 2, rec, rows=rows)
    for rows in ((0, 21), (21, 5), (26, 24)):                # the staged kernel, tiles cut by the row ranges
        assert b.run_pass_paths(0, 1, k, "velocity", 3, rec, rows=rows)
    want = oracle.convolve(tex, u, v, kernel=k, boundaries=WALLS["closed"], iterations=3)
    assert_array_equal(b.unpad_texture(1), want)


def test_the_record_does_not_depend_on_the_texture():
    """The claim the whole scheme rests on, checked on the record itself."""
    tex, u, v, k = random_case((30, 30), np.float64, 25, seed=9)
    records = []
    for texture in (tex, np.full_like(tex, np.nan), np.zeros_like(tex)):
        b = ke.Buffers(np.float64, 30, 30, _core.wall_codes(WALLS["periodic"]), 25)
        b.pack_field(u, v)
        b.pad_texture(texture, 0)
        rec = b.path_record(25)
        b.run_pass_paths(0, 1, k, "polarization", 1, rec)
        records.append(rec)
    assert_array_equal(records[0], records[1])
    assert_array_equal(records[0], records[2])


def test_record_size_and_layout():
    # 16 bytes (four planes) per cell and group of 32 steps, forward groups then backward groups
    cells = _core.padded_cells(10, 20)
    assert _core.path_record_bytes(10, 20, 65) == 2 * 16 * cells          # 32 + 32 steps
    assert _core.path_record_bytes(10, 20, 64) == 2 * 16 * cells          # 31 + 32
    assert _core.path_record_bytes(10, 20, 66) == 3 * 16 * cells          # 32 + 33
    assert _core.path_record_bytes(10, 20, 1) == 0
    assert _core.path_record_bytes(10, 20, 2) == 1 * 16 * cells           # 0 + 1
    assert _core.path_record_bytes(10, 20, 129) == 4 * 16 * cells
    # a uniform +x field, closed walls: every step of the forward half moves along x upwards
    tex = np.random.default_rng(1).random((6, 80))
    b = ke.Buffers(np.float64, 6, 80, _core.wall_codes(WALLS["closed"]), 9)
    b.pack_field(np.ones_like(tex), np.zeros_like(tex))
    b.pad_texture(tex, 0)
    rec = b.path_record(9)
    b.run_pass_paths(0, 1, np.ones(9), "velocity", 1, rec)
    entries = rec.reshape(2, -1, 4)                                           # (group, cell, plane)
    pitch = 82
    cell = pitch + 2 * pitch + 40                                             # pixel (2, 40): far from the walls
    axis, sign, rare, extra = (int(x) for x in entries[0, cell])              # forward: +x, +x, +x, +x
    assert (axis, sign, rare, extra) == (0, 0, 0, 0)
    axis, sign, rare, extra = (int(x) for x in entries[1, cell])              # backward: -x four times
    assert (axis, sign, rare, extra) == (0, 0xF0000000, 0, 0)
    # pixel (2, 79) at the closed right wall: every forward step crosses the wall again (RARE from
    # the second step on: the step that finds the walker on the wall cell)
    axis, sign, rare, extra = (int(x) for x in entries[0, pitch + 2 * pitch + 79])
    assert (axis, sign, rare, extra) == (0, 0, 0x70000000, 0)


def test_options_surface():
    assert rlic_b200.get_paths() in ("replay", "recompute")
    assert "paths" in rlic_b200.effective_options()
    with rlic_b200.options(paths="recompute"):
        assert rlic_b200.effective_options()["paths"] == "recompute"
        with rlic_b200.options(arithmetic="fma"):       # the other overrides leave it alone
            assert rlic_b200.effective_options()["paths"] == "recompute"
    assert rlic_b200.effective_options()["paths"] == rlic_b200.get_paths()
    with pytest.raises(ValueError):
        rlic_b200.options(paths="sometimes")
    seen = {}

    def other():
        seen["other"] = rlic_b200.effective_options()["paths"]

    with rlic_b200.options(paths="recompute"):
        t = threading.Thread(target=other)
        t.start()
        t.join()
    assert seen["other"] == rlic_b200.get_paths()        # a thread's override is its own


# ---------------------------------------------------------------------------------- GPU
def gpu_both(tex, u, v, kernel, mode="velocity", walls="closed", iterations=3):
    bnd = {"x": WALLS[walls][0][0], "y": WALLS[walls][1][0]}
    with rlic_b200.options(paths="replay"):
        got = rlic_b200.convolve(tex, u, v, kernel=kernel, uv_mode=mode, boundaries=bnd, iterations=iterations)
    with rlic_b200.options(paths="recompute"):
        walked = rlic_b200.convolve(tex, u, v, kernel=kernel, uv_mode=mode, boundaries=bnd, iterations=iterations)
    assert_array_equal(got, walked)
    want = oracle.convolve(np.ascontiguousarray(tex), np.ascontiguousarray(u), np.ascontiguousarray(v),
                           kernel=kernel, uv_mode=mode, boundaries=WALLS[walls], iterations=iterations)
    assert_array_equal(got, want)
    return got


@pytest.mark.gpu
@pytest.mark.parametrize("walls", WALLS)
@pytest.mark.parametrize("mode", ["velocity", "polarization"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_gpu_replay_with_special_pixels(dtype, mode, walls):
    tex, u, v, k = random_case((45, 70), dtype, 23, seed=11)
    gpu_both(np.abs(tex), u, v, k, mode=mode, walls=walls)


@pytest.mark.gpu
@pytest.mark.parametrize("klen", [1, 2, 3, 8, 33, 64, 65, 66, 129, 130, 200, 257, 1001])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_gpu_replay_kernel_lengths(dtype, klen):
    tex, u, v, k = random_case((33, 30), dtype, klen, seed=klen)
    gpu_both(np.abs(tex), u, v, k, mode="polarization", walls="x-periodic")
    gpu_both(np.abs(tex), u, v, k, walls="periodic", iterations=2)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(16))
def test_gpu_replay_fuzz(seed):
    tex, u, v, kernel, mode, walls, iterations = fuzz_case(seed)
    gpu_both(np.abs(tex), u, v, kernel, mode=mode, walls=walls, iterations=max(2, iterations))


@pytest.mark.gpu
def test_gpu_replay_stagnant_and_nan_fields():
    rng = np.random.default_rng(3)
    tex = rng.random((40, 50)).astype(np.float32)
    one, zero = np.ones_like(tex), np.zeros_like(tex)
    k = np.linspace(0.1, 1, 15).astype(np.float32)
    for u, v in ((one, zero), (zero, -one), (zero, zero), (-one, one), (np.full_like(tex, np.nan), one)):
        for walls in WALLS:
            gpu_both(tex, u, v, k, walls=walls)
            gpu_both(tex, u, v, k, mode="polarization", walls=walls)


@pytest.mark.gpu
def test_gpu_replay_large_image_both_schedules_and_device_entry():
    """A banded host call (the wavefront and the trailing schedule record band by band), the
    device entry, and many iterations: all equal to walking every pass."""
    w = workloads.vortex_noise(2048, rows=1536)
    args = dict(kernel=w.kernel, uv_mode=w.uv_mode, boundaries=w.boundaries)
    with rlic_b200.options(paths="recompute"):
        want = rlic_b200.convolve(w.texture, w.u, w.v, iterations=4, **args)
    for schedule in ("wavefront", "trailing"):
        with rlic_b200.options(paths="replay", schedule=schedule):
            got = rlic_b200.convolve(w.texture, w.u, w.v, iterations=4, **args)
        assert_array_equal(got, want)
    import torch
    from rlic_b200 import device

    t, u, v = (torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (w.texture, w.u, w.v))
    with rlic_b200.options(paths="replay"):
        got = device.convolve_device(t, u, v, iterations=4, **args).cpu().numpy()
    assert_array_equal(got, want)
    band = oracle.pass_rows(w.texture, w.u, w.v, kernel=w.kernel, rows=(700, 764), uv_mode=w.uv_mode,
                            boundaries=_bnd(w.boundaries))
    with rlic_b200.options(paths="replay"):
        once = rlic_b200.convolve(w.texture, w.u, w.v, iterations=1, **args)
    assert_array_equal(once[700:764], band)


@pytest.mark.gpu
def test_gpu_replay_through_the_raw_slab_abi():
    """rlic_b200_pass_slab_paths_* on caller-owned buffers: record once, replay twice."""
    import torch

    ny, nx, klen = 96, 80, 33
    tex, u, v, k = random_case((ny, nx), np.float32, klen, seed=31)
    tex = np.abs(tex)
    walls = _core.wall_codes(WALLS["x-periodic"])
    cells = _core.padded_cells(ny, nx)
    dev = torch.device("cuda")
    d_u, d_v, d_t = (torch.from_numpy(a).to(dev) for a in (u, v, tex))
    field = torch.zeros(4 * cells, dtype=torch.float32, device=dev)
    bufs = [torch.zeros(cells, dtype=torch.float32, device=dev) for _ in range(2)]
    rec = torch.empty(_core.path_record_bytes(ny, nx, klen) // 4, dtype=torch.int32, device=dev)
    s = int(torch.cuda.current_stream().cuda_stream)
    slab = (ny, nx, 0, ny, 0, 0)
    lib = _core.lib
    _core.check(lib.rlic_b200_slab_pack_field_f32(d_u.data_ptr(), d_v.data_ptr(), *slab, *walls, field.data_ptr(), s))
    _core.check(lib.rlic_b200_slab_pad_texture_f32(d_t.data_ptr(), *slab, *walls, bufs[0].data_ptr(), s))
    kp = k.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    for it in range(3):
        src, dst = bufs[it % 2], bufs[(it + 1) % 2]
        pieces = ((0, 32), (32, 64)) if it < 2 else ((0, 96),)
        for a, n in pieces:
            _core.check(lib.rlic_b200_pass_slab_paths_f32(
                src.data_ptr(), field.data_ptr() if it == 0 else None, dst.data_ptr(), *slab, a, n, kp, klen, 0,
                *walls, None, 0, _core.PASS_RECORD if it == 0 else _core.PASS_REPLAY, rec.data_ptr(), s))
    out = torch.empty((ny, nx), dtype=torch.float32, device=dev)
    _core.check(lib.rlic_b200_slab_unpad_texture_f32(bufs[1].data_ptr(), *slab, *walls, out.data_ptr(), s))
    want = oracle.convolve(tex, u, v, kernel=k, boundaries=WALLS["x-periodic"], iterations=3)
    assert_array_equal(out.cpu().numpy(), want)
    # the `fma` arithmetic has no recording walk: refused, not silently walked
    with rlic_b200.options(arithmetic="fma"):
        rc = lib.rlic_b200_pass_slab_paths_f32(
            bufs[0].data_ptr(), field.data_ptr(), bufs[1].data_ptr(), *slab, 0, ny, kp, klen, 0, *walls, None, 0,
            _core.PASS_RECORD, rec.data_ptr(), s)
    assert rc == _core.EINVAL
    torch.cuda.synchronize()


@pytest.mark.gpu
def test_gpu_replay_batch_and_concurrent_calls():
    rng = np.random.default_rng(5)
    nf, ny, nx = 6, 64, 48
    tex = rng.random((nf, ny, nx)).astype(np.float32)
    u = (rng.random((nf, ny, nx)) - 0.5).astype(np.float32)
    v = (rng.random((nf, ny, nx)) - 0.5).astype(np.float32)
    k = workloads.triangle_kernel(33, np.float32)
    got = rlic_b200.convolve_batch(tex, u, v, kernel=k, iterations=3)
    for f in range(nf):
        assert_array_equal(got[f], oracle.convolve(tex[f], u[f], v[f], kernel=k, iterations=3))
    results = {}

    def call(name, paths):
        with rlic_b200.options(paths=paths):
            results[name] = rlic_b200.convolve(tex[0], u[0], v[0], kernel=k, iterations=4)

    threads = [threading.Thread(target=call, args=(i, "replay" if i % 2 else "recompute")) for i in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for i in range(1, 6):
        assert_array_equal(results[i], results[0])
