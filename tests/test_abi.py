"""The C-ABI shared library: loads, exports what include/rlic_b200.h declares,
and refuses to compute without a GPU (no CPU fallback).  No compute calls here
unless a device is present."""

import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

from rlic_b200 import _core

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "rlic_b200.h").read_text()


def declared_symbols() -> list[str]:
    # every function prototype in the header: "<ret> rlic_b200_name("
    names = re.findall(r"^\s*(?:const\s+)?[A-Za-z_0-9]+\s*\*?\s*(rlic_b200_[a-z0-9_]+)\s*\(", HEADER, re.M)
    assert "rlic_b200_debug_force_wide_index" in names
    return sorted(set(names))


def test_header_declares_the_expected_surface():
    names = declared_symbols()
    for base in ("convolve", "convolve_checked", "convolve_device", "convolve_packed", "pack_field",
                 "pass_slab", "slab_pack_field", "slab_pad_texture", "slab_unpad_texture", "convolve_batch"):
        for sfx in ("f32", "f64"):
            assert f"rlic_b200_{base}_{sfx}" in names
    for misc in ("abi_version", "last_error", "device_count", "launch_count", "set_device", "padded_cells",
                 "result_alloc", "result_free"):
        assert f"rlic_b200_{misc}" in names


@pytest.mark.parametrize("symbol", declared_symbols())
def test_library_exports_every_declared_symbol(symbol):
    assert hasattr(_core.lib, symbol), f"{symbol} declared in the header but not exported"


def test_no_torch_or_cuda_types_in_signatures():
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)   # prototypes only, comments dropped
    for forbidden in ("at::", "torch", "cudaStream_t", "CUstream", "Tensor"):
        assert forbidden not in body


def test_abi_version_matches_binding():
    assert _core.lib.rlic_b200_abi_version() == _core.ABI_VERSION
    m = re.search(r"#define RLIC_B200_ABI_VERSION (\d+)", HEADER)
    assert int(m.group(1)) == _core.ABI_VERSION


def test_error_codes_match_header():
    for name, value in (("OK", 0), ("EINVAL", 1), ("ENODEVICE", 2), ("ECUDA", 3), ("ESHARD", 4)):
        m = re.search(rf"#define RLIC_B200_{name} (\d+)", HEADER)
        assert int(m.group(1)) == value == getattr(_core, name)


def test_bad_arguments_are_reported_not_fatal():
    # argument checks run before any CUDA call, so this works without a GPU
    p = ctypes.POINTER(ctypes.c_float)
    a = np.zeros((4, 4), dtype=np.float32)
    out = np.empty_like(a)
    k = np.ones(3, dtype=np.float32)
    ptr = lambda x: x.ctypes.data_as(p)  # noqa: E731
    f = _core.lib.rlic_b200_convolve_f32
    # empty kernel: the reference aborts the interpreter here (lib.rs:371,378)
    assert f(ptr(a), ptr(a), ptr(a), 4, 4, ptr(k), 0, 0, 0, 0, 0, 0, 1, ptr(out)) == _core.EINVAL
    assert b"empty" in _core.lib.rlic_b200_last_error()
    assert f(ptr(a), ptr(a), ptr(a), 4, 4, ptr(k), 3, 7, 0, 0, 0, 0, 1, ptr(out)) == _core.EINVAL
    assert f(ptr(a), ptr(a), ptr(a), 4, 4, ptr(k), 3, 0, 0, 9, 0, 0, 1, ptr(out)) == _core.EINVAL
    assert f(ptr(a), ptr(a), ptr(a), -1, 4, ptr(k), 3, 0, 0, 0, 0, 0, 1, ptr(out)) == _core.EINVAL
    with pytest.raises(ValueError, match="empty convolution kernel"):
        _core.convolve_f32(a, (a, a, "velocity"), np.ones(0, dtype=np.float32),
                           (("closed", "closed"), ("closed", "closed")), 1)


def test_core_rejects_wrong_dtype_like_the_pyo3_signature():
    a64 = np.zeros((4, 4))
    k64 = np.ones(3)
    walls = (("closed", "closed"), ("closed", "closed"))
    with pytest.raises(TypeError, match="texture"):
        _core.convolve_f32(a64, (a64, a64, "velocity"), k64, walls, 1)
    with pytest.raises(TypeError, match="kernel"):
        _core.convolve_f64(a64, (a64, a64, "velocity"), np.ones((3, 3)), walls, 1)


def test_empty_images_need_no_device():
    walls = (("closed", "closed"), ("closed", "closed"))
    for shape in ((0, 5), (5, 0), (0, 0)):
        a = np.zeros(shape, dtype=np.float32)
        out = _core.convolve_f32(a, (a, a, "velocity"), np.ones(3, dtype=np.float32), walls, 2)
        assert out.shape == shape and out.dtype == np.float32


@pytest.mark.skipif(_core.device_count() > 0, reason="a GPU is present")
def test_without_a_gpu_the_product_fails_loudly():
    import rlic_b200

    img = np.random.default_rng(0).random((8, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rlic_b200.convolve(img, img, img, kernel=np.ones(3))


def test_result_blocks_fall_back_without_a_device_and_recycle_with_one():
    import gc

    first = _core.new_result((600, 700), np.float32)  # 1.6 MB: above the pinned threshold
    assert first.flags.owndata                        # a size seen for the first time stays on the heap
    a = _core.new_result((600, 700), np.float32)
    assert a.shape == (600, 700) and a.dtype == np.float32 and a.flags.c_contiguous and a.flags.writeable
    a[:] = 3.0
    assert float(a.sum()) == 3.0 * 600 * 700
    if _core.device_count() == 0:
        assert a.flags.owndata                       # ordinary memory: the pool is unavailable
        return
    assert not a.flags.owndata                       # second request of the size: page-locked block
    ptr = a.ctypes.data
    view = a[10:20]                                  # views keep the block alive
    del a
    gc.collect()
    assert view[0, 0] == 3.0
    del view
    gc.collect()
    b = _core.new_result((600, 700), np.float32)     # same size class: the block comes back
    assert b.ctypes.data == ptr


def test_product_never_imports_the_oracle():
    # the oracle is test infrastructure; the package must not reference it
    pkg = ROOT / "rlic_b200"
    for path in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")):
        text = path.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), path
        assert "liblic_oracle" not in text, path


def test_arithmetic_selection_round_trips_and_rejects_unknown_builds():
    import os
    import subprocess
    import sys

    import rlic_b200

    for name, code in _core.ARITHMETICS.items():
        m = re.search(rf"#define RLIC_B200_ARITH_{name.upper().replace('+', '_')} (\d+)", HEADER)
        assert int(m.group(1)) == code
    assert rlic_b200.get_arithmetic() == "fma+branchless"          # the crate default
    try:
        rlic_b200.set_arithmetic("fma")
        assert rlic_b200.get_arithmetic() == "fma"
        with pytest.raises(ValueError, match="unknown arithmetic"):
            rlic_b200.set_arithmetic("branchless")
        assert _core.lib.rlic_b200_set_arithmetic(7) == _core.EINVAL
        assert rlic_b200.get_arithmetic() == "fma"
    finally:
        rlic_b200.set_arithmetic("fma+branchless")
    # the environment variable is read when the library is first loaded
    code = "import rlic_b200; print(rlic_b200.get_arithmetic())"
    env = dict(os.environ, RLIC_B200_ARITHMETIC="fma", PYTHONPATH=str(ROOT))
    assert subprocess.run([sys.executable, "-c", code], env=env, capture_output=True,
                          text=True).stdout.strip() == "fma"
    env["RLIC_B200_ARITHMETIC"] = "nonsense"
    bad = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    assert bad.returncode != 0 and "RLIC_B200_ARITHMETIC" in bad.stderr


def test_schedule_selection_round_trips():
    import rlic_b200

    for name, code in _core.SCHEDULES.items():
        m = re.search(rf"#define RLIC_B200_SCHEDULE_{name.upper()} (\d+)", HEADER)
        assert int(m.group(1)) == code
    assert rlic_b200.get_schedule() == "wavefront"     # measured on a B200: the faster order
    try:
        rlic_b200.set_schedule("trailing")
        assert rlic_b200.get_schedule() == "trailing"
        with pytest.raises(ValueError, match="unknown schedule"):
            rlic_b200.set_schedule("diagonal")
        assert _core.lib.rlic_b200_set_schedule(9) == _core.EINVAL
    finally:
        rlic_b200.set_schedule("wavefront")


def test_walk_selection_round_trips_and_reads_the_environment():
    import os
    import subprocess
    import sys

    import rlic_b200

    for name, code in _core.WALKS.items():
        m = re.search(rf"#define RLIC_B200_WALK_{name.upper().replace('-', '_')} (\d+)", HEADER)
        assert int(m.group(1)) == code
    assert rlic_b200.get_walk() == "grouped"       # measured on a B200: the faster formulation
    try:
        rlic_b200.set_walk("per-step")
        assert rlic_b200.get_walk() == "per-step"
        with pytest.raises(ValueError, match="unknown walk"):
            rlic_b200.set_walk("sideways")
        assert _core.lib.rlic_b200_set_walk(5) == _core.EINVAL
        assert rlic_b200.get_walk() == "per-step"
    finally:
        rlic_b200.set_walk("grouped")
    code = "import rlic_b200; print(rlic_b200.get_walk())"
    env = dict(os.environ, RLIC_B200_WALK="per-step", PYTHONPATH=str(ROOT))
    assert subprocess.run([sys.executable, "-c", code], env=env, capture_output=True,
                          text=True).stdout.strip() == "per-step"
    env["RLIC_B200_WALK"] = "nonsense"
    bad = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    assert bad.returncode != 0 and "RLIC_B200_WALK" in bad.stderr


def test_thread_options_override_the_defaults_for_one_thread_only():
    """rlic_b200.options: a thread's choices for its own calls (rlic_b200_set_thread_options);
    other threads and the process-wide defaults are untouched, and -1 / None inherit."""
    import threading

    import rlic_b200

    base = rlic_b200.effective_options()
    assert base == {"arithmetic": "fma+branchless", "schedule": "wavefront", "walk": "grouped", "paths": "replay"}
    seen = {}

    def other_thread():
        seen["other"] = rlic_b200.effective_options()
        with rlic_b200.options(walk="per-step"):
            seen["other_inside"] = rlic_b200.effective_options()

    with rlic_b200.options(arithmetic="fma", schedule="trailing"):
        mine = rlic_b200.effective_options()
        t = threading.Thread(target=other_thread)
        t.start()
        t.join()
        with rlic_b200.options(walk="per-step"):          # nests: keeps the outer overrides
            nested = rlic_b200.effective_options()
        assert rlic_b200.effective_options() == mine
    assert mine == {"arithmetic": "fma", "schedule": "trailing", "walk": "grouped", "paths": "replay"}
    assert nested == {"arithmetic": "fma", "schedule": "trailing", "walk": "per-step", "paths": "replay"}
    assert seen["other"] == base
    assert seen["other_inside"] == dict(base, walk="per-step")
    assert rlic_b200.effective_options() == base
    assert rlic_b200.get_arithmetic() == "fma+branchless"   # the defaults never moved
    with pytest.raises(ValueError, match="unknown walk"):
        rlic_b200.options(walk="sideways")
    assert _core.lib.rlic_b200_set_thread_options(7, -1, -1) == _core.EINVAL
