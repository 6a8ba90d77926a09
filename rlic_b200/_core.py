"""ctypes binding of the C ABI (``include/rlic_b200.h``) with the call shapes of
the reference's native module ``rlic._core``
(``/root/reference/src/lib.rs:451-482``, ``src/rlic/_core.pyi:8-29``):

    convolve_f32(texture, (u, v, uv_mode), kernel, ((xl, xr), (yl, yr)), iterations)
    convolve_f64(...)

The shared library is loaded on first use (so that the package can be imported
to build it); a missing or stale library is an ImportError and a missing GPU a
RuntimeError at call time.  Nothing here computes on the CPU.
"""

from __future__ import annotations

__all__ = [
    "convolve_batch",
    "convolve_f32",
    "convolve_f64",
    "device_count",
    "launch_count",
    "lib",
]

import ctypes
import os
from pathlib import Path

import numpy as np

_LIB_PATH = Path(__file__).resolve().parent / "librlic_b200.so"

OK, EINVAL, ENODEVICE, ECUDA, ESHARD = range(5)
ABI_VERSION = 4
ARITHMETICS = {"fma+branchless": 0, "fma": 1}   # RLIC_B200_ARITH_* in include/rlic_b200.h
SCHEDULES = {"trailing": 0, "wavefront": 1}      # RLIC_B200_SCHEDULE_*
WALKS = {"per-step": 0, "grouped": 1}            # RLIC_B200_WALK_*
PATHS = {"recompute": 0, "replay": 1}            # RLIC_B200_PATHS_*
PASS_WALK, PASS_RECORD, PASS_REPLAY = range(3)   # RLIC_B200_PASS_* (paths_mode of pass_slab_paths)

_MODE_CODE = {"velocity": 0, "polarization": 1}
_WALL_CODE = {"closed": 0, "periodic": 1}

_i64 = ctypes.c_int64
_int = ctypes.c_int
_vp = ctypes.c_void_p


def _signatures(real) -> dict[str, list]:
    """argtypes per entry point for one scalar type; mirrors include/rlic_b200.h."""
    p = ctypes.POINTER(real)
    walls = [_int] * 4
    slab = [_i64] * 6   # ny, nx, row0, nrows, halo_lo, halo_hi
    return {
        "convolve": [p, p, p, _i64, _i64, p, _i64, _int, *walls, _i64, p],
        "convolve_checked": [p, p, p, _i64, _i64, p, _i64, _int, *walls, _i64, p,
                             ctypes.POINTER(_int)],
        "convolve_batch": [p, p, p, _i64, _i64, _i64, p, _i64, _int, *walls, _i64,
                           ctypes.POINTER(_int), _int, p],
        "convolve_batch_checked": [p, p, p, _i64, _i64, _i64, p, _i64, _int, *walls, _i64,
                                   ctypes.POINTER(_int), _int, p, ctypes.POINTER(_int)],
        "convolve_device": [_vp, _vp, _vp, _i64, _i64, p, _i64, _int, *walls, _i64, _vp, _vp],
        "convolve_device_batch": [_vp, _vp, _vp, _i64, _i64, _i64, p, _i64, _int, *walls, _i64, _vp, _vp],
        "pack_field": [_vp, _vp, _i64, _i64, *walls, _vp, _vp],
        "convolve_packed": [_vp, _vp, _i64, _i64, p, _i64, _int, *walls, _i64, _vp, _vp],
        "slab_pack_field": [_vp, _vp, *slab, *walls, _vp, _vp],
        "slab_pad_texture": [_vp, *slab, *walls, _vp, _vp],
        "slab_unpad_texture": [_vp, *slab, *walls, _vp, _vp],
        "slab_pack_field_rows": [_vp, _vp, *slab, _i64, _i64, *walls, _vp, _vp],
        "slab_pad_texture_rows": [_vp, *slab, _i64, _i64, *walls, _vp, _vp],
        "slab_unpad_texture_rows": [_vp, *slab, _i64, _i64, *walls, _vp, _vp],
        "measure_gather_ceiling": [_vp, _vp, _vp, _i64, _i64, p, _i64, _int, _vp],
        "pass_slab": [_vp, _vp, _vp, *slab, _i64, _i64, p, _i64, _int, *walls, _vp],
        "pass_slab_peer": [_vp, _vp, _vp, *slab, _i64, _i64, p, _i64, _int, *walls, _vp, _i64, _vp],
        "pass_slab_paths": [_vp, _vp, _vp, *slab, _i64, _i64, p, _i64, _int, *walls, _vp, _i64, _int, _vp, _vp],
    }


def _load() -> ctypes.CDLL:
    if not _LIB_PATH.exists():
        raise ImportError(
            f"{_LIB_PATH} is missing: build it with `python -m rlic_b200._build` "
            "(nvcc, sm_100a). rlic_b200 has no CPU fallback."
        )
    cdll = ctypes.CDLL(os.fspath(_LIB_PATH))
    cdll.rlic_b200_abi_version.restype = _int
    if cdll.rlic_b200_abi_version() != ABI_VERSION:
        raise ImportError(f"{_LIB_PATH} has a different ABI version; rebuild it")
    cdll.rlic_b200_last_error.restype = ctypes.c_char_p
    cdll.rlic_b200_device_count.restype = _int
    cdll.rlic_b200_launch_count.restype = _i64
    cdll.rlic_b200_padded_cells.argtypes = [_i64, _i64]
    cdll.rlic_b200_padded_cells.restype = _i64
    cdll.rlic_b200_debug_wall_cell.argtypes = [_i64] * 6 + [_int] * 4 + [_i64, ctypes.POINTER(_i64)]
    cdll.rlic_b200_debug_wall_cell.restype = _int
    cdll.rlic_b200_debug_geometry.argtypes = [_i64] * 6 + [_int] * 4 + [_i64, ctypes.POINTER(_i64)]
    cdll.rlic_b200_debug_geometry.restype = _int
    cdll.rlic_b200_result_alloc.argtypes = [_i64]
    cdll.rlic_b200_result_alloc.restype = _vp
    cdll.rlic_b200_result_free.argtypes = [_vp]
    cdll.rlic_b200_result_free.restype = None
    cdll.rlic_b200_set_device.argtypes = [_int]
    cdll.rlic_b200_set_device.restype = _int
    cdll.rlic_b200_set_arithmetic.argtypes = [_int]
    cdll.rlic_b200_set_arithmetic.restype = _int
    cdll.rlic_b200_get_arithmetic.restype = _int
    cdll.rlic_b200_set_schedule.argtypes = [_int]
    cdll.rlic_b200_set_schedule.restype = _int
    cdll.rlic_b200_get_schedule.restype = _int
    cdll.rlic_b200_debug_wavefront_order.argtypes = [_i64, _i64, ctypes.POINTER(ctypes.c_int32), _i64]
    cdll.rlic_b200_debug_wavefront_order.restype = _i64
    cdll.rlic_b200_debug_band_plan.argtypes = [_i64, _i64, _i64, _i64, ctypes.POINTER(_i64), _i64]
    cdll.rlic_b200_debug_band_plan.restype = _i64
    # peer memory and flags of the fused halo exchange (rlic_b200/sharded.py)
    handle = ctypes.POINTER(ctypes.c_ubyte)
    cdll.rlic_b200_peer_alloc.argtypes = [_i64, ctypes.POINTER(_vp), handle]
    cdll.rlic_b200_peer_open.argtypes = [handle, ctypes.POINTER(_vp)]
    cdll.rlic_b200_peer_close.argtypes = [_vp]
    cdll.rlic_b200_peer_free.argtypes = [_vp]
    cdll.rlic_b200_peer_signal.argtypes = [_vp, ctypes.c_uint32, _vp]
    cdll.rlic_b200_peer_wait.argtypes = [_vp, ctypes.c_uint32, _i64, _vp, _vp]
    cdll.rlic_b200_peer_signal2.argtypes = [_vp, _vp, ctypes.c_uint32, _vp]
    cdll.rlic_b200_peer_wait4.argtypes = [_vp, ctypes.c_uint32] * 4 + [_i64, _vp, _vp]
    for name in ("alloc", "open", "close", "free", "signal", "wait", "signal2", "wait4"):
        getattr(cdll, f"rlic_b200_peer_{name}").restype = _int
    cdll.rlic_b200_set_thread_options.argtypes = [_int, _int, _int]
    cdll.rlic_b200_set_thread_options.restype = _int
    _pi = ctypes.POINTER(ctypes.c_int)
    cdll.rlic_b200_get_thread_options.argtypes = [_pi, _pi, _pi]
    cdll.rlic_b200_get_thread_options.restype = None
    cdll.rlic_b200_get_effective_options.argtypes = [_pi, _pi, _pi]
    cdll.rlic_b200_get_effective_options.restype = None
    cdll.rlic_b200_set_walk.argtypes = [_int]
    cdll.rlic_b200_set_walk.restype = _int
    cdll.rlic_b200_get_walk.restype = _int
    cdll.rlic_b200_set_paths.argtypes = [_int]
    cdll.rlic_b200_set_thread_paths.argtypes = [_int]
    for name in ("set_paths", "get_paths", "set_thread_paths", "get_thread_paths", "get_effective_paths"):
        getattr(cdll, f"rlic_b200_{name}").restype = _int
    cdll.rlic_b200_path_record_bytes.argtypes = [_i64, _i64, _i64]
    cdll.rlic_b200_path_record_bytes.restype = _i64
    cdll.rlic_b200_debug_replay_staging.argtypes = [_int]
    cdll.rlic_b200_debug_replay_staging.restype = None
    staging = os.environ.get("RLIC_B200_REPLAY_STAGING")
    if staging:
        if staging not in ("0", "1"):
            raise ImportError(f"RLIC_B200_REPLAY_STAGING={staging!r}: expected 0 or 1")
        cdll.rlic_b200_debug_replay_staging(int(staging))
    paths = os.environ.get("RLIC_B200_PATHS")
    if paths:
        if paths not in PATHS:
            raise ImportError(f"RLIC_B200_PATHS={paths!r}: expected one of {sorted(PATHS)}")
        cdll.rlic_b200_set_paths(PATHS[paths])
    walk = os.environ.get("RLIC_B200_WALK")
    if walk:
        if walk not in WALKS:
            raise ImportError(f"RLIC_B200_WALK={walk!r}: expected one of {sorted(WALKS)}")
        cdll.rlic_b200_set_walk(WALKS[walk])
    wanted = os.environ.get("RLIC_B200_SCHEDULE")
    if wanted:
        if wanted not in SCHEDULES:
            raise ImportError(f"RLIC_B200_SCHEDULE={wanted!r}: expected one of {sorted(SCHEDULES)}")
        cdll.rlic_b200_set_schedule(SCHEDULES[wanted])
    requested = os.environ.get("RLIC_B200_ARITHMETIC")
    if requested:
        if requested not in ARITHMETICS:
            raise ImportError(f"RLIC_B200_ARITHMETIC={requested!r}: expected one of {sorted(ARITHMETICS)}")
        cdll.rlic_b200_set_arithmetic(ARITHMETICS[requested])
    for sfx, real in (("f32", ctypes.c_float), ("f64", ctypes.c_double)):
        pr = ctypes.POINTER(real)
        eq = getattr(cdll, f"rlic_b200_equalize_histogram_{sfx}")
        eq.argtypes, eq.restype = [pr, _i64, _i64, _i64, pr], _int
        eqd = getattr(cdll, f"rlic_b200_equalize_histogram_device_{sfx}")
        eqd.argtypes, eqd.restype = [_vp, _i64, _i64, _i64, _vp, _vp], _int
        for name, argtypes in _signatures(real).items():
            fn = getattr(cdll, f"rlic_b200_{name}_{sfx}")
            fn.argtypes = argtypes
            fn.restype = _int
    return cdll


class _LazyLib:
    """Loads librlic_b200.so the first time an entry point is looked up."""

    _cdll = None

    def __getattr__(self, name):
        if _LazyLib._cdll is None:
            _LazyLib._cdll = _load()
        return getattr(_LazyLib._cdll, name)


lib = _LazyLib()


def set_arithmetic(name: str) -> None:
    """Choose which build of the reference the kernels reproduce bit for bit
    (include/rlic_b200.h): ``"fma+branchless"`` -- the crate default, i.e. source
    builds and the aarch64 wheels, the default here -- or ``"fma"``, the x86-64
    wheels.  The process-wide default; also settable with ``RLIC_B200_ARITHMETIC``.  For one
    thread's calls use ``options``."""
    try:
        code = ARITHMETICS[name]
    except KeyError:
        raise ValueError(f"unknown arithmetic {name!r}: expected one of {sorted(ARITHMETICS)}") from None
    check(lib.rlic_b200_set_arithmetic(code))


def get_arithmetic() -> str:
    code = int(lib.rlic_b200_get_arithmetic())
    return next(name for name, c in ARITHMETICS.items() if c == code)


def set_schedule(name: str) -> None:
    """How the host path orders uploads, passes and downloads of one large image
    (include/rlic_b200.h): ``"wavefront"`` (default since round 2) or ``"trailing"``.  Same
    results either way.  The process-wide default; also settable with ``RLIC_B200_SCHEDULE``.
    For one thread's calls use ``options``."""
    try:
        code = SCHEDULES[name]
    except KeyError:
        raise ValueError(f"unknown schedule {name!r}: expected one of {sorted(SCHEDULES)}") from None
    check(lib.rlic_b200_set_schedule(code))


def get_schedule() -> str:
    code = int(lib.rlic_b200_get_schedule())
    return next(name for name, c in SCHEDULES.items() if c == code)


def set_walk(name: str) -> None:
    """Which formulation of the pass kernels runs (include/rlic_b200.h): ``"grouped"``
    (default since round 2: fewer instructions per step, same bits, 1.43 against 1.59 ms per
    4096^2 x 65-tap pass on a B200) or ``"per-step"`` (round 1's kernels).  The process-wide
    default; also settable with ``RLIC_B200_WALK``.  For one thread's calls use ``options``."""
    try:
        code = WALKS[name]
    except KeyError:
        raise ValueError(f"unknown walk {name!r}: expected one of {sorted(WALKS)}") from None
    check(lib.rlic_b200_set_walk(code))


def get_walk() -> str:
    code = int(lib.rlic_b200_get_walk())
    return next(name for name, c in WALKS.items() if c == code)


def set_paths(name: str) -> None:
    """What the passes of a call after its first do (include/rlic_b200.h): ``"replay"`` (default)
    -- the first pass records which way every walker went at every step, the others replay
    the record (a streamline's path depends on the field, never on the texture; same bits,
    about a fifth of the instructions) -- or ``"recompute"``, every pass walks as the reference
    does.  The process-wide default; also settable with ``RLIC_B200_PATHS``.  For one thread's
    calls use ``options``."""
    try:
        code = PATHS[name]
    except KeyError:
        raise ValueError(f"unknown paths choice {name!r}: expected one of {sorted(PATHS)}") from None
    check(lib.rlic_b200_set_paths(code))


def get_paths() -> str:
    code = int(lib.rlic_b200_get_paths())
    return next(name for name, c in PATHS.items() if c == code)


def path_record_bytes(rows: int, nx: int, klen: int) -> int:
    """Bytes of the path record of a padded buffer holding `rows` x `nx` pixels, `klen` taps."""
    return int(lib.rlic_b200_path_record_bytes(rows, nx, klen))


class options:
    """Context manager: choices for the calls THIS THREAD makes inside the block, leaving the
    process-wide defaults (and every other thread) alone::

        with rlic_b200.options(arithmetic="fma"):
            out = rlic_b200.convolve(...)          # the bits of rLIC's x86-64 wheels

    ``arithmetic``, ``schedule``, ``walk``, ``paths`` take the names ``set_arithmetic`` /
    ``set_schedule`` / ``set_walk`` / ``set_paths`` take; ``None`` keeps whatever is in force.
    Backed by ``rlic_b200_set_thread_options`` and ``rlic_b200_set_thread_paths``
    (include/rlic_b200.h): nothing shared is written, so concurrent threads with different
    choices do not race."""

    def __init__(self, *, arithmetic: str | None = None, schedule: str | None = None,
                 walk: str | None = None, paths: str | None = None):
        def code(table, name, what):
            if name is None:
                return None
            if name not in table:
                raise ValueError(f"unknown {what} {name!r}: expected one of {sorted(table)}")
            return table[name]

        self._want = (code(ARITHMETICS, arithmetic, "arithmetic"), code(SCHEDULES, schedule, "schedule"),
                      code(WALKS, walk, "walk"))
        self._want_paths = code(PATHS, paths, "paths choice")
        self._saved = None
        self._saved_paths = -1

    def __enter__(self):
        saved = [ctypes.c_int(), ctypes.c_int(), ctypes.c_int()]
        lib.rlic_b200_get_thread_options(*(ctypes.byref(x) for x in saved))
        self._saved = tuple(x.value for x in saved)
        self._saved_paths = int(lib.rlic_b200_get_thread_paths())
        check(lib.rlic_b200_set_thread_options(*(s if w is None else w for s, w in zip(self._saved, self._want))))
        if self._want_paths is not None:
            check(lib.rlic_b200_set_thread_paths(self._want_paths))
        return self

    def __exit__(self, *exc):
        check(lib.rlic_b200_set_thread_options(*self._saved))
        check(lib.rlic_b200_set_thread_paths(self._saved_paths))
        return False


def effective_options() -> dict:
    """What a call made now by this thread would use:
    ``{"arithmetic", "schedule", "walk", "paths"}``."""
    got = [ctypes.c_int(), ctypes.c_int(), ctypes.c_int()]
    lib.rlic_b200_get_effective_options(*(ctypes.byref(x) for x in got))
    names = []
    for table, x in zip((ARITHMETICS, SCHEDULES, WALKS), got):
        names.append(next(name for name, c in table.items() if c == x.value))
    code = int(lib.rlic_b200_get_effective_paths())
    names.append(next(name for name, c in PATHS.items() if c == code))
    return dict(zip(("arithmetic", "schedule", "walk", "paths"), names))


def device_count() -> int:
    return int(lib.rlic_b200_device_count())


def launch_count() -> int:
    return int(lib.rlic_b200_launch_count())


def padded_cells(rows: int, nx: int) -> int:
    """Cells of the kernels' padded buffer for `rows` x `nx` pixels (include/rlic_b200.h)."""
    return int(lib.rlic_b200_padded_cells(rows, nx))


def check(rc: int) -> None:
    """Turn a non-zero return code into the matching Python exception."""
    if rc == OK:
        return
    msg = (lib.rlic_b200_last_error() or b"").decode("utf-8", "replace")
    if rc == EINVAL or rc == ESHARD:
        raise ValueError(msg)
    if rc == ENODEVICE:
        raise RuntimeError(f"rlic_b200 needs a CUDA device and has no CPU fallback: {msg}")
    raise RuntimeError(msg)


def wall_codes(boundaries) -> tuple[int, int, int, int]:
    (xl, xr), (yl, yr) = boundaries
    try:
        return (_WALL_CODE[xl], _WALL_CODE[xr], _WALL_CODE[yl], _WALL_CODE[yr])
    except KeyError as exc:
        raise ValueError(f"unknown boundary {exc.args[0]!r}") from None


def mode_code(uv_mode: str) -> int:
    try:
        return _MODE_CODE[uv_mode]
    except KeyError:
        raise ValueError(f"unknown uv_mode {uv_mode!r}") from None


class _ResultBlock:
    """Page-locked storage of one result array; returns to the library's pool when the
    last array viewing it is garbage collected."""

    __slots__ = ("ptr", "nbytes", "__weakref__")

    def __init__(self, ptr: int, nbytes: int):
        self.ptr, self.nbytes = ptr, nbytes

    @property
    def __array_interface__(self):
        return {"shape": (self.nbytes,), "typestr": "|u1", "data": (self.ptr, False), "version": 3}

    def __del__(self):
        try:
            lib.rlic_b200_result_free(self.ptr)
        except Exception:   # interpreter shutdown
            pass


# results at least this large are placed in page-locked memory (faster download)
_PINNED_RESULT_MIN_BYTES = 1 << 20


def new_result(shape, dtype) -> np.ndarray:
    """A freshly allocated, writable, C-contiguous array for a result."""
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    if nbytes >= _PINNED_RESULT_MIN_BYTES:
        ptr = lib.rlic_b200_result_alloc(nbytes)
        if ptr:
            return np.asarray(_ResultBlock(ptr, nbytes)).view(dtype).reshape(shape)
    return np.empty(shape, dtype=dtype)


def _as_image(name: str, arr, dtype: np.dtype, ndim: int) -> np.ndarray:
    # the PyO3 signature rejects anything but an ndarray of the right dtype/rank
    if not isinstance(arr, np.ndarray) or arr.dtype != dtype or arr.ndim != ndim:
        got = f"{type(arr).__name__}" + (
            f"[{arr.dtype}, ndim={arr.ndim}]" if isinstance(arr, np.ndarray) else ""
        )
        raise TypeError(f"argument '{name}': expected a {ndim}-D {dtype} ndarray, got {got}")
    # any strides are accepted (stride-0 broadcasts, F-order views): normalise before upload
    return np.ascontiguousarray(arr)


NEGATIVE_TEXTURE_MESSAGE = "Found invalid texture element(s). Expected only positive values."


def _convolve(sfx: str, real, texture, uv, kernel, boundaries, iterations, check_texture=False):
    dtype = np.dtype(real)
    u, v, uv_mode = uv
    texture = _as_image("texture", texture, dtype, 2)
    u = _as_image("u", u, dtype, 2)
    v = _as_image("v", v, dtype, 2)
    kernel = _as_image("kernel", kernel, dtype, 1)
    if u.shape != texture.shape or v.shape != texture.shape:
        raise ValueError("texture, u and v must have identical shapes")
    ny, nx = texture.shape
    out = new_result((ny, nx), dtype)
    p = ctypes.POINTER(real)
    args = [
        texture.ctypes.data_as(p),
        u.ctypes.data_as(p),
        v.ctypes.data_as(p),
        ny,
        nx,
        kernel.ctypes.data_as(p),
        kernel.size,
        mode_code(uv_mode),
        *wall_codes(boundaries),
        int(iterations),
        out.ctypes.data_as(p),
    ]
    if check_texture:
        negative = _int(0)
        check(getattr(lib, f"rlic_b200_convolve_checked_{sfx}")(*args, ctypes.byref(negative)))
        if negative.value:
            raise ValueError(NEGATIVE_TEXTURE_MESSAGE)
    else:
        check(getattr(lib, f"rlic_b200_convolve_{sfx}")(*args))
    return out


def convolve_batch(textures, uv, kernel, boundaries, iterations=1, devices=None, *, check_texture=False):
    """``nfields`` independent images stored back to back (``(nfields, ny, nx)`` arrays of
    one dtype), split whole-image over ``devices`` (default: every visible device).
    Binding of ``rlic_b200_convolve_batch_*``; ``check_texture=True`` performs the reference's
    "no negative texture values" validation on the devices, during the uploads
    (``rlic_b200_convolve_batch_checked_*``)."""
    u, v, uv_mode = uv
    dtype = textures.dtype
    if dtype == np.dtype("float32"):
        sfx, real = "f32", ctypes.c_float
    elif dtype == np.dtype("float64"):
        sfx, real = "f64", ctypes.c_double
    else:
        raise TypeError(f"argument 'textures': expected float32 or float64, got {dtype}")
    textures = _as_image("textures", textures, dtype, 3)
    u = _as_image("u", u, dtype, 3)
    v = _as_image("v", v, dtype, 3)
    kernel = _as_image("kernel", kernel, dtype, 1)
    if u.shape != textures.shape or v.shape != textures.shape:
        raise ValueError("textures, u and v must have identical shapes")
    nf, ny, nx = textures.shape
    out = new_result(textures.shape, dtype)
    p = ctypes.POINTER(real)
    dev_arr, ndev = None, 0
    if devices is not None:
        devices = [int(d) for d in devices]
        dev_arr, ndev = (ctypes.c_int * len(devices))(*devices), len(devices)
    args = (textures.ctypes.data_as(p), u.ctypes.data_as(p), v.ctypes.data_as(p), nf, ny, nx,
            kernel.ctypes.data_as(p), kernel.size, mode_code(uv_mode), *wall_codes(boundaries),
            int(iterations), dev_arr, ndev, out.ctypes.data_as(p))
    if check_texture:
        negative = _int(0)
        check(getattr(lib, f"rlic_b200_convolve_batch_checked_{sfx}")(*args, ctypes.byref(negative)))
        if negative.value:
            raise ValueError(NEGATIVE_TEXTURE_MESSAGE)
    else:
        check(getattr(lib, f"rlic_b200_convolve_batch_{sfx}")(*args))
    return out


def _equalize(sfx: str, real, image, nbins: int) -> np.ndarray:
    dtype = np.dtype(real)
    image = _as_image("image", image, dtype, 2)
    if not isinstance(nbins, (int, np.integer)) or isinstance(nbins, bool):
        raise TypeError(f"argument 'nbins': expected an int, got {type(nbins).__name__}")
    ny, nx = image.shape
    out = new_result((ny, nx), dtype)
    p = ctypes.POINTER(real)
    check(getattr(lib, f"rlic_b200_equalize_histogram_{sfx}")(
        image.ctypes.data_as(p), ny, nx, int(nbins), out.ctypes.data_as(p)))
    return out


def equalize_histogram_f32(image, nbins):
    """The reference's stub `rlic._core.equalize_histogram_f32(image, nbins)` (_core.pyi:30-33),
    with the semantics of include/rlic_b200.h; runs on the GPU."""
    return _equalize("f32", ctypes.c_float, image, nbins)


def equalize_histogram_f64(image, nbins):
    return _equalize("f64", ctypes.c_double, image, nbins)


def convolve_f32(texture, uv, kernel, boundaries, iterations=1, *, check_texture=False):
    """``check_texture=True`` additionally performs the reference's "no negative
    texture values" validation on the device, during the upload."""
    return _convolve("f32", ctypes.c_float, texture, uv, kernel, boundaries, iterations, check_texture)


def convolve_f64(texture, uv, kernel, boundaries, iterations=1, *, check_texture=False):
    return _convolve("f64", ctypes.c_double, texture, uv, kernel, boundaries, iterations, check_texture)
