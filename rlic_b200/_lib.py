"""Public entry point ``convolve``.

Drop-in for ``rlic.convolve``: behaviour (argument kinds, order and text of
every validation error, ``ExceptionGroup`` aggregation, ``iterations == 0``
copy, dtype dispatch) mirrors ``/root/reference/src/rlic/_lib.py:37-235``.
The computation itself happens on the GPU through ``rlic_b200._core``.
"""

from __future__ import annotations

__all__ = [
    "convolve",
    "convolve_batch",
    "equalize_histogram",
]

from typing import TYPE_CHECKING

import numpy as np

from rlic_b200 import _core
from rlic_b200._boundaries import BoundarySet

if TYPE_CHECKING:
    from numpy import dtype, ndarray

    from rlic_b200._boundaries import BoundarySpec
    from rlic_b200._typing import F, UVMode

_KNOWN_UV_MODES = ["velocity", "polarization"]
_SUPPORTED_DTYPES: list[np.dtype] = [np.dtype("float32"), np.dtype("float64")]


_NEGATIVE_TEXTURE = _core.NEGATIVE_TEXTURE_MESSAGE
# textures at least this large have their sign check done on the GPU, during the upload
_DEVICE_CHECK_MIN_SIZE = 1 << 16


def _check_inputs(texture, u, v, kernel, uv_mode, boundaries, iterations, defer_sign_check=False):
    """Run every validator, in the reference's order; return (problems, walls, deferred).

    With ``defer_sign_check`` the O(pixels) ``texture < 0`` scan is skipped and
    ``deferred`` is the position its error would take in ``problems``: the caller
    either lets the GPU perform it (when nothing else is wrong) or calls again
    without deferring, so the outcome is the same as checking eagerly."""
    problems: list[Exception] = []
    add = problems.append
    deferred = None

    if iterations < 0:
        add(
            ValueError(
                f"Invalid number of iterations: {iterations}\n"
                "Expected a strictly positive integer."
            )
        )

    if uv_mode not in _KNOWN_UV_MODES:
        add(ValueError(f"Invalid uv_mode {uv_mode!r}. Expected one of {_KNOWN_UV_MODES}"))

    named = (("texture", texture), ("u", u), ("v", v), ("kernel", kernel))

    def expectation() -> str:      # built only when it is needed: four dtype reprs cost more than the rest
        return (
            "Expected texture, u, v and kernel with identical dtype, "
            f"from {_SUPPORTED_DTYPES}. Got "
            + ", ".join(f"{name}.dtype={arr.dtype!r}" for name, arr in named)
        )

    seen = {arr.dtype for _, arr in named}
    if rejected := seen.difference(_SUPPORTED_DTYPES):
        add(TypeError(f"Found unsupported data type(s): {list(rejected)}. {expectation()}"))
    if len(seen) != 1:
        add(TypeError(f"Data types mismatch. {expectation()}"))

    if texture.ndim != 2:
        add(
            ValueError(
                f"Expected a texture with exactly two dimensions. Got texture.ndim={texture.ndim}"
            )
        )
    if defer_sign_check:
        deferred = len(problems)
    elif np.any(texture < 0):
        add(ValueError(_NEGATIVE_TEXTURE))
    if u.shape != texture.shape or v.shape != texture.shape:
        add(
            ValueError(
                "Shape mismatch: expected texture, u and v with identical shapes. "
                f"Got texture.shape={texture.shape}, u.shape={u.shape}, v.shape={v.shape}"
            )
        )

    if kernel.ndim != 1:
        add(
            ValueError(
                f"Expected a kernel with exactly one dimension. Got kernel.ndim={kernel.ndim}"
            )
        )
    if np.any(~np.isfinite(kernel)):
        add(ValueError("Found non-finite value(s) in kernel."))

    walls = BoundarySet.from_spec(boundaries)
    if walls is None:
        add(TypeError(f"Invalid boundary specification {boundaries}"))
    else:
        problems.extend(walls.collect_exceptions())
    return problems, walls, deferred


def convolve(
    texture: ndarray[tuple[int, int], dtype[F]],
    /,
    u: ndarray[tuple[int, int], dtype[F]],
    v: ndarray[tuple[int, int], dtype[F]],
    *,
    kernel: ndarray[tuple[int], dtype[F]],
    uv_mode: UVMode = "velocity",
    boundaries: BoundarySpec = "closed",
    iterations: int = 1,
) -> ndarray[tuple[int, int], dtype[F]]:
    """2-dimensional line integral convolution (GPU).

    Convolve ``texture`` along the streamlines of the vector field ``(u, v)``
    with the 1D ``kernel``.

    Arguments
    ---------
    texture: 2D numpy array, positional-only
      The image that is smeared along the field lines (usually noise).
      Must not contain negative values.

    u, v: 2D numpy arrays, same shape as ``texture``
      Horizontal (along axis 1) and vertical (along axis 0) field components.

    kernel: 1D numpy array, keyword-only
      Weights along the field line; the first half applies upstream of the
      starting pixel, the second half downstream.  Must be finite; may be
      negative; odd lengths balance both directions.

    uv_mode: 'velocity' (default) or 'polarization', keyword-only
      With 'polarization' only the orientation of (u, v) matters, not its sign.

    boundaries: 'closed' (default), 'periodic', or ``{'x': ..., 'y': ...}`` whose
      values are one of these names or a ``(left, right)`` pair of them.
      A 'periodic' side requires 'periodic' on the opposite side.

    iterations: int >= 0 (default 1), keyword-only
      Number of passes; each pass takes the previous result as its texture.
      Intermediate results never leave the GPU.

    Returns
    -------
    A newly allocated 2D array with the dtype of the inputs (a copy of
    ``texture`` when ``iterations == 0``).  Inputs are never modified.

    Raises
    ------
    TypeError, ValueError for invalid arguments; an ``ExceptionGroup`` carrying
    all of them when there is more than one.  RuntimeError if the CUDA library
    or a GPU is unavailable (there is no CPU fallback).

    Notes
    -----
    All arrays must share one dtype, float32 or float64; the computation is
    carried out in that dtype.  Streamlines stop at pixels where u or v is NaN.
    Arrays of any memory layout are accepted (they are made C-contiguous before
    upload).  Results are bit-identical to rLIC built with its default
    features (fma + branchless); see DESIGN.md.
    """
    # The sign scan of a large texture costs more on the host than a GPU pass
    # (SURVEY.md 8(f).3): when every other check passes it is fused into the upload.
    defer = iterations > 0 and getattr(texture, "size", 0) >= _DEVICE_CHECK_MIN_SIZE
    problems, walls, deferred = _check_inputs(
        texture, u, v, kernel, uv_mode, boundaries, iterations, defer_sign_check=defer
    )
    if problems and deferred is not None:
        # something else is wrong: evaluate the sign check here so that it takes
        # its place in the group
        problems, walls, deferred = _check_inputs(
            texture, u, v, kernel, uv_mode, boundaries, iterations
        )
    if len(problems) == 1:
        raise problems[0]
    if problems:
        raise ExceptionGroup("Invalid inputs were received.", problems)
    assert walls is not None

    if iterations == 0:
        return texture.copy()

    if texture.dtype == np.dtype("float32"):
        run = _core.convolve_f32
    elif texture.dtype == np.dtype("float64"):
        run = _core.convolve_f64
    else:
        raise AssertionError
    return run(texture, (u, v, uv_mode), kernel, (walls.x, walls.y), iterations,
               check_texture=deferred is not None)


def convolve_batch(
    textures,
    /,
    u,
    v,
    *,
    kernel,
    uv_mode: UVMode = "velocity",
    boundaries: BoundarySpec = "closed",
    iterations: int = 1,
    devices=None,
):
    """``convolve`` for a stack of independent fields (extension; the reference accepts 2D only).

    ``textures``, ``u``, ``v``: arrays of shape ``(nfields, ny, nx)`` and one dtype.
    ``out[f]`` equals ``convolve(textures[f], u[f], v[f], ...)`` bit for bit; the
    fields are split whole-image over ``devices`` (default: all visible GPUs) and
    their uploads, passes and downloads overlap.  Arguments are validated with
    the same rules and messages as :func:`convolve`, applied to the first field
    for the per-image checks (every field shares shape and dtype).
    """
    for name, arr in (("textures", textures), ("u", u), ("v", v)):
        if getattr(arr, "ndim", None) != 3:
            raise ValueError(
                f"Expected {name} with exactly three dimensions (nfields, ny, nx). "
                f"Got {name}.ndim={getattr(arr, 'ndim', None)}"
            )
    if u.shape != textures.shape or v.shape != textures.shape:
        raise ValueError(
            "Shape mismatch: expected textures, u and v with identical shapes. "
            f"Got textures.shape={textures.shape}, u.shape={u.shape}, v.shape={v.shape}"
        )
    if textures.shape[0] == 0:
        return textures.copy()
    # dtype / kernel / mode / boundary / iteration rules are those of a single image
    problems, walls, _ = _check_inputs(
        textures[0], u[0], v[0], kernel, uv_mode, boundaries, iterations, defer_sign_check=True
    )
    # The sign scan of the whole stack (4 GiB at BASELINE config 5: about a second on the host)
    # rides on the uploads when nothing else is wrong; otherwise it runs here, on the first
    # field that has a negative value, so that it takes its place in the group.
    on_device = not problems and iterations > 0 and textures.size >= _DEVICE_CHECK_MIN_SIZE
    if not on_device:
        negative = np.flatnonzero((textures < 0).any(axis=(1, 2)))
        if negative.size:
            problems, walls, _ = _check_inputs(
                textures[negative[0]], u[0], v[0], kernel, uv_mode, boundaries, iterations
            )
    if len(problems) == 1:
        raise problems[0]
    if problems:
        raise ExceptionGroup("Invalid inputs were received.", problems)
    assert walls is not None
    if iterations == 0:
        return textures.copy()
    return _core.convolve_batch(
        textures, (u, v, uv_mode), kernel, (walls.x, walls.y), iterations, devices,
        check_texture=on_device,
    )


def convolve_sharded(
    texture,
    /,
    u,
    v,
    *,
    kernel,
    uv_mode: UVMode = "velocity",
    boundaries: BoundarySpec = "closed",
    iterations: int = 1,
    devices=None,
):
    """``convolve`` with the image split into row slabs over several GPUs of this process
    (extension; BASELINE config 4 without ``torchrun``).

    Same arguments, validation, messages and result as :func:`convolve` -- bit for bit: the
    walkers run in global row numbers -- plus ``devices`` (default: all visible GPUs, fewer
    when the image is too short for that many slabs).  Each pass computes the edge strips of
    every slab first, copies them into the neighbours' halos device to device and computes
    the interiors meanwhile (``rlic_b200/multi.py``).  Not yet run on GPUs.
    """
    problems, walls, _ = _check_inputs(texture, u, v, kernel, uv_mode, boundaries, iterations)
    if len(problems) == 1:
        raise problems[0]
    if problems:
        raise ExceptionGroup("Invalid inputs were received.", problems)
    if iterations == 0:
        return texture.copy()
    from rlic_b200.multi import MultiDeviceConvolver   # torch is needed only here

    mc = MultiDeviceConvolver(*texture.shape, kernel=kernel, uv_mode=uv_mode, boundaries=boundaries,
                              devices=devices)
    mc.set_field(u, v)
    return mc.convolve(texture, iterations)


def equalize_histogram(image, /, *, nbins: int = 256):
    """Histogram equalisation of a 2D float32 / float64 image on the GPU: the usual step between
    a line integral convolution and its rendering (extension: rLIC only declares the operation,
    ``rlic._core.equalize_histogram_f32/_f64(image, nbins)``, and points to its sister project
    ``ahe`` for an implementation).

    Every pixel is replaced by the fraction of the (non-NaN) pixels that fall into its own or a
    lower one of ``nbins`` equal-width bins between the image's minimum and maximum; NaN pixels
    stay NaN.  Returns a new array of the image's dtype with values in (0, 1].  The exact
    arithmetic is stated in ``include/rlic_b200.h``.
    """
    if not isinstance(image, np.ndarray) or image.dtype not in _SUPPORTED_DTYPES:
        raise TypeError(f"Expected a float32 or float64 numpy array, got {getattr(image, 'dtype', type(image))}")
    if image.ndim != 2:
        raise ValueError(f"Expected an image with exactly two dimensions. Got image.ndim={image.ndim}")
    if isinstance(nbins, bool) or not isinstance(nbins, (int, np.integer)) or not 1 <= nbins <= 1 << 24:
        raise ValueError(f"Invalid number of bins: {nbins!r}. Expected an integer between 1 and 2**24.")
    if image.size == 0:
        return image.copy()
    run = _core.equalize_histogram_f32 if image.dtype == np.dtype("float32") else _core.equalize_histogram_f64
    return run(image, int(nbins))
