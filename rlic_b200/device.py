"""Device-resident entry points.

``convolve`` in ``rlic_b200._lib`` mirrors the reference's NumPy contract and
therefore pays two PCIe trips per call.  Callers whose data already lives on
the GPU use these functions instead: same arithmetic, same C ABI underneath
(``rlic_b200_pack_field_*`` / ``rlic_b200_convolve_packed_*`` /
``rlic_b200_pass_slab_*``), torch tensors in and out.  Arrays of other libraries
are taken without a copy through ``__cuda_array_interface__`` (CuPy, Numba) or
DLPack; the result is a torch tensor, which exports both protocols.  torch is
used for device memory and streams only.
"""

from __future__ import annotations

__all__ = ["PackedField", "convolve_device", "convolve_device_batch", "equalize_histogram_device", "pack_field"]

import ctypes
from dataclasses import dataclass

import numpy as np
import torch

from rlic_b200 import _core
from rlic_b200._boundaries import BoundarySet

_SFX = {torch.float32: ("f32", ctypes.c_float, np.float32), torch.float64: ("f64", ctypes.c_double, np.float64)}


def _kind(t: torch.Tensor):
    try:
        return _SFX[t.dtype]
    except KeyError:
        raise TypeError(f"expected float32 or float64 tensors, got {t.dtype}") from None


def _stream_handle(stream) -> int:
    s = torch.cuda.current_stream() if stream is None else stream
    return int(s.cuda_stream)


def _walls(boundaries) -> tuple[int, int, int, int]:
    bs = BoundarySet.from_spec(boundaries)
    if bs is None:
        raise TypeError(f"Invalid boundary specification {boundaries}")
    bs.validate()
    return _core.wall_codes((bs.x, bs.y))


def _host_taps(kernel, np_dtype) -> np.ndarray:
    if isinstance(kernel, torch.Tensor):
        kernel = kernel.detach().cpu().numpy()
    kernel = np.ascontiguousarray(kernel, dtype=np_dtype)
    if kernel.ndim != 1:
        raise ValueError(f"Expected a kernel with exactly one dimension. Got kernel.ndim={kernel.ndim}")
    if not np.isfinite(kernel).all():
        raise ValueError("Found non-finite value(s) in kernel.")
    return kernel


def _as_tensor(x):
    """View a foreign device array as a torch tensor (zero copy); tensors and
    ``None`` pass through.  Host arrays come out as CPU tensors and are rejected
    by ``_check_image``."""
    if x is None or isinstance(x, torch.Tensor):
        return x
    if hasattr(x, "__cuda_array_interface__"):
        return torch.as_tensor(x)
    if hasattr(x, "__dlpack__"):
        return torch.from_dlpack(x)
    return x


def _check_image(name: str, t: torch.Tensor, like: torch.Tensor | None = None) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError(f"{name} must be a CUDA tensor or an array exporting "
                        "__cuda_array_interface__ / DLPack from device memory")
    if t.dim() != 2 or not t.is_contiguous():
        raise ValueError(f"{name} must be a contiguous 2-D tensor")
    if like is not None and (t.shape != like.shape or t.dtype != like.dtype or t.device != like.device):
        raise ValueError(f"{name} must match the texture's shape, dtype and device")


@dataclass
class PackedField:
    """A vector field in the kernels' private layout (``include/rlic_b200.h``): one
    ``(u, v, ru, rv)`` record per cell of the padded buffer, wall sentinels included.
    Tied to the image shape and boundary kinds it was packed for."""

    data: torch.Tensor   # flat, 4 * padded_cells(ny, nx) scalars
    shape: tuple[int, int]
    walls: tuple[int, int, int, int]


def pack_field(u: torch.Tensor, v: torch.Tensor, *, boundaries="closed", stream=None) -> PackedField:
    """Build the packed field from two planar components (one streaming kernel)."""
    u, v = _as_tensor(u), _as_tensor(v)
    _check_image("u", u)
    _check_image("v", v, u)
    sfx, _, _ = _kind(u)
    walls = _walls(boundaries)
    ny, nx = u.shape
    data = torch.empty(4 * _core.padded_cells(ny, nx), dtype=u.dtype, device=u.device)
    with torch.cuda.device(u.device):
        rc = getattr(_core.lib, f"rlic_b200_pack_field_{sfx}")(
            u.data_ptr(), v.data_ptr(), ny, nx, *walls, data.data_ptr(), _stream_handle(stream))
    _core.check(rc)
    return PackedField(data, (ny, nx), walls)


def convolve_device(texture: torch.Tensor, u: torch.Tensor | None = None, v: torch.Tensor | None = None,
                    *, kernel, field: PackedField | None = None, uv_mode: str = "velocity",
                    boundaries="closed", iterations: int = 1, out: torch.Tensor | None = None,
                    stream=None) -> torch.Tensor:
    """``rlic_b200.convolve`` for CUDA tensors; returns a new CUDA tensor (or ``out``).

    Pass either planar ``u, v`` or a pre-packed ``field`` (saves the packing when
    the same field is reused; it must have been packed with the same
    ``boundaries``).  Work is enqueued on ``stream`` (default: torch's current
    stream) without synchronising.
    """
    texture, u, v, out = _as_tensor(texture), _as_tensor(u), _as_tensor(v), _as_tensor(out)
    _check_image("texture", texture)
    sfx, real, np_dtype = _kind(texture)
    if iterations < 0:
        raise ValueError(
            f"Invalid number of iterations: {iterations}\nExpected a strictly positive integer.")
    taps = _host_taps(kernel, np_dtype)
    walls = _walls(boundaries)
    mode = _core.mode_code(uv_mode)
    if iterations == 0:
        return texture.clone()
    if out is None:
        out = torch.empty_like(texture)
    else:
        _check_image("out", out, texture)
    ny, nx = texture.shape
    tap_ptr = taps.ctypes.data_as(ctypes.POINTER(real))
    with torch.cuda.device(texture.device):
        if field is None:
            if u is None or v is None:
                raise TypeError("pass u and v, or field=")
            _check_image("u", u, texture)
            _check_image("v", v, texture)
            rc = getattr(_core.lib, f"rlic_b200_convolve_device_{sfx}")(
                texture.data_ptr(), u.data_ptr(), v.data_ptr(), ny, nx, tap_ptr, taps.size, mode,
                *walls, int(iterations), out.data_ptr(), _stream_handle(stream))
        else:
            if field.shape != (ny, nx) or field.data.dtype != texture.dtype or field.walls != walls:
                raise ValueError("field was packed for another shape, dtype or boundary kinds")
            rc = getattr(_core.lib, f"rlic_b200_convolve_packed_{sfx}")(
                texture.data_ptr(), field.data.data_ptr(), ny, nx, tap_ptr, taps.size, mode,
                *walls, int(iterations), out.data_ptr(), _stream_handle(stream))
    _core.check(rc)
    return out


def convolve_device_batch(textures: torch.Tensor, u: torch.Tensor, v: torch.Tensor, *, kernel,
                          uv_mode: str = "velocity", boundaries="closed", iterations: int = 1,
                          out: torch.Tensor | None = None, stream=None) -> torch.Tensor:
    """``convolve_device`` for a stack of independent fields: CUDA tensors of shape
    ``(nfields, ny, nx)``; ``out[f]`` equals ``convolve_device(textures[f], u[f], v[f], ...)`` bit
    for bit.  One launch of each kernel covers the whole stack (``rlic_b200_convolve_device_batch_*``):
    the device-resident form of BASELINE config 5."""
    textures, u, v, out = _as_tensor(textures), _as_tensor(u), _as_tensor(v), _as_tensor(out)
    for name, t in (("textures", textures), ("u", u), ("v", v)) + ((("out", out),) if out is not None else ()):
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise TypeError(f"{name} must be a CUDA tensor or an array exporting "
                            "__cuda_array_interface__ / DLPack from device memory")
        if t.dim() != 3 or not t.is_contiguous():
            raise ValueError(f"{name} must be a contiguous 3-D tensor (nfields, ny, nx)")
        if t.shape != textures.shape or t.dtype != textures.dtype or t.device != textures.device:
            raise ValueError(f"{name} must match the textures' shape, dtype and device")
    sfx, real, np_dtype = _kind(textures)
    if iterations < 0:
        raise ValueError(
            f"Invalid number of iterations: {iterations}\nExpected a strictly positive integer.")
    taps = _host_taps(kernel, np_dtype)
    walls = _walls(boundaries)
    mode = _core.mode_code(uv_mode)
    if iterations == 0:
        return textures.clone()
    if out is None:
        out = torch.empty_like(textures)
    nf, ny, nx = textures.shape
    with torch.cuda.device(textures.device):
        rc = getattr(_core.lib, f"rlic_b200_convolve_device_batch_{sfx}")(
            textures.data_ptr(), u.data_ptr(), v.data_ptr(), nf, ny, nx,
            taps.ctypes.data_as(ctypes.POINTER(real)), taps.size, mode, *walls, int(iterations),
            out.data_ptr(), _stream_handle(stream))
    _core.check(rc)
    return out


def equalize_histogram_device(image: torch.Tensor, *, nbins: int = 256, out: torch.Tensor | None = None,
                              stream=None) -> torch.Tensor:
    """``rlic_b200.equalize_histogram`` for a CUDA tensor (e.g. the result of ``convolve_device``,
    without a host round trip): four streaming kernels on ``stream``; returns a new tensor or
    ``out``.  Semantics: ``include/rlic_b200.h``."""
    image, out = _as_tensor(image), _as_tensor(out)
    _check_image("image", image)
    sfx, _, _ = _kind(image)
    if isinstance(nbins, bool) or not isinstance(nbins, (int, np.integer)) or not 1 <= nbins <= 1 << 24:
        raise ValueError(f"Invalid number of bins: {nbins!r}. Expected an integer between 1 and 2**24.")
    if out is None:
        out = torch.empty_like(image)
    else:
        _check_image("out", out, image)
    ny, nx = image.shape
    with torch.cuda.device(image.device):
        rc = getattr(_core.lib, f"rlic_b200_equalize_histogram_device_{sfx}")(
            image.data_ptr(), ny, nx, int(nbins), out.data_ptr(), _stream_handle(stream))
    _core.check(rc)
    return out
