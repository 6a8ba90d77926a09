"""Device arrays of other libraries at the device entry points:
``__cuda_array_interface__`` and DLPack producers are accepted without a copy."""
from __future__ import annotations

import numpy as np
import pytest


torch = pytest.importorskip("torch")


class CudaArrayOnly:
    """What a CuPy / Numba array looks like to a consumer: only the interface dict."""

    def __init__(self, tensor):
        self._keep = tensor
        self.__cuda_array_interface__ = tensor.__cuda_array_interface__


class DLPackOnly:
    def __init__(self, tensor):
        self._t = tensor

    def __dlpack__(self, **kw):
        return self._t.__dlpack__(**kw)

    def __dlpack_device__(self):
        return self._t.__dlpack_device__()


def test_tensors_pass_through_unchanged():
    from rlic_b200.device import _as_tensor

    t = torch.zeros(3, 4)
    assert _as_tensor(t) is t
    assert _as_tensor(None) is None


def test_host_arrays_are_rejected_with_a_type_error():
    from rlic_b200.device import _as_tensor, _check_image

    host = _as_tensor(np.zeros((4, 4), dtype=np.float32))     # numpy exports DLPack (CPU)
    with pytest.raises(TypeError, match="CUDA tensor"):
        _check_image("texture", host)
    with pytest.raises(TypeError, match="CUDA tensor"):
        _check_image("texture", [[0.0]])


@pytest.mark.gpu
@pytest.mark.parametrize("wrap", [CudaArrayOnly, DLPackOnly], ids=["cuda_array_interface", "dlpack"])
@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_foreign_device_arrays_match_the_oracle(wrap, dtype):
    import oracle
    from rlic_b200.device import convolve_device, pack_field

    rng = np.random.default_rng(11)
    shape = (96, 130)
    tex = rng.random(shape).astype(dtype)
    u = (rng.random(shape) - 0.5).astype(dtype)
    v = (rng.random(shape) - 0.5).astype(dtype)
    taps = np.linspace(0.1, 1.0, 21).astype(dtype)
    walls = (("periodic", "periodic"), ("closed", "closed"))
    want = oracle.convolve(tex, u, v, kernel=taps, boundaries=walls, iterations=2)

    dev = [torch.from_numpy(a).cuda() for a in (tex, u, v)]
    got = convolve_device(*(wrap(t) for t in dev), kernel=taps,
                          boundaries={"x": "periodic", "y": "closed"}, iterations=2)
    assert isinstance(got, torch.Tensor) and got.is_cuda
    np.testing.assert_array_equal(got.cpu().numpy(), want)

    # zero copy: the packed field and an out= buffer see the producer's memory
    field = pack_field(wrap(dev[1]), wrap(dev[2]), boundaries={"x": "periodic", "y": "closed"})
    out = torch.empty_like(dev[0])
    ret = convolve_device(wrap(dev[0]), kernel=taps, field=field,
                          boundaries={"x": "periodic", "y": "closed"}, iterations=2, out=wrap(out))
    assert ret.data_ptr() == out.data_ptr()
    np.testing.assert_array_equal(out.cpu().numpy(), want)
