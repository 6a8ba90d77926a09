"""Catalogue of calls used to pin the argument validation of ``convolve`` against
the reference's own Python layer (see tests/golden/make_validation_golden.py).
Each case is a dict of overrides on top of a valid baseline call."""

from __future__ import annotations

import numpy as np

BOUNDARY_SPECS = [
    "closed", "periodic", "open", {"x": "closed", "y": "periodic"}, {"x": ("closed", "periodic"), "y": "closed"},
    {"x": ("periodic", "closed"), "y": ("periodic", "closed")}, {"x": ["a", "b"], "y": "closed"},
    {"x": "closed"}, {"x": "closed", "y": "closed", "z": "closed"}, 7, None, ["closed", "closed"],
    {"x": ("closed",), "y": "closed"}, {"x": ("closed", 1), "y": "closed"}, {"x": "nope", "y": ("periodic", "nope")},
]

CASES: dict[str, dict] = {
    "valid_f64": {},
    "valid_f32": {"dtype": "float32"},
    "valid_polarization_periodic": {"uv_mode": "polarization", "boundaries": "periodic"},
    "valid_dict_boundaries": {"boundaries": {"x": ("closed", "closed"), "y": "periodic"}},
    "iterations_zero": {"iterations": 0},
    "iterations_negative": {"iterations": -3},
    "bad_uv_mode": {"uv_mode": "astral"},
    "texture_complex": {"texture_dtype": "complex128"},
    "texture_int": {"texture_dtype": "int64"},
    "kernel_complex": {"kernel_dtype": "complex128"},
    "u_float32_rest_float64": {"u_dtype": "float32"},
    "all_float16": {"dtype": "float16"},
    "texture_3d": {"texture_shape": (4, 4, 4)},
    "texture_1d": {"texture_shape": (16,)},
    "texture_negative": {"texture_fill": -1.0},
    "texture_negative_zero_and_nan": {"texture_special": True},
    "u_shape": {"u_shape": (9, 8)},
    "v_shape": {"v_shape": (8, 7)},
    "kernel_2d": {"kernel_shape": (3, 3)},
    "kernel_nan": {"kernel_poison": float("nan")},
    "kernel_inf": {"kernel_poison": float("inf")},
    "kernel_negative_values": {"kernel_negative": True},
    "boundaries_none": {"boundaries": None},
    "boundaries_int": {"boundaries": 3},
    "boundaries_missing_key": {"boundaries": {"x": "closed"}},
    "boundaries_unknown": {"boundaries": "open"},
    "boundaries_mixed_periodic": {"boundaries": {"x": ("periodic", "closed"), "y": "closed"}},
    "boundaries_two_axes_wrong": {"boundaries": {"x": ("periodic", "closed"), "y": ("closed", "periodic")}},
    "many_problems": {"iterations": -1, "uv_mode": "x", "texture_fill": -1.0, "kernel_poison": float("nan"),
                      "boundaries": {"x": ("periodic", "nope"), "y": ("closed", "periodic")}},
    "dtype_and_shape_problems": {"texture_dtype": "complex128", "u_shape": (3, 3), "kernel_shape": (2, 2)},
    "zero_iterations_with_bad_kernel": {"iterations": 0, "kernel_poison": float("nan")},
    "zero_iterations_with_bad_boundaries": {"iterations": 0, "boundaries": None},
}


def build_args(spec: dict):
    rng = np.random.default_rng(0)
    dtype = spec.get("dtype", "float64")
    shape = (8, 8)
    tshape = spec.get("texture_shape", shape)
    texture = rng.random(tshape).astype(spec.get("texture_dtype", dtype))
    if "texture_fill" in spec:
        texture = np.full(tshape, spec["texture_fill"]).astype(spec.get("texture_dtype", dtype))
    if spec.get("texture_special"):
        texture[0, 0] = -0.0
        texture[1, 1] = np.nan
    u = rng.random(spec.get("u_shape", shape)).astype(spec.get("u_dtype", dtype))
    v = rng.random(spec.get("v_shape", shape)).astype(spec.get("v_dtype", dtype))
    kernel = np.linspace(0, 1, int(np.prod(spec.get("kernel_shape", (5,))))).reshape(spec.get("kernel_shape", (5,)))
    kernel = kernel.astype(spec.get("kernel_dtype", dtype))
    if "kernel_poison" in spec:
        kernel.flat[1] = spec["kernel_poison"]
    if spec.get("kernel_negative"):
        kernel = -kernel
    kwargs = {"kernel": kernel}
    for key in ("uv_mode", "boundaries", "iterations"):
        if key in spec:
            kwargs[key] = spec[key]
    return (texture, u, v), kwargs
