"""Synthetic inputs of the five BASELINE.json configurations.

Shared by bench.py, the tests and ``__graft_entry__.smoke()`` so that all three
run the very same data.  Pure NumPy; nothing here touches the GPU.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass(frozen=True)
class Workload:
    name: str
    texture: np.ndarray
    u: np.ndarray
    v: np.ndarray
    kernel: np.ndarray
    uv_mode: str = "velocity"
    boundaries: object = "closed"
    iterations: int = 1
    description: str = ""
    extra: dict = field(default_factory=dict)

    @property
    def pixels(self) -> int:
        return int(self.texture.size)

    @property
    def steps_per_pixel(self) -> int:
        return int(self.kernel.size) - 1

    @property
    def pixel_steps(self) -> int:
        return self.pixels * self.steps_per_pixel * self.iterations

    @property
    def gather_bytes_per_pixel(self) -> int:
        """Algorithmic gather bytes per pixel per iteration, BASELINE.md section 2:
        u and v once per step, texture once per step plus the centre, one store."""
        return (3 * self.steps_per_pixel + 2) * self.texture.dtype.itemsize

    def kwargs(self) -> dict:
        return dict(
            kernel=self.kernel,
            uv_mode=self.uv_mode,
            boundaries=self.boundaries,
            iterations=self.iterations,
        )


def triangle_kernel(length: int, dtype) -> np.ndarray:
    """``1 - |linspace(-1, 1, L)|`` (reference README.md:78)."""
    return (1 - np.abs(np.linspace(-1, 1, length))).astype(dtype)


def path_probe(shape, dtype, taps: int, seed: int = 0):
    """Inputs that turn one convolution pass into an exact signature of the pixels each
    walker visited (the path-divergence fraction): a texture of random
    integers below 2^17 and a kernel of ones.  A pixel's result is then the plain sum of at
    most ``taps`` such integers -- below 2^24, so it is exact in f32 and f64 whatever the
    order or fusing of the additions -- and two implementations return the same number iff
    their walkers visited the same pixels the same number of times (up to a 2^-17 chance of
    a collision).  Comparing two such images counts diverging *paths*, not rounding."""
    if taps * (1 << 17) > (1 << 24):
        raise ValueError("too many taps for an exact f32 sum")
    rng = np.random.default_rng(seed)
    texture = rng.integers(0, 1 << 17, size=shape).astype(dtype)
    return texture, np.ones(taps, dtype=dtype)


def vortex(ny: int, nx: int, dtype, drift: float = 0.0):
    """Solid-body rotation about the image centre; even sizes have no exact zeros."""
    y = np.linspace(-1, 1, ny, dtype=np.float64)
    x = np.linspace(-1, 1, nx, dtype=np.float64)
    u = np.broadcast_to((-y)[:, None] + drift, (ny, nx)).astype(dtype)
    v = np.broadcast_to(x[None, :] + drift, (ny, nx)).astype(dtype)
    return u, v


def readme_example(n: int = 256, iterations: int = 1) -> Workload:
    """C1 — reference README.md:50-56,78,84-91: f64, U=cos(2x), V=sin(x), periodic."""
    rng = np.random.default_rng(0)
    shape = (n, n)
    texture = rng.random(shape)
    x = np.linspace(0, np.pi, n)
    u = np.broadcast_to(np.cos(2 * x), shape)
    v = np.broadcast_to(np.sin(x).T, shape)
    return Workload(
        name="c1_readme_256_f64",
        texture=texture,
        u=u,
        v=v,
        kernel=triangle_kernel(65, np.float64),
        boundaries="periodic",
        iterations=iterations,
        description=f"README example {n}x{n} f64, U=cos(2x), V=sin(x), 65 taps, periodic",
    )


def vortex_noise(n: int = 4096, *, rows: int | None = None, dtype=np.float32, taps: int = 65,
                 iterations: int = 5, seed: int = 0) -> Workload:
    """C2 (n=4096, it=5) and C4 (n=16384, it=20): noise texture, vortex, closed walls."""
    ny = n if rows is None else rows
    rng = np.random.default_rng(seed)
    texture = rng.random((ny, n), dtype=dtype)
    u, v = vortex(ny, n, dtype)
    return Workload(
        name=f"vortex_{ny}x{n}_{np.dtype(dtype).name}_L{taps}_it{iterations}",
        texture=texture,
        u=u,
        v=v,
        kernel=triangle_kernel(taps, dtype),
        boundaries="closed",
        iterations=iterations,
        description=(
            f"{ny}x{n} {np.dtype(dtype).name} noise, analytic vortex, {taps}-tap triangle "
            f"kernel, closed, iterations={iterations}"
        ),
    )


def polarization_split(n: int = 2048, taps: int = 129, iterations: int = 1) -> Workload:
    """C3 — f64 polarization, u flips sign at mid-width, v = 0 (README.md:125-128 form)."""
    rng = np.random.default_rng(0)
    shape = (n, n)
    texture = rng.random(shape)
    col = np.broadcast_to(np.arange(n), shape)
    u = np.where(col < n / 2, -1.0, 1.0)
    v = np.zeros(shape)
    return Workload(
        name=f"c3_polarization_{n}_f64_L{taps}",
        texture=texture,
        u=u,
        v=v,
        kernel=triangle_kernel(taps, np.float64),
        uv_mode="polarization",
        boundaries={"x": "periodic", "y": "closed"},
        iterations=iterations,
        description=f"{n}x{n} f64 polarization, sign-flipping U, x periodic / y closed, {taps} taps",
    )


def fourier_field(n: int, seed: int, dtype=np.float32, modes: int = 3):
    """Smooth random field from a few low-order Fourier modes (no NaN, no exact zeros)."""
    rng = np.random.default_rng(seed)
    y, x = np.meshgrid(np.linspace(0, 2 * np.pi, n, endpoint=False),
                       np.linspace(0, 2 * np.pi, n, endpoint=False), indexing="ij")
    out = []
    for _ in range(2):
        f = np.full((n, n), 0.05 * (rng.random() + 0.1))
        for ky in range(modes):
            for kx in range(modes):
                a, ph = rng.normal(), rng.random() * 2 * np.pi
                f = f + a * np.cos(ky * y + kx * x + ph)
        out.append(f.astype(dtype))
    return out[0], out[1]


def snapshot_batch(nfields: int = 4096, n: int = 512, taps: int = 33, iterations: int = 3,
                   dtype=np.float32, distinct: int = 16) -> Workload:
    """C5 — batch of independent fields.  ``distinct`` different (u, v) pairs are
    generated and cycled (generating 4096 analytic fields costs minutes of host
    time and does not change the GPU work); textures are all different."""
    rng = np.random.default_rng(0)
    texture = rng.random((nfields, n, n), dtype=dtype)
    base = [fourier_field(n, seed, dtype) for seed in range(min(distinct, nfields))]
    u = np.empty((nfields, n, n), dtype=dtype)
    v = np.empty((nfields, n, n), dtype=dtype)
    for f in range(nfields):
        u[f], v[f] = base[f % len(base)]
    return Workload(
        name=f"c5_batch_{nfields}x{n}_L{taps}_it{iterations}",
        texture=texture,
        u=u,
        v=v,
        kernel=triangle_kernel(taps, dtype),
        boundaries="closed",
        iterations=iterations,
        description=f"{nfields} independent {n}x{n} f32 fields, {taps} taps, iterations={iterations}",
    )
