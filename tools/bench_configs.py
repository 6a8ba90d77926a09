#!/usr/bin/env python
"""Times every BASELINE.json configuration on one GPU (kernel-only and end to end)
and prints one JSON line per configuration.  Companion to bench.py, which is the
contract benchmark for the headline configuration only.

    python tools/bench_configs.py [--configs c1,c2,c3,c4,c5] [--c5-fields 1024]
"""

from __future__ import annotations

import argparse
import ctypes
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import rlic_b200  # noqa: E402
from rlic_b200 import _core, workloads  # noqa: E402
from rlic_b200.device import convolve_device, pack_field  # noqa: E402

PEAK = 6456.2
try:
    PEAK = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    pass


def time_device(w, reps=5):
    dev = torch.device("cuda", 0)
    tex = torch.from_numpy(np.ascontiguousarray(w.texture)).to(dev)
    u = torch.from_numpy(np.ascontiguousarray(w.u)).to(dev)
    v = torch.from_numpy(np.ascontiguousarray(w.v)).to(dev)
    field = pack_field(u, v, boundaries=w.boundaries)
    del u, v
    out = torch.empty_like(tex)
    kw = dict(kernel=w.kernel, uv_mode=w.uv_mode, boundaries=w.boundaries, iterations=w.iterations)
    for _ in range(2):
        convolve_device(tex, field=field, out=out, **kw)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        convolve_device(tex, field=field, out=out, **kw)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def time_device_abi(w, reps=200):
    """Small images: the same device call through the raw C ABI (rlic_b200_convolve_packed_*) with
    its arguments prepared once, so that what is timed is the library and the GPU rather than
    the Python wrapper's per-call validation (tens of microseconds, more than the kernels)."""
    dev = torch.device("cuda", 0)
    tex = torch.from_numpy(np.ascontiguousarray(w.texture)).to(dev)
    u = torch.from_numpy(np.ascontiguousarray(w.u)).to(dev)
    v = torch.from_numpy(np.ascontiguousarray(w.v)).to(dev)
    field = pack_field(u, v, boundaries=w.boundaries)
    out = torch.empty_like(tex)
    sfx, real = ("f32", ctypes.c_float) if w.texture.dtype == np.float32 else ("f64", ctypes.c_double)
    from rlic_b200._boundaries import BoundarySet

    bs = BoundarySet.from_spec(w.boundaries)
    walls = _core.wall_codes((bs.x, bs.y))
    taps = np.ascontiguousarray(w.kernel)
    fn = getattr(_core.lib, f"rlic_b200_convolve_packed_{sfx}")
    ny, nx = w.texture.shape
    args = (tex.data_ptr(), field.data.data_ptr(), ny, nx, taps.ctypes.data_as(ctypes.POINTER(real)), taps.size,
            _core.mode_code(w.uv_mode), *walls, w.iterations, out.data_ptr(),
            int(torch.cuda.current_stream().cuda_stream))
    for _ in range(10):
        _core.check(fn(*args))
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(reps):
        fn(*args)
    b.record()
    enqueue_us = (time.perf_counter() - t0) / reps * 1e6
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, enqueue_us


def time_host(w, reps=3):
    kw = w.kwargs()
    pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory().numpy()  # noqa: E731
    tex, u, v = pin(w.texture), pin(w.u), pin(w.v)
    for _ in range(3):   # the result's page-locked block is set up on the second call of a size
        out = rlic_b200.convolve(tex, u, v, **kw)
    del out
    t0 = time.perf_counter()
    for _ in range(reps):
        rlic_b200.convolve(tex, u, v, **kw)
    return (time.perf_counter() - t0) / reps * 1e3


def time_batch(w, reps=2):
    p = ctypes.POINTER(ctypes.c_float)
    nf, ny, nx = w.texture.shape
    out = np.empty_like(w.texture)
    args = (w.texture.ctypes.data_as(p), w.u.ctypes.data_as(p), w.v.ctypes.data_as(p), nf, ny, nx,
            w.kernel.ctypes.data_as(p), w.kernel.size, 0, 0, 0, 0, 0, w.iterations, None, 0,
            out.ctypes.data_as(p))
    _core.check(_core.lib.rlic_b200_convolve_batch_f32(*args))
    t0 = time.perf_counter()
    for _ in range(reps):
        _core.check(_core.lib.rlic_b200_convolve_batch_f32(*args))
    return (time.perf_counter() - t0) / reps * 1e3, out


def report(name, w, dev_ms, host_ms, extra=None):
    pix = w.pixels
    gather = w.gather_bytes_per_pixel * pix * w.iterations
    line = {
        "config": name, "description": w.description,
        "pixels": pix, "iterations": w.iterations, "taps": int(w.kernel.size),
        "dtype": str(w.texture.dtype),
        "device_ms": dev_ms, "device_Mpix_s": pix * w.iterations / dev_ms / 1e3 if dev_ms else None,
        "device_Gsteps_s": w.pixel_steps / dev_ms / 1e6 if dev_ms else None,
        "gather_GBps": gather / dev_ms / 1e6 if dev_ms else None,
        "roofline_frac": gather / dev_ms / 1e6 / PEAK if dev_ms else None,
        "host_e2e_ms": host_ms, "host_e2e_Mpix_s": pix * w.iterations / host_ms / 1e3 if host_ms else None,
    }
    if extra:
        line.update(extra)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c1,c2,c3,c4,c5")
    ap.add_argument("--c5-fields", type=int, default=1024)
    args = ap.parse_args()
    want = set(args.configs.split(","))
    if "c1" in want:
        w = workloads.readme_example()
        abi_ms, enqueue_us = time_device_abi(w)
        report("c1", w, time_device(w, 20), time_host(w, 20),
               {"device_ms_raw_abi": abi_ms, "host_enqueue_us_per_call_raw_abi": enqueue_us,
                "note": "device_ms goes through the Python wrapper (its per-call validation outlasts the kernels); "
                        "device_ms_raw_abi is the same call through ctypes with prepared arguments"})
    if "c2" in want:
        w = workloads.vortex_noise(4096, iterations=5)
        report("c2", w, time_device(w), time_host(w))
    if "c3" in want:
        w = workloads.polarization_split(2048, taps=129)
        report("c3", w, time_device(w), time_host(w))
    if "c4" in want:
        w = workloads.vortex_noise(16384, iterations=20)
        report("c4", w, time_device(w, 2), time_host(w, 1))
    if "c5" in want:
        w = workloads.snapshot_batch(args.c5_fields)
        ms, _ = time_batch(w)
        report("c5", w, None, ms, {"note": f"{args.c5_fields} of 4096 fields, host batch entry point, pageable host memory"})


if __name__ == "__main__":
    main()
