#!/usr/bin/env python
"""BASELINE config 5 -- a batch of 4096 independent 512 x 512 f32 fields (33 taps, 3
iterations, closed walls), whole fields split over the GPUs of one box.  No collective:
the path shards by field.

    python tools/bench_c5_batch.py [--fields 4096] [--reps 3] [--check 8]

One process drives every visible GPU (that is the public batch API's own model: two host
lanes per device inside rlic_b200_convolve_batch_*).  Prints one JSON line:

* `device_resident` -- every GPU holds its share of the stack in HBM; one
  `convolve_device_batch` call per GPU, all issued before any is awaited; CUDA events per
  device, the slowest device counts.  The figure BASELINE.md section 2 sets 30.6 ms (8 GPUs,
  80 % of the nominal gather roofline) against;
* `host_pinned` / `host_pageable` -- `rlic_b200.convolve_batch` end to end, NumPy in, NumPy out
  (12 GiB in, 4 GiB out at full size), inputs page-locked or ordinary;
* `parity` -- `--check` sampled fields of the batch result against the CPU oracle, and the
  device-resident result against the host path's, bit for bit.
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import rlic_b200  # noqa: E402
from rlic_b200 import workloads  # noqa: E402
from rlic_b200.device import convolve_device_batch  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--fields", type=int, default=4096)
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--taps", type=int, default=33)
    ap.add_argument("--iterations", type=int, default=3)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--check", type=int, default=8)
    ap.add_argument("--skip-pageable", action="store_true")
    args = ap.parse_args()
    ndev = torch.cuda.device_count()
    if ndev == 0:
        raise SystemExit("needs CUDA devices (rlic_b200 has no CPU fallback)")
    w = workloads.snapshot_batch(args.fields, args.n, args.taps, args.iterations)
    nf = args.fields
    pix = nf * args.n * args.n
    kw = dict(kernel=w.kernel, boundaries="closed", iterations=args.iterations)
    line = {"config": "c5", "fields": nf, "field": [args.n, args.n], "taps": args.taps,
            "iterations": args.iterations, "n_gpus": ndev, "scaling": "whole fields per GPU, no collective",
            **rlic_b200.effective_options()}

    # ---- device-resident: each GPU's share in HBM, one batched call per GPU -----------------
    share = [(nf * d // ndev, nf * (d + 1) // ndev) for d in range(ndev)]
    dev_in, dev_out, events = [], [], []
    for d, (a, b) in enumerate(share):
        dv = torch.device("cuda", d)
        dev_in.append(tuple(torch.from_numpy(x[a:b]).to(dv) for x in (w.texture, w.u, w.v)))
        dev_out.append(torch.empty_like(dev_in[-1][0]))
    def run_all(timed: bool):
        evs = []
        for d in range(ndev):
            with torch.cuda.device(d):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                convolve_device_batch(*dev_in[d], out=dev_out[d], **kw)
                e1.record()
                evs.append((e0, e1))
        for d in range(ndev):
            torch.cuda.synchronize(d)
        return max(e0.elapsed_time(e1) for e0, e1 in evs)
    run_all(False)
    t0 = time.perf_counter()
    ms = [run_all(True) for _ in range(args.reps)]
    wall_ms = (time.perf_counter() - t0) / args.reps * 1e3
    best = min(ms)
    gather_bytes = (3 * (args.taps - 1) + 2) * 4 * pix * args.iterations
    line["device_resident"] = {
        "ms_per_call": best, "wall_ms_per_call_all_devices": wall_ms,
        "Mpix_s": pix * args.iterations / best / 1e3,
        "G_pixel_steps_s": pix * args.iterations * (args.taps - 1) / best / 1e6,
        "gather_GBps_per_gpu": gather_bytes / ndev / best / 1e6,
        "baseline_md_target_ms_8gpu_80pct": 30.6,
        "timing": "CUDA events around the call on every device, max over devices, best of reps",
    }
    device_result = np.concatenate([o.cpu().numpy() for o in dev_out])
    del dev_in, dev_out
    torch.cuda.empty_cache()

    # ---- host end to end through the public batch API -----------------------------------------
    def time_host(tex, u, v):
        out = rlic_b200.convolve_batch(tex, u, v, **kw)          # warm-up (pools, result blocks)
        ts = []
        for _ in range(args.reps):
            t0 = time.perf_counter()
            out = rlic_b200.convolve_batch(tex, u, v, **kw)
            ts.append(time.perf_counter() - t0)
        dt = min(ts)
        return out, {"ms_per_call": dt * 1e3, "Mpix_s": pix * args.iterations / dt / 1e6,
                     "h2d_bytes": 3 * tex.nbytes, "d2h_bytes": tex.nbytes}

    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()  # noqa: E731
    p_tex, p_u, p_v = pin(w.texture), pin(w.u), pin(w.v)
    host_result, line["host_pinned"] = time_host(p_tex, p_u, p_v)
    del p_tex, p_u, p_v
    if not args.skip_pageable:
        _, line["host_pageable"] = time_host(w.texture, w.u, w.v)

    # ---- parity ---------------------------------------------------------------------------------
    import oracle

    rng = np.random.default_rng(1)
    picks = sorted({0, nf - 1, *rng.integers(0, nf, size=max(0, args.check - 2)).tolist()})
    bad = []
    for f in picks:
        want = oracle.convolve(w.texture[f], w.u[f], w.v[f], kernel=w.kernel, iterations=args.iterations,
                               threads=oracle.max_threads())
        if not np.array_equal(host_result[f], want):
            bad.append(f)
    line["parity"] = {
        "fields_checked_against_cpu_oracle": picks, "mismatching_fields": bad,
        "device_resident_equals_host_path": bool(np.array_equal(device_result, host_result)),
        "bit_equal": not bad and bool(np.array_equal(device_result, host_result)),
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
