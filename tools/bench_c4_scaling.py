#!/usr/bin/env python
"""BASELINE config 4 — strong scaling of ONE 16384 x 16384 f32 image (65 taps,
20 iterations, closed walls) over the ranks of a torchrun job: row slabs with a
32-row NCCL halo exchange per iteration (rlic_b200.sharded).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        tools/bench_c4_scaling.py [--n 16384] [--iterations 20] [--reps 3]

Prints one JSON line on rank 0.  Run it under `timeout` on shared machines.  Device-resident timing (CUDA events, max over
ranks); the result of the N-rank run is checked against checksums that do not
depend on N (sum and sum of squares of every rank's slab, reduced).
"""

from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from rlic_b200 import _core, workloads  # noqa: E402
from rlic_b200.sharded import ShardedConvolver  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=16384)
    ap.add_argument("--iterations", type=int, default=20)
    ap.add_argument("--taps", type=int, default=65)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    real_stdout = os.dup(1)
    os.dup2(2, 1)   # keep library chatter (NCCL version line) off stdout
    # env:// rendezvous in every case (torchrun provides it; a bare `python` run gets
    # defaults).  NB: an explicit tcp:// init_method under torchrun waits forever.
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29577")
    os.environ.setdefault("RANK", "0")
    os.environ.setdefault("WORLD_SIZE", "1")
    dist.init_process_group("nccl", device_id=dev)
    _core.check(_core.lib.rlic_b200_set_device(local))

    n = args.n
    kernel = workloads.triangle_kernel(args.taps, np.float32)
    sc = ShardedConvolver(n, n, kernel=kernel, boundaries="closed",
                          exchange=os.environ.get("RLIC_B200_EXCHANGE", "nccl"))
    r0, r1 = sc.plan.row0, sc.plan.row1
    # the same global image whatever the rank count: per-row seeds
    rng = np.random.default_rng(1234)
    seeds = rng.integers(0, 2**31, size=n)
    tex = np.empty((r1 - r0, n), dtype=np.float32)
    for k, r in enumerate(range(r0, r1)):
        tex[k] = np.random.default_rng(int(seeds[r])).random(n, dtype=np.float32)
    y = np.linspace(-1, 1, n)[r0:r1]
    x = np.linspace(-1, 1, n)
    u = np.broadcast_to((-y)[:, None], tex.shape).astype(np.float32)
    v = np.broadcast_to(x[None, :], tex.shape).astype(np.float32)
    d_tex, d_u, d_v = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (tex, u, v))
    sc.set_field(d_u, d_v)
    del d_u, d_v

    out = sc.convolve(d_tex, iterations=args.iterations)   # warm-up
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.reps):
        out = sc.convolve(d_tex, iterations=args.iterations)
    b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / args.reps], device=dev, dtype=torch.float64)
    sums = torch.stack([out.double().sum(), (out.double() ** 2).sum()])
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums)
    ms = float(ms.item())
    pix = n * n
    line = {
        "config": "c4", "image": [n, n], "taps": args.taps, "iterations": args.iterations,
        "n_gpus": world, "scaling": "strong", "ms_per_call": ms,
        "Mpix_s": pix * args.iterations / ms / 1e3,
        "G_pixel_steps_s": pix * args.iterations * (args.taps - 1) / ms / 1e6,
        "checksum": [float(sums[0]), float(sums[1])],
        "halo_bytes_per_side_per_iteration": (args.taps // 2) * (n + 2) * 4,
        "exchange": sc.exchange,
    }
    sc.close()
    dist.barrier()
    dist.destroy_process_group()
    os.dup2(real_stdout, 1)
    if rank == 0:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
