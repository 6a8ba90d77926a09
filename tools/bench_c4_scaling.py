#!/usr/bin/env python
"""BASELINE config 4 -- strong scaling of ONE 16384 x 16384 f32 image (65 taps,
20 iterations, closed walls) over the ranks of a torchrun job: row slabs with a
32-row halo exchange per iteration (rlic_b200.sharded; RLIC_B200_EXCHANGE=peer|nccl).

    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node N \
        --master-addr 127.0.0.1 --master-port P tools/bench_c4_scaling.py [--n 16384] [--iterations 20]

Prints one JSON line on rank 0 (append the lines of N = 1, 2, 4, 8 to
profiles/r2_c4_scaling.jsonl).  Device-resident timing: CUDA events around `reps` calls,
max over ranks.  Parity, after the timing:

* `checksum` -- two EXACT integer checksums of the whole result (the f32 bit patterns summed
  modulo 2^64, plain and weighted by the global row number): they do not depend on how the
  rows are split, so the lines of different N must carry identical numbers;
* `band_parity` -- every rank compares the first and last 32 rows of its slab after
  `--parity-iterations` passes with the CPU oracle run on a sub-image wide enough that its
  artificial cuts cannot reach the band (a walker moves at most 32 rows per pass,
  /root/reference/src/lib.rs:325-329).
"""

from __future__ import annotations

import argparse
import datetime
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
from rlic_b200.sharded import PeerMemoryUnavailable, ShardedConvolver  # noqa: E402


def image_rows(n: int, r0: int, r1: int):
    """Rows [r0, r1) of the global image -- the same whatever the rank count: per-row seeds."""
    seeds = np.random.default_rng(1234).integers(0, 2**31, size=n)
    tex = np.empty((r1 - r0, n), dtype=np.float32)
    for k, r in enumerate(range(r0, r1)):
        tex[k] = np.random.default_rng(int(seeds[r])).random(n, dtype=np.float32)
    y = np.linspace(-1, 1, n)[r0:r1]
    x = np.linspace(-1, 1, n)
    u = np.broadcast_to((-y)[:, None], tex.shape).astype(np.float32)
    v = np.broadcast_to(x[None, :], tex.shape).astype(np.float32)
    return tex, np.ascontiguousarray(u), np.ascontiguousarray(v)


def band_parity(n: int, kernel, iterations: int, a: int, rows: int, mine: np.ndarray, threads: int) -> dict:
    import oracle

    m = iterations * (kernel.size // 2)
    g0, g1 = max(0, a - m), min(n, a + rows + m)
    tex, u, v = image_rows(n, g0, g1)
    want = oracle.convolve(tex, u, v, kernel=kernel, iterations=iterations, threads=threads)[a - g0:a - g0 + rows]
    bad = int(np.count_nonzero(mine.view(np.uint32) != want.view(np.uint32)))
    return {"rows": [a, a + rows], "mismatches": bad, "bit_equal": bad == 0}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=16384)
    ap.add_argument("--iterations", type=int, default=20)
    ap.add_argument("--taps", type=int, default=65)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--parity-iterations", type=int, default=3)
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
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
    _core.check(_core.lib.rlic_b200_set_device(local))

    n = args.n
    kernel = workloads.triangle_kernel(args.taps, np.float32)
    sc = ShardedConvolver(n, n, kernel=kernel, boundaries="closed",
                          exchange=os.environ.get("RLIC_B200_EXCHANGE", "peer"))
    r0, r1 = sc.plan.row0, sc.plan.row1
    tex, u, v = image_rows(n, r0, r1)
    d_tex, d_u, d_v = (torch.from_numpy(a).to(dev) for a in (tex, u, v))
    try:
        sc.set_field(d_u, d_v)
    except PeerMemoryUnavailable as exc:      # raised on every rank alike: go on with NCCL messages
        if rank == 0:
            print(f"peer exchange unavailable ({exc}); using exchange='nccl'", file=sys.stderr)
        sc = ShardedConvolver(n, n, kernel=kernel, boundaries="closed", exchange="nccl")
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
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())

    # exact, split-independent checksums of the full result (arithmetic modulo 2^64)
    bits = out.view(torch.int32).to(torch.int64)
    weight = torch.arange(r0 + 1, r1 + 1, device=dev, dtype=torch.int64)[:, None]
    sums = torch.stack([bits.sum(), (bits * weight).sum()])
    dist.all_reduce(sums)
    del bits

    # band parity against the CPU oracle after a few passes (every halo exchange is exercised)
    import oracle

    few = sc.convolve(d_tex, iterations=args.parity_iterations)
    threads = max(1, oracle.max_threads() // world)
    rows = 32
    mine = [band_parity(n, kernel, args.parity_iterations, r0 + first, rows,
                        few[first:first + rows].cpu().numpy(), threads)
            for first in (0, (r1 - r0) // 2, r1 - r0 - rows)]
    bands = [None] * world
    dist.all_gather_object(bands, mine)

    pix = n * n
    line = {
        "config": "c4", "image": [n, n], "taps": args.taps, "iterations": args.iterations,
        "n_gpus": world, "scaling": "strong", "ms_per_call": ms,
        "Mpix_s": pix * args.iterations / ms / 1e3,
        "G_pixel_steps_s": pix * args.iterations * (args.taps - 1) / ms / 1e6,
        "checksum": [int(sums[0]), int(sums[1])],
        "band_parity": {"iterations": args.parity_iterations,
                        "bit_equal": all(x["bit_equal"] for per_rank in bands for x in per_rank),
                        "per_rank": bands},
        "halo_bytes_per_side_per_iteration": (args.taps // 2) * (n + 2) * 4,
        "exchange": sc.exchange, **{k: v for k, v in _effective().items()},
    }
    sc.close()
    dist.barrier()
    dist.destroy_process_group()
    os.dup2(real_stdout, 1)
    if rank == 0:
        print(json.dumps(line), flush=True)


def _effective() -> dict:
    import rlic_b200

    return rlic_b200.effective_options()


if __name__ == "__main__":
    main()
