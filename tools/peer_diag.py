#!/usr/bin/env python
"""Where a sharded call's time goes, rank by rank (diagnostics for the fused peer exchange).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P tools/peer_diag.py [--calls 10] [--check sync|lazy|off]

bench.py's weak-scaling workload (a 4096 x 4096 f32 slab per rank, 65 taps, 5 iterations).  Every
operation of ShardedConvolver._peer_passes is followed by a timing event; the table printed by
rank 0 gives, per rank, the milliseconds per call spent up to each kind of event (a wait's
share is the time the stream sat in the one-thread polling kernel, i.e. waiting for a neighbour).
"""
from __future__ import annotations

import argparse
import collections
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from rlic_b200 import _core  # noqa: E402
from rlic_b200.sharded import ShardedConvolver  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--calls", type=int, default=10)
    ap.add_argument("--check", default=os.environ.get("RLIC_B200_PEER_CHECK", "lazy"))
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=dev)
    _core.check(_core.lib.rlic_b200_set_device(local))
    texture, u, v, kernel = bench.make_slab(rank, world)
    sc = ShardedConvolver(bench.N_SIDE * world, bench.N_SIDE, kernel=kernel, boundaries="closed",
                          exchange=os.environ.get("RLIC_B200_EXCHANGE", "peer"))
    sc.check_peer_timeouts = {"sync": True, "lazy": "lazy", "off": False}[args.check]
    d_tex, d_u, d_v = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (texture, u, v))
    sc.set_field(d_u, d_v)
    for _ in range(3):
        sc.convolve(d_tex, iterations=bench.ITERATIONS)
    torch.cuda.synchronize()
    dist.barrier()
    totals = collections.OrderedDict()
    wall0 = time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    traces = []
    for _ in range(args.calls):
        sc.trace = []
        sc.convolve(d_tex, iterations=bench.ITERATIONS)
        traces.append(sc.trace)
    sc.trace = None
    b.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - wall0) / args.calls * 1e3
    for trace in traces:
        for (_, e0), (label, e1) in zip(trace, trace[1:]):
            totals[label] = totals.get(label, 0.0) + e0.elapsed_time(e1) / args.calls
    mine = {"rank": rank, "ms_per_call_events": a.elapsed_time(b) / args.calls, "ms_per_call_wall": wall,
            "phases_ms_per_call": {k: round(x, 4) for k, x in totals.items()}}
    everyone = [None] * world
    dist.all_gather_object(everyone, mine)
    sc.close()
    dist.barrier()
    dist.destroy_process_group()
    os.dup2(real_stdout, 1)
    if rank == 0:
        print(json.dumps({"n_gpus": world, "exchange": sc.exchange, "check": args.check, "ranks": everyone}))


if __name__ == "__main__":
    main()
