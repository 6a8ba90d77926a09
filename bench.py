#!/usr/bin/env python
"""Benchmark of the rlic.convolve hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one full `convolve` call on the headline configuration
(BASELINE.json configs[1]: 4096x4096 f32 noise, analytic vortex, 65-tap triangle
kernel, closed boundaries, iterations=5).  With N > 1 ranks the image is N times
taller ((N*4096) x 4096), split into row slabs with a per-iteration halo
exchange: per-GPU work is fixed, i.e. weak scaling.

Prints ONE JSON line on rank 0.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_SIDE = 4096
TAPS = 65
ITERATIONS = 5
METRIC = "Mpix/s"


# ----------------------------------------------------------------------------
def measured_peak_gbs() -> tuple[float, str]:
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        try:
            return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons; `mark_begin`/`mark_end` bracket the
    timed region and only samples that arrived inside it are summarised (nvidia-smi takes
    longer to start than a short timed region lasts, so it is started before the warm-up)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    PERIOD_MS = 50

    def __init__(self, index):
        """`index`: a GPU ordinal, or a comma-separated list of them (ONE nvidia-smi process for all:
        a poller per rank, every 20 ms, slowed the launches of eight ranks measurably); None
        disables the sampler (ranks other than 0)."""
        self.index = index
        self.proc = None
        self.lines: list[tuple[float, str]] = []
        self.t0 = self.t1 = None

    def __enter__(self):
        if self.index is None:
            return self
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", str(self.PERIOD_MS), "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.05)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self) -> dict:
        sm, smax, power, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        inside = [l for t, l in self.lines if self.t0 is not None and self.t0 <= t <= (self.t1 or t) + 0.06]
        where = "timed region"
        if not inside:   # region shorter than one sampling period: fall back to the whole loaded run
            inside = [l for _, l in self.lines]
            where = "warm-up + timed region"
        per_gpu: dict[str, list[float]] = {}
        for line in inside:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
                power.append(float(parts[3]))
                per_gpu.setdefault(parts[0], []).append(float(parts[1]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[4:8]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        if self.index is None or not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax),
               "power_w_max": max(power), "reasons": sorted(reasons), "samples": len(sm),
               "sampled_during": where}
        if len(per_gpu) > 1:    # one figure per GPU of the job: a slow one must not hide in the median
            out["sm_mhz_per_gpu"] = {k: statistics.median(v) for k, v in sorted(per_gpu.items())}
        return out


def make_slab(rank: int, world: int):
    """This rank's rows of the (world*4096) x 4096 weak-scaling image."""
    from rlic_b200 import workloads

    ny = N_SIDE * world
    rng = np.random.default_rng(rank)
    texture = rng.random((N_SIDE, N_SIDE), dtype=np.float32)
    y = np.linspace(-1, 1, ny, dtype=np.float64)[rank * N_SIDE:(rank + 1) * N_SIDE]
    x = np.linspace(-1, 1, N_SIDE, dtype=np.float64)
    u = np.broadcast_to((-y)[:, None], (N_SIDE, N_SIDE)).astype(np.float32)
    v = np.broadcast_to(x[None, :], (N_SIDE, N_SIDE)).astype(np.float32)
    return texture, u, v, workloads.triangle_kernel(TAPS, np.float32)


def gather_bytes_per_pixel() -> int:
    # u, v once per step, texture once per step + centre, one store
    return (3 * (TAPS - 1) + 2) * 4


# ----------------------------------------------------------------------------
def _timed(fn) -> tuple[float, object]:
    t0 = time.perf_counter()
    out = fn()
    return time.perf_counter() - t0, out


def cpu_single_thread(texture, u, v, kernel, reps: int = 3) -> dict:
    """The reference's own execution model (src/lib.rs:364-406 is one thread under the GIL):
    one pass over a fixed 256-row band in the middle of the image, walkers seeing the whole
    image, median of `reps`."""
    import oracle

    r0, rows = N_SIDE // 2 - 128, 256
    oracle.pass_rows(texture, u, v, kernel=kernel, rows=(r0, r0 + 8), threads=1)      # warm
    secs = sorted(_timed(lambda: oracle.pass_rows(texture, u, v, kernel=kernel, rows=(r0, r0 + rows),
                                                  threads=1))[0] for _ in range(reps))
    dt = secs[len(secs) // 2]
    return {"value": rows * N_SIDE / dt / 1e6, "unit": METRIC, "cores": 1,
            "ns_per_pixel_step": dt / (rows * N_SIDE * (TAPS - 1)) * 1e9,
            "sample": f"one pass over rows [{r0}, {r0 + rows}) of the same image, 1 thread, median of {reps}",
            "seconds": dt}


def cpu_baseline(threads: int | None, reps: int = 5) -> dict:
    """Times the CPU oracle (restatement of the reference's Rust core) on WHOLE passes of the
    same workload -- the full 4096 x 4096 image, every thread the host gives us, helper threads
    pinned -- after one untimed whole pass; the median of `reps` is reported.  The pass-1
    output is kept (`_pass1`) so that the run's parity figures cover the whole image."""
    import oracle

    texture, u, v, kernel = make_slab(0, 1)
    threads = threads or oracle.max_threads()
    run = lambda: oracle.convolve(texture, u, v, kernel=kernel, iterations=1, threads=threads)  # noqa: E731
    _, pass1 = _timed(run)                                   # warm-up: threads, page faults, caches
    secs = sorted(_timed(run)[0] for _ in range(reps))
    dt = secs[len(secs) // 2]
    mpix = N_SIDE * N_SIDE / dt / 1e6
    return {
        "value": mpix, "unit": METRIC, "cores": threads, "kind": "port",
        "sample": (f"whole passes over the 4096x4096 f32 65-tap vortex workload "
                   f"({N_SIDE * N_SIDE * (TAPS - 1) / 1e6:.0f} M pixel-steps each): median of {reps} after one "
                   f"untimed pass, {dt:.2f} s per pass on {threads} pinned thread(s); C restatement of the "
                   "reference's Rust core (oracle/lic_oracle.c, fma+branchless); the reference itself "
                   "is single-threaded, see single_thread"),
        "ns_per_pixel_step": dt / (N_SIDE * N_SIDE * (TAPS - 1)) * 1e9 * threads,
        "seconds": dt, "spread": [secs[0], secs[-1]],
        "_pass1": pass1, "_inputs": (texture, u, v, kernel),
    }


def parity_of_sample(one_pass: np.ndarray, r0: int, band: np.ndarray) -> dict:
    """parity figures reported with the timing.  `band` is the
    oracle's pass-1 output for rows [r0, r0+len(band)) (the CPU-baseline pass: the whole
    image), `one_pass` the CUDA path's iterations=1 result for the whole image."""
    mine = one_pass[r0:r0 + band.shape[0]]
    span = float(band.max() - band.min()) or 1.0
    return {
        "against": "CPU oracle (restatement of src/lib.rs), pass 1",
        "rows": [r0, r0 + band.shape[0]], "pixels": int(band.size),
        "bit_equal_fraction": float(np.mean(mine.view(np.uint32) == band.view(np.uint32))),
        "max_abs_err_over_range": float(np.max(np.abs(mine.astype(np.float64) - band)) / span),
        "tolerance": 1e-5,
    }


def path_divergence(rlic_b200, u: np.ndarray, v: np.ndarray, r0: int, rows: int) -> dict:
    """Fraction of the sample's pixels whose walker visits different pixels on the GPU than
    in the CPU oracle: both run one pass over an exact path-signature texture
    (workloads.path_probe), so any difference is a different path, never rounding."""
    import oracle
    from rlic_b200 import workloads

    probe, ones = workloads.path_probe((N_SIDE, N_SIDE), np.float32, TAPS)
    mine = rlic_b200.convolve(probe, u, v, kernel=ones, boundaries="closed", iterations=1)
    want = oracle.pass_rows(probe, u, v, kernel=ones, rows=(r0, r0 + rows), threads=oracle.max_threads())
    return {"path_divergence_fraction": float(np.mean(mine[r0:r0 + rows] != want)),
            "path_probe": "one pass over random integers < 2^17 with a kernel of ones: exact sums, "
                          "equal iff the same pixels were visited"}


def slab_parity_band(rank: int, world: int, result: np.ndarray, first: int, rows: int = 64) -> dict:
    """Multi-GPU parity, run by every rank after the timing: rows [first, first + rows) of this
    rank's slab of the FINAL result (`result` holds just those rows; all ITERATIONS passes;
    at a slab edge they depend on the neighbour through every halo exchange) against the CPU oracle.  The oracle runs on
    the sub-image of global rows [a - m, a + rows + m), m = ITERATIONS * (TAPS // 2): a walker
    moves at most TAPS // 2 rows per pass (src/lib.rs:325-329), so rows at least m away from
    an artificial cut are exactly the whole image's; the image's real top wall is kept."""
    import oracle

    ny = N_SIDE * world
    a = rank * N_SIDE + first
    m = ITERATIONS * (TAPS // 2)
    g0, g1 = max(0, a - m), min(ny, a + rows + m)
    tex = np.empty((g1 - g0, N_SIDE), dtype=np.float32)
    u = np.empty_like(tex)
    v = np.empty_like(tex)
    for r in sorted({g0 // N_SIDE, (g1 - 1) // N_SIDE}):
        t_r, u_r, v_r, kernel = make_slab(r, world)
        lo, hi = max(g0, r * N_SIDE), min(g1, (r + 1) * N_SIDE)
        src = slice(lo - r * N_SIDE, hi - r * N_SIDE)
        dst = slice(lo - g0, hi - g0)
        tex[dst], u[dst], v[dst] = t_r[src], u_r[src], v_r[src]
    want = oracle.convolve(tex, u, v, kernel=kernel, iterations=ITERATIONS,
                           threads=max(1, oracle.max_threads() // world))[a - g0:a - g0 + rows]
    mine = result
    return {"rows": [a, a + rows], "pixels": int(want.size),
            "bit_equal": bool(np.array_equal(mine.view(np.uint32), want.view(np.uint32))),
            "mismatches": int(np.count_nonzero(mine.view(np.uint32) != want.view(np.uint32)))}


def run_reference(args) -> dict:
    """--impl reference: the CPU implementation of the path on the host cores.  One step is
    one whole `convolve` of the headline workload -- all ITERATIONS passes over the full
    4096 x 4096 image on every host thread (helper threads pinned) -- unless K + W such steps
    would not end within a few minutes on this host, in which case a step is ONE whole pass
    and `ms_per_step` is ITERATIONS times its duration (the passes of this workload cost the
    same: the field does not change and no walker stops early), and the line says so."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return {}
    import oracle

    texture, u, v, kernel = make_slab(0, 1)
    threads = oracle.max_threads()
    one_pass = lambda: oracle.convolve(texture, u, v, kernel=kernel, iterations=1, threads=threads)  # noqa: E731
    one_pass()                                               # untimed: threads, page faults
    t_pass = _timed(one_pass)[0]
    total = args.warmup + args.steps
    whole = total * ITERATIONS * t_pass <= 240.0
    passes = ITERATIONS if whole else 1
    step = lambda: oracle.convolve(texture, u, v, kernel=kernel, iterations=passes, threads=threads)  # noqa: E731
    steps = args.steps if total * passes * t_pass <= 240.0 else max(3, int(240.0 / (passes * t_pass)) - args.warmup)
    for _ in range(args.warmup):
        step()
    secs = sorted(_timed(step)[0] * (ITERATIONS / passes) for _ in range(steps))
    dt = secs[len(secs) // 2]
    mpix = N_SIDE * N_SIDE * ITERATIONS / dt / 1e6
    single = cpu_single_thread(texture, u, v, kernel)
    info = {
        "value": mpix, "unit": METRIC, "cores": threads, "kind": "port",
        "sample": ((f"whole convolve calls ({ITERATIONS} passes over the full 4096x4096 image)" if whole else
                    f"whole single passes over the full 4096x4096 image, scaled by {ITERATIONS}")
                   + f": median of {steps} after {args.warmup} untimed, {dt:.2f} s per {ITERATIONS}-pass step on "
                   f"{threads} pinned thread(s); C restatement of the reference's Rust core "
                   "(oracle/lic_oracle.c, fma+branchless)"),
        "ns_per_pixel_step": dt / (N_SIDE * N_SIDE * (TAPS - 1) * ITERATIONS) * 1e9 * threads,
        "seconds": dt, "spread": [secs[0], secs[-1]], "steps_timed": steps,
        "single_thread": single,
    }
    return {
        "impl": "reference", "metric": METRIC, "value": mpix, "unit": METRIC, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(1),
        "pixel_steps_per_s": mpix * 1e6 * (TAPS - 1),
        "single_thread": {"value": single["value"], "unit": METRIC,
                          "note": "the reference runs on one thread (src/lib.rs:364-406); `value` uses all "
                                  "host threads, which flatters the CPU side"},
        "cpu_baseline": info,
        "e2e": {"value": mpix, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def workload_config(world: int) -> dict:
    return {
        "workload": (f"BASELINE configs[1] per GPU: {N_SIDE * world}x{N_SIDE} f32 noise, analytic vortex, "
                     f"{TAPS}-tap triangle kernel, closed boundaries, iterations={ITERATIONS}"
                     + (f", row slabs over {world} GPUs with a {TAPS // 2}-row halo exchange per iteration"
                        if world > 1 else "")),
        "image": [N_SIDE * world, N_SIDE], "taps": TAPS, "iterations": ITERATIONS,
        "boundaries": "closed", "uv_mode": "velocity",
        "l2": "inputs larger than L2 (padded texture 2 x 64 MiB + packed field 256 MiB + dense in/out 128 MiB per GPU vs 126 MB)",
        "step": "pack field + pad texture + 5 passes + un-pad" if world == 1 else "5 x (edge strips, halo exchange, interior)",
        "passes": "pass 1 walks the streamlines and records each walker's moves; passes 2-5 replay the record "
                  "(a path depends on the field, never on the texture): same bits as walking every pass; "
                  "every step records afresh, nothing is kept between steps (RLIC_B200_PATHS=recompute walks every pass)",
    }


# ----------------------------------------------------------------------------
def run_ours(args) -> dict:
    import torch

    import rlic_b200
    from rlic_b200 import _core

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run for --gpus > 1")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (rlic_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _core.check(_core.lib.rlic_b200_set_device(local))
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    texture, u, v, kernel = make_slab(rank, world)
    exchange_fallback = None
    pixels_local = texture.size
    h2d = texture.nbytes + u.nbytes + v.nbytes + kernel.nbytes
    d2h = texture.nbytes

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident: `value` and the kernel roofline -------
    d_tex = torch.from_numpy(texture).to(dev)
    d_u = torch.from_numpy(np.ascontiguousarray(u)).to(dev)
    d_v = torch.from_numpy(np.ascontiguousarray(v)).to(dev)
    d_out = torch.empty_like(d_tex)

    if world == 1:
        # The same work `rlic_b200_convolve_device_*` does, issued through the
        # C ABI's building blocks so that the pass launches can be bracketed by
        # their own events: pack field, pad texture, [5 passes], un-pad.
        import ctypes

        lib = _core.lib
        p_f32 = ctypes.POINTER(ctypes.c_float)
        slab = (N_SIDE, N_SIDE, 0, N_SIDE, 0, 0)
        closed = (0, 0, 0, 0)
        cells = _core.padded_cells(N_SIDE, N_SIDE)
        pad_a = torch.empty(cells, dtype=torch.float32, device=dev)
        pad_b = torch.empty(cells, dtype=torch.float32, device=dev)
        field = torch.empty(4 * cells, dtype=torch.float32, device=dev)
        taps_ptr = kernel.ctypes.data_as(p_f32)
        # what rlic_b200_convolve_device_* does with the library's options in force: the first pass
        # records the streamline paths, the others replay them -- or, with
        # RLIC_B200_PATHS=recompute (or the per-step walk / `fma` arithmetic), every pass walks
        opts = rlic_b200.effective_options()
        replay = (opts["paths"] == "replay" and opts["walk"] == "grouped" and opts["arithmetic"] == "fma+branchless"
                  and ITERATIONS >= 2)
        record = (torch.empty(_core.path_record_bytes(N_SIDE, N_SIDE, TAPS) // 4, dtype=torch.int32, device=dev)
                  if replay else None)

        def step(events=None, replay=replay):
            st = int(torch.cuda.current_stream().cuda_stream)
            _core.check(lib.rlic_b200_slab_pack_field_f32(d_u.data_ptr(), d_v.data_ptr(), *slab, *closed,
                                                          field.data_ptr(), st))
            _core.check(lib.rlic_b200_slab_pad_texture_f32(d_tex.data_ptr(), *slab, *closed,
                                                           pad_a.data_ptr(), st))
            if events is not None:
                events[0].record()
            src, dst = pad_a, pad_b
            for it in range(ITERATIONS):
                if replay:
                    _core.check(lib.rlic_b200_pass_slab_paths_f32(
                        src.data_ptr(), field.data_ptr(), dst.data_ptr(), *slab, 0, N_SIDE, taps_ptr, kernel.size,
                        0, *closed, None, 0, _core.PASS_RECORD if it == 0 else _core.PASS_REPLAY,
                        record.data_ptr(), st))
                else:
                    _core.check(lib.rlic_b200_pass_slab_f32(src.data_ptr(), field.data_ptr(), dst.data_ptr(),
                                                            *slab, 0, N_SIDE, taps_ptr, kernel.size, 0,
                                                            *closed, st))
                if it == 0 and events is not None:
                    events[2].record()           # between the first pass and the others
                src, dst = dst, src
            if events is not None:
                events[1].record()
            _core.check(lib.rlic_b200_slab_unpad_texture_f32(src.data_ptr(), *slab, *closed,
                                                             d_out.data_ptr(), st))
            return d_out
    else:
        from rlic_b200 import sharded

        # the halo exchange fused into the edge-strip kernels (peer stores over NVLink, counters in
        # the neighbours' memory); RLIC_B200_EXCHANGE=nccl selects point-to-point NCCL messages
        wanted = os.environ.get("RLIC_B200_EXCHANGE", "peer")
        sc = sharded.ShardedConvolver(N_SIDE * world, N_SIDE, kernel=kernel, boundaries="closed", exchange=wanted)
        try:
            sc.set_field(d_u, d_v)
        except sharded.PeerMemoryUnavailable as exc:
            # raised on every rank alike (no peer access between these GPUs, CUDA IPC refused ...):
            # the job goes on with the NCCL exchange, and the line says so
            exchange_fallback = f"{wanted} -> nccl: {exc}"
            sc = sharded.ShardedConvolver(N_SIDE * world, N_SIDE, kernel=kernel, boundaries="closed", exchange="nccl")
            sc.set_field(d_u, d_v)

        replay = False

        def step(events=None):
            if events is not None:
                events[0].record()
                events[2].record()
            out = sc.convolve(d_tex, iterations=ITERATIONS)
            if events is not None:
                events[1].record()
            return out

    ev = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(3)) for _ in range(args.steps)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # one sampler for the job: rank 0 watches every GPU the job uses
    with ClockSampler(",".join(str(i) for i in range(world)) if rank == 0 else None) as clocks:
        for _ in range(max(args.warmup, 3)):
            step()
        barrier()
        launches0 = _core.launch_count()
        clocks.mark_begin()
        t_begin.record()
        for k in range(args.steps):
            result = step(ev[k])
        t_end.record()
        barrier()
        clocks.mark_end()
        launches = _core.launch_count() - launches0
    total_ms = t_begin.elapsed_time(t_end)
    pass_ms = sum(a.elapsed_time(b) for a, b, _ in ev)   # the 5 pass launches of every step
    every_pass_walks = None
    if world == 1 and replay:
        # The same step with every pass WALKING its streamlines, as the reference does in every
        # iteration (RLIC_B200_PATHS=recompute), timed the same way right after the timed region:
        # the line then says how much of `value` is the recorded paths and how much the walk kernel.
        try:
            expected = result.clone()
            for _ in range(3):
                step(replay=False)
            w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            w0.record()
            for _ in range(args.steps):
                walked = step(replay=False)
            w1.record()
            w1.synchronize()
            walk_ms = w0.elapsed_time(w1) / args.steps
            every_pass_walks = {
                "value": pixels_local * ITERATIONS / (walk_ms * 1e-3) / 1e6, "unit": METRIC, "ms_per_step": walk_ms,
                "steps": args.steps, "result_bit_equal_to_the_timed_step": bool(torch.equal(walked, expected)),
                "note": "the same step with the record/replay switched off (five launches of the walking "
                        "lic_pass_kernel), device-resident, CUDA events; not the headline, for comparison",
            }
        except Exception as exc:  # noqa: BLE001 -- a reporting extra
            every_pass_walks = {"error": f"{type(exc).__name__}: {exc}"}
        else:
            result.copy_(expected)       # `result` is the timed step's own output for everything below
            del expected
    first_ms = sum(a.elapsed_time(m) for a, _, m in ev) / args.steps      # world == 1: the first pass alone
    later_ms = sum(m.elapsed_time(b) for _, b, m in ev) / args.steps / max(ITERATIONS - 1, 1)   # each of the others
    per_rank_ms = solo_per_rank_ms = None
    if dist is not None:
        t = torch.tensor([total_ms, pass_ms], device=dev, dtype=torch.float64)
        # every rank's own device time per step travels with the line: `value` is their maximum,
        # and a spread between ranks (or between runs) can be attributed to a GPU
        mine = [torch.zeros(1, device=dev, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(mine, t[:1] / args.steps)
        per_rank_ms = [float(x.item()) for x in mine]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, pass_ms = t.tolist()
        n_l = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(n_l)
        launches = int(n_l.item())
        # The ranks of a sharded call advance in lock step (each pass waits for the neighbours'
        # halos), so their times per step above agree whichever GPU sets the pace.  What tells a
        # slow GPU apart is its speed ALONE: one walking pass over its own slab, no exchange, a
        # few repetitions, every rank at once after the barrier that ended the timed region.
        solo_ms = -1.0
        try:
            pads = sc._peer.bufs if sc._peer is not None else sc._work[1:3]
            run_solo = lambda: sc.ops.pass_rows(pads[0], sc.field, pads[1], sc.plan, 0, sc.plan.nrows,  # noqa: E731
                                                sc.taps, sc.mode, sc.walls)
            run_solo()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for _ in range(3):
                run_solo()
            s1.record()
            s1.synchronize()
            solo_ms = s0.elapsed_time(s1) / 3
        except Exception:  # noqa: BLE001 -- a diagnostic; the gather below must still be entered by every rank
            solo_ms = -1.0
        mine = [torch.zeros(1, device=dev, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(mine, torch.tensor([solo_ms], device=dev, dtype=torch.float64))
        solo_per_rank_ms = [float(x.item()) for x in mine]
        barrier()
    ms_per_step = total_ms / args.steps
    pixels_all = pixels_local * world
    value = pixels_all * ITERATIONS / (ms_per_step * 1e-3) / 1e6

    passes = args.steps * ITERATIONS
    pass_avg_ms = pass_ms / passes
    peak, peak_src = measured_peak_gbs()
    if world == 1 and replay:
        # `roofline` proper describes the walking pass -- the kernel SURVEY.md section 8(d)'s
        # algorithmic bytes are defined for, and the longest launch of the step; the replay kernel
        # (four launches per step) has its own object, roofline["replay"], below
        pass_avg_ms = first_ms
    achieved = gather_bytes_per_pixel() * pixels_local / (pass_avg_ms * 1e-3) / 1e9
    traffic = None
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists() and world == 1:
        try:
            traffic = json.loads(tf.read_text()).get("lic_pass_kernel_dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {
        "bound": "hbm", "kernel": "lic_pass_kernel<float,...> (one pass over this GPU's pixels)",
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "peak_source": peak_src, "traffic": traffic,
        "algorithmic_bytes_per_launch": gather_bytes_per_pixel() * pixels_local,
        "launch_ms": pass_avg_ms,
        "note": "achieved = (3*(L-1)+2)*4 gather bytes per pixel x pixels per launch / mean launch time",
    }

    if world > 1:
        # one figure for the whole call here: the strips, the exchange and the interior of a pass
        # are several launches, so the call's passes are timed together and averaged
        if getattr(sc, "_call_paths", False):
            roofline["kernel"] = (f"mean over the {ITERATIONS} passes of a call on this GPU's slab: pass 1 walks and "
                                  "records (lic_pass_kernel<float,...,REC>), the others replay (lic_replay_kernel), "
                                  "each as two edge strips + interior with the halo exchange fused into the strips' "
                                  "stores; the N = 1 line has one roofline object per kernel")
            roofline["note"] = ("achieved = section 8(d)'s bytes of the REFERENCE pass, (3*(L-1)+2)*4 per pixel, over "
                                "the mean pass time: a replayed pass moves fewer (the N = 1 line's roofline.replay)")
        else:
            roofline["kernel"] = (f"mean over the {ITERATIONS} passes of a call on this GPU's slab "
                                  "(lic_pass_kernel<float,...>: two edge strips + interior per pass, halo exchange included)")
    if world == 1 and replay:
        groups = 2 * ((TAPS // 2 + 31) // 32)
        replay_bytes = (TAPS - 1) * 4 + 4 + 4 + 3 * 4 * groups      # gathers + centre + store + three planes per group
        roofline["kernel"] = ("lic_pass_kernel<float,...,REC> (pass 1: the walk, recording the paths: the longest "
                              f"launch of the step, {first_ms:.3f} of the {pass_ms / args.steps:.3f} ms its passes take; "
                              f"the four replay launches take {(ITERATIONS - 1) * later_ms:.3f} ms: roofline.replay)")
        roofline["algorithmic_bytes_per_launch"] = gather_bytes_per_pixel() * pixels_local
        roofline["step_share"] = {"walk_record_ms": first_ms, "replay_ms_each": later_ms,
                                  "replay_launches": ITERATIONS - 1, "passes_ms": pass_ms / args.steps}
        roofline["replay"] = {
            "bound": "hbm", "kernel": "lic_replay_kernel<float,...> (passes 2-5: one launch each)",
            "achieved": replay_bytes * pixels_local / (later_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
            "frac": replay_bytes * pixels_local / (later_ms * 1e-3) / 1e9 / peak,
            "algorithmic_bytes_per_launch": replay_bytes * pixels_local, "launch_ms": later_ms,
            "note": f"algorithmic bytes per pixel = (L-1)*4 texture gathers + centre + store + {3 * groups} plane words "
                    f"x 4 = {replay_bytes} (the replay reads no field); nominal like the fraction above: the gathers "
                    "are served by L1/L2",
            "vs_walking_pass": first_ms / later_ms,
            # the same pass in section 8(d)'s accounting of the REFERENCE algorithm (u, v and texture per step)
            "reference_pass_equivalent_GBps": gather_bytes_per_pixel() * pixels_local / (later_ms * 1e-3) / 1e9,
        }
        try:
            tr = json.loads(tf.read_text())
            roofline["traffic"] = tr.get("lic_pass_record_kernel_dram_bytes_per_launch", traffic)
            roofline["replay"]["traffic"] = tr.get("lic_replay_kernel_dram_bytes_per_launch")
        except Exception:
            roofline["replay"]["traffic"] = None
    roofline["frac_note"] = ("nominal: section 8(d)'s algorithmic gather bytes over the HBM copy peak; the gathers "
                             "are served by L1/L2 (DRAM traffic per launch is the compulsory 0.4 GB), so this "
                             "fraction exceeds 1 -- the bound that holds is gather_peak below")
    if world == 1:
        # The memory system's own limit for this access pattern, measured now, on this box
        # (SURVEY.md section 8(d)): the library's gather-ceiling probe -- the loads of a pass and
        # the tap FMA, nothing else -- on the very buffers the timed passes used.  Best of the
        # two forms (address chain through the loaded record, or free), best of 5 launches each.
        try:
            best = {}
            for dependent in (1, 0):
                times = []
                for _ in range(6):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    _core.check(lib.rlic_b200_measure_gather_ceiling_f32(
                        pad_a.data_ptr(), field.data_ptr(), pad_b.data_ptr(), N_SIDE, N_SIDE, taps_ptr,
                        kernel.size, dependent, int(torch.cuda.current_stream().cuda_stream)))
                    b.record()
                    b.synchronize()
                    times.append(a.elapsed_time(b))
                best["dependent" if dependent else "free"] = min(times[1:])
            ceil_ms = min(best.values())
            peak_g = gather_bytes_per_pixel() * pixels_local / (ceil_ms * 1e-3) / 1e9
            roofline["gather_peak"] = {
                "GBps": peak_g, "launch_ms": ceil_ms, "forms_ms": best,
                "how": "rlic_b200_measure_gather_ceiling_f32 (gather_ceiling_kernel in lic_walk.cuh): same loads "
                       "per step as the walk (one 16-byte field record, one texture value) + the tap FMA, "
                       "staircase walkers, same image, same launch shape; CUDA events, best of 5, this run",
            }
            roofline["frac_of_gather_peak"] = achieved / peak_g
            if replay:
                # the replay kernel's own ceiling: the same probe without the field record
                times = []
                for _ in range(6):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    _core.check(lib.rlic_b200_measure_gather_ceiling_f32(
                        pad_a.data_ptr(), field.data_ptr(), pad_b.data_ptr(), N_SIDE, N_SIDE, taps_ptr,
                        kernel.size, 2, int(torch.cuda.current_stream().cuda_stream)))
                    b.record()
                    b.synchronize()
                    times.append(a.elapsed_time(b))
                rp = roofline["replay"]
                peak_r = rp["algorithmic_bytes_per_launch"] / (min(times[1:]) * 1e-3) / 1e9
                rp["gather_peak"] = {"GBps": peak_r, "launch_ms": min(times[1:]),
                                     "how": "the same probe with dependent=2: one texture value per step and the "
                                            "tap FMA, no field record; best of 5, this run"}
                rp["frac_of_gather_peak"] = rp["achieved"] / peak_r
        except Exception as exc:  # noqa: BLE001 -- a reporting extra
            roofline["gather_peak"] = {"error": f"{type(exc).__name__}: {exc}"}

    # Secondary ceiling: instruction issue.  The pass kernel is
    # issue-bound, not memory-bound: warp instructions per pixel-step
    # as ncu counted them for the kernel of the walk in force, against 4 issue slots per SM per
    # clock at the SM clock sampled during the timed region.  Arithmetic on recorded figures only.
    try:
        walk = rlic_b200.effective_options()["walk"]
        src = {"per-step": "r1_pass_kernel_ncu_summary.json", "grouped": "r2_pass_kernel_ncu_summary.json"}[walk]
        if world == 1 and replay and (ROOT / "profiles" / "r2_record_kernel_ncu_summary.json").exists():
            src = "r2_record_kernel_ncu_summary.json"
        prof = json.loads((ROOT / "profiles" / src).read_text())
        ips = float(prof["warp_instructions_per_pixel_step"])
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        mhz = clocks.summary().get("sm_mhz") or clocks.summary().get("sm_max_mhz")
        if world == 1 and mhz:
            warp_instr = pixels_local * (TAPS - 1) / 32 * ips
            ceiling_ms = warp_instr / (4 * sms * mhz * 1e6) * 1e3
            roofline["issue"] = {
                "warp_instructions_per_pixel_step": ips,
                "source": f"profiles/{src} (smsp__inst_executed.sum / pixel-steps)",
                "sms": sms, "sm_mhz": mhz, "ceiling_ms": ceiling_ms, "frac": ceiling_ms / pass_avg_ms,
                "note": "launch time if every issue slot of every SM issued a warp instruction of this kernel",
            }
    except Exception as exc:  # noqa: BLE001 -- a reporting extra, never allowed to fail the line
        roofline["issue"] = {"error": f"{type(exc).__name__}: {exc}"}
    if world == 1 and replay:
        try:
            src = "r2_replay_kernel_ncu_summary.json"
            ips = float(json.loads((ROOT / "profiles" / src).read_text())["warp_instructions_per_pixel_step"])
            mhz = clocks.summary().get("sm_mhz") or clocks.summary().get("sm_max_mhz")
            ceiling_ms = pixels_local * (TAPS - 1) / 32 * ips / (4 * sms * mhz * 1e6) * 1e3
            roofline["replay"]["issue"] = {"warp_instructions_per_pixel_step": ips, "source": f"profiles/{src}",
                                           "ceiling_ms": ceiling_ms, "frac": ceiling_ms / later_ms}
        except Exception as exc:  # noqa: BLE001
            roofline["replay"]["issue"] = {"error": f"{type(exc).__name__}: {exc}"}

    # ---------------- end to end through the public API, host buffers ---------
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
    h_tex, h_u, h_v = pin(texture), pin(u), pin(v)
    e2e = None
    if world == 1:
        for _ in range(3):   # warm-up, holding the result exactly as the timed loop does
            out = rlic_b200.convolve(h_tex, h_u, h_v, kernel=kernel, boundaries="closed",
                                     iterations=ITERATIONS)
        torch.cuda.synchronize()
        n_e2e = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        laps = [t0]
        for _ in range(n_e2e):
            out = rlic_b200.convolve(h_tex, h_u, h_v, kernel=kernel, boundaries="closed",
                                     iterations=ITERATIONS)
            laps.append(time.perf_counter())
        dt = (laps[-1] - t0) / n_e2e             # the figure reported: the mean over the timed calls
        e2e_val = pixels_all * ITERATIONS / dt / 1e6
        e2e_ms = dt * 1e3
        e2e_calls_ms = sorted((b - a) * 1e3 for a, b in zip(laps, laps[1:]))
        assert np.array_equal(out, result.cpu().numpy()), "host and device paths disagree"
        # the same call as a drop-in user makes it: ordinary (pageable) NumPy arrays
        p_tex, p_u, p_v = texture.copy(), np.ascontiguousarray(u).copy(), np.ascontiguousarray(v).copy()
        for _ in range(2):
            out_p = rlic_b200.convolve(p_tex, p_u, p_v, kernel=kernel, boundaries="closed", iterations=ITERATIONS)
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            out_p = rlic_b200.convolve(p_tex, p_u, p_v, kernel=kernel, boundaries="closed", iterations=ITERATIONS)
        dt_p = (time.perf_counter() - t0) / n_e2e
        assert np.array_equal(out_p, out)
        pageable = {"value": pixels_all * ITERATIONS / dt_p / 1e6, "unit": METRIC, "ms_per_step": dt_p * 1e3,
                    "api": "rlic_b200.convolve(ordinary pageable numpy arrays)"}
    else:
        # each rank: its slab of texture, u and v from page-locked host memory through
        # ShardedConvolver.convolve_host (band pipeline: uploads, passes and downloads overlap;
        # the field is re-packed and its halos re-sent every step, as a convolve(texture, u, v)
        # call implies) into a page-locked host slab; the call returns when the slab is there
        from rlic_b200.sharded import pinned_empty

        pinned_out = pinned_empty(texture.shape, np.float32)

        def e2e_step():
            sc.convolve_host(h_tex, h_u, h_v, iterations=ITERATIONS, out=pinned_out)

        for _ in range(2):
            e2e_step()
        assert np.array_equal(pinned_out, result.cpu().numpy()), "host and device paths disagree"
        barrier()
        n_e2e = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_step()
        dt = time.perf_counter() - t0
        barrier()
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = t.item() / n_e2e * 1e3
        e2e_val = pixels_all * ITERATIONS / (e2e_ms * 1e-3) / 1e6
    # The floor the host link sets for such a step: nothing but the step's copies -- the three
    # input slabs up on one stream, a result-sized slab down on another, page-locked memory,
    # every rank at once -- no kernels.  On a box whose GPUs share host links or host memory
    # bandwidth this, not the compute, bounds the end-to-end figure at N > 1.
    up_s, down_s = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    staging = [torch.empty(texture.shape, dtype=torch.float32, device=dev) for _ in range(4)]
    host_in = [torch.from_numpy(a) for a in (h_tex, h_u, h_v)]
    host_out = torch.empty(texture.shape, dtype=torch.float32).pin_memory()

    def copies_only():
        with torch.cuda.stream(up_s):
            for dst, src in zip(staging, host_in):
                dst.copy_(src, non_blocking=True)
        with torch.cuda.stream(down_s):
            host_out.copy_(staging[3], non_blocking=True)
        up_s.synchronize()
        down_s.synchronize()

    copies_only()
    barrier()
    t0 = time.perf_counter()
    for _ in range(5):
        copies_only()
    dt_copy = (time.perf_counter() - t0) / 5
    barrier()
    if dist is not None:
        t = torch.tensor([dt_copy], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt_copy = t.item()
    del staging
    host_link = {"copies_only_ms_per_step": dt_copy * 1e3,
                 "aggregate_GBps": (h2d + d2h) * world / dt_copy / 1e9,
                 "note": "the step's host<->device copies alone (3 slabs up, 1 down, pinned, all ranks at once): "
                         "the floor the host link sets for ms_per_step"}
    e2e = {"value": e2e_val, "unit": METRIC, "h2d_bytes_per_step": h2d * world,
           "d2h_bytes_per_step": d2h * world, "ms_per_step": e2e_ms,
           "api": "rlic_b200.convolve(numpy arrays in pinned host memory)" if world == 1
                  else "per rank: ShardedConvolver.convolve_host(texture, u, v slabs in pinned host memory) "
                       "-> pinned host slab",
           **rlic_b200.effective_options()}
    e2e["host_link"] = host_link
    if world > 1:
        e2e["exchange"] = sc.exchange
        if exchange_fallback:
            e2e["exchange_fallback"] = exchange_fallback
    else:
        e2e["pageable"] = pageable
        e2e["calls"] = {"n": len(e2e_calls_ms), "min_ms": e2e_calls_ms[0],
                        "median_ms": statistics.median(e2e_calls_ms), "max_ms": e2e_calls_ms[-1]}

    line = {
        "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(world),
        "pixel_steps_per_s": value * 1e6 * (TAPS - 1),
        "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
    }
    if per_rank_ms is not None:
        line["ms_per_step_per_rank"] = per_rank_ms
        line["solo_walking_pass_ms_per_rank"] = solo_per_rank_ms
    if every_pass_walks is not None:
        line["every_pass_walks"] = every_pass_walks
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = cpu_baseline(None)
        pass1 = cb.pop("_pass1")
        cb.pop("_inputs")
        line["cpu_baseline"] = cb
        try:   # a reporting extra: never allowed to take the bench line down with it
            one_pass = rlic_b200.convolve(h_tex, h_u, h_v, kernel=kernel, boundaries="closed", iterations=1)
            line["parity"] = parity_of_sample(one_pass, 0, pass1)
            line["parity"].update(path_divergence(rlic_b200, h_u, h_v, 0, N_SIDE))
        except Exception as exc:  # noqa: BLE001
            line["parity"] = {"error": f"{type(exc).__name__}: {exc}"}
        try:
            # ... and the timed step's own result -- all ITERATIONS passes, i.e. the recorded paths
            # replayed four times -- on three bands (the top wall, the middle, the bottom wall)
            bands = [slab_parity_band(0, 1, result[first:first + 64].cpu().numpy(), first)
                     for first in (0, N_SIDE // 2 - 32, N_SIDE - 64)]
            line["parity"]["all_passes"] = {
                "against": f"CPU oracle, all {ITERATIONS} passes (which walks every pass), on three 64-row bands of "
                           "the device-resident step's result",
                "bit_equal": all(b["bit_equal"] for b in bands), "bands": bands}
        except Exception as exc:  # noqa: BLE001
            line.setdefault("parity", {})["all_passes"] = {"error": f"{type(exc).__name__}: {exc}"}
        cb["single_thread"] = cpu_single_thread(texture, u, v, kernel)
    elif rank == 0:
        line["cpu_baseline"] = None
    if world > 1:
        # multi-GPU parity, every rank: a band of the final result whose value went through
        # every halo exchange, against the CPU oracle (see slab_parity_band)
        mine = []
        for first in (0, N_SIDE - 64):
            try:
                mine.append(slab_parity_band(rank, world, result[first:first + 64].cpu().numpy(), first))
            except Exception as exc:  # noqa: BLE001
                mine.append({"error": f"{type(exc).__name__}: {exc}", "bit_equal": False})
        bands = [None] * world
        dist.all_gather_object(bands, mine)
        if rank == 0:
            line["parity"] = {
                "against": f"CPU oracle (restatement of src/lib.rs), all {ITERATIONS} passes, on the first and "
                           "the last 64 rows of every rank's slab (rows that depend on the neighbours' halos)",
                "bit_equal": all(b.get("bit_equal") for pair in bands for b in pair), "per_rank": bands,
            }
    if dist is not None:
        sc.close()           # peer mappings of the fused exchange, if any (collective)
        dist.barrier()
        dist.destroy_process_group()
    return line if rank == 0 else {}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    # Only the JSON line may reach stdout: libraries (NCCL prints its version at
    # communicator creation) get stderr instead, at the file-descriptor level.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        line = run_reference(args) if args.impl == "reference" else run_ours(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    if line:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
