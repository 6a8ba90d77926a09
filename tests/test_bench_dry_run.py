"""bench.py's single-GPU arm, run from end to end WITHOUT a GPU.

The timed path needs a B200, but everything bench.py does around it -- sequencing the C-ABI
building blocks, the comparison leg with the recorded paths switched off, the rooflines, the
end-to-end loops, the parity blocks, the JSON line the driver records -- is host logic, and a
slip there (a name defined on one branch only, a key missing from the line) would cost the
round's measurement.  Here the CUDA side is replaced by stand-ins: ``torch.cuda`` by inert
objects, the compute entry points of the library by stubs that deliver the CPU oracle's result
where the real ones would deliver the kernels', and ``rlic_b200.convolve`` by the oracle.  The
numbers in the resulting line mean nothing; its shape, and that every block ran without an
error entry, is what is checked.  (The oracle is used as the checker's stand-in inside a test,
which is where it is allowed; nothing here is a product path.)
"""
from __future__ import annotations

import argparse
import contextlib
import ctypes
import json
import sys
import time
import types
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import oracle  # noqa: E402
import rlic_b200  # noqa: E402
from rlic_b200 import _core  # noqa: E402

SIDE = 512


class _Event:
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def synchronize(self):
        pass

    def elapsed_time(self, other):
        return max((other.t - self.t) * 1e3, 1e-3)


class _Stream:
    cuda_stream = 0

    def __init__(self, device=None):
        pass

    def synchronize(self):
        pass

    def wait_event(self, ev):
        pass


class _StubLib:
    """The entry points run_ours() calls, with the real library behind everything else (the option
    getters, padded_cells, path_record_bytes, launch_count: host-only functions)."""

    def __init__(self, real, final):
        self._real = real
        self._final = np.ascontiguousarray(final)
        self.calls: list[str] = []
        self.passes_since_pad = 0
        self.pass_kinds: list[list[int]] = []

    def __getattr__(self, name):
        return getattr(self._real, name)

    def rlic_b200_set_device(self, device):
        return 0

    def rlic_b200_slab_pack_field_f32(self, *a):
        assert len(a) == 2 + 6 + 4 + 2
        self.calls.append("pack")
        return 0

    def rlic_b200_slab_pad_texture_f32(self, *a):
        assert len(a) == 1 + 6 + 4 + 2
        self.calls.append("pad")
        self.passes_since_pad = 0
        self.pass_kinds.append([])
        return 0

    def rlic_b200_pass_slab_paths_f32(self, *a):
        assert len(a) == 3 + 6 + 2 + 2 + 1 + 4 + 2 + 1 + 2     # see _core._signatures("pass_slab_paths")
        self.passes_since_pad += 1
        self.pass_kinds[-1].append(int(a[-3]))
        return 0

    def rlic_b200_pass_slab_f32(self, *a):
        assert len(a) == 3 + 6 + 2 + 2 + 1 + 4 + 1
        self.passes_since_pad += 1
        self.pass_kinds[-1].append(_core.PASS_WALK)
        return 0

    def rlic_b200_slab_unpad_texture_f32(self, *a):
        assert self.passes_since_pad == bench.ITERATIONS
        ctypes.memmove(int(a[-2]), self._final.ctypes.data, self._final.nbytes)    # what five passes leave
        self.calls.append("unpad")
        return 0

    def rlic_b200_measure_gather_ceiling_f32(self, *a):
        assert len(a) == 3 + 2 + 2 + 1 + 1
        return 0


@pytest.fixture
def cpu_stand_ins(monkeypatch):
    monkeypatch.setattr(bench, "N_SIDE", SIDE)
    for name in ("WORLD_SIZE", "RANK", "LOCAL_RANK"):
        monkeypatch.delenv(name, raising=False)
    texture, u, v, kernel = bench.make_slab(0, 1)
    final = oracle.convolve(texture, u, v, kernel=kernel, iterations=bench.ITERATIONS, threads=oracle.max_threads())

    cpu = torch.device("cpu")
    monkeypatch.setattr(torch, "device", lambda *a, **k: cpu)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    cuda = torch.cuda
    monkeypatch.setattr(cuda, "is_available", lambda: True)
    monkeypatch.setattr(cuda, "set_device", lambda d: None)
    monkeypatch.setattr(cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(cuda, "Event", _Event)
    monkeypatch.setattr(cuda, "Stream", _Stream)
    monkeypatch.setattr(cuda, "current_stream", lambda *a: _Stream())
    monkeypatch.setattr(cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(cuda, "get_device_properties", lambda d: types.SimpleNamespace(multi_processor_count=148))

    # no nvidia-smi here: the sampler starts nothing (OSError) and would report no clock
    monkeypatch.setattr(bench.ClockSampler, "summary", lambda self: {
        "sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "power_w_max": 300.0, "reasons": [], "samples": 2,
        "sampled_during": "timed region"})

    stub = _StubLib(_core.lib, final)
    monkeypatch.setattr(_core, "lib", stub)

    def convolve(texture, u, v, *, kernel, boundaries, iterations):
        assert boundaries == "closed"
        return oracle.convolve(np.ascontiguousarray(texture), np.ascontiguousarray(u), np.ascontiguousarray(v),
                               kernel=kernel, iterations=iterations, threads=oracle.max_threads())

    monkeypatch.setattr(rlic_b200, "convolve", convolve)
    return stub


def test_single_gpu_arm_builds_its_line_end_to_end(cpu_stand_ins):
    stub = cpu_stand_ins
    args = argparse.Namespace(gpus=1, steps=2, warmup=1, impl="ours", no_cpu_baseline=False)
    line = bench.run_ours(args)
    line = json.loads(json.dumps(line))                      # what the driver parses

    # the contract's keys
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in line, key
    assert line["metric"] == line["unit"] == "Mpix/s" and line["n_gpus"] == 1 and line["steps"] == 2
    assert line["warmup"] >= 3 and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["dtype"] == "f32" and line["data"] == "synthetic" and "workload" in line["config"]
    assert line["value"] > 0 and line["ms_per_step"] > 0

    # the default options: pass 1 records, the others replay -- in every step of the timed region
    # and of the warm-up; the comparison leg walks every pass
    kinds = stub.pass_kinds
    recorded = [k for k in kinds if k[0] == _core.PASS_RECORD]
    walked = [k for k in kinds if k[0] == _core.PASS_WALK]
    assert len(recorded) == 3 + 2 and len(walked) == 3 + 2
    assert all(k == [_core.PASS_RECORD] + [_core.PASS_REPLAY] * (bench.ITERATIONS - 1) for k in recorded)
    assert all(k == [_core.PASS_WALK] * bench.ITERATIONS for k in walked)
    assert "replay" in line["config"]["passes"]

    roof = line["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic", "launch_ms", "algorithmic_bytes_per_launch",
                "step_share", "replay", "gather_peak", "frac_of_gather_peak", "issue"):
        assert key in roof, key
    assert "error" not in roof["gather_peak"] and "error" not in roof["issue"]
    assert roof["bound"] == "hbm" and roof["unit"] == "GB/s"
    assert roof["algorithmic_bytes_per_launch"] == 776 * SIDE * SIDE
    rep = roof["replay"]
    assert rep["algorithmic_bytes_per_launch"] == 288 * SIDE * SIDE
    assert "error" not in rep.get("issue", {}) and "gather_peak" in rep and "frac_of_gather_peak" in rep
    assert roof["traffic"] and rep["traffic"]                 # profiles/traffic.json is read

    walks = line["every_pass_walks"]
    assert "error" not in walks, walks
    assert walks["result_bit_equal_to_the_timed_step"] is True and walks["steps"] == 2 and walks["value"] > 0

    e2e = line["e2e"]
    for key in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step", "ms_per_step", "api", "paths",
                "host_link", "pageable", "calls"):
        assert key in e2e, key
    assert e2e["h2d_bytes_per_step"] == 3 * 4 * SIDE * SIDE + 4 * bench.TAPS and e2e["d2h_bytes_per_step"] == 4 * SIDE * SIDE
    assert e2e["calls"]["n"] == 3 and e2e["calls"]["min_ms"] <= e2e["calls"]["median_ms"] <= e2e["calls"]["max_ms"]
    assert abs(e2e["ms_per_step"] * 1e-3 * e2e["value"] * 1e6 - bench.ITERATIONS * SIDE * SIDE) < 1

    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and "single_thread" in cb and "_pass1" not in cb
    parity = line["parity"]
    assert "error" not in parity, parity
    assert parity["bit_equal_fraction"] == 1.0 and parity["path_divergence_fraction"] == 0.0
    assert "error" not in parity["all_passes"], parity["all_passes"]
    assert parity["all_passes"]["bit_equal"] is True and len(parity["all_passes"]["bands"]) == 3


def test_single_gpu_arm_with_every_pass_walking(cpu_stand_ins, monkeypatch):
    """RLIC_B200_PATHS=recompute (here: the per-thread override): no record, no replay object, no comparison leg."""
    stub = cpu_stand_ins
    args = argparse.Namespace(gpus=1, steps=1, warmup=1, impl="ours", no_cpu_baseline=True)
    with rlic_b200.options(paths="recompute"):
        line = json.loads(json.dumps(bench.run_ours(args)))
    assert all(k == [_core.PASS_WALK] * bench.ITERATIONS for k in stub.pass_kinds) and len(stub.pass_kinds) == 3 + 1
    assert "every_pass_walks" not in line and "replay" not in line["roofline"]
    assert line["cpu_baseline"] is None and "parity" not in line
    assert line["e2e"]["paths"] == "recompute"


def test_reference_arm_line(monkeypatch):
    """--impl reference: whole 5-pass steps of the CPU path on the host cores, the contract's keys
    plus `impl`, and an `e2e` that repeats the line's own value with no copies."""
    monkeypatch.setattr(bench, "N_SIDE", SIDE)
    monkeypatch.delenv("RANK", raising=False)
    args = argparse.Namespace(gpus=1, steps=3, warmup=1, impl="reference", no_cpu_baseline=False)
    line = json.loads(json.dumps(bench.run_reference(args)))
    assert line["impl"] == "reference" and line["metric"] == line["unit"] == "Mpix/s" and line["n_gpus"] == 1
    assert line["steps"] == 3 and line["warmup"] == 1 and line["higher_is_better"] is True and line["gpu_launches"] == 0
    assert line["config"] == json.loads(json.dumps(bench.workload_config(1)))          # the GPU arm's config
    assert line["e2e"] == {"value": line["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == oracle.max_threads() and cb["value"] == line["value"]
    assert "whole convolve calls" in cb["sample"] and cb["steps_timed"] == 3
    assert abs(line["ms_per_step"] * 1e-3 * line["value"] * 1e6 - bench.ITERATIONS * SIDE * SIDE) < 1
    assert line["single_thread"]["value"] > 0
    # the other ranks of a torchrun launch do nothing
    monkeypatch.setenv("RANK", "1")
    assert bench.run_reference(args) == {}


# ---------------------------------------------------------------------------------------------
# The multi-GPU arm (bench.py --gpus N under torchrun), two ranks over gloo: the sharded driver
# is the real one, with the kernel source compiled for the CPU as its compute steps and named
# shared memory for the peer mappings (the stand-ins tests/test_sharded_gloo.py uses).
def _multi_rank_worker(rank, world, port, exchange, queue):
    import os
    import traceback

    try:
        refused = exchange == "peer refused"                 # a box whose GPUs cannot map each other's memory
        exchange = "peer" if refused else exchange
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank),
                          WORLD_SIZE=str(world), RLIC_B200_EXCHANGE=exchange)
        import torch.distributed as dist

        import rlic_b200.sharded as sharded
        import test_sharded_gloo as stand_ins

        bench.N_SIDE = 256                                    # slabs of four kernel half-widths and more
        cpu = torch.device("cpu")
        torch.device = lambda *a, **k: cpu
        torch.Tensor.pin_memory = lambda self: self
        cuda = torch.cuda
        cuda.is_available = lambda: True
        cuda.set_device = lambda d: None
        cuda.current_device = lambda: 0
        cuda.synchronize = lambda *a: None
        cuda.Event, cuda.Stream = _Event, _Stream
        cuda.current_stream = lambda *a: _Stream()
        cuda.stream = lambda s: contextlib.nullcontext()
        cuda.get_device_properties = lambda d: types.SimpleNamespace(multi_processor_count=148)
        bench.ClockSampler.summary = lambda self: {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": [], "samples": 1}

        real_init = dist.init_process_group
        dist.init_process_group = lambda backend, **kw: real_init("gloo", rank=rank, world_size=world)

        class Lib:
            def __getattr__(self, name, real=_core.lib):
                return getattr(real, name)

            def rlic_b200_set_device(self, device):
                return 0

        _core.lib = Lib()

        class Convolver(sharded.ShardedConvolver):
            def __init__(self, *a, **kw):
                ops = stand_ins._make_ops("emulated kernels, grouped walk")
                ops.poison_halos = kw.get("exchange") != "peer"
                peers = None
                if kw.get("exchange") == "peer":
                    peers = (stand_ins._BrokenPeers(rank, world - 1, "open") if refused
                             else stand_ins.SharedMemoryPeers())
                super().__init__(*a, ops=ops, peers=peers, **kw)

        sharded.ShardedConvolver = Convolver
        sharded.pinned_empty = lambda shape, dtype: np.empty(shape, dtype=dtype)

        args = argparse.Namespace(gpus=world, steps=2, warmup=1, impl="ours", no_cpu_baseline=True)
        line = bench.run_ours(args)
        if rank == 0:
            queue.put(("ok", json.dumps(line)))
        else:
            assert line == {}
    except Exception:
        queue.put(("error", f"rank {rank}:\n{traceback.format_exc()}"))
        raise


@pytest.mark.parametrize("exchange", ["peer", "nccl", "peer refused"])
def test_multi_gpu_arm_builds_its_line_over_gloo(exchange):
    import socket

    import torch.multiprocessing as mp

    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    procs = [ctx.Process(target=_multi_rank_worker, args=(r, world, port, exchange, queue)) for r in range(world)]
    for pr in procs:
        pr.start()
    try:
        status, payload = queue.get(timeout=300)
    finally:
        for pr in procs:
            pr.join(timeout=120)
            if pr.is_alive():
                pr.terminate()
    assert status == "ok", payload
    line = json.loads(payload)
    assert line["n_gpus"] == 2 and line["scaling"] == "weak" and line["steps"] == 2 and line["value"] > 0
    assert line["config"]["image"] == [512, 256] and "row slabs over 2 GPUs" in line["config"]["workload"]
    assert len(line["ms_per_step_per_rank"]) == 2 and max(line["ms_per_step_per_rank"]) <= line["ms_per_step"] * 1.0001
    assert len(line["solo_walking_pass_ms_per_rank"]) == 2 and min(line["solo_walking_pass_ms_per_rank"]) > 0
    if exchange == "peer refused":      # every rank fell back to the NCCL exchange, and the line says why
        assert line["e2e"]["exchange"] == "nccl"
        assert line["e2e"]["exchange_fallback"].startswith("peer -> nccl: rank 1: RuntimeError: no peer access")
    else:
        assert line["e2e"]["exchange"] == exchange and "exchange_fallback" not in line["e2e"]
    assert line["e2e"]["value"] > 0 and "host_link" in line["e2e"]
    assert line["e2e"]["h2d_bytes_per_step"] == 2 * (3 * 4 * 256 * 256 + 4 * bench.TAPS)
    assert line["cpu_baseline"] is None and "every_pass_walks" not in line
    # every rank's first and last 64 rows of the five-pass result, against the oracle
    parity = line["parity"]
    assert parity["bit_equal"] is True and len(parity["per_rank"]) == 2
    assert [b["rows"] for pair in parity["per_rank"] for b in pair] == [[0, 64], [192, 256], [256, 320], [448, 512]]
    assert all(b["bit_equal"] and b["mismatches"] == 0 for pair in parity["per_rank"] for b in pair)
    assert "error" not in line["roofline"].get("issue", {})
    assert "pass 1 walks and records" in line["roofline"]["kernel"] and "replay" not in line["roofline"]
