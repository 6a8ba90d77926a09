"""bench.py's reporting helpers, exercised without a GPU (the timed path itself needs one):
the JSON they produce must be serialisable and mean what DESIGN.md section 6 says."""
from __future__ import annotations

import json
import sys
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import oracle  # noqa: E402


def test_parity_of_sample_counts_bits_and_scales_errors_by_the_range():
    rng = np.random.default_rng(0)
    band = rng.random((64, 4096), dtype=np.float32)
    image = np.zeros((4096, 4096), dtype=np.float32)
    image[100:164] = band
    p = bench.parity_of_sample(image, 100, band)
    assert p["bit_equal_fraction"] == 1.0 and p["max_abs_err_over_range"] == 0.0
    assert p["rows"] == [100, 164] and p["pixels"] == band.size
    image[100, 0] = np.nextafter(image[100, 0], np.float32(2))
    image[101, 5] += 0.5
    p = bench.parity_of_sample(image, 100, band)
    assert p["bit_equal_fraction"] == 1.0 - 2 / band.size
    assert abs(p["max_abs_err_over_range"] - 0.5 / float(band.max() - band.min())) < 1e-6
    json.dumps(p)


def test_path_divergence_is_zero_for_the_oracle_itself_and_sees_the_other_build():
    u, v = bench.make_slab(0, 1)[1:3]

    def stub(variant):
        def convolve(texture, u, v, *, kernel, boundaries, iterations):
            assert boundaries == "closed" and iterations == 1
            return oracle.convolve(texture, u, v, kernel=kernel, variant=variant, threads=oracle.max_threads())
        return types.SimpleNamespace(convolve=convolve)

    same = bench.path_divergence(stub(oracle.VARIANT_DEFAULT), u, v, 2000, 96)
    assert same["path_divergence_fraction"] == 0.0
    other = bench.path_divergence(stub(oracle.VARIANT_FMA), u, v, 2000, 96)
    assert 0.0 < other["path_divergence_fraction"] < 0.01
    json.dumps(other)


def test_cpu_baseline_times_whole_passes_of_the_headline_workload():
    info = bench.cpu_baseline(None, reps=1)
    pass1 = info.pop("_pass1")
    tex, u, v, kernel = info.pop("_inputs")
    assert info["kind"] == "port" and info["cores"] == oracle.max_threads() and info["unit"] == "Mpix/s"
    assert pass1.shape == (bench.N_SIDE, bench.N_SIDE) and pass1.dtype == np.float32
    assert "whole passes" in info["sample"] and info["spread"][0] <= info["seconds"] <= info["spread"][1]
    assert abs(info["value"] - bench.N_SIDE ** 2 / info["seconds"] / 1e6) < 1e-6
    json.dumps(info)
    # what was timed is pass 1 of the workload bench.py times on the GPU
    want = oracle.pass_rows(tex, u, v, kernel=kernel, rows=(1000, 1008), threads=oracle.max_threads())
    np.testing.assert_array_equal(pass1[1000:1008], want)
    single = bench.cpu_single_thread(tex, u, v, kernel, reps=1)
    assert single["cores"] == 1 and single["value"] > 0
    json.dumps(single)


def test_slab_parity_band_checks_slab_edges_against_a_whole_image_run(monkeypatch):
    """bench.py --gpus N: every rank compares the first and last rows of its slab of the final
    result with the oracle run on a sub-image wide enough that the artificial cuts cannot
    reach them.  Here the 'ranks' results are cut from a whole-image oracle run at a reduced
    size: every band agrees, and a perturbed one is reported."""
    monkeypatch.setattr(bench, "N_SIDE", 384)
    world = 3
    slabs = [bench.make_slab(r, world) for r in range(world)]
    tex, u, v = (np.concatenate([s[k] for s in slabs]) for k in range(3))
    full = oracle.convolve(tex, u, v, kernel=slabs[0][3], iterations=bench.ITERATIONS, threads=oracle.max_threads())
    for rank in range(world):
        for first in (0, bench.N_SIDE - 64):
            rows = full[rank * bench.N_SIDE + first:rank * bench.N_SIDE + first + 64]
            got = bench.slab_parity_band(rank, world, rows, first)
            assert got["bit_equal"] and got["mismatches"] == 0
            assert got["rows"] == [rank * bench.N_SIDE + first, rank * bench.N_SIDE + first + 64]
            json.dumps(got)
    wrong = full[bench.N_SIDE:bench.N_SIDE + 64].copy()
    wrong[3, 7] = np.nextafter(wrong[3, 7], np.float32(9))
    got = bench.slab_parity_band(1, world, wrong, 0)
    assert not got["bit_equal"] and got["mismatches"] == 1


def test_clock_sampler_summary_keeps_the_timed_region_and_every_gpu_of_the_job():
    """One nvidia-smi poller watches all GPUs of a job: the summary is the median over the samples
    that arrived inside the timed region, throttle reasons are collected, and with several GPUs
    each one's own median is listed (a slow GPU must not hide in the job's median)."""
    s = bench.ClockSampler("0,1")
    s.t0, s.t1 = 10.0, 11.0
    ok = "Not Active, Not Active, Not Active, Not Active"
    s.lines = [
        (9.0, f"0, 1000, 1965, 150.0, {ok}"),                       # warm-up: ignored
        (10.1, f"0, 1965, 1965, 300.5, {ok}"),
        (10.1, f"1, 1500, 1965, 280.0, Not Active, Not Active, Not Active, Active"),
        (10.6, f"0, 1965, 1965, 301.0, {ok}"),
        (10.6, f"1, 1510, 1965, 281.0, {ok}"),
        (10.7, "garbage"),
        (12.0, f"0, 500, 1965, 90.0, {ok}"),                        # after the region: ignored
    ]
    got = s.summary()
    assert got["samples"] == 4 and got["sampled_during"] == "timed region"
    assert got["sm_max_mhz"] == 1965.0 and got["power_w_max"] == 301.0
    assert got["reasons"] == ["sw_power_cap"]
    assert got["sm_mhz_per_gpu"] == {"0": 1965.0, "1": 1505.0}
    json.dumps(got)
    one = bench.ClockSampler("0")
    one.t0, one.t1 = 10.0, 11.0
    one.lines = [(10.2, f"0, 1965, 1965, 300.0, {ok}")]
    assert "sm_mhz_per_gpu" not in one.summary() and one.summary()["sm_mhz"] == 1965.0
    assert bench.ClockSampler(None).summary()["samples"] == 0
