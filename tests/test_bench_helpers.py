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


def test_cpu_baseline_sample_is_a_band_of_the_headline_workload():
    info = bench.cpu_baseline(None, 0.5)
    r0, band = info.pop("_band")
    assert info["kind"] == "port" and info["cores"] == oracle.max_threads() and info["unit"] == "Mpix/s"
    assert band.shape[1] == bench.N_SIDE and band.dtype == np.float32 and 0 <= r0 < bench.N_SIDE
    assert f"rows [{r0}, {r0 + band.shape[0]})" in info["sample"]
    json.dumps(info)
    # the band is pass 1 of the workload bench.py times on the GPU
    tex, u, v, kernel = bench.make_slab(0, 1)
    want = oracle.pass_rows(tex, u, v, kernel=kernel, rows=(r0, r0 + 8), threads=oracle.max_threads())
    np.testing.assert_array_equal(band[:8], want)
