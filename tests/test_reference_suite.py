"""The reference's own test files, unmodified and read where they lie, against this package
imported under the name ``rlic`` (tools/run_reference_tests.py).

Runs in the development container only: /root/reference does not travel to the GPU box, and
nothing of it is copied into this repository.  Three backends behind ``rlic.convolve``:
``native`` (without a GPU: the files whose cases are decided before the native call), the C
``oracle`` (pins the oracle to the reference's property tests as the reference states them), and
``emulation`` (the CUDA kernel source compiled for the CPU: grouped walk, recorded paths)."""
from __future__ import annotations

import re
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REFERENCE = Path("/root/reference")

pytestmark = pytest.mark.skipif(not (REFERENCE / "tests" / "test_convolution.py").exists(),
                                reason="the reference tree is not present on this machine")


@pytest.mark.parametrize("backend, expected", [("native", 38), ("oracle", 63), ("emulation", 63)])
def test_the_references_own_tests_pass_against_this_package(backend, expected):
    from rlic_b200 import _core

    if backend == "native" and _core.device_count() >= 1:
        expected = 63                     # with a GPU the convolution and regression tests run on it as well
    run = subprocess.run([sys.executable, str(ROOT / "tools" / "run_reference_tests.py"), "--backend", backend],
                         capture_output=True, text=True, timeout=600)
    tail = run.stdout[-2000:] + run.stderr[-2000:]
    assert run.returncode == 0, tail
    counted = re.search(r"(\d+) passed", run.stdout)
    assert counted and int(counted.group(1)) == expected and "failed" not in run.stdout, tail
