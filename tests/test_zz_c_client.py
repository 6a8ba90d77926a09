"""The C ABI used from plain C (examples/convolve_from_c.c): the header is valid C99, a C
program links against librlic_b200.so with nothing else, fails loudly without a device and
reproduces the oracle bit for bit with one."""
from __future__ import annotations

import math
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest


ROOT = Path(__file__).resolve().parents[1]
LIBDIR = ROOT / "rlic_b200"
N, TAPS = 256, 65

pytestmark = pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")


@pytest.fixture(scope="module")
def client(tmp_path_factory) -> Path:
    exe = tmp_path_factory.mktemp("c_client") / "convolve_from_c"
    subprocess.run(
        ["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", f"-I{ROOT / 'include'}",
         str(ROOT / "examples" / "convolve_from_c.c"), "-o", str(exe),
         f"-L{LIBDIR}", "-lrlic_b200", f"-Wl,-rpath,{LIBDIR}", "-lm"],
        check=True, capture_output=True, text=True)
    return exe


def test_header_is_plain_c99_and_cxx(tmp_path):
    for std, lang in (("-std=c99", "c"), ("-std=c++17", "c++")):
        subprocess.run(["gcc", std, "-pedantic", "-Wall", "-Wextra", "-Werror", "-fsyntax-only",
                        "-x", lang, str(ROOT / "include" / "rlic_b200.h")],
                       check=True, capture_output=True, text=True)


def _has_gpu() -> bool:
    import torch

    return torch.cuda.is_available()


def test_c_client_fails_loudly_without_a_device(client):
    if _has_gpu():
        pytest.skip("a GPU is present")
    run = subprocess.run([str(client)], capture_output=True, text=True)
    assert run.returncode == 2, (run.returncode, run.stderr)     # RLIC_B200_ENODEVICE
    assert "rlic_b200_convolve_f64 failed (2)" in run.stderr
    assert run.stdout.splitlines() == [f"inputs {inputs_checksum():016x}"]


def example_inputs():
    """The arrays examples/convolve_from_c.c builds, operation for operation."""
    state, mask = 42, (1 << 64) - 1
    texture = np.empty(N * N)
    for i in range(N * N):
        state = (state * 6364136223846793005 + 1442695040888963407) & mask
        texture[i] = (state >> 11) / 9007199254740992.0
    xs = [math.pi * j / (N - 1) for j in range(N)]
    u = np.tile(np.array([math.cos(2 * x) for x in xs]), (N, 1))
    v = np.tile(np.array([math.sin(x) for x in xs]), (N, 1))
    taps = np.array([1.0 - abs(-1.0 + 2.0 * k / (TAPS - 1)) for k in range(TAPS)])
    return texture.reshape(N, N), u, v, taps


def bit_sum(a: np.ndarray) -> int:
    return int(np.ascontiguousarray(a).view(np.uint64).sum(dtype=np.uint64))


def inputs_checksum() -> int:
    return sum(bit_sum(a) for a in example_inputs()) & ((1 << 64) - 1)


def test_python_regenerates_the_c_inputs_exactly(client):
    first = subprocess.run([str(client)], capture_output=True, text=True).stdout.splitlines()[0]
    assert first == f"inputs {inputs_checksum():016x}"


@pytest.mark.gpu
def test_c_client_matches_the_oracle(client):
    import oracle

    run = subprocess.run([str(client)], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stderr
    words = run.stdout.splitlines()[1].split()
    assert words[0] == "devices" and int(words[1]) >= 1 and int(words[3]) >= 5
    texture, u, v, taps = example_inputs()
    periodic = (("periodic", "periodic"), ("periodic", "periodic"))
    want = oracle.convolve(texture, u, v, kernel=taps, boundaries=periodic, iterations=5)
    assert words[5] == f"{bit_sum(want):016x}"
