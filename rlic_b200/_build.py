"""Builds ``librlic_b200.so`` in-tree with nvcc for sm_100a.

    python -m rlic_b200._build [--force]

The flags matter for parity: ``-fmad=false`` forbids contraction outside the
explicit fused operations, and no fast-math option is ever passed (IEEE
division, denormals kept).
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "librlic_b200.so"
INCLUDE = PKG_DIR.parent / "include"

NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",
    "-Xcompiler",
    "-fPIC",
    "-shared",
    "-cudart",
    "static",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (set NVCC or add /usr/local/cuda/bin to PATH)")


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def is_stale() -> bool:
    if not LIB_PATH.exists():
        return True
    built = LIB_PATH.stat().st_mtime
    deps = list(CSRC.glob("*")) + list(INCLUDE.glob("*.h")) + [Path(__file__)]
    return any(d.stat().st_mtime > built for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not is_stale():
        return LIB_PATH
    cmd = [find_nvcc(), *NVCC_FLAGS, f"-I{INCLUDE}", "-o", str(LIB_PATH), *map(str, sources())]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), file=sys.stderr)
    # gcc wrappers in some images need a plain host compiler
    env = dict(os.environ)
    proc = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if verbose or proc.returncode:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode:
        raise RuntimeError(f"nvcc failed with exit code {proc.returncode}")
    return LIB_PATH


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(LIB_PATH)
