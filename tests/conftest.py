"""Test harness: marker registration and one-time builds.

* ``-m "not gpu"`` runs everywhere (oracle vs known answers, host logic, ABI
  surface, gloo sharding logic);
* ``-m gpu`` needs a B200: parity of the CUDA path against the oracle, through
  the public API and the C ABI.
"""

from __future__ import annotations

import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config: pytest.Config) -> None:
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # build artefacts are git-ignored; make sure both exist before collection
    from rlic_b200 import _build

    _build.build()
    import oracle

    oracle.build()


def pytest_report_header(config: pytest.Config) -> list[str]:
    from rlic_b200 import _core

    return [f"rlic_b200: {_core._LIB_PATH.name}, CUDA devices visible: {_core.device_count()}"]
