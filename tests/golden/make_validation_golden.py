"""Regenerates tests/golden/validation_golden.json by running the REFERENCE's own
Python layer (pure Python: /root/reference/src/rlic/_lib.py and _boundaries.py)
on a catalogue of valid and invalid calls, with its native module replaced by a
stub (the Rust core cannot be built here and is not needed for validation).

For every case the file records what the reference does: the exception type(s)
and exact message(s), in order, or "reached the native call" for accepted input.
tests/test_validation_golden.py replays the catalogue against rlic_b200.

    python tests/golden/make_validation_golden.py      (needs /root/reference)
"""

from __future__ import annotations

import json
import sys
import types
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))

from validation_cases import BOUNDARY_SPECS, CASES, build_args  # noqa: E402


def load_reference():
    core = types.ModuleType("rlic._core")
    core.convolve_f32 = lambda *a, **k: "NATIVE_CALL_f32"
    core.convolve_f64 = lambda *a, **k: "NATIVE_CALL_f64"
    sys.path.insert(0, "/root/reference/src")
    sys.modules["rlic._core"] = core
    import rlic  # noqa: F401  (the reference package, with the stubbed native module)
    from rlic._boundaries import BoundarySet
    from rlic._lib import convolve

    return convolve, BoundarySet


def describe(exc: BaseException):
    if isinstance(exc, BaseExceptionGroup):
        return {"group": str(exc.message), "members": [describe(e) for e in exc.exceptions]}
    return {"type": type(exc).__name__, "message": str(exc)}


def main() -> None:
    convolve, BoundarySet = load_reference()
    out = {"convolve": {}, "boundary_sets": {}}
    for name, spec in CASES.items():
        args, kwargs = build_args(spec)
        try:
            res = convolve(*args, **kwargs)
        except BaseException as exc:  # noqa: BLE001
            out["convolve"][name] = {"raises": describe(exc)}
        else:
            if isinstance(res, str):
                out["convolve"][name] = {"returns": res}
            else:
                out["convolve"][name] = {"returns": "COPY_OF_TEXTURE", "equal": bool(np.array_equal(res, args[0])),
                                         "is_input": res is args[0]}
    # boundary specs on their own
    specs = BOUNDARY_SPECS
    for i, spec in enumerate(specs):
        bs = BoundarySet.from_spec(spec)
        if bs is None:
            out["boundary_sets"][str(i)] = {"spec": repr(spec), "parsed": None}
        else:
            out["boundary_sets"][str(i)] = {
                "spec": repr(spec), "parsed": [list(bs.x), list(bs.y)],
                "problems": [describe(e) for e in bs.collect_exceptions()],
            }
    (HERE / "validation_golden.json").write_text(json.dumps(out, indent=1, sort_keys=True) + "\n")
    print("wrote", len(out["convolve"]), "convolve cases and", len(out["boundary_sets"]), "boundary specs")


if __name__ == "__main__":
    main()
