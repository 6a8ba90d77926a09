"""Runs the reference's OWN, unmodified test files against this package imported under the name
`rlic` -- the drop-in claim, checked by the reference's tests rather than by ours.

    python tools/run_reference_tests.py [--reference /root/reference] [--backend native|oracle|emulation] [pytest args]

The files are read where they lie (nothing of the reference is copied); its conftest.py is not
loaded (it imports `runtime_introspect`, a reporting-only dependency that is not installed here).
tests/test_regressions.py compares with vectorplot, which is not installable here: unless the real
package is importable, a stand-in `vectorplot.lic_internal.line_integral_convolution` is put in its
place -- the exact rational-arithmetic streamline tracer of tests/regression_cases.py, which sums
the way vectorplot's loop does -- so that the reference's file still runs as written, at its own
tolerances (rtol 1.5e-7, atol 1e-6); the output says which of the two was used.

What computes behind `rlic.convolve` (everything in front of it -- signature, validation, error
messages, boundary handling, dtype dispatch -- is rlic_b200's own Python layer in all cases):

  native     librlic_b200.so: needs a B200.  Without a GPU only test_exceptions.py and
             test_boundaries.py are run (every one of their cases is decided before the native
             call), which is what `tests/test_reference_suite.py` does on the CPU.
  oracle     the C restatement in oracle/ (a checker: pins the ORACLE to the reference's tests).
  emulation  the CUDA kernel source itself, compiled for the CPU (tests/kernel_emulation): the
             walk, the recording walk and the replay kernel of lic_walk.cuh under the
             reference's property tests.  Test infrastructure, never a product path.

/root/reference exists in the development container only; on the GPU box this script reports that
and exits 0.
"""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--backend", choices=["native", "oracle", "emulation"], default="native")
    args, rest = ap.parse_known_args()
    tests = Path(args.reference) / "tests"
    if not tests.is_dir():
        print(f"no reference tests at {tests}: nothing to run")
        return 0

    import numpy as np
    import pytest

    import rlic_b200
    import rlic_b200._boundaries
    import rlic_b200._lib
    import rlic_b200._typing
    from rlic_b200 import _core

    files = ["test_exceptions.py", "test_boundaries.py", "test_convolution.py"]
    if args.backend == "native":
        if _core.device_count() < 1:
            files.remove("test_convolution.py")
            print("no CUDA device: running the files whose cases are decided before the native call")
    else:
        if args.backend == "oracle":
            import oracle

            def compute(texture, u, v, kernel, mode, boundaries, iterations):
                return oracle.convolve(texture, u, v, kernel=kernel, uv_mode=mode, boundaries=boundaries,
                                       iterations=iterations, threads=oracle.max_threads())
        else:
            import kernel_emulation

            def compute(texture, u, v, kernel, mode, boundaries, iterations):
                # what the library does: the grouped walk, and recorded paths when there is a second pass
                return kernel_emulation.convolve(texture, u, v, kernel=kernel, uv_mode=mode, boundaries=boundaries,
                                                 iterations=iterations, walk=1, paths=iterations >= 2)

        def stand_in(texture, uv, kernel, boundaries, iterations=1, *, check_texture=False):
            u, v, mode = uv
            if check_texture and np.any(texture < 0):
                raise ValueError(_core.NEGATIVE_TEXTURE_MESSAGE)
            c = np.ascontiguousarray
            return compute(c(texture), c(u), c(v), c(kernel), mode, boundaries, int(iterations))

        _core.convolve_f32 = _core.convolve_f64 = stand_in

    if args.backend != "native" or _core.device_count() >= 1:
        files.append("test_regressions.py")
        try:
            import vectorplot.lic_internal  # noqa: F401
            print("test_regressions.py: against the real vectorplot")
        except ImportError:
            import types

            import regression_cases

            def line_integral_convolution(u, v, texture, kernel, polarization):
                return regression_cases.exact_streamline_sum(texture, u, v, kernel,
                                                             "polarization" if polarization else "velocity")

            package, module = types.ModuleType("vectorplot"), types.ModuleType("vectorplot.lic_internal")
            module.line_integral_convolution = line_integral_convolution
            package.lic_internal = module
            sys.modules["vectorplot"], sys.modules["vectorplot.lic_internal"] = package, module
            print("test_regressions.py: vectorplot is not installed; the exact-arithmetic tracer of "
                  "tests/regression_cases.py stands in for its line_integral_convolution")

    sys.modules["rlic"] = rlic_b200
    for sub in ("_boundaries", "_lib", "_typing"):
        sys.modules[f"rlic.{sub}"] = getattr(rlic_b200, sub)
    print(f"backend: {args.backend}; files: {', '.join(files)}")
    return int(pytest.main(["--noconftest", "-p", "no:cacheprovider", "-q", f"--rootdir={ROOT / 'tools'}",
                            *[str(tests / f) for f in files], *rest]))


if __name__ == "__main__":
    sys.exit(main())
