"""Markers that describe how far a test has been exercised."""
import pytest

# GPU tests written after the round's GPU budget was spent.  They have only ever been
# collected, never run on a B200; until they have, a failure is reported (xfail) without
# failing the suite, and a pass shows up as XPASS.  Remove the marker once they are green
# on hardware.
first_gpu_run = pytest.mark.xfail(
    strict=False, reason="written after the round-1 GPU budget was spent: not yet run on a B200")
