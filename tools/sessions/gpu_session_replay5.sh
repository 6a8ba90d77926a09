#!/usr/bin/env bash
# One-GPU check of the band plan: schedule tests, the bench line, the host trace.
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/${SESSION_NAME:-replay5}
mkdir -p "$OUT"
step() { local limit=$1 name=$2; shift 2; echo "=== $name" | tee -a "$OUT/summary.txt"; local t0=$SECONDS
    timeout "$limit" "$@" >"$OUT/$name.log" 2>&1; echo "    exit $? after $((SECONDS - t0)) s" | tee -a "$OUT/summary.txt"
    tail -n 3 "$OUT/$name.log" | cut -c1-600 | sed 's/^/    | /' | tee -a "$OUT/summary.txt"; }
step 300 pytest_some python -m pytest tests/test_zz_schedule.py tests/test_parity.py tests/test_path_replay.py -q -m gpu -x
step 240 bench python bench.py --steps 20 --warmup 5 --no-cpu-baseline
step 200 e2e_probe env RLIC_B200_TRACE=1 python tools/e2e_probe.py
echo "=== done" | tee -a "$OUT/summary.txt"
