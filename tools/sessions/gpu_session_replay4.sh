#!/usr/bin/env bash
# One-GPU check of the tightened wavefront order: the GPU suite, the bench line, the host trace.
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/${SESSION_NAME:-replay4}
mkdir -p "$OUT"
step() { local limit=$1 name=$2; shift 2; echo "=== $name" | tee -a "$OUT/summary.txt"; local t0=$SECONDS
    timeout "$limit" "$@" >"$OUT/$name.log" 2>&1; echo "    exit $? after $((SECONDS - t0)) s" | tee -a "$OUT/summary.txt"
    tail -n 3 "$OUT/$name.log" | cut -c1-600 | sed 's/^/    | /' | tee -a "$OUT/summary.txt"; }
step 700 pytest_gpu python -m pytest tests -q -m gpu -rxXs
step 240 bench python bench.py --steps 20 --warmup 5
step 200 e2e_probe env RLIC_B200_TRACE=1 python tools/e2e_probe.py
step 300 configs python tools/bench_configs.py --configs c1,c2,c3,c4
echo "=== done" | tee -a "$OUT/summary.txt"
