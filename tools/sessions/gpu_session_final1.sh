#!/usr/bin/env bash
# One-GPU check of the late additions: the whole GPU suite (histogram equalisation, small-image
# kernel), the configurations (C1 through the raw ABI too), the launch list of a C1 call.
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/final1
mkdir -p "$OUT"
step() { local limit=$1 name=$2; shift 2; echo "=== $name" | tee -a "$OUT/summary.txt"; local t0=$SECONDS
    timeout "$limit" "$@" >"$OUT/$name.log" 2>&1; echo "    exit $? after $((SECONDS - t0)) s" | tee -a "$OUT/summary.txt"
    tail -n 3 "$OUT/$name.log" | cut -c1-400 | sed 's/^/    | /' | tee -a "$OUT/summary.txt"; }
step 700 pytest_gpu python -m pytest tests -q -m gpu -rxXs
step 120 smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
step 300 configs python tools/bench_configs.py --configs c1,c2,c3
step 200 ncu_c1 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file "$OUT/launches_c1.csv" python tools/bench_configs.py --configs c1
step 240 bench python bench.py --steps 20 --warmup 5
echo "=== done" | tee -a "$OUT/summary.txt"
