#!/usr/bin/env bash
# One-GPU check after the host-path changes (tapered row bands, conversions beside the copies):
# the whole GPU suite, the bench line, the host path's trace, every configuration.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/sessions/gpu_session_replay3.sh'
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/${SESSION_NAME:-replay3}
mkdir -p "$OUT"
step() { local limit=$1 name=$2; shift 2; echo "=== $name" | tee -a "$OUT/summary.txt"; local t0=$SECONDS
    timeout "$limit" "$@" >"$OUT/$name.log" 2>&1; echo "    exit $? after $((SECONDS - t0)) s" | tee -a "$OUT/summary.txt"
    tail -n 3 "$OUT/$name.log" | cut -c1-600 | sed 's/^/    | /' | tee -a "$OUT/summary.txt"; }
step 700 pytest_gpu python -m pytest tests -q -m gpu -rxXs
step 120 smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
step 240 bench python bench.py --steps 20 --warmup 5
step 200 e2e_probe env RLIC_B200_TRACE=1 python tools/e2e_probe.py
step 300 configs python tools/bench_configs.py --configs c1,c2,c3,c4
step 300 c5 python tools/bench_c5_batch.py --fields 512
step 240 bench_reference python bench.py --impl reference --steps 3 --warmup 1
step 300 ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file "$OUT/launches.csv" python bench.py --steps 2 --warmup 1 --no-cpu-baseline
step 200 sanitizer_init compute-sanitizer --tool initcheck python -c "import __graft_entry__ as g; g.smoke()"
step 200 sanitizer compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()"
echo "=== done" | tee -a "$OUT/summary.txt"
