#!/usr/bin/env bash
# First GPU call of a session: everything that was written or changed while no GPU was
# available gets run once, with its own time limit per step, and leaves its output under
# gpurun_out/ (merged back by gpurun).  Nothing here changes clocks or kills by pattern.
#
#   python __graft_entry__.py && sh tools/build_lab.sh      # here, without a GPU: the built files travel
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/sessions/gpu_session.sh'
#
# Steps (each independent; a failure is logged and the script goes on):
#   1  the -m gpu suite
#   2  smoke() of __graft_entry__
#   3  bench.py as the driver runs it, then with the grouped walk, the wavefront schedule and
#      the fma-only arithmetic (the JSON line's e2e.walk / e2e.schedule / e2e.arithmetic say
#      which), and the kernel lab's sweep of the grouped-walk formulations (tools/kernel_lab)
#   4  every BASELINE configuration (tools/bench_configs.py), default and wavefront
#   5  ncu launch list of one bench step, and a full capture of the f64 pass kernel
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/session
mkdir -p "$OUT"
step() {   # step <seconds> <name> <command...>
    local limit=$1 name=$2
    shift 2
    echo "=== $name" | tee -a "$OUT/summary.txt"
    local t0=$SECONDS
    timeout "$limit" "$@" >"$OUT/$name.log" 2>&1
    local rc=$?
    echo "    exit $rc after $((SECONDS - t0)) s" | tee -a "$OUT/summary.txt"
    tail -n 3 "$OUT/$name.log" | sed 's/^/    | /' | tee -a "$OUT/summary.txt"
}

nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv >"$OUT/gpu.csv" 2>&1

step 900 pytest_gpu python -m pytest tests -q -m gpu -rxX
step 120 smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
step 240 bench_default python bench.py --steps 10 --warmup 3
step 240 bench_grouped env RLIC_B200_WALK=grouped python bench.py --steps 10 --warmup 3 --no-cpu-baseline
[ -x tools/kernel_lab ] || step 200 build_lab sh tools/build_lab.sh     # normally built before the call
if [ -x tools/kernel_lab ]; then
    step 300 lab_grouped_f32_f64 tools/kernel_lab 4096 65 grouped
    step 120 lab_shipped tools/kernel_lab 4096 65 shipped
    step 120 lab_gather_ceiling tools/kernel_lab 4096 65 ceiling
    step 300 lab_grouped_c3_field tools/kernel_lab 4096 65 grouped 1
fi
step 240 bench_wavefront env RLIC_B200_SCHEDULE=wavefront python bench.py --steps 10 --warmup 3 --no-cpu-baseline
step 240 bench_fma env RLIC_B200_ARITHMETIC=fma python bench.py --steps 10 --warmup 3 --no-cpu-baseline
step 300 configs_default python tools/bench_configs.py --configs c1,c2,c3,c4
step 300 configs_grouped env RLIC_B200_WALK=grouped python tools/bench_configs.py --configs c1,c2,c3,c4
step 300 configs_wavefront env RLIC_B200_SCHEDULE=wavefront python tools/bench_configs.py --configs c2,c4
if command -v ncu >/dev/null; then
    step 300 ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
        --log-file "$OUT/launches.csv" python bench.py --steps 2 --warmup 1 --no-cpu-baseline
    step 400 ncu_f64 ncu --set full --clock-control none --import-source on -k regex:lic_pass_kernel -s 2 -c 1 \
        -o "$OUT/f64_pass" -f python tools/bench_configs.py --configs c3
fi
grep -h '^{' "$OUT"/bench_*.log >"$OUT/bench_lines.jsonl" 2>/dev/null
echo "=== done" | tee -a "$OUT/summary.txt"
