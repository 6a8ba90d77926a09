#!/usr/bin/env bash
# One-GPU session: the whole GPU suite, the bench line, the lab's latest candidates, and the ncu
# evidence of the kernels that ship (launch list + one full capture each of the f32 velocity,
# f64 velocity (C1) and f64 polarization (C3) pass kernels).  Output under gpurun_out/ncu/.
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash tools/sessions/gpu_session_ncu.sh'
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/ncu
mkdir -p "$OUT"
step() { local limit=$1 name=$2; shift 2; echo "=== $name" | tee -a "$OUT/summary.txt"; local t0=$SECONDS
    timeout "$limit" "$@" >"$OUT/$name.log" 2>&1; echo "    exit $? after $((SECONDS - t0)) s" | tee -a "$OUT/summary.txt"
    tail -n 2 "$OUT/$name.log" | cut -c1-300 | sed 's/^/    | /' | tee -a "$OUT/summary.txt"; }
step 600 pytest_gpu python -m pytest tests -q -m gpu -rxXs
step 120 smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
step 240 bench python bench.py --steps 20 --warmup 5
step 240 bench_reference python bench.py --impl reference --steps 5 --warmup 1
step 200 lab_packed tools/kernel_lab 4096 65 packed
step 200 lab_grouped tools/kernel_lab 4096 65 "grouped tuned"
step 300 configs python tools/bench_configs.py --configs c1,c2,c3,c4
step 300 ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv \
    --log-file "$OUT/launches.csv" python bench.py --steps 2 --warmup 1 --no-cpu-baseline
step 400 ncu_f32 ncu --set full --clock-control none --import-source on -k regex:lic_pass_kernel -s 3 -c 1 \
    -o "$OUT/f32_pass" -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline
step 400 ncu_c1 ncu --set full --clock-control none --import-source on -k regex:lic_pass_kernel -s 2 -c 1 \
    -o "$OUT/f64_vel_c1_pass" -f python tools/bench_configs.py --configs c1
step 400 ncu_c3 ncu --set full --clock-control none --import-source on -k regex:lic_pass_kernel -s 2 -c 1 \
    -o "$OUT/f64_pol_c3_pass" -f python tools/bench_configs.py --configs c3
step 300 sanitizer compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()"
step 300 sanitizer_init compute-sanitizer --tool initcheck python -c "import __graft_entry__ as g; g.smoke()"
step 300 sanitizer_race compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()"
echo "=== done" | tee -a "$OUT/summary.txt"
