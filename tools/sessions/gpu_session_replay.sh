#!/usr/bin/env bash
# One-GPU session for the recorded-path replay (PathPlanes in lic_walk.cuh): the whole GPU suite,
# the bench line with and without replay, the lab's record / replay variants, every configuration,
# and the ncu evidence of the two kernels of a step (launch list + one full capture each).
#   /usr/local/graft/bin/gpurun --timeout 1100 -- 'bash tools/sessions/gpu_session_replay.sh'
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/${SESSION_NAME:-replay}
mkdir -p "$OUT"
step() { local limit=$1 name=$2; shift 2; echo "=== $name" | tee -a "$OUT/summary.txt"; local t0=$SECONDS
    timeout "$limit" "$@" >"$OUT/$name.log" 2>&1; echo "    exit $? after $((SECONDS - t0)) s" | tee -a "$OUT/summary.txt"
    tail -n 3 "$OUT/$name.log" | cut -c1-600 | sed 's/^/    | /' | tee -a "$OUT/summary.txt"; }
step 300 pytest_replay python -m pytest tests/test_path_replay.py -q -m gpu -x
step 120 smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
step 240 bench python bench.py --steps 20 --warmup 5
step 120 lab_f32 tools/replay_lab 4096 65
step 120 lab_c5 tools/replay_lab 4096 33 g1
step 700 pytest_gpu python -m pytest tests -q -m gpu -rxXs
step 240 bench_recompute env RLIC_B200_PATHS=recompute python bench.py --steps 20 --warmup 5 --no-cpu-baseline
step 240 bench_unstaged env RLIC_B200_REPLAY_STAGING=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
step 300 configs python tools/bench_configs.py --configs c1,c2,c3,c4
step 300 c5 python tools/bench_c5_batch.py --fields 512 --skip-pageable
step 300 ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv \
    --log-file "$OUT/launches.csv" python bench.py --steps 2 --warmup 1 --no-cpu-baseline
step 300 ncu_record ncu --set full --clock-control none --import-source on -k regex:lic_pass_kernel -s 2 -c 1 \
    -o "$OUT/f32_record_pass" -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline
step 300 ncu_replay ncu --set full --clock-control none --import-source on -k regex:lic_replay_kernel -s 9 -c 1 \
    -o "$OUT/f32_replay_pass" -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline
step 200 sanitizer compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()"
step 200 sanitizer_race compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()"
step 200 sanitizer_init compute-sanitizer --tool initcheck python -c "import __graft_entry__ as g; g.smoke()"
echo "=== done" | tee -a "$OUT/summary.txt"
