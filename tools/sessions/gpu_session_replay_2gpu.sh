#!/usr/bin/env bash
# Two-GPU session for the recorded-path replay: the two-GPU tests, bench.py --gpus 2 with both
# exchanges, the per-rank phase table, C4 strong scaling at 1 and 2 GPUs, plus -- on one GPU of
# the box -- the bench line, the host path's trace and the ncu capture of the replay kernel.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash tools/sessions/gpu_session_replay_2gpu.sh'
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/${SESSION_NAME:-replay_2gpu}
mkdir -p "$OUT"
step() { local limit=$1 name=$2; shift 2; echo "=== $name" | tee -a "$OUT/summary.txt"; local t0=$SECONDS
    timeout "$limit" "$@" >"$OUT/$name.log" 2>&1; echo "    exit $? after $((SECONDS - t0)) s" | tee -a "$OUT/summary.txt"
    tail -n 3 "$OUT/$name.log" | cut -c1-600 | sed 's/^/    | /' | tee -a "$OUT/summary.txt"; }
runN() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$1" --master-addr 127.0.0.1 --master-port "$2" "${@:3}"; }
export -f runN
step 300 pytest_multi_gpu python -m pytest tests/test_slab.py tests/test_multi_device.py tests/test_path_replay.py -q -m gpu -rxXs
step 200 bench2_peer bash -c 'runN 2 29512 bench.py --gpus 2 --steps 20 --warmup 5'
step 200 bench1 python bench.py --steps 20 --warmup 5
step 200 bench2_nccl env RLIC_B200_EXCHANGE=nccl bash -c 'runN 2 29511 bench.py --gpus 2 --steps 20 --warmup 5'
step 150 diag2_peer bash -c 'runN 2 29554 tools/peer_diag.py --check lazy'
step 200 e2e_probe env RLIC_B200_TRACE=1 python tools/e2e_probe.py
step 240 c4_n1 bash -c 'runN 1 29513 tools/bench_c4_scaling.py'
step 240 c4_n2_peer bash -c 'runN 2 29514 tools/bench_c4_scaling.py'
step 300 ncu_replay ncu --set full --clock-control none --import-source on -k regex:lic_replay_kernel -s 9 -c 1 \
    -o "$OUT/f32_replay_pass" -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline
step 300 ncu_record ncu --set full --clock-control none --import-source on -k regex:lic_pass_kernel -s 2 -c 1 \
    -o "$OUT/f32_record_pass" -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline
step 300 ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file "$OUT/launches.csv" python bench.py --steps 2 --warmup 1 --no-cpu-baseline
grep -h '^{' "$OUT"/bench*.log "$OUT"/diag*.log "$OUT"/c4_*.log >"$OUT/lines.jsonl" 2>/dev/null
echo "=== done" | tee -a "$OUT/summary.txt"
