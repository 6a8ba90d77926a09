#!/usr/bin/env bash
# Eight-GPU session for the recorded-path replay: bench.py --gpus 8 as the driver runs it (twice:
# spread), --gpus 4 and the same box's --gpus 1, C4 strong scaling on 8 GPUs, C5 on 8 GPUs.
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 400 -- 'bash tools/sessions/gpu_session_replay_8gpu.sh'
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/${SESSION_NAME:-replay_8gpu}
mkdir -p "$OUT"
step() { local limit=$1 name=$2; shift 2; echo "=== $name" | tee -a "$OUT/summary.txt"; local t0=$SECONDS
    timeout "$limit" "$@" >"$OUT/$name.log" 2>&1; echo "    exit $? after $((SECONDS - t0)) s" | tee -a "$OUT/summary.txt"
    tail -n 2 "$OUT/$name.log" | cut -c1-500 | sed 's/^/    | /' | tee -a "$OUT/summary.txt"; }
runN() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$1" --master-addr 127.0.0.1 --master-port "$2" "${@:3}"; }
export -f runN
step 100 bench8_peer_a bash -c 'runN 8 29551 bench.py --gpus 8 --steps 20 --warmup 5'
step 60 bench1 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline
step 100 c4_n8 bash -c 'runN 8 29538 tools/bench_c4_scaling.py'
step 100 bench8_peer_b bash -c 'runN 8 29553 bench.py --gpus 8 --steps 20 --warmup 5'
step 100 c5_batch python tools/bench_c5_batch.py --skip-pageable
step 80 bench4_peer bash -c 'runN 4 29556 bench.py --gpus 4 --steps 20 --warmup 5'
step 80 diag8_peer bash -c 'runN 8 29554 tools/peer_diag.py --check lazy'
grep -h '^{' "$OUT"/bench*.log "$OUT"/diag*.log "$OUT"/c4_*.log "$OUT"/c5_*.log >"$OUT/lines.jsonl" 2>/dev/null
echo "=== done" | tee -a "$OUT/summary.txt"
