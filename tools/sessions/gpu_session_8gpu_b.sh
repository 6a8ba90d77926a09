#!/usr/bin/env bash
# Second eight-GPU session: bench.py --gpus 8 as the driver runs it (fused peer exchange, lazy
# time-out check, one clock sampler per job), the NCCL exchange beside it, the per-rank phase
# table, and the host-link floor (in the bench lines: e2e.host_link).
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/session8b
mkdir -p "$OUT"
step() { local limit=$1 name=$2; shift 2; echo "=== $name" | tee -a "$OUT/summary.txt"; local t0=$SECONDS
    timeout "$limit" "$@" >"$OUT/$name.log" 2>&1; echo "    exit $? after $((SECONDS - t0)) s" | tee -a "$OUT/summary.txt"
    tail -n 2 "$OUT/$name.log" | cut -c1-300 | sed 's/^/    | /' | tee -a "$OUT/summary.txt"; }
runN() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$1" --master-addr 127.0.0.1 --master-port "$2" "${@:3}"; }
export -f runN
step 150 bench8_peer_a bash -c 'runN 8 29551 bench.py --gpus 8 --steps 20 --warmup 5'
step 150 bench8_nccl env RLIC_B200_EXCHANGE=nccl bash -c 'runN 8 29552 bench.py --gpus 8 --steps 20 --warmup 5'
step 150 bench8_peer_b bash -c 'runN 8 29553 bench.py --gpus 8 --steps 20 --warmup 5'
step 150 diag8_peer bash -c 'runN 8 29554 tools/peer_diag.py --check lazy'
step 150 bench8_peer_c bash -c 'runN 8 29555 bench.py --gpus 8 --steps 20 --warmup 5'
step 150 bench4_peer bash -c 'runN 4 29556 bench.py --gpus 4 --steps 20 --warmup 5'
step 150 bench2_peer bash -c 'runN 2 29557 bench.py --gpus 2 --steps 20 --warmup 5'
grep -h '^{' "$OUT"/bench*.log "$OUT"/diag*.log >"$OUT/lines.jsonl" 2>/dev/null
echo "=== done" | tee -a "$OUT/summary.txt"
