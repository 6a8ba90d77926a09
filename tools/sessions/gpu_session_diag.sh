#!/usr/bin/env bash
# Diagnostics for the peer exchange at N GPUs (default 4) + the lab's new candidates on GPU 0.
set -u
cd "$(dirname "$0")/../.."
N=${1:-4}
OUT=gpurun_out/diag$N
mkdir -p "$OUT"
step() { local limit=$1 name=$2; shift 2; echo "=== $name" | tee -a "$OUT/summary.txt"; local t0=$SECONDS
    timeout "$limit" "$@" >"$OUT/$name.log" 2>&1; echo "    exit $? after $((SECONDS - t0)) s" | tee -a "$OUT/summary.txt"
    tail -n 2 "$OUT/$name.log" | cut -c1-300 | sed 's/^/    | /' | tee -a "$OUT/summary.txt"; }
runN() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$1" --master-addr 127.0.0.1 --master-port "$2" "${@:3}"; }
export -f runN
export N
step 120 diag_sync bash -c 'runN $N 29541 tools/peer_diag.py --check sync'
step 120 diag_lazy bash -c 'runN $N 29542 tools/peer_diag.py --check lazy'
step 120 diag_off bash -c 'runN $N 29543 tools/peer_diag.py --check off'
step 120 diag_nccl env RLIC_B200_EXCHANGE=nccl bash -c 'runN $N 29544 tools/peer_diag.py'
step 150 bench_peer env RLIC_B200_EXCHANGE=peer bash -c 'runN $N 29545 bench.py --gpus $N --steps 20 --warmup 5'
step 150 bench_nccl env RLIC_B200_EXCHANGE=nccl bash -c 'runN $N 29546 bench.py --gpus $N --steps 20 --warmup 5'
step 200 lab_packed env CUDA_VISIBLE_DEVICES=0 tools/kernel_lab 4096 65 packed
step 200 lab_staged env CUDA_VISIBLE_DEVICES=0 tools/kernel_lab 4096 65 staged
step 100 lab_shipped env CUDA_VISIBLE_DEVICES=0 tools/kernel_lab 4096 65 shipped
step 300 pytest_full_size env CUDA_VISIBLE_DEVICES=0 python -m pytest tests/test_parity_full_size.py -q -m gpu
grep -h '^{' "$OUT"/diag_*.log "$OUT"/bench_*.log >"$OUT/lines.jsonl" 2>/dev/null
echo "=== done" | tee -a "$OUT/summary.txt"
