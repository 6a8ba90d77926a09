#!/usr/bin/env bash
# Two-GPU companion of tools/gpu_session.sh: what was written for more than one GPU while none
# was available.
#
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_session_2gpu.sh'
#
#   1  the two-GPU tests: NCCL halo exchange (seen green in round 1) and the fused peer
#      exchange (first run: XPASS = good)
#   2  bench.py --gpus 2 with the NCCL exchange, then with RLIC_B200_EXCHANGE=peer, then both
#      with the grouped walk: four JSON lines to compare (e2e.exchange / e2e.walk say which)
#   3  C4 strong scaling on two GPUs (tools/bench_c4_scaling.py), both exchanges
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/session2
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
run2() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port "$1" "${@:2}"; }
export -f run2

step 600 pytest_two_gpu python -m pytest tests/test_slab.py -q -m gpu -rxX
step 200 bench_nccl bash -c 'run2 29511 bench.py --gpus 2 --steps 10 --warmup 3'
step 200 bench_peer env RLIC_B200_EXCHANGE=peer bash -c 'run2 29512 bench.py --gpus 2 --steps 10 --warmup 3'
step 200 bench_nccl_grouped env RLIC_B200_WALK=grouped bash -c 'run2 29513 bench.py --gpus 2 --steps 10 --warmup 3'
step 200 bench_peer_grouped env RLIC_B200_WALK=grouped RLIC_B200_EXCHANGE=peer bash -c 'run2 29514 bench.py --gpus 2 --steps 10 --warmup 3'
step 300 c4_nccl bash -c 'run2 29515 tools/bench_c4_scaling.py'
step 300 c4_peer env RLIC_B200_EXCHANGE=peer bash -c 'run2 29516 tools/bench_c4_scaling.py'
grep -h '^{' "$OUT"/bench_*.log "$OUT"/c4_*.log >"$OUT/lines.jsonl" 2>/dev/null
echo "=== done" | tee -a "$OUT/summary.txt"
