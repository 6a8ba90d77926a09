#!/usr/bin/env bash
# Two-GPU session (round 2): the multi-GPU paths on real hardware, each step with its own
# time limit, output under gpurun_out/session2/ (merged back by gpurun).
#
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash tools/sessions/gpu_session_2gpu.sh'
#
#   1  the two-GPU tests: NCCL halo exchange, fused peer exchange (device and host-slab
#      pipeline), single-process multi-device slabs, batch over every visible device
#   2  bench.py --gpus 2 with the NCCL exchange, then with RLIC_B200_EXCHANGE=peer
#   3  C4 strong scaling on one and two GPUs (tools/bench_c4_scaling.py), both exchanges
set -u
cd "$(dirname "$0")/../.."
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
    tail -n 3 "$OUT/$name.log" | cut -c1-600 | sed 's/^/    | /' | tee -a "$OUT/summary.txt"
}
runN() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$1" --master-addr 127.0.0.1 --master-port "$2" "${@:3}"; }
export -f runN

nvidia-smi topo -m >"$OUT/topo.txt" 2>&1
step 400 pytest_multi_gpu python -m pytest tests/test_slab.py tests/test_multi_device.py -q -m gpu -rxXs
step 200 bench_nccl env RLIC_B200_EXCHANGE=nccl bash -c 'runN 2 29511 bench.py --gpus 2 --steps 10 --warmup 3'
step 200 bench_peer env RLIC_B200_EXCHANGE=peer bash -c 'runN 2 29512 bench.py --gpus 2 --steps 10 --warmup 3'
step 240 c4_n1 bash -c 'runN 1 29513 tools/bench_c4_scaling.py'
step 240 c4_n2_peer env RLIC_B200_EXCHANGE=peer bash -c 'runN 2 29514 tools/bench_c4_scaling.py'
step 240 c4_n2_nccl env RLIC_B200_EXCHANGE=nccl bash -c 'runN 2 29515 tools/bench_c4_scaling.py'
grep -h '^{' "$OUT"/bench_*.log "$OUT"/c4_*.log >"$OUT/lines.jsonl" 2>/dev/null
echo "=== done" | tee -a "$OUT/summary.txt"
