#!/usr/bin/env bash
# Eight-GPU session (round 2): the scaling runs the north star names, each step with its own
# time limit, output under gpurun_out/session8/ (merged back by gpurun).  An 8-GPU call is
# charged 8x: everything here is sized to finish in about five minutes.
#
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 700 -- 'bash tools/sessions/gpu_session_8gpu.sh'
#
#   1  bench.py --gpus 8 three times with the fused peer exchange (spread), once with NCCL,
#      and --gpus 4 with the peer exchange
#   2  C4 strong scaling at 1, 2, 4 and 8 GPUs (tools/bench_c4_scaling.py, exact checksums +
#      band parity in every line)
#   3  C5: 4096 fields over 8 GPUs (tools/bench_c5_batch.py)
#   4  the multi-GPU tests on this box
set -u
cd "$(dirname "$0")/../.."
OUT=gpurun_out/session8
mkdir -p "$OUT"
step() {   # step <seconds> <name> <command...>
    local limit=$1 name=$2
    shift 2
    echo "=== $name" | tee -a "$OUT/summary.txt"
    local t0=$SECONDS
    timeout "$limit" "$@" >"$OUT/$name.log" 2>&1
    local rc=$?
    echo "    exit $rc after $((SECONDS - t0)) s" | tee -a "$OUT/summary.txt"
    tail -n 2 "$OUT/$name.log" | cut -c1-400 | sed 's/^/    | /' | tee -a "$OUT/summary.txt"
}
runN() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$1" --master-addr 127.0.0.1 --master-port "$2" "${@:3}"; }
export -f runN
: "${EXCHANGE:=peer}"

nvidia-smi topo -m >"$OUT/topo.txt" 2>&1
nproc >"$OUT/host.txt"; free -g >>"$OUT/host.txt"
step 120 bench8_a env RLIC_B200_EXCHANGE=$EXCHANGE bash -c 'runN 8 29521 bench.py --gpus 8 --steps 20 --warmup 5'
step 120 bench8_b env RLIC_B200_EXCHANGE=$EXCHANGE bash -c 'runN 8 29522 bench.py --gpus 8 --steps 20 --warmup 5'
step 120 bench8_c env RLIC_B200_EXCHANGE=$EXCHANGE bash -c 'runN 8 29523 bench.py --gpus 8 --steps 20 --warmup 5'
step 120 bench8_nccl env RLIC_B200_EXCHANGE=nccl bash -c 'runN 8 29524 bench.py --gpus 8 --steps 20 --warmup 5'
step 120 bench4 env RLIC_B200_EXCHANGE=$EXCHANGE bash -c 'runN 4 29525 bench.py --gpus 4 --steps 20 --warmup 5'
step 120 bench1 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline
for n in 8 4 2 1; do
    step 150 c4_n$n env RLIC_B200_EXCHANGE=$EXCHANGE bash -c "runN $n 2953$n tools/bench_c4_scaling.py"
done
step 300 c5_batch python tools/bench_c5_batch.py
step 240 pytest_multi python -m pytest tests/test_slab.py tests/test_parity_full_size.py -q -m gpu -k "two_gpu or every_visible or c5_shape" -rs
grep -h '^{' "$OUT"/bench*.log "$OUT"/c4_*.log "$OUT"/c5_*.log >"$OUT/lines.jsonl" 2>/dev/null
echo "=== done" | tee -a "$OUT/summary.txt"
