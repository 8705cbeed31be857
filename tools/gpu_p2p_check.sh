#!/usr/bin/env bash
# First hardware run of the gradient exchange over NVLink peer memory (csrc/p2p_kernels.cu, dist.PeerGradExchange):
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_p2p_check.sh'
# Outputs under gpurun_out/p2p/.  Every step has its own timeout; the kernels' waits are bounded (NaN, not a hang).
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/p2p
mkdir -p "$out"
nvidia-smi topo -m > "$out/topo.txt" 2>&1
# 1. parity against the NCCL path (params / gradient / loss after a few steps, bitwise agreement across ranks)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29541 \
  tests/multi_gpu_p2p_check.py > "$out/check.log" 2>&1
echo "exit=$?" >> "$out/check.log"
# 2. the bench with both exchanges, same box, back to back
for mode in nccl p2p; do
  flag=""; [ "$mode" = p2p ] && flag="--p2p-grad"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29542 \
    bench.py --gpus 2 --steps 100 --warmup 5 $flag > "$out/bench_$mode.json" 2> "$out/bench_$mode.err"
done
echo done > "$out/finished"
