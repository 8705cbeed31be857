#!/usr/bin/env bash
# Round-2, 2 GPUs: peer-memory gradient exchange check, the 2-GPU tests, bench with both exchanges
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/r2_multi2
mkdir -p "$out"
nvidia-smi topo -m > "$out/topo.txt" 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29541 \
  tests/multi_gpu_p2p_check.py > "$out/check.log" 2>&1
echo "exit=$?" >> "$out/check.log"
APG_TEST_P2P=1 timeout 600 python -m pytest tests -q -m gpu -k "two_gpu or p2p" > "$out/pytest_2gpu.log" 2>&1
echo "exit=$?" >> "$out/pytest_2gpu.log"
for mode in nccl p2p; do
  flag=""; [ "$mode" = p2p ] && flag="--p2p-grad"
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29542 \
    bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline $flag > "$out/bench_$mode.json" 2> "$out/bench_$mode.err"
done
echo done > "$out/finished"
