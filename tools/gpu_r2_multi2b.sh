#!/usr/bin/env bash
# Round-2 second session, 2 GPUs: peer-exchange check, the 2-GPU tests, the bench line exactly as the driver launches it
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/r2_multi2b
mkdir -p "$out"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29541 \
  tests/multi_gpu_p2p_check.py > "$out/check.log" 2>&1
echo "exit=$?" >> "$out/check.log"
APG_TEST_P2P=1 timeout 600 python -m pytest tests -q -m gpu -k "two_gpu or p2p" > "$out/pytest_2gpu.log" 2>&1
echo "exit=$?" >> "$out/pytest_2gpu.log"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29542 \
  bench.py --gpus 2 --steps 50 --warmup 5 > "$out/bench_g2.json" 2> "$out/bench_g2.err"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29543 \
  bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > "$out/bench_ref_g2.json" 2> "$out/bench_ref_g2.err"
echo done > "$out/finished"
