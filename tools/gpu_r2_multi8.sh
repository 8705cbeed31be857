#!/usr/bin/env bash
# Round-2, 8 GPUs of one box: scaling of the headline workload at N = 1, 2, 4, 8 with both gradient exchanges
# (NCCL all-reduce / the package's own peer-memory kernels), the peer-memory check on 8 ranks, and the configurations
# BASELINE.json quotes on more than one GPU (fixed wing: 131072 drones on 4 GPUs; LSTM: 262144 drones on 8 GPUs).
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/r2_multi8
mkdir -p "$out"
nvidia-smi topo -m > "$out/topo.txt" 2>&1
run() { # name nproc args...
  local name=$1 np=$2; shift 2
  if [ "$np" = 1 ]; then
    timeout 500 python bench.py --gpus 1 "$@" > "$out/$name.json" 2> "$out/$name.err"
  else
    timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$np --master-addr 127.0.0.1 --master-port 29551 \
      bench.py --gpus $np "$@" > "$out/$name.json" 2> "$out/$name.err"
  fi
}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29550 \
  tests/multi_gpu_p2p_check.py > "$out/p2p_check_8.log" 2>&1
echo "exit=$?" >> "$out/p2p_check_8.log"
for np in 1 2 4 8; do
  run bench_quad_concurrent_g${np}_nccl $np --steps 50 --warmup 5 --no-cpu-baseline
  [ "$np" = 1 ] || run bench_quad_concurrent_g${np}_p2p $np --steps 50 --warmup 5 --no-cpu-baseline --p2p-grad
done
run bench_wing_concurrent_g4 4 --workload wing_concurrent --drones-per-gpu 32768 --steps 20 --warmup 5 --no-cpu-baseline
run bench_wing_concurrent_g4_weak 4 --workload wing_concurrent --steps 20 --warmup 5 --no-cpu-baseline
run bench_quad_lstm_g8 8 --workload quad_lstm --steps 20 --warmup 5 --no-cpu-baseline
run bench_quad_autoregressive_g8 8 --workload quad_autoregressive --steps 20 --warmup 5 --no-cpu-baseline
APG_TEST_P2P=1 timeout 600 python -m pytest tests -q -m gpu -k "two_gpu or p2p or multi" > "$out/pytest_multi.log" 2>&1
echo "exit=$?" >> "$out/pytest_multi.log"
echo done > "$out/finished"
