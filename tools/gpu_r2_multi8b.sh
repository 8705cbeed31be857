#!/usr/bin/env bash
# Round-2, 8 GPUs, second pass: the default gradient exchange (peer-memory kernels on the tcgen05 path, launched with
# programmatic serialization) at N = 1, 2, 4, 8, and the fixed-wing configuration of BASELINE.json (131072 drones on 4 GPUs)
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/r2_multi8b
mkdir -p "$out"
run() { # name nproc args...
  local name=$1 np=$2; shift 2
  if [ "$np" = 1 ]; then
    timeout 400 python bench.py --gpus 1 "$@" > "$out/$name.json" 2> "$out/$name.err"
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$np --master-addr 127.0.0.1 --master-port 29561 \
      bench.py --gpus $np "$@" > "$out/$name.json" 2> "$out/$name.err"
  fi
}
for np in 8 4 2 1; do
  run bench_quad_concurrent_g${np} $np --steps 50 --warmup 5 --no-cpu-baseline
done
run bench_wing_concurrent_g4 4 --workload wing_concurrent --drones-per-gpu 32768 --steps 20 --warmup 5 --no-cpu-baseline
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29560 \
  tests/multi_gpu_p2p_check.py > "$out/p2p_check_8.log" 2>&1
echo "exit=$?" >> "$out/p2p_check_8.log"
echo done > "$out/finished"
