#!/usr/bin/env bash
# Runs every variant of tcgen05_gemm in its own process (a rejected descriptor poisons the CUDA context)
# and writes one JSON line per run to gpurun_out/tcgen05_gemm.jsonl. Usage on the GPU box:
#   gpurun --timeout 300 -- 'bash tools/micro/run_tcgen05.sh'
set -u
cd "$(dirname "$0")"
bin=./tcgen05_gemm
if [ ! -x "$bin" ]; then
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tcgen05_gemm tcgen05_gemm.cu || exit 1
fi
out=../../gpurun_out/tcgen05_gemm.jsonl
mkdir -p ../../gpurun_out
: > "$out"
run() { timeout 60 "$bin" "$@" | tee -a "$out"; echo "# exit=$? args=$*" | tee -a "$out"; }
for v in 0 1 2 3 4 5 6; do
  run $v 128 64 64 64 0
done
# B operand MN-major without swizzle = the K-major image of the transposed matrix (dX from the forward weight images)
for v in 7 8; do run $v 128 64 64 64 0; run $v 128 64 64 64 1; done
run 7 128 224 64 64 0
run 8 128 48 64 64 0
# encoding fall-backs: LBO/SBO swapped
for v in 1 2 3; do run $v 128 64 64 64 1; done
# shapes the policy layers need: wider N (two layers' worth), K = 32, the M = 64 accumulator lane mapping
run 1 128 128 64 64 0
run 1 128 256 64 64 0
run 1 128 64 32 64 0
run 1 128 64 128 64 0
run 1 64 64 64 64 0
run 3 64 64 128 64 0
run 4 128 256 64 64 0
# layer chain through TMEM (TS form): one and two tiles in flight
if [ ! -x ./tcgen05_mlp ]; then
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tcgen05_mlp tcgen05_mlp.cu || exit 1
fi
for args in "64 1 4" "64 2 4" "64 2 3" "7 2 4"; do
  timeout 60 ./tcgen05_mlp $args | tee -a "$out"; echo "# exit=$? mlp args=$args" | tee -a "$out"
done
# CTA-pair (cta_group::2) GEMM: each CTA holds half of B
if [ ! -x ./tcgen05_gemm2sm ]; then
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tcgen05_gemm2sm tcgen05_gemm2sm.cu || exit 1
fi
for args in "64 64 64" "128 64 64" "256 64 64" "64 32 64"; do
  timeout 60 ./tcgen05_gemm2sm $args | tee -a "$out"; echo "# exit=$? gemm2sm args=$args" | tee -a "$out"
done
# full policy forward (quad concurrent net) on tcgen05: small ragged case first, then the bench size
if [ ! -x ./tcgen05_policy ]; then
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tcgen05_policy tcgen05_policy.cu || exit 1
fi
for args in "1000 3 2" "65536 20 0"; do
  timeout 120 ./tcgen05_policy $args | tee -a "$out"; echo "# exit=$? policy args=$args" | tee -a "$out"
done
# policy backward (dX chain through TMEM with TMA-streamed W^T, dW = dZ^T X from MN-major smem images)
if [ ! -x ./tcgen05_policy_bwd ]; then
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tcgen05_policy_bwd tcgen05_policy_bwd.cu || exit 1
fi
for args in "300 1 2" "8192 5 0" "65536 5 0"; do
  timeout 180 ./tcgen05_policy_bwd $args | tee -a "$out"; echo "# exit=$? policy_bwd args=$args" | tee -a "$out"
done
