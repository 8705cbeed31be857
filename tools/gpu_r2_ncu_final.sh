#!/usr/bin/env bash
# ncu evidence for the kernels as they stand at the end of round 2: launch list of one bench run, one --set full capture
# of the five tcgen05-path kernels, and a launch list of the GPU tests that drive the kernels added in the second session
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/r2_ncu_final
mkdir -p "$out"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/launches_bench.csv" \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-raw-e2e --no-live-traffic > "$out/bench_under_ncu.log" 2>&1
timeout 400 ncu --set full --clock-control none --import-source on \
  -k regex:'tq_fwd_kernel|tq_dx_kernel|tq_dw_kernel|tq_dyn_kernel|apg_reduce4' -s 14 -c 5 -o "$out/tq_kernels" \
  python tools/quick_bench.py 65536 > "$out/ncu_tq.log" 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$out/launches_new_tests.csv" \
  python -m pytest tests/test_zz_new_paths_gpu.py -q -m gpu -k "learnt_dynamics_fused_vs_oracle or lstm_policy_batched" > "$out/new_tests.log" 2>&1
echo done > "$out/finished"
