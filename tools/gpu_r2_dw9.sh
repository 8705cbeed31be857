#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/r2_dw9
mkdir -p "$out"
timeout 600 python -m pytest tests/test_zz_new_paths_gpu.py -q -m gpu -x -k "tc" > "$out/pytest_tc.log" 2>&1
echo "exit=$?" >> "$out/pytest_tc.log"
timeout 300 python tools/tq_kernel_times.py 65536 > "$out/times_default.log" 2>&1
APG_B200_LIB=$PWD/apg_trajectory_tracking_b200/libapg_b200_prof.so timeout 300 python tools/tq_profile.py > "$out/tq_profile.log" 2>&1
timeout 300 python tools/raw_vs_prepared.py 65536 > "$out/raw_vs_prepared.log" 2>&1
timeout 900 python bench.py --no-cpu-baseline > "$out/bench_quad_concurrent.json" 2> "$out/bench_quad_concurrent.err"
echo done > "$out/finished"
