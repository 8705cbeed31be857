#!/usr/bin/env bash
# raw-input mode on hardware + tanh variant timing + kernel timelines
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/r2_raw1
mkdir -p "$out"
timeout 900 python -m pytest tests -q -m gpu -x > "$out/pytest_gpu.log" 2>&1
echo "exit=$?" >> "$out/pytest_gpu.log"
timeout 300 python tools/tq_kernel_times.py 65536 > "$out/times_default.log" 2>&1
APG_B200_LIB=$PWD/apg_trajectory_tracking_b200/libapg_b200_sfutanh.so timeout 300 python tools/tq_kernel_times.py 65536 > "$out/times_sfutanh.log" 2>&1
APG_B200_LIB=$PWD/apg_trajectory_tracking_b200/libapg_b200_prof.so timeout 300 python tools/tq_profile.py > "$out/tq_profile.log" 2>&1
timeout 900 python bench.py --no-cpu-baseline > "$out/bench_quad_concurrent.json" 2> "$out/bench_quad_concurrent.err"
echo done > "$out/finished"
