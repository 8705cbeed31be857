#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/r2_call11
mkdir -p "$out"
timeout 600 python tools/tq_kernel_times.py 18944 65536 > "$out/times.log" 2>&1
APG_B200_LIB=$PWD/apg_trajectory_tracking_b200/libapg_b200_prof.so timeout 300 python tools/tq_profile.py > "$out/tq_profile.log" 2>&1
timeout 900 python -m pytest tests -q -m gpu -x > "$out/pytest_gpu.log" 2>&1
echo "exit=$?" >> "$out/pytest_gpu.log"
echo done > "$out/finished"
