#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/r2_dyn1
mkdir -p "$out"
timeout 600 python -m pytest tests/test_zz_new_paths_gpu.py -q -m gpu -x -k "tc" > "$out/pytest_tc.log" 2>&1
echo "exit=$?" >> "$out/pytest_tc.log"
timeout 300 python tools/tq_kernel_times.py 65536 > "$out/times_default.log" 2>&1
APG_B200_LIB=$PWD/apg_trajectory_tracking_b200/libapg_b200_noprefetch.so timeout 300 python tools/tq_kernel_times.py 65536 > "$out/times_noprefetch.log" 2>&1
timeout 300 python tools/tq_kernel_times.py 65536 >> "$out/times_default.log" 2>&1
echo done > "$out/finished"
