#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/r2_call7
mkdir -p "$out"
timeout 300 python tools/quick_bench.py 65536 > "$out/quick_65536.log" 2>&1
APG_B200_LIB=$PWD/apg_trajectory_tracking_b200/libapg_b200_prof.so timeout 300 python tools/tq_profile.py > "$out/tq_profile.log" 2>&1
timeout 900 python -m pytest tests -q -m gpu -x > "$out/pytest_gpu.log" 2>&1
echo "exit=$?" >> "$out/pytest_gpu.log"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 48 --csv --log-file "$out/launches.csv" \
  python tools/quick_bench.py 65536 > "$out/launches.log" 2>&1
echo done > "$out/finished"
