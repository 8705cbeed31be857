#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/r2_det
mkdir -p "$out"
timeout 300 python tools/raw_vs_prepared.py 65536 > "$out/raw_vs_prepared.log" 2>&1
timeout 300 python tools/raw_vs_prepared.py 37893 >> "$out/raw_vs_prepared.log" 2>&1
timeout 600 python -m pytest tests/test_zz_new_paths_gpu.py -q -m gpu -k "tc" > "$out/pytest_tc.log" 2>&1
echo "exit=$?" >> "$out/pytest_tc.log"
timeout 300 python tools/tq_kernel_times.py 65536 > "$out/times_default.log" 2>&1
echo done > "$out/finished"
