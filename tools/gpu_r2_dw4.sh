#!/usr/bin/env bash
# dW kernel: where does the time go (probes without MMAs / without feeders), PDL launches
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/r2_dw8
mkdir -p "$out"
timeout 600 python -m pytest tests/test_zz_new_paths_gpu.py -q -m gpu -x -k "tc1 or tc2 or tc3" > "$out/pytest_tc.log" 2>&1
echo "exit=$?" >> "$out/pytest_tc.log"
timeout 300 python tools/tq_kernel_times.py 65536 > "$out/times_default.log" 2>&1
for v in nomma nofeed; do
APG_B200_LIB=$PWD/apg_trajectory_tracking_b200/libapg_b200_$v.so timeout 300 python tools/tq_kernel_times.py 65536 > "$out/times_$v.log" 2>&1
done
APG_B200_LIB=$PWD/apg_trajectory_tracking_b200/libapg_b200_prof.so timeout 300 python tools/tq_profile.py > "$out/tq_profile.log" 2>&1
timeout 900 python bench.py --no-cpu-baseline > "$out/bench_quad_concurrent.json" 2> "$out/bench_quad_concurrent.err"
timeout 900 python -m pytest tests -q -m gpu > "$out/pytest_gpu.log" 2>&1
echo "exit=$?" >> "$out/pytest_gpu.log"
echo done > "$out/finished"
