#!/usr/bin/env bash
# Round-2 call 5: restructured tq kernels (column-split chains, separate dynamics kernel, deep raw ring in dW)
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/r2_call5
mkdir -p "$out"
timeout 300 python tools/quick_bench.py 1000 > "$out/quick_1000.log" 2>&1
timeout 300 python tools/quick_bench.py 65536 > "$out/quick_65536.log" 2>&1
timeout 900 python -m pytest tests -q -m gpu -x > "$out/pytest_gpu.log" 2>&1
echo "exit=$?" >> "$out/pytest_gpu.log"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 48 --csv --log-file "$out/launches.csv" \
  python tools/quick_bench.py 65536 > "$out/launches.log" 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'tq_fwd_kernel|tq_dx_kernel|tq_dw_kernel|tq_dyn_kernel' -s 12 -c 4 -o "$out/tq_kernels" \
  python tools/quick_bench.py 65536 > "$out/ncu_tq.log" 2>&1
echo done > "$out/finished"
