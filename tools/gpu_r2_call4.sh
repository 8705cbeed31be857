#!/usr/bin/env bash
# Round-2 call 4: ncu evidence for the tq kernels (launch list + full set, source-level)
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/r2_call4
mkdir -p "$out"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file "$out/launches.csv" \
  python tools/quick_bench.py 65536 > "$out/launches.log" 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'tq_fwd_kernel|tq_dx_kernel|tq_dw_kernel' -s 9 -c 3 -o "$out/tq_kernels" \
  python tools/quick_bench.py 65536 > "$out/ncu_tq.log" 2>&1
echo done > "$out/finished"
