#!/usr/bin/env bash
# Round-2 call 2: suite again (expect green), descriptor address-function probes, ncu captures of the tcgen05 kernels.
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/r2_call2
mkdir -p "$out"
timeout 900 python -m pytest tests -q -m gpu > "$out/pytest_gpu.log" 2>&1
echo "exit=$?" >> "$out/pytest_gpu.log"
P=tools/micro/tcgen05_probe
{
  # controls: K-major unswizzled / 128B-swizzled (known good)
  timeout 30 $P B 0 0 128 256
  timeout 30 $P B 0 2 16 1024
  timeout 30 $P A 0 0 128 256
  # MN-major unswizzled: (lbo, sbo) candidates
  timeout 30 $P B 1 0 2048 128
  timeout 30 $P B 1 0 128 2048
  timeout 30 $P B 1 0 256 128
  timeout 30 $P B 1 0 128 256
  # MN-major swizzled
  timeout 30 $P B 1 2 8192 1024
  timeout 30 $P B 1 2 1024 8192
  timeout 30 $P B 1 4 4096 512
  timeout 30 $P B 1 6 2048 256
  timeout 30 $P A 1 0 2048 128
  timeout 30 $P A 1 2 8192 1024
  timeout 30 $P B 1 2 4096 1024 32
  timeout 30 $P B 1 0 2048 128 32
} > "$out/probe.jsonl" 2>&1
# ncu: one launch of each tcgen05 kernel + the default pair (full set, source-level)
APG_TC_FWD=1 APG_TC_DW=1 timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:'hutter_fwd_tc_kernel|adj_dw_tc_kernel|hutter_adj_dx_kernel' -s 9 -c 3 -o "$out/tc_kernels" \
  python tools/quick_bench.py 65536 > "$out/ncu_tc.log" 2>&1
echo done > "$out/finished"
