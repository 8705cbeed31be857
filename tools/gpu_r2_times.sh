#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/r2_times
mkdir -p "$out"
timeout 600 python tools/tq_kernel_times.py 18944 37888 56832 65536 75776 151552 > "$out/times.log" 2>&1
