#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/r2_rate
mkdir -p "$out"
R=tools/micro/tcgen05_rate
: > "$out/rate.jsonl"
for ts in 0 1; do for sw in 0 1; do
  for cfg in "64 1" "64 2" "64 4" "48 1" "48 4" "128 1" "128 2" "256 1" "16 1" "16 4" "32 1" "32 4"; do
    set -- $cfg
    timeout 60 $R $ts $1 $2 2048 $sw 148 >> "$out/rate.jsonl" 2>&1
  done
done; done
timeout 60 $R 1 64 1 2048 1 1 >> "$out/rate.jsonl" 2>&1
timeout 60 $R 1 64 1 64 1 148 >> "$out/rate.jsonl" 2>&1
timeout 60 $R 1 64 1 12 1 148 >> "$out/rate.jsonl" 2>&1
echo done > "$out/finished"
