#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/r2_pp
mkdir -p "$out"
: > "$out/pingpong.jsonl"
for mode in 0 1; do for depth in 1 2 4 7; do for nbusy in 0 8; do for w in 0 1; do
  timeout 30 tools/micro/mbar_pingpong $mode $depth 4096 $nbusy $w >> "$out/pingpong.jsonl" 2>&1
done; done; done; done
echo done > "$out/finished"
