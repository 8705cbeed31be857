#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/r2_b1
mkdir -p "$out"
timeout 900 python bench.py --no-cpu-baseline > "$out/bench_quad_concurrent.json" 2> "$out/bench_quad_concurrent.err"
timeout 900 python bench.py --no-cpu-baseline --steps 20 > "$out/bench_quad_concurrent_b.json" 2> "$out/bench_quad_concurrent_b.err"
echo done > "$out/finished"
