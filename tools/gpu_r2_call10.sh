#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/r2_call10
mkdir -p "$out"
timeout 900 python -m pytest tests -q -m gpu -x > "$out/pytest_gpu.log" 2>&1
echo "exit=$?" >> "$out/pytest_gpu.log"
timeout 600 python tools/tq_kernel_times.py 65536 > "$out/times.log" 2>&1
timeout 900 python bench.py > "$out/bench.json" 2> "$out/bench.err"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > "$out/bench_reference.json" 2> "$out/bench_reference.err"
echo done > "$out/finished"
