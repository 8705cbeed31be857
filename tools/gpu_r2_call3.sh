#!/usr/bin/env bash
# Round-2 call 3: first hardware run of the second-generation tcgen05 kernels (default path for quad concurrent)
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/r2_call3
mkdir -p "$out"
timeout 300 python tools/quick_bench.py 1000 > "$out/quick_1000.log" 2>&1
timeout 300 python tools/quick_bench.py 65536 > "$out/quick_65536.log" 2>&1
APG_LEGACY_MMA=1 timeout 300 python tools/quick_bench.py 65536 > "$out/quick_65536_legacy.log" 2>&1
timeout 900 python -m pytest tests -q -m gpu -x > "$out/pytest_gpu.log" 2>&1
echo "exit=$?" >> "$out/pytest_gpu.log"
timeout 600 python bench.py --steps 30 --no-cpu-baseline --no-raw-e2e > "$out/bench.json" 2> "$out/bench.err"
echo done > "$out/finished"
