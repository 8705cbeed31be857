#!/usr/bin/env bash
# Round-2 closing run on 1 GPU: smoke, the whole GPU suite, the bench line the driver will ask for (live ncu traffic),
# the other workloads
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/r2_final5
mkdir -p "$out"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > "$out/smoke.log" 2>&1
echo "exit=$?" >> "$out/smoke.log"
timeout 1200 python -m pytest tests -q -m gpu > "$out/pytest_gpu.log" 2>&1
echo "exit=$?" >> "$out/pytest_gpu.log"
timeout 900 python bench.py > "$out/bench_quad_concurrent.json" 2> "$out/bench_quad_concurrent.err"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > "$out/bench_reference_arm.json" 2> "$out/bench_reference_arm.err"
for w in wing_concurrent quad_autoregressive quad_lstm cartpole_concurrent; do
  timeout 400 python bench.py --workload $w --steps 20 --no-cpu-baseline > "$out/bench_$w.json" 2> "$out/bench_$w.err"
done
timeout 300 python tools/tq_kernel_times.py 65536 > "$out/kernel_times.log" 2>&1
echo done > "$out/finished"
