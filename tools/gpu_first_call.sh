#!/usr/bin/env bash
# First GPU call after a GPU-less stretch of work: everything that was only compile-verified, in order of importance,
# each step under its own timeout and with its own log, so that one failure does not hide the rest.
#   gpurun --timeout 1500 -- 'bash tools/gpu_first_call.sh'
# Outputs under gpurun_out/first_call/ (merged back by gpurun).
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/first_call
mkdir -p "$out"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > "$out/gpu.txt" 2>&1

# 1. the previously verified parity suite (must stay green), then the new input-side / evaluation / learnt tests
timeout 900 python -m pytest tests -q -m gpu -x --deselect tests/test_zz_new_paths_gpu.py > "$out/pytest_verified.log" 2>&1
echo "exit=$?" >> "$out/pytest_verified.log"
APG_TEST_TC=1 timeout 600 python -m pytest tests/test_zz_new_paths_gpu.py -q -m gpu > "$out/pytest_new.log" 2>&1
echo "exit=$?" >> "$out/pytest_new.log"

# 2. bench, default workload (raw-sample e2e arm included), then the other workloads without the CPU baseline
timeout 600 python bench.py > "$out/bench_quad_concurrent.json" 2> "$out/bench_quad_concurrent.err"
for w in wing_concurrent quad_autoregressive quad_lstm cartpole_concurrent; do
  timeout 300 python bench.py --workload $w --steps 20 --no-cpu-baseline > "$out/bench_$w.json" 2> "$out/bench_$w.err"
done

# 2b. the optional tcgen05 forward in the product bench (only meaningful if its parity test passed above)
timeout 300 python bench.py --tc-forward --steps 30 --no-cpu-baseline --no-raw-e2e > "$out/bench_tc_forward.json" 2> "$out/bench_tc_forward.err"

timeout 300 python bench.py --tc-dw --steps 30 --no-cpu-baseline --no-raw-e2e > "$out/bench_tc_dw.json" 2> "$out/bench_tc_dw.err"
timeout 300 python bench.py --tc-forward --tc-dw --steps 30 --no-cpu-baseline --no-raw-e2e > "$out/bench_tc_both.json" 2> "$out/bench_tc_both.err"
timeout 300 python bench.py --tc-forward --tc-dx --steps 30 --no-cpu-baseline --no-raw-e2e > "$out/bench_tc_all.json" 2> "$out/bench_tc_all.err"

# 3. launch list of a short bench run (shares of the step, not absolute times)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/launches.csv" \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-raw-e2e > "$out/bench_under_ncu.log" 2>&1

# 4. tcgen05 / TMEM microbenchmarks and prototypes (tools/micro), one process per variant
timeout 900 bash tools/micro/run_tcgen05.sh > "$out/tcgen05.log" 2>&1
cp -f gpurun_out/tcgen05_gemm.jsonl "$out/" 2>/dev/null
echo done > "$out/finished"
