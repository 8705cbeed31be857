#!/usr/bin/env bash
# Round-2 first GPU call: whole GPU suite WITHOUT -x, the opt-in tcgen05 tests, the tcgen05 microbenchmarks, benches.
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/r2_call1
mkdir -p "$out"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > "$out/gpu.txt" 2>&1
timeout 900 python -m pytest tests -q -m gpu > "$out/pytest_gpu.log" 2>&1
echo "exit=$?" >> "$out/pytest_gpu.log"
timeout 600 bash tools/micro/run_tcgen05.sh > "$out/tcgen05.log" 2>&1
cp -f gpurun_out/tcgen05_gemm.jsonl "$out/" 2>/dev/null
APG_TEST_TC=1 timeout 600 python -m pytest tests/test_zz_new_paths_gpu.py -q -m gpu -k "tc1 or tc2 or tc3" > "$out/pytest_tc.log" 2>&1
echo "exit=$?" >> "$out/pytest_tc.log"
timeout 600 python bench.py > "$out/bench_quad_concurrent.json" 2> "$out/bench_quad_concurrent.err"
for w in wing_concurrent quad_autoregressive quad_lstm cartpole_concurrent; do
  timeout 300 python bench.py --workload $w --steps 20 --no-cpu-baseline > "$out/bench_$w.json" 2> "$out/bench_$w.err"
done
timeout 300 python bench.py --tc-forward --steps 30 --no-cpu-baseline --no-raw-e2e > "$out/bench_tc_forward.json" 2> "$out/bench_tc_forward.err"
timeout 300 python bench.py --tc-dw --steps 30 --no-cpu-baseline --no-raw-e2e > "$out/bench_tc_dw.json" 2> "$out/bench_tc_dw.err"
timeout 300 python bench.py --tc-forward --tc-dw --steps 30 --no-cpu-baseline --no-raw-e2e > "$out/bench_tc_both.json" 2> "$out/bench_tc_both.err"
timeout 300 python bench.py --tc-forward --tc-dx --tc-dw --steps 30 --no-cpu-baseline --no-raw-e2e > "$out/bench_tc_all.json" 2> "$out/bench_tc_all.err"
echo done > "$out/finished"
