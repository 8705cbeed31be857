#!/usr/bin/env bash
# Round-2 evidence run on 1 GPU: suite, benches of all workloads, reference arm, ncu launch list + full captures
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/r2_final2
mkdir -p "$out"
timeout 900 python -m pytest tests -q -m gpu > "$out/pytest_gpu.log" 2>&1
echo "exit=$?" >> "$out/pytest_gpu.log"
timeout 900 python bench.py > "$out/bench_quad_concurrent.json" 2> "$out/bench_quad_concurrent.err"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > "$out/bench_reference_arm.json" 2> "$out/bench_reference_arm.err"
timeout 600 python bench.py --legacy-mma --steps 50 --no-cpu-baseline --no-raw-e2e > "$out/bench_quad_concurrent_legacy_mma.json" 2> "$out/bench_legacy.err"
for w in wing_concurrent quad_autoregressive quad_lstm cartpole_concurrent; do
  timeout 400 python bench.py --workload $w --steps 20 --no-cpu-baseline > "$out/bench_$w.json" 2> "$out/bench_$w.err"
done
timeout 300 python tools/tq_kernel_times.py 65536 > "$out/kernel_times.log" 2>&1
APG_B200_LIB=$PWD/apg_trajectory_tracking_b200/libapg_b200_prof.so timeout 300 python tools/tq_profile.py > "$out/tq_profile.log" 2>&1
timeout 300 python tools/raw_vs_prepared.py 65536 > "$out/raw_vs_prepared.log" 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/launches_bench.csv" \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-raw-e2e --no-live-traffic > "$out/bench_under_ncu.log" 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'tq_fwd_kernel|tq_dx_kernel|tq_dw_kernel|tq_dyn_kernel|apg_reduce4' -s 14 -c 5 -o "$out/tq_kernels" \
  python tools/quick_bench.py 65536 > "$out/ncu_tq.log" 2>&1
echo done > "$out/finished"
