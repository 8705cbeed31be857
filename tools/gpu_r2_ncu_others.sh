#!/usr/bin/env bash
# ncu evidence for the workloads besides the headline one: launch list + one --set full capture of each main kernel
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/r2_ncu_others
mkdir -p "$out"
for w in wing_concurrent quad_autoregressive quad_lstm cartpole_concurrent; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file "$out/launches_$w.csv" \
    python tools/one_step.py $w > "$out/one_step_$w.log" 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fwd_kernel|adj_kernel' -s 4 -c 2 \
    -o "$out/kernels_$w" python tools/one_step.py $w >> "$out/one_step_$w.log" 2>&1
done
# evaluation / input-side kernels: one launch list of the GPU tests that drive them
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/launches_eval_prep_tests.csv" \
  python -m pytest tests/test_zz_new_paths_gpu.py -q -m gpu -k "eval_rollout_batched or wing_fly_to_points_batched or cartpole_balance_batched or prepare_quad_random or sample_windows_and_poly or learnt_dynamics_many_tiles" > "$out/eval_prep_tests.log" 2>&1
echo done > "$out/finished"
