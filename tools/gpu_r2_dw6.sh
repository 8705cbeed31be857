#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/r2_dw6
mkdir -p "$out"
for v in profnomma profnofeed; do
APG_B200_LIB=$PWD/apg_trajectory_tracking_b200/libapg_b200_$v.so timeout 300 python tools/tq_profile.py > "$out/tq_profile_$v.log" 2>&1
done
echo done > "$out/finished"
