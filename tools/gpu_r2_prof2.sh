#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/r2_prof2
mkdir -p "$out"
APG_B200_LIB=$PWD/apg_trajectory_tracking_b200/libapg_b200_prof.so timeout 300 python tools/tq_profile.py > "$out/tq_profile.log" 2>&1
