#!/bin/bash
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
O=gpurun_out
for opt in "tune_steps=32" "tune_steps=0" "tune_steps=64"; do
(time MMMFE_SETUP_PROFILE=1 MMMFE_PROBE_NOSOLVE=1 timeout 300 python tests/gpu_solve_probe.py cfg2 $opt) > $O/c13_cfg2_$opt.log 2>&1; grep -E "set-up:|host:|config|rror" $O/c13_cfg2_$opt.log | cut -c1-420
done
