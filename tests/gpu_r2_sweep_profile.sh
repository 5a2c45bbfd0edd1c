#!/bin/bash
# new profile of k_sweep after the signal pruning: ncu full capture + per-phase cycle counters + trace
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-solve --no-north-star"
MMMFE_ENGINE_OPTIONS=sweep_autotune=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep --launch-skip 2 -c 1 -o $O/c23_ksweep -f $B > $O/c23_ncu_full.log 2>&1; tail -2 $O/c23_ncu_full.log | cut -c1-200
cd tests
MMMFE_SWEEP_AUTOTUNE=0 MMMFE_TRACE_ALL=../$O/c23_trace.npz timeout 300 python gpu_perf_probe.py 1024 256 512 128 256 64 3 > ../$O/c23_probe.log 2>&1
grep -E "^apply|sweep kernel:|sweep batches|rror|cycles/CTA" ../$O/c23_probe.log | cut -c1-1200
python gpu_trace_report.py ../$O/c23_trace.npz > ../$O/c23_trace.txt 2>&1; tail -8 ../$O/c23_trace.txt
