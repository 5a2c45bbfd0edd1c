#!/bin/bash
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
O=gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > $O/s5_pytest.log 2>&1; tail -4 $O/s5_pytest.log
cd tests
MMMFE_SWEEP_AUTOTUNE=0 MMMFE_TRACE_ALL=../$O/s5_trace.npz timeout 300 python gpu_perf_probe.py 1024 256 512 128 256 64 3 > ../$O/s5_probe.log 2>&1
grep -E "^apply|sweep kernel:|sweep batches|cycles/CTA|rror" ../$O/s5_probe.log | cut -c1-900
python gpu_trace_report.py ../$O/s5_trace.npz > ../$O/s5_trace.txt 2>&1; tail -6 ../$O/s5_trace.txt
bash gpu_tune.sh x:conc:SWEEP_CONCURRENT=1 x:k5:SWEEP_CHUNK=5120 x:k7:SWEEP_CHUNK=7168 x:m2048:SWEEP_MICRO_TILE=2048 x:sp50:SWEEP_SPIN_NS=50 x:sp0:SWEEP_SPIN_NS=0 2>&1 | grep -v "cycles/CTA" | cut -c1-200
