#!/bin/bash
cd "$(dirname "$0")"; mkdir -p ../gpurun_out
(timeout 900 python -m pytest test_engine_gpu.py -m gpu -q -x) 2>&1 | tail -2
MMMFE_SWEEP_AUTOTUNE=0 MMMFE_TRACE_ALL=../gpurun_out/s10_trace.npz timeout 300 python gpu_perf_probe.py 1024 256 512 128 256 64 3 > ../gpurun_out/s10_probe.log 2>&1
grep -E "^apply|sweep kernel:|sweep batches|rror" ../gpurun_out/s10_probe.log | cut -c1-300
python gpu_trace_report.py ../gpurun_out/s10_trace.npz > ../gpurun_out/s10_trace.txt 2>&1; grep -E "^sub|per CTA|sub fwd|sub bwd" ../gpurun_out/s10_trace.txt | cut -c1-220

