#!/bin/bash
# watchdog test (bounded), engine tests, perf probe of k_sweep after the counter prefetch
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
O=gpurun_out
(time timeout 120 python -m pytest tests/test_engine_gpu.py -q -x -k watchdog) > $O/c18_watchdog.log 2>&1; tail -5 $O/c18_watchdog.log | cut -c1-300
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
(time timeout 600 python -m pytest tests/test_engine_gpu.py -q -x --durations=3) > $O/c18_pytest.log 2>&1; tail -5 $O/c18_pytest.log | cut -c1-250
(time MMMFE_PROBE_NOSOLVE=1 timeout 300 python tests/gpu_solve_probe.py cfg2) > $O/c18_cfg2.log 2>&1; grep -E "config|rror" $O/c18_cfg2.log | cut -c1-900
(time timeout 300 python tests/gpu_solve_probe.py cfg4 operator=0) > $O/c18_cfg4_sweeps.log 2>&1; grep -E "config|rror" $O/c18_cfg4_sweeps.log | cut -c1-900
