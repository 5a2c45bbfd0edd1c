#!/bin/bash
# one signal per (CTA, front): engine + full-size parity tests, cfg2 / cfg4-sweeps perf
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
O=gpurun_out
(time timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_fullsize_gpu.py -q -x --durations=3) > $O/c21_pytest.log 2>&1; tail -6 $O/c21_pytest.log | cut -c1-250
(time MMMFE_PROBE_NOSOLVE=1 timeout 300 python tests/gpu_solve_probe.py cfg2) > $O/c21_cfg2.log 2>&1; grep -E "config|rror" $O/c21_cfg2.log | cut -c1-900
(time MMMFE_PROBE_NOSOLVE=1 timeout 300 python tests/gpu_solve_probe.py cfg4 operator=0) > $O/c21_cfg4_sweeps.log 2>&1; grep -E "config|rror" $O/c21_cfg4_sweeps.log | cut -c1-900
