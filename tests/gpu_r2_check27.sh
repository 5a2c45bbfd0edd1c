#!/bin/bash
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
O=gpurun_out
(time timeout 600 python -m pytest tests/test_engine_gpu.py -q -x -k "rectangular or watchdog or bar_and_final" --durations=3) > $O/c27_pytest.log 2>&1; tail -25 $O/c27_pytest.log | cut -c1-300
(time MMMFE_SETUP_PROFILE=1 MMMFE_PROBE_NOSOLVE=1 timeout 300 python tests/gpu_solve_probe.py cfg2) > $O/c27_cfg2_setup.log 2>&1; grep -E "host:|config|rror" $O/c27_cfg2_setup.log | cut -c1-300
