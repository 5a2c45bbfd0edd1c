#!/bin/bash
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
O=gpurun_out
nproc; free -g | head -2
(time MMMFE_SETUP_PROFILE=1 MMMFE_PROBE_NOSOLVE=1 timeout 300 python tests/gpu_solve_probe.py cfg2) > $O/c9_cfg2_setup.log 2>&1; grep -E "setup profile|config|rror" $O/c9_cfg2_setup.log | cut -c1-1200
(time MMMFE_SETUP_PROFILE=1 timeout 300 python tests/gpu_solve_probe.py cfg4) > $O/c9_cfg4.log 2>&1; grep -E "setup profile|config|rror" $O/c9_cfg4.log | cut -c1-1500
(time timeout 600 python -m pytest tests/test_engine_gpu.py -q -x -k "variable or full_solve or neumann") > $O/c9_engine.log 2>&1; tail -4 $O/c9_engine.log | cut -c1-300
