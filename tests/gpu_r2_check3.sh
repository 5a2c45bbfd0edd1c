#!/bin/bash
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
O=gpurun_out
(time timeout 900 python -m pytest tests/test_basis_gpu.py -q --durations=5) > $O/c3_basis.log 2>&1; tail -15 $O/c3_basis.log | cut -c1-250
(time timeout 600 python -m pytest tests/test_engine_gpu.py -q -x -k "saddle or full_solve or checkerboard") > $O/c3_engine.log 2>&1; tail -5 $O/c3_engine.log | cut -c1-250
(time MMMFE_BASIS_PROFILE=1 timeout 300 python tests/gpu_solve_probe.py cfg4 operator=1) > $O/c3_cfg4_basis.log 2>&1; grep -E "basis profile|config" $O/c3_cfg4_basis.log | cut -c1-1600
(time MMMFE_BASIS_PROFILE=2 timeout 300 python tests/gpu_solve_probe.py cfg2 operator=1) > $O/c3_cfg2_prof.log 2>&1; grep -E "basis profile" $O/c3_cfg2_prof.log | cut -c1-1600
(time MMMFE_BASIS_PROFILE=2 timeout 300 python tests/gpu_solve_probe.py cfg2 operator=1,basis_columns=64) > $O/c3_cfg2_prof64.log 2>&1; grep -E "basis profile" $O/c3_cfg2_prof64.log | cut -c1-1600
