#!/bin/bash
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
O=gpurun_out
(time timeout 900 python -m pytest tests/test_basis_gpu.py -q -x --durations=3) > $O/c4_basis.log 2>&1; tail -6 $O/c4_basis.log | cut -c1-250
(time MMMFE_BASIS_PROFILE=1 timeout 300 python tests/gpu_solve_probe.py cfg4 operator=1) > $O/c4_cfg4_basis.log 2>&1; grep -E "basis profile|config|rror" $O/c4_cfg4_basis.log | cut -c1-1700
(time MMMFE_BASIS_PROFILE=2 timeout 300 python tests/gpu_solve_probe.py cfg2 operator=1,basis_columns=64) > $O/c4_cfg2_prof64.log 2>&1; grep -E "basis profile" $O/c4_cfg2_prof64.log | cut -c1-1600
(time timeout 900 python -m pytest tests/test_darcyvt_gpu.py -q -x --durations=3) > $O/c4_darcy.log 2>&1; tail -6 $O/c4_darcy.log | cut -c1-250
