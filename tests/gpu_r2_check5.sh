#!/bin/bash
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
O=gpurun_out
(time timeout 900 python -m pytest tests/test_basis_gpu.py -q -x --durations=3) > $O/c5_basis.log 2>&1; tail -6 $O/c5_basis.log | cut -c1-250
(time MMMFE_BASIS_PROFILE=1 timeout 300 python tests/gpu_solve_probe.py cfg4 operator=1) > $O/c5_cfg4_basis.log 2>&1; grep -E "basis profile|config|rror" $O/c5_cfg4_basis.log | cut -c1-1700
(time MMMFE_BASIS_PROFILE=2 timeout 300 python tests/gpu_solve_probe.py cfg2 operator=1,basis_columns=64) > $O/c5_cfg2_prof64.log 2>&1; grep -E "basis profile" $O/c5_cfg2_prof64.log | cut -c1-1600
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wide_subtree_forward --launch-skip 40 -c 1 -o $O/c5_subtree_fwd -f python tests/gpu_solve_probe.py cfg4 operator=1,use_graphs=0 > $O/c5_ncu1.log 2>&1; tail -2 $O/c5_ncu1.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_toeplitz_apply --launch-skip 5 -c 1 -o $O/c5_toeplitz -f python tests/gpu_solve_probe.py cfg4 operator=1 > $O/c5_ncu2.log 2>&1; tail -2 $O/c5_ncu2.log | cut -c1-200
