#!/bin/bash
# pipelined GMRES + grid-wide Arnoldi kernels, device data path, launch list of the GMRES loop without graphs
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
O=gpurun_out
(time timeout 1200 python -m pytest tests/test_darcyvt_gpu.py tests/test_engine_gpu.py tests/test_basis_gpu.py -q -x --durations=5) > $O/c12_pytest.log 2>&1; tail -12 $O/c12_pytest.log | cut -c1-250
(time MMMFE_SETUP_PROFILE=1 MMMFE_PROBE_NOSOLVE=1 timeout 300 python tests/gpu_solve_probe.py cfg2) > $O/c12_cfg2_setup.log 2>&1; grep -E "config|rror" $O/c12_cfg2_setup.log | cut -c1-700
(time timeout 300 python tests/gpu_solve_probe.py cfg4) > $O/c12_cfg4.log 2>&1; grep -E "config|rror" $O/c12_cfg4.log | cut -c1-700
(time timeout 300 python tests/gpu_solve_probe.py cfg4 gmres_lookahead=0) > $O/c12_cfg4_lag0.log 2>&1; grep -E "config|rror" $O/c12_cfg4_lag0.log | cut -c1-700
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-solve --no-north-star"
MMMFE_ENGINE_OPTIONS=use_graphs=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:"k_sweep|k_spmv|k_multi_dot|k_cgs|k_exchange|k_scale|k_final|k_axpy" --launch-skip 4 -c 120 --csv --log-file $O/c12_launches_gmres.csv $B > $O/c12_ncu_launch.log 2>&1; tail -1 $O/c12_ncu_launch.log | cut -c1-200
grep -c k_sweep $O/c12_launches_gmres.csv
