#!/bin/bash
# final sweep through k_sweep, set-up profile with release timing, launch list of the GMRES loop with the persistent kernel forced
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
O=gpurun_out
(time timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_darcyvt_gpu.py -q -x --durations=3) > $O/c15_pytest.log 2>&1; tail -8 $O/c15_pytest.log | cut -c1-250
(time MMMFE_SETUP_PROFILE=1 MMMFE_PROBE_NOSOLVE=1 timeout 300 python tests/gpu_solve_probe.py cfg2) > $O/c15_cfg2_setup.log 2>&1; grep -E "setup profile|config|rror" $O/c15_cfg2_setup.log | cut -c1-700
(time MMMFE_SETUP_PROFILE=1 timeout 300 python tests/gpu_solve_probe.py cfg4) > $O/c15_cfg4.log 2>&1; grep -E "set-up:|host:|config|rror" $O/c15_cfg4.log | cut -c1-700
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-solve --no-north-star"
MMMFE_ENGINE_OPTIONS=sweep_autotune=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:"k_sweep|k_spmv|k_multi_dot|k_cgs|k_exchange|k_scale|k_final|k_axpy" -c 100 --csv --log-file $O/c15_launches_gmres.csv $B > $O/c15_ncu_launch.log 2>&1; tail -1 $O/c15_ncu_launch.log | cut -c1-200
grep -c k_sweep $O/c15_launches_gmres.csv
