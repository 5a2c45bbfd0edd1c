#!/bin/bash
# full GPU suite, bench (N = 1), set-up profile, launch list of the GMRES loop (graphs off so that ncu sees the cooperative launches)
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
O=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -x --durations=5) > $O/c14_pytest.log 2>&1; tail -12 $O/c14_pytest.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
(time MMMFE_SETUP_PROFILE=1 MMMFE_PROBE_NOSOLVE=1 timeout 300 python tests/gpu_solve_probe.py cfg2) > $O/c14_cfg2_setup.log 2>&1; grep -E "setup profile|config|rror" $O/c14_cfg2_setup.log | cut -c1-700
(time MMMFE_SETUP_PROFILE=1 timeout 300 python tests/gpu_solve_probe.py cfg4) > $O/c14_cfg4.log 2>&1; grep -E "set-up:|host:|config|rror" $O/c14_cfg4.log | cut -c1-700
(time timeout 600 python bench.py --steps 5 --warmup 3) > $O/c14_bench.log 2> $O/c14_bench.err; grep '^{"metric"' $O/c14_bench.log | cut -c1-600; tail -3 $O/c14_bench.err
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-solve --no-north-star"
MMMFE_ENGINE_OPTIONS=use_graphs=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:"k_sweep|k_spmv|k_multi_dot|k_cgs|k_exchange|k_scale|k_final|k_axpy" --launch-skip 4 -c 120 --csv --log-file $O/c14_launches_gmres.csv $B > $O/c14_ncu_launch.log 2>&1; tail -1 $O/c14_ncu_launch.log | cut -c1-200
grep -c k_sweep $O/c14_launches_gmres.csv
