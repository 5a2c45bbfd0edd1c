#!/bin/bash
# round-2 final state on one B200: full GPU suite, smoke, bench (both arms), launch list of the GMRES loop, full capture of k_sweep
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
O=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -x --durations=5) > $O/f1_pytest.log 2>&1; tail -10 $O/f1_pytest.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
(time timeout 900 python bench.py --steps 5 --warmup 3) > $O/f1_bench.log 2> $O/f1_bench.err; grep '^{"metric"' $O/f1_bench.log | cut -c1-500; tail -3 $O/f1_bench.err
(time timeout 600 python bench.py --impl reference --steps 2 --warmup 1) > $O/f1_bench_ref.log 2> $O/f1_bench_ref.err; cut -c1-400 $O/f1_bench_ref.log | tail -2
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-solve --no-north-star"
MMMFE_ENGINE_OPTIONS=sweep_autotune=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:"k_sweep|k_spmv|k_multi_dot|k_cgs|k_exchange|k_scale|k_final|k_axpy" -c 100 --csv --log-file $O/f1_launches_gmres.csv $B > $O/f1_ncu_launch.log 2>&1; tail -1 $O/f1_ncu_launch.log | cut -c1-200
MMMFE_ENGINE_OPTIONS=sweep_autotune=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep --launch-skip 2 -c 1 -o $O/f1_ksweep -f $B > $O/f1_ncu_full.log 2>&1; tail -2 $O/f1_ncu_full.log | cut -c1-200
