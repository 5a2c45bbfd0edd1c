#!/bin/bash
# round-2 ncu evidence: launch list that reaches the GMRES loop, full capture of k_sweep, DMMA utilisation of the factorisation GEMM
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-solve --no-north-star"
(time timeout 300 $B) > $O/c10_bench.log 2> $O/c10_bench.err; grep '^{"metric"' $O/c10_bench.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:"k_sweep|k_spmv|k_multi_dot|k_cgs|k_exchange|k_scale|k_final|k_norm|k_axpy|k_copy|k_pack|k_unpack" -c 400 --csv --log-file $O/c10_launches_gmres.csv $B > $O/c10_ncu_launch.log 2>&1; tail -1 $O/c10_ncu_launch.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep -c 2 -o $O/c10_ksweep -f $B > $O/c10_ncu_full.log 2>&1; tail -2 $O/c10_ncu_full.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_tensor_subpipe_dmma.sum,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:"k_grouped_gemm|k_small_chol|k_extend|k_assemble|k_front" -c 3000 --csv --log-file $O/c10_factor_kernels.csv env MMMFE_PROBE_NOSOLVE=1 python tests/gpu_solve_probe.py cfg2 > $O/c10_ncu_factor.log 2>&1; tail -1 $O/c10_ncu_factor.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_grouped_gemm --launch-skip 300 -c 12 -o $O/c10_gemm -f env MMMFE_PROBE_NOSOLVE=1 python tests/gpu_solve_probe.py cfg2 > $O/c10_ncu_gemm.log 2>&1; tail -2 $O/c10_ncu_gemm.log | cut -c1-200
ls -la $O/c10*
