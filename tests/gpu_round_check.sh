#!/bin/bash
# final state of the round: tests, smoke, bench (N = 1), perf probe with trace, ncu launch list + full capture of k_sweep
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
O=gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > $O/rc_pytest.log 2>&1; tail -4 $O/rc_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
(time timeout 500 python bench.py --steps 5 --warmup 3) > $O/rc_bench.log 2> $O/rc_bench.err; grep '^{"metric"' $O/rc_bench.log | cut -c1-400; tail -3 $O/rc_bench.err
(time timeout 500 python bench.py --impl reference --steps 2 --warmup 1) > $O/rc_bench_ref.log 2> $O/rc_bench_ref.err; cut -c1-300 $O/rc_bench_ref.log | tail -2
cd tests
MMMFE_SWEEP_AUTOTUNE=0 MMMFE_TRACE_ALL=../$O/rc_trace.npz timeout 300 python gpu_perf_probe.py 1024 256 512 128 256 64 3 > ../$O/rc_probe.log 2>&1
grep -E "^apply|sweep kernel:|sweep batches|rror" ../$O/rc_probe.log | cut -c1-300
python gpu_trace_report.py ../$O/rc_trace.npz > ../$O/rc_trace.txt 2>&1; tail -6 ../$O/rc_trace.txt
cd ..
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file $O/rc_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/rc_ncu_launch.log 2>&1; tail -1 $O/rc_ncu_launch.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep -c 2 -o $O/rc_ksweep -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/rc_ncu_full.log 2>&1; tail -2 $O/rc_ncu_full.log | cut -c1-200
