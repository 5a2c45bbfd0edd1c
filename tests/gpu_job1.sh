#!/bin/bash
# first GPU job of this session: tests, bench, perf probe with trace, tuning variants, ncu launch list + full capture
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L
(time timeout 900 python -m pytest tests -m gpu -x -q) > $O/s4_pytest.log 2>&1; tail -4 $O/s4_pytest.log
(time timeout 500 python bench.py --steps 5 --warmup 3) > $O/s4_bench.log 2> $O/s4_bench.err; tail -c 3000 $O/s4_bench.log; tail -3 $O/s4_bench.err
cd tests
MMMFE_TRACE_ALL=../$O/s4_trace.npz timeout 300 python gpu_perf_probe.py 1024 256 512 128 256 64 3 > ../$O/s4_probe.log 2>&1
grep -E "^apply|sweep kernel:|sweep batches|cycles/CTA|rror" ../$O/s4_probe.log
python gpu_trace_report.py ../$O/s4_trace.npz > ../$O/s4_trace.txt 2>&1; tail -12 ../$O/s4_trace.txt
bash gpu_tune.sh conc p4 k6 c2
cd ..
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file $O/s4_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/s4_ncu_launch.log 2>&1; tail -2 $O/s4_ncu_launch.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep -c 2 -o $O/s4_ksweep -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/s4_ncu_full.log 2>&1; tail -3 $O/s4_ncu_full.log | cut -c1-300
ls -la $O | tail -20
