#!/bin/bash
# weak-scaling bench at N GPUs ($1), one line into gpurun_out/s6_bench_n$1.log
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
N=$1
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus $N --steps 5 --warmup 3 $2) > gpurun_out/s6_bench_n$N.log 2> gpurun_out/s6_bench_n$N.err
echo "rc=$?"; grep '^{"metric"' gpurun_out/s6_bench_n$N.log | cut -c1-330; tail -4 gpurun_out/s6_bench_n$N.err | cut -c1-300
