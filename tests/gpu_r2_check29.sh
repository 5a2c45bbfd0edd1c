#!/bin/bash
# bench under torchrun at N = 4 (both arms), as the driver launches it
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
O=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
(time timeout 600 $T --master-port 29571 bench.py --gpus 4 --steps 5 --warmup 3) > $O/c29_bench_n4.log 2> $O/c29_bench_n4.err; echo "rc=$?"; grep '^{"metric"' $O/c29_bench_n4.log | cut -c1-300; tail -2 $O/c29_bench_n4.err | cut -c1-200
(time timeout 300 $T --master-port 29572 bench.py --impl reference --gpus 4 --steps 1 --warmup 1) > $O/c29_ref_n4.log 2> $O/c29_ref_n4.err; echo "rc=$?"; cut -c1-300 $O/c29_ref_n4.log | tail -1
