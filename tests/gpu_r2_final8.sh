#!/bin/bash
# 8-GPU call: multi-rank parity on 8 ranks, then the bench at N = 8 (weak-scaled configs[1] + nested north-star record)
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L | head -8; nproc
cd tests
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 gpu_multirank_check.py ../$O/f8_multirank.json) > ../$O/f8_multirank.log 2>&1
echo "multirank rc=$?"; grep -E "^(default|checker|nonmatching)|rror|Traceback" ../$O/f8_multirank.log | cut -c1-600
cd ..
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 10 --warmup 3) > $O/f8_bench_n8.log 2> $O/f8_bench_n8.err
echo "bench n8 rc=$?"; tail -c 5000 $O/f8_bench_n8.log; tail -3 $O/f8_bench_n8.err | cut -c1-300
