#!/bin/bash
# 2-GPU job: multi-rank parity check against the oracle, then the weak-scaling bench at N = 2 (and N = 1 beside it)
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L
cd tests
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 gpu_multirank_check.py ../$O/s4_multirank.json) > ../$O/s4_multirank.log 2>&1
echo "multirank rc=$?"; grep -E "^(default|checker|nonmatching)|rror|Traceback" ../$O/s4_multirank.log | cut -c1-300
cd ..
if [ "$1" = "bench" ]; then
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 5 --warmup 3) > $O/s4_bench_n2.log 2> $O/s4_bench_n2.err
echo "bench n2 rc=$?"; tail -c 1500 $O/s4_bench_n2.log; tail -5 $O/s4_bench_n2.err | cut -c1-300
fi
(time timeout 600 python -m pytest tests/test_darcyvt_gpu.py -q -k two_gpus) 2>&1 | tail -5
