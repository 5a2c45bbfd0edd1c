#!/bin/bash
# 2-GPU call: multi-rank parity (forced k_sweep cases, shared multiscale basis on the others), cfg4 probe, bench N = 2
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L
cd tests
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 gpu_multirank_check.py ../$O/c7_multirank.json) > ../$O/c7_multirank.log 2>&1
echo "multirank rc=$?"; grep -E "^(default|checker|nonmatching)|rror|Traceback" ../$O/c7_multirank.log | cut -c1-600
cd ..
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tests/gpu_solve_probe.py cfg4) > $O/c7_cfg4_n2.log 2>&1; grep -E "config|rror" $O/c7_cfg4_n2.log | cut -c1-1500
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 5 --warmup 3) > $O/c7_bench_n2.log 2> $O/c7_bench_n2.err
echo "bench n2 rc=$?"; tail -c 1800 $O/c7_bench_n2.log; tail -3 $O/c7_bench_n2.err | cut -c1-300
(time timeout 600 python -m pytest tests/test_engine_gpu.py -q -x -k "full_solve or neumann") 2>&1 | tail -3
