#!/bin/bash
# 2 GPUs: multirank parity with the helper-thread NCCL warm-up, cfg4 + cfg3 probes (auto operator)
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
O=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
cd tests
(time timeout 600 $T --master-port 29561 gpu_multirank_check.py ../$O/c20_multirank.json) > ../$O/c20_multirank.log 2>&1
echo "multirank rc=$?"; grep -E "^(default|checker|nonmatching)|rror|Traceback" ../$O/c20_multirank.log | cut -c1-500
cd ..
(time MMMFE_SETUP_PROFILE=1 timeout 400 $T --master-port 29562 tests/gpu_solve_probe.py cfg4) > $O/c20_cfg4.log 2>&1; grep -E "config|rror|set-up:" $O/c20_cfg4.log | cut -c1-900
(time timeout 400 $T --master-port 29563 tests/gpu_solve_probe.py cfg3) > $O/c20_cfg3.log 2>&1; grep -E "config|rror" $O/c20_cfg3.log | cut -c1-900
