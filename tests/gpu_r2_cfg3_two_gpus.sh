#!/bin/bash
# configs[2] (4x4 checkerboard 512^2x128 / 256^2x64, mortar 128x128x32) on 2 GPUs: star sweeps vs multiscale basis, converged solves
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
O=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
(time timeout 400 $T --master-port 29551 tests/gpu_solve_probe.py cfg3) > $O/c19_cfg3_sweeps.log 2>&1; grep -E "config|rror" $O/c19_cfg3_sweeps.log | cut -c1-900
(time MMMFE_BASIS_PROFILE=1 timeout 400 $T --master-port 29552 tests/gpu_solve_probe.py cfg3 operator=1) > $O/c19_cfg3_basis.log 2>&1; grep -E "config|rror|basis profile" $O/c19_cfg3_basis.log | cut -c1-900
