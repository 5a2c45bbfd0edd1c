#!/bin/bash
cd "$(dirname "$0")"; mkdir -p ../gpurun_out
(timeout 900 python -m pytest test_engine_gpu.py -m gpu -q -x) 2>&1 | tail -3
bash gpu_tune.sh x:direct1: x:direct0:SWEEP_DIRECT=0 2>&1 | cut -c1-900
