#!/bin/bash
cd "$(dirname "$0")"; mkdir -p ../gpurun_out
(timeout 900 python -m pytest . -m gpu -q -x) 2>&1 | tail -3
bash gpu_tune.sh x:base3: x:k8:SWEEP_CHUNK=8192 x:k9:SWEEP_CHUNK=9216 x:m896:SWEEP_MICRO_TILE=896 x:conc3:SWEEP_CONCURRENT=1 x:l6:LEAF_CELLS=6 x:l12:LEAF_CELLS=12 2>&1 | grep -v "cycles/CTA" | cut -c1-200
