#!/bin/bash
cd "$(dirname "$0")"; mkdir -p ../gpurun_out
(timeout 600 python -m pytest test_engine_gpu.py -q -x -k "packs or persistent") 2>&1 | tail -3
bash gpu_tune.sh x:p8:SWEEP_PACK=8 x:p8k7:SWEEP_PACK=8,SWEEP_CHUNK=7168 x:l8k7:LEAF_CELLS=8,SWEEP_CHUNK=7168 x:p8l8:SWEEP_PACK=8,LEAF_CELLS=8 2>&1 | grep -v "cycles/CTA" | cut -c1-200
grep "cycles/CTA" ../gpurun_out/tune_p8.log | head -1 | cut -c1-700
