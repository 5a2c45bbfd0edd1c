#!/bin/bash
# re-tuning of the plan parameters after the signal pruning (configs[1] shapes, one operator application each)
cd "$(dirname "$0")"
bash gpu_tune.sh base x:k6:SWEEP_CHUNK=6144 x:k8:SWEEP_CHUNK=8192 x:k5:SWEEP_CHUNK=5120 x:sp0:SWEEP_SPIN_NS=0 x:p8:SWEEP_PACK=8 x:noconc:SWEEP_CONCURRENT=0 x:e09:SWEEP_SHARE_EXPONENT=0.9 x:e11:SWEEP_SHARE_EXPONENT=1.1 x:m1792:SWEEP_MICRO_TILE=1792 x:m448:SWEEP_MICRO_TILE=448 2>&1 | grep -E "apply|sweep kernel" | cut -c1-200
