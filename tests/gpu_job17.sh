#!/bin/bash
cd "$(dirname "$0")"; mkdir -p ../gpurun_out
bash gpu_tune.sh x:e110:SWEEP_SHARE_EXPONENT=1.10 x:e120:SWEEP_SHARE_EXPONENT=1.20 x:e090:SWEEP_SHARE_EXPONENT=0.90 2>&1 | grep -v "cycles/CTA" | cut -c1-200
