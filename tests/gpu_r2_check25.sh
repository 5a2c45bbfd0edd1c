#!/bin/bash
# set-up changes (parallel tree build, parallel subtree classes, descriptor arena in the factorisation): tests + profile
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
O=gpurun_out
(time timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_fullsize_gpu.py tests/test_basis_gpu.py -q -x --durations=3) > $O/c25_pytest.log 2>&1; tail -5 $O/c25_pytest.log | cut -c1-250
for i in 1 2; do
(time MMMFE_SETUP_PROFILE=1 MMMFE_PROBE_NOSOLVE=1 timeout 300 python tests/gpu_solve_probe.py cfg2) > $O/c25_cfg2_setup$i.log 2>&1; grep -E "setup profile|config|rror" $O/c25_cfg2_setup$i.log | cut -c1-420
done
