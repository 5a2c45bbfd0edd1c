#!/bin/bash
# bar + final sweeps through k_sweep, retained memory pool: full GPU suite, set-up profile, solve probes
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
O=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -x --durations=5) > $O/c16_pytest.log 2>&1; tail -10 $O/c16_pytest.log | cut -c1-250
(time MMMFE_SETUP_PROFILE=1 MMMFE_PROBE_NOSOLVE=1 timeout 300 python tests/gpu_solve_probe.py cfg2) > $O/c16_cfg2_setup.log 2>&1; grep -E "setup profile|config|rror" $O/c16_cfg2_setup.log | cut -c1-700
(time MMMFE_SETUP_PROFILE=1 timeout 300 python tests/gpu_solve_probe.py cfg4) > $O/c16_cfg4.log 2>&1; grep -E "set-up:|host:|config|rror" $O/c16_cfg4.log | cut -c1-700
(time timeout 300 python tests/gpu_solve_probe.py cfg4 sweep_bar=0,sweep_final=0) > $O/c16_cfg4_levels.log 2>&1; grep -E "config|rror" $O/c16_cfg4_levels.log | cut -c1-700
