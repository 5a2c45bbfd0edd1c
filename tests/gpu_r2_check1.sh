#!/bin/bash
# round 2, first GPU call: whole gpu test suite (incl. the new full-size parity tests), converged configs[1] solve
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L
(time timeout 1500 python -m pytest tests -m gpu -x -q --durations=8) > $O/c1_pytest.log 2>&1; tail -25 $O/c1_pytest.log
(time timeout 600 python tests/gpu_solve_probe.py cfg2) > $O/c1_solve_cfg2.log 2>&1; tail -3 $O/c1_solve_cfg2.log | cut -c1-1200
