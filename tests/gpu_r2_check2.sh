#!/bin/bash
# round 2, second GPU call: basis operator tests first (new code), then the whole suite, then timing probes
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L
(time timeout 900 python -m pytest tests/test_basis_gpu.py -x -q --durations=5) > $O/c2_basis.log 2>&1; tail -30 $O/c2_basis.log
(time timeout 1500 python -m pytest tests -m gpu -q --durations=8 --deselect tests/test_basis_gpu.py) > $O/c2_pytest.log 2>&1; tail -25 $O/c2_pytest.log
(time timeout 300 python tests/gpu_solve_probe.py cfg4 operator=1) > $O/c2_solve_cfg4_basis.log 2>&1; tail -3 $O/c2_solve_cfg4_basis.log | cut -c1-1500
(time timeout 300 python tests/gpu_solve_probe.py cfg4 operator=0) > $O/c2_solve_cfg4_sweep.log 2>&1; tail -3 $O/c2_solve_cfg4_sweep.log | cut -c1-1500
