#!/bin/bash
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -x --durations=3) > gpurun_out/f2_pytest.log 2>&1; tail -6 gpurun_out/f2_pytest.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
