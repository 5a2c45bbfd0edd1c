#!/bin/bash
# final committed state: full GPU suite, smoke, bench N = 1
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
O=gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -x --durations=3) > $O/c26_pytest.log 2>&1; tail -6 $O/c26_pytest.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
(time timeout 900 python bench.py --steps 5 --warmup 3) > $O/c26_bench.log 2> $O/c26_bench.err; grep '^{"metric"' $O/c26_bench.log | cut -c1-400; tail -3 $O/c26_bench.err
