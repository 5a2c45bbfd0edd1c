#!/bin/bash
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
O=gpurun_out
(time timeout 900 python bench.py --steps 5 --warmup 3) > $O/c6_bench.log 2> $O/c6_bench.err; tail -c 6000 $O/c6_bench.log; tail -5 $O/c6_bench.err | cut -c1-300
