#!/bin/bash
O=gpurun_out
set -x
timeout 900 python bench.py --steps 3 --no-cpu-baseline --index-rows 0 --e2e-sweep 16x4,8x4,8x8,12x6,16x8,24x8,32x8,16x12 > $O/c5_bench_sweep.json 2> $O/c5_bench_sweep.err
grep "e2e sweep" $O/c5_bench_sweep.err
