#!/bin/bash
O=gpurun_out
set -x
BENCH_DEBUG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --index-rows 0 --people 2000000 > $O/c7_bench_n2.json 2> $O/c7_bench_n2.err; echo "rc=$?" >> $O/c7_bench_n2.err
grep "phases\|per-step" $O/c7_bench_n2.err | tail -n 12
