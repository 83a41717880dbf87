#!/bin/bash
# round-end style validation: the whole GPU suite, smoke(), the default bench line, the reference arm, an ncu launch list
O=gpurun_out
set -x
timeout 1500 python -m pytest tests -x -q -m gpu > $O/f_tests.log 2>&1; echo "rc=$?" >> $O/f_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/f_smoke.log 2>&1; echo "rc=$?" >> $O/f_smoke.log
timeout 900 python bench.py > $O/f_bench_n1.json 2> $O/f_bench_n1.err; echo "rc=$?" >> $O/f_bench_n1.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/f_bench_ref.json 2> $O/f_bench_ref.err; echo "rc=$?" >> $O/f_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/r2b_launches_full.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --index-rows 0 --people 20000000 > $O/f_bench_under_ncu.log 2>&1
tail -n 3 $O/f_tests.log; tail -n 3 $O/f_smoke.log; tail -n 2 $O/f_bench_n1.err; tail -c 600 $O/f_bench_ref.json
