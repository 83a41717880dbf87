#!/bin/bash
# round-2 profile pass (run under gpurun, one GPU): launch list + DRAM traffic of the bench step, ncu --set full of the top kernels
set -x
O=gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --index-rows 0 --people 20000000"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r2_launches_full.csv $B > $O/r2_bench_under_ncu.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:csv_scan -c 12 --csv --log-file $O/r2_traffic_csv_scan.csv $B > /dev/null 2>&1
for m in filter orders; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:csv_scan -s 2 -c 1 -o $O/r2_scan_$m -f python tools/prof_parse.py 20000000 $m > /dev/null 2>&1
done
PROF_ITERS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"slot_copy|join_probe16|hash16_insert|slot_fill|scan_lens" -c 8 -o $O/r2_join -f python tools/prof_join.py 100000000 125000000 join3 > $O/r2_join_under_ncu.log 2>&1
PROF_ITERS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"radix_scatter|radix_count|hist8|gather_copy" -c 6 -o $O/r2_sort -f python tools/prof_join.py 100000000 0 sort > $O/r2_sort_under_ncu.log 2>&1
PROF_ITERS=3 python tools/prof_join.py 100000000 125000000 sort > $O/r2_sort_times.log 2>&1
PROF_ITERS=3 python tools/prof_join.py 100000000 125000000 join3 > $O/r2_join_times.log 2>&1
ls -la $O/*.ncu-rep | tail -5
