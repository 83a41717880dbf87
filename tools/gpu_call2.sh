#!/bin/bash
O=gpurun_out
set -x
CPB_LIB=$PWD/csvplus_b200/_var/lean_c2.so timeout 300 python tools/time_parse.py 40000000 > $O/c2_time_c2.log 2>&1
for m in filter orders; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:csv_scan_lean -s 2 -c 1 -o /tmp/lean_$m -f python tools/prof_parse.py 20000000 $m > /dev/null 2>&1
  python tools/ncu_lines.py /tmp/lean_$m.ncu-rep 100 > $O/c2_lines_$m.txt 2>&1
  ncu -i /tmp/lean_$m.ncu-rep --page details --csv > $O/c2_details_$m.csv 2>&1
  ncu -i /tmp/lean_$m.ncu-rep --page raw --csv > $O/c2_raw_$m.csv 2>&1
done
CPB_SCAN=general timeout 300 ncu --set full --clock-control none -k regex:csv_scan -s 2 -c 1 -o /tmp/gen_orders -f python tools/prof_parse.py 20000000 orders > /dev/null 2>&1
ncu -i /tmp/gen_orders.ncu-rep --page raw --csv > $O/c2_raw_gen_orders.csv 2>&1
grep "GB/s" $O/c2_time_c2.log; du -sh $O
