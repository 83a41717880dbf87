#!/bin/bash
# lean-kernel bring-up: parity tests (auto + forced lean), timings of the variants, sanitizer, ncu
O=gpurun_out
set -x
timeout 600 python -m pytest tests/test_gpu_parse.py tests/test_gpu_plain_inputs.py tests/test_zz_gpu_golden.py tests/test_gpu_shards.py tests/test_gpu_plan_ops.py -x -q -m gpu > $O/c1_tests_auto.log 2>&1; echo "rc=$?" >> $O/c1_tests_auto.log
CPB_SCAN=lean timeout 600 python -m pytest tests/test_gpu_parse.py tests/test_gpu_plain_inputs.py tests/test_zz_gpu_golden.py tests/test_gpu_index_join.py -x -q -m gpu > $O/c1_tests_lean.log 2>&1; echo "rc=$?" >> $O/c1_tests_lean.log
timeout 900 python -m pytest tests/test_gpu_scale.py -x -q -m gpu > $O/c1_tests_scale.log 2>&1; echo "rc=$?" >> $O/c1_tests_scale.log
CPB_SCAN_DEBUG=1 timeout 300 python tools/time_parse.py 40000000 > $O/c1_time_default.log 2>&1
CPB_SCAN=general timeout 300 python tools/time_parse.py 40000000 > $O/c1_time_general.log 2>&1
CPB_LIB=$PWD/csvplus_b200/_var/lean_c2.so timeout 300 python tools/time_parse.py 40000000 > $O/c1_time_c2.log 2>&1
CPB_LIB=$PWD/csvplus_b200/_var/lean_dp4a.so timeout 300 python tools/time_parse.py 40000000 > $O/c1_time_dp4a.log 2>&1
CPB_SCAN=lean timeout 300 compute-sanitizer --tool memcheck python tools/prof_parse.py 300000 orders > $O/c1_sanitizer.log 2>&1
for m in filter orders; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:csv_scan_lean -s 2 -c 1 -o $O/c1_lean_$m -f python tools/prof_parse.py 20000000 $m > /dev/null 2>&1
done
tail -n 3 $O/c1_tests_*.log; grep "GB/s" $O/c1_time_*.log
