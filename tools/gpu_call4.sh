#!/bin/bash
O=gpurun_out
set -x
timeout 900 python -m pytest tests/test_gpu_parse.py tests/test_gpu_plain_inputs.py tests/test_zz_gpu_golden.py tests/test_gpu_shards.py -x -q -m gpu > $O/c4_tests.log 2>&1; echo "rc=$?" >> $O/c4_tests.log
timeout 300 python tools/time_parse.py 40000000 > $O/c4_time_default.log 2>&1
CPB_GENERAL_PATH=dfa timeout 300 python tools/time_parse.py 20000000 > $O/c4_time_dfa.log 2>&1
BENCH_DEBUG=1 timeout 600 python bench.py > $O/c4_bench.json 2> $O/c4_bench.err
BENCH_DEBUG=1 timeout 600 python bench.py --steps 3 --no-cpu-baseline --index-rows 0 --e2e-workers 6 --e2e-batches 24 > $O/c4_bench_w6.json 2> $O/c4_bench_w6.err
tail -n 4 $O/c4_tests.log; grep "GB/s" $O/c4_time_*.log; grep "e2e step" $O/c4_bench.err | tail -n 3; grep "e2e step" $O/c4_bench_w6.err | tail -n 3
python - <<'PY'
import json
for f in ['gpurun_out/c4_bench.json','gpurun_out/c4_bench_w6.json']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms/step', round(d['ms_per_step'],2), 'e2e ms', round(d['e2e']['ms_per_step'],1), 'syncs', d.get('host_syncs_per_step'), 'gap', round(d.get('host_gap_ms_per_step',0),2), 'parse GB/s', round(d['csv_parse']['value'],1))
    except Exception as e: print(f, 'ERR', e)
PY
