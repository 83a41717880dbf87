#!/bin/bash
O=gpurun_out
set -x
nvidia-smi -L | head -4
timeout 600 python -m pytest tests/test_gpu_scale.py -x -q -m gpu -k "two_rank or gathered" > $O/c6_tests_n2.log 2>&1; echo "rc=$?" >> $O/c6_tests_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > $O/c6_bench_n2.json 2> $O/c6_bench_n2.err; echo "rc=$?" >> $O/c6_bench_n2.err
tail -n 4 $O/c6_tests_n2.log; tail -n 5 $O/c6_bench_n2.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/c6_bench_n2.json').read().strip().splitlines()[-1])
    print('N=2 ms/step', round(d['ms_per_step'],2), 'value', d['value'], 'e2e ms', round(d['e2e']['ms_per_step'],1), 'syncs', d.get('host_syncs_per_step'), 'gap', d.get('host_gap_ms_per_step'))
except Exception as e: print('ERR', e)
PY
