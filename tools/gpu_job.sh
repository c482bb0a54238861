#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -n 4
timeout 600 python bench.py --configs cfg3shard,cfg4 --no-e2e --no-cpu 2> gpurun_out/bench_mk.err | tail -n 1 > gpurun_out/bench_mk.json
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_mk.json').read())
for c in d['configs']: print(c['name'], c['value'], c['ms_per_step'], c['stage_ms'], c['bit_exact'], c['bit_exact_checked_images'])
P
(timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/r2g_launches_cfg2_bench.csv python bench.py --configs none --no-e2e --no-cpu --steps 1 --warmup 0 > gpurun_out/ncu_a.log 2>&1)
awk -F'","' 'NR>2{print $5, $NF}' gpurun_out/r2g_launches_cfg2_bench.csv | sed -n 3,6p
JSGPU_MARKER=0 timeout 300 python bench.py --configs cfg4 --no-e2e --no-cpu 2>/dev/null | tail -n 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); [print('old scan', c['name'], c['value'], c['stage_ms']) for c in d['configs']]"
