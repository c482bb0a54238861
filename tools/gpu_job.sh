#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -n 5
timeout 600 python bench.py --config cfg5 --configs cfg4 --no-e2e 2> gpurun_out/bench_c5.err | tail -n 1 > gpurun_out/bench_c5.json
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_c5.json').read())
for c in d['configs']: print(c['name'], c['value'], c['stage_ms'], c['bit_exact'], c['bit_exact_checked_images'])
P
(timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2f_launches_cfg5.csv python bench.py --config cfg5 --configs none --no-e2e --no-cpu --steps 1 --warmup 1 > gpurun_out/ncu_d.log 2>&1)
grep -E "k_ph_sync|k_huff_lane" gpurun_out/r2f_launches_cfg5.csv | awk -F'","' '{print $5, $NF}' | tail -n 8
