#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -n 8
timeout 1500 python bench.py --no-e2e 2> gpurun_out/bench_un.err | tail -n 1 > gpurun_out/bench_un.json
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_un.json').read())
for c in d['configs']: print(c['name'], c['value'], c['stage_ms'], c['bit_exact'], c['bit_exact_checked_images'])
P
tail -n 3 gpurun_out/bench_un.err
