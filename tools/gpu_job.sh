#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -n 5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
timeout 600 python bench.py --configs cfg3shard,cfg4 --no-e2e 2> gpurun_out/bench_fk.err | tail -n 1 > gpurun_out/bench_fk.json
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_fk.json').read())
for c in d['configs']: print(c['name'], c['value'], c['ms_per_step'], c['stage_ms'], c['bit_exact'], c['bit_exact_checked_images'])
P
tail -n 2 gpurun_out/bench_fk.err
