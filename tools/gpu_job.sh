#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -n 5
timeout 600 python bench.py --no-e2e --configs cfg3shard,cfg4,cfg1 2> gpurun_out/bench_st.err | tail -n 1 > gpurun_out/bench_st.json
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_st.json').read())
for c in d['configs']: print(c['name'], c['value'], c['stage_ms'], c['k2_frac_of_hbm_peak'], c['bit_exact'], c['bit_exact_checked_images'])
P
tail -n 3 gpurun_out/bench_st.err
