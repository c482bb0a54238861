#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_detail.py -q 2>&1 | tail -n 25
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "unusual" 2>&1 | tail -n 12
PYTHONPATH=/root/repo:/root/repo/tests python - <<'P'
import numpy as np, jpeg_cases as JC
from oracle_util import Oracle
from jpegsnoop_b200 import CimgDecode
orc=Oracle("ref_fixed"); dec=CimgDecode()
for name,j in JC.mini_cases()[5:]:
    want=orc.decode(j); got=dec.decode(j); bad=JC.compare(want,got)
    print(name, bad)
    for k in ("pix_cb","pix_cr"):
        a=getattr(want,k); b=getattr(got,k)
        if a is not None and not np.array_equal(a,b):
            ys,xs=np.nonzero(a!=b); print("   ",k,len(ys),"cols mod mcu", sorted(set((xs%int(want.geom[0])).tolist()))[:40], "rows mod", sorted(set((ys%int(want.geom[1])).tolist()))[:40])
P
