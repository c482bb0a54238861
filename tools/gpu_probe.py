"""Scratch GPU probe: smoke + small batch timing (not a bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import __graft_entry__ as g
g.smoke()
import jpeg_cases as JC
from jpegsnoop_b200 import BatchDecoder
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
t = time.time()
base = [JC.enc(JC.synth_rgb(1920, 1080, 100 + i), quality=85, subsampling=2, restart_marker_blocks=4) for i in range(4)]
jpegs = [base[i % 4] for i in range(n)]
print("encode s", time.time() - t, "bytes/img", len(base[0]))
for ik in (1,):
    bd = BatchDecoder(idct_kernel=ik)
    t = time.time(); bd.set_batch(jpegs); print("set_batch s", time.time() - t)
    for it in range(3):
        bd.decode(); bd.sync()
        ms = bd.stage_ms()
        mpix = bd.nsof_pixels / 1e6
        print(f"idct_kernel={ik} it={it} stage ms {np.round(ms,3)}  -> {mpix/ (ms[4]/1e3)/1e3:.2f} GPix/s  launches {bd.launches()}")
    st = [l.status for l in bd.refresh_layout()]
    print("status", set(st))
