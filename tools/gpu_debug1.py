import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import jpeg_cases as JC
from oracle_util import Oracle
from jpegsnoop_b200 import CimgDecode
cases = dict(JC.small_cases())
orc = Oracle("ref_fixed")
for hk, ik in ((1, 1), (2, 2), (1, 2), (2, 1)):
    dec = CimgDecode(huff_kernel=hk, idct_kernel=ik)
    for name in cases:
        j = cases[name]
        w = orc.decode(j); g = dec.decode(j)
        bad = JC.compare(w, g)
        print(hk, ik, name, "bad", bad, "nerr", g.nerr, "status", g.status, "ms", np.round(g.stage_ms, 3))
        if name != "gray_dri3" and "pix_y" not in bad: continue
        for f in bad:
            if f in ("pix_y", "pix_cb", "pix_cr", "dib"):
                a, b = getattr(w, f), getattr(g, f)
                idx = np.argwhere(a != b)
                print("   ", f, "n mismatch", len(idx), "first", idx[:4].tolist(), [(int(a[tuple(i)]), int(b[tuple(i)])) for i in idx[:4]])
        if "mcu_map" in bad and hk == 1:
            idx = np.nonzero(w.mcu_map != g.mcu_map)[0]
            print(" n mismatch", idx.size, "of", w.mcu_map.size, "first", idx[:12])
            for i in idx[:8]:
                print("   mcu", i, "want %x.%d got %x.%d" % (w.mcu_map[i] >> 4, w.mcu_map[i] & 15, g.mcu_map[i] >> 4, g.mcu_map[i] & 15))
            d = np.frombuffer(j, np.uint8); s = w.scan_start
            rst = [i for i in range(s, len(d) - 1) if d[i] == 0xFF and 0xD0 <= d[i + 1] <= 0xD7]
            print(" scan start", hex(s), "RSTs", [hex(x) for x in rst[:8]])
            for r in rst[:8]:
                print("   bytes before rst", [hex(x) for x in d[r - 5:r + 2]])
