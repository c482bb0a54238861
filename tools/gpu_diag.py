#!/usr/bin/env python
"""GPU-box diagnostics (not product code): decode test cases through the C-ABI, compare every output with the CPU
oracle and print WHERE they differ (first index, count), decoder status words, self-synchronising-pass statistics and
stage times.  Usage: python tools/gpu_diag.py [long|small|mini] [huff_kernel]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import jpeg_cases as JC
from oracle_util import Oracle, ref_available
from jpegsnoop_b200 import BatchDecoder


def diff(name, a, b):
    if a is None and b is None:
        return None
    a = np.asarray(a).ravel(); b = np.asarray(b).ravel()
    if a.shape != b.shape:
        return f"{name}: shape {a.shape} vs {b.shape}"
    bad = np.flatnonzero(a != b)
    if bad.size == 0:
        return None
    i = int(bad[0])
    return f"{name}: {bad.size} of {a.size} differ, first at {i} (want {a[i]}, got {b[i]}), last at {int(bad[-1])}"


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "long"
    huff = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    cases = {"long": JC.long_cases, "small": JC.small_cases, "mini": JC.mini_cases}[which]()
    orc = Oracle("ref_fixed") if ref_available("fixed") else Oracle("port", idct_fixed=True)
    for name, j in cases:
        bd = BatchDecoder(huff_kernel=huff, idct_kernel=0)
        bd.set_batch([j]); bd.decode(); bd.sync()
        got = bd.fetch(0); want = orc.decode(j)
        msgs = [diff("pix_y", want.pix_y, got.pix_y), diff("pix_cb", want.pix_cb, got.pix_cb), diff("pix_cr", want.pix_cr, got.pix_cr),
                diff("dib", want.dib, got.dib), diff("mcu_map", want.mcu_map, got.mcu_map), diff("histo", want.dht_histo, got.dht_histo)]
        for c in range(3):
            msgs.append(diff(f"blk_dc{c}", want.blk_dc[c], got.blk_dc[c]))
        msgs = [m for m in msgs if m]
        print(f"{name}: status={got.status:#x} selfsync={bd.selfsync_info()} stage_ms={np.round(bd.stage_ms(), 3).tolist()} launches={bd.launches()} "
              f"{'OK' if not msgs else 'MISMATCH'}")
        for m in msgs:
            print("    " + m)
        bd.close()


if __name__ == "__main__":
    main()
