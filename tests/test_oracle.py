"""CPU: the oracle is pinned.  (1) the plain-C port reproduces every golden fixture that the compiled
reference produced (tests/golden/make_golden.py); (2) where the compiled reference is available
(this container; prebuilt .so on the GPU box) port == reference on fresh inputs, all outputs."""
import glob
import os
import numpy as np
import pytest

import jpeg_cases as JC
from oracle_util import Oracle, ref_available

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
GOLD = [g for g in GOLD if not g.endswith("idct_tables.npz")]


def test_port_idct_tables_match_golden(built):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "idct_tables.npz"))
    lf, li = Oracle("port").idct_tables()
    assert np.array_equal(li, g["li"])
    assert np.array_equal(lf.view(np.uint32), g["lf"].view(np.uint32))      # bit pattern, not value


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
@pytest.mark.parametrize("fixed", [True, False], ids=["fixed", "float"])
def test_port_matches_golden(built, path, fixed):
    g = np.load(path)
    d = Oracle("port", idct_fixed=fixed).decode(g["jpeg"].tobytes())
    tag = "fixed" if fixed else "float"
    assert d.nerr == 0
    assert np.array_equal(d.geom, g["geom"])
    assert np.array_equal(d.pix_y, g[f"pix_y_{tag}"])
    assert np.array_equal(d.dib, g[f"dib_{tag}"])
    if f"pix_cb_{tag}" in g:
        assert np.array_equal(d.pix_cb, g[f"pix_cb_{tag}"]) and np.array_equal(d.pix_cr, g[f"pix_cr_{tag}"])
    assert np.array_equal(d.mcu_map, g["mcu_map"])
    assert np.array_equal(d.dht_histo, g["dht_histo"])
    assert np.array_equal(d.stats, g[f"stats_{tag}"])
    for c in range(3):
        if f"blk_dc{c}" in g:
            assert np.array_equal(d.blk_dc[c], g[f"blk_dc{c}"])


@pytest.mark.skipif(not ref_available("fixed"), reason="compiled reference not present")
@pytest.mark.parametrize("fixed", [True, False], ids=["fixed", "float"])
def test_port_matches_compiled_reference(built, fixed):
    ref = Oracle("ref_fixed" if fixed else "ref_float"); port = Oracle("port", idct_fixed=fixed)
    for name, j in JC.small_cases()[:7] + JC.mini_cases():
        a, b = ref.decode(j), port.decode(j)
        assert a.nerr == 0 and b.nerr == 0
        assert JC.compare(a, b) == [], name
        assert np.array_equal(a.mcu_map, b.mcu_map) and np.array_equal(a.stats, b.stats), name


def test_synth_generator_is_deterministic_and_decodable(built):
    from jpegsnoop_b200 import synth
    a = synth.encode(160, 96, "420", 85, 4, False, seed=7)
    b = synth.encode(160, 96, "420", 85, 4, False, seed=7)
    c = synth.encode(160, 96, "420", 85, 4, False, seed=8)
    assert a == b and a != c
    d = Oracle("port").decode(a)
    assert d.nerr == 0 and tuple(d.geom[6:8]) == (160, 96)
    assert int(d.stats[10]) == (10 * 6 + 3) // 4 - 1        # RST markers read = intervals - 1
    # batch form equals one-by-one form
    specs = [dict(width=64, height=32, subsampling=s, quality=80, restart_interval=2, optimize=o, seed=i)
             for i, (s, o) in enumerate([("444", False), ("422", True), ("gray", False)])]
    buf, offs = synth.encode_batch(specs, threads=2)
    for i, sp in enumerate(specs):
        one = synth.encode(sp["width"], sp["height"], sp["subsampling"], sp["quality"], sp["restart_interval"], sp["optimize"], sp["seed"])
        assert buf[int(offs[i]):int(offs[i + 1])].tobytes() == one
