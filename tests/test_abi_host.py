"""CPU: the C-ABI library loads and exports every symbol include/*.h declares; host-side logic
(setters, validation, marker walk, table export) behaves like the reference's; and without a CUDA
device every decode entry point FAILS LOUDLY (there is no CPU fallback)."""
import ctypes as C
import os
import re
import numpy as np
import pytest

import jpeg_cases as JC

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(js(?:gpu|img)_[A-Za-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported(built):
    from jpegsnoop_b200 import _lib as B
    L = B.load()
    decl = _declared("jsgpu.h") + _declared("jsimg.h")
    assert len(decl) >= 55
    missing = [s for s in decl if not hasattr(L, s)]
    assert missing == []
    assert sorted(set(B.JSGPU_SYMBOLS + B.JSIMG_SYMBOLS)) == sorted(set(decl)), "bindings list out of sync with the headers"


def test_struct_layouts_match_the_header(built):
    from jpegsnoop_b200 import _lib as B
    assert C.sizeof(B.jsgpu_tables) == 4 * 64 * 2 + 2 * 4 * 4 + 2 * 4 * 260 * 4 + 2 * 2 * 4 * 260
    assert C.sizeof(B.jsgpu_image_desc) == 136 and C.sizeof(B.jsgpu_image_layout) == 72 and C.sizeof(B.jsgpu_options) == 32


def test_ctypes_structs_have_the_compilers_sizes(built, tmp_path):
    """The ctypes mirrors of the larger C-ABI structures against what gcc makes of include/jsgpu.h."""
    import subprocess
    from jpegsnoop_b200 import _lib as B
    names = ["jsgpu_tables", "jsgpu_image_desc", "jsgpu_image_layout", "jsgpu_options", "jsgpu_pools", "jsgpu_host_outputs",
             "jsgpu_scan_errors", "jsgpu_preview", "jsgpu_ycc_warn", "jsgpu_colour_stats"]
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "jsgpu.h"\nint main(void){' +
                   "".join('printf("%s %%zu\\n", sizeof(%s));' % (n, n) for n in names) + "return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for n in names:
        assert C.sizeof(getattr(B, n)) == int(out[n]), (n, C.sizeof(getattr(B, n)), out[n])


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful without a GPU")
def test_no_cpu_fallback_without_gpu(built):
    from jpegsnoop_b200 import BatchDecoder, CimgDecode, JsgpuError
    with pytest.raises(JsgpuError):
        BatchDecoder()
    dec = CimgDecode()
    d = dec.decode(JC.small_cases()[7][1])            # 8x8 image
    assert d.nerr >= 1 and "GPU scan decoder unavailable" in " ".join(dec.log_lines(3))
    assert not dec.IsPreviewReady()
    assert d.dib is None or not d.dib.any()            # nothing was decoded anywhere else


def test_setters_follow_reference_conventions(built):
    from jpegsnoop_b200 import CimgDecode
    d = CimgDecode()
    assert d.SetDqtEntry(0, 5, 2, 17) and d.GetDqtEntry(0, 5) == 17
    assert not d.SetDqtEntry(4, 0, 0, 1)                               # ref :426 range check, no log line
    assert d.num_err_lines() == 0
    assert not d.SetDqtTables(256, 0) and d.num_err_lines() == 1        # ref :507-518
    assert not d.SetDhtTables(0, 0, 0) and not d.SetDhtTables(5, 0, 0) and not d.SetDhtTables(1, 4, 0)   # ref :539
    assert d.SetDhtTables(4, 3, 3)
    assert not d.SetDhtEntry(4, 0, 0, 1, 0, 0, 0) and not d.SetDhtEntry(0, 2, 0, 1, 0, 0, 0) and not d.SetDhtEntry(0, 0, 260, 1, 0, 0, 0)
    assert not d.SetDhtSize(0, 0, 260) and d.SetDhtSize(0, 0, 259)
    assert "out of indexed range" in d.log_lines(3)[0]
    # DecodeScanImg before SetImageDetails: error line, early return (ref :2755-2758)
    d2 = CimgDecode(); d2.set_file(b"\x00" * 16); d2.DecodeScanImg(0, True, True)
    assert d2.log_lines(3) == ["*** ERROR: Decoding image before Image components defined ***"]
    # unsupported component count: warning + return (ref :2764-2769)
    d3 = CimgDecode(); d3.set_file(b"\x00" * 16); d3.SetImageDetails(8, 8, 4, 4, False, 0); d3.DecodeScanImg(0, True, True)
    assert d3.num_err_lines() == 0 and any("Number of SOS components not supported [4]" in s for s in d3.log_lines(2))
    # tables not selected (ref :3047-3054)
    d4 = CimgDecode(); d4.set_file(b"\x00" * 16); d4.SetSofSampFactors(1, 1, 1); d4.SetImageDetails(8, 8, 1, 1, False, 0); d4.DecodeScanImg(0, True, True)
    assert d4.log_lines(3) == ["*** ERROR: Decoding image before DQT Table Selection via JFIF_SOF ***"]


def test_marker_walk_exports_reference_state(built):
    from jpegsnoop_b200 import parse_jpeg
    from oracle_util import Oracle
    name, j = JC.small_cases()[1]                      # 4:2:2, optimised DHT, DRI=5
    t, d, start = parse_jpeg(j)
    assert (d.dim_x, d.dim_y, d.num_sos_comps, d.restart_interval, d.restart_en) == (640, 480, 3, 5, 1)
    assert list(d.samp_h)[:3] == [2, 1, 1] and list(d.samp_v)[:3] == [1, 1, 1]
    assert start == Oracle("port").decode(j).scan_start
    # canonical code order and left-justified bits, as SetDhtEntry receives them (JfifDecode.cpp:3577-3582)
    n = t.dht_size[1][0]
    lens = [t.dht_len[1][0][i] for i in range(n)]
    assert lens == sorted(lens) and 1 <= lens[0] and lens[-1] <= 16
    bits = [t.dht_bits[1][0][i] >> (32 - lens[i]) for i in range(n)]
    for i in range(1, n):
        assert (bits[i] == bits[i - 1] + 1) if lens[i] == lens[i - 1] else bits[i] == (bits[i - 1] + 1) << (lens[i] - lens[i - 1])
    with pytest.raises(ValueError):
        parse_jpeg(b"not a jpeg at all")


def test_idct_tables_host_equals_golden(built):
    from jpegsnoop_b200 import CimgDecode
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "idct_tables.npz"))
    lf, li = CimgDecode().idct_tables()
    assert np.array_equal(li, g["li"]) and np.array_equal(lf.view(np.uint32), g["lf"].view(np.uint32))
