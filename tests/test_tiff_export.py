"""Export-to-TIFF consumer: the file writer against the reference's own FileTiff.cpp (compiled in place
into oracle/_ref), and — on the GPU — the device-packed sample arrays + complete files against
CJPEGsnoopDoc::OnToolsExporttiff's loops (JPEGsnoopDoc.cpp:2098-2170, restated in oracle/ref_harness.cpp)."""
import ctypes as C
import os

import numpy as np
import pytest

import jpeg_cases as JC
from oracle_util import Oracle, ref_available

needs_ref = pytest.mark.skipif(not ref_available("fixed"), reason="needs oracle/_ref (the compiled reference)")


@needs_ref
@pytest.mark.parametrize("mode", [(0, 0), (0, 1), (1, 0)], ids=["rgb8", "rgb16", "ycc8"])
@pytest.mark.parametrize("size", [(8, 8), (640, 480), (1920, 1088), (3840, 2160), (70000, 2)], ids=lambda s: f"{s[0]}x{s[1]}")
def test_tiff_file_equals_the_references(built, tmp_path, mode, size):
    """Header, IFD (dimensions and strip offset are SHORTs there: 70000 columns wrap like the reference's), out-of-line values
    and sample data, byte for byte."""
    from jpegsnoop_b200 import _lib
    L = _lib.load()
    ycc, b16 = mode; w, h = size
    rng = np.random.default_rng(w * 31 + h)
    data = rng.integers(0, 256, w * h * (6 if b16 else 3), dtype=np.uint8)
    ref = Oracle("ref_fixed")
    a, b = str(tmp_path / "ref.tif").encode(), str(tmp_path / "new.tif").encode()
    ref.lib.ref_tiff_write.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_uint, C.c_uint]
    ref.lib.ref_tiff_write(a, ycc, b16, data.ctypes.data, w, h)
    assert L.jsimg_tiff_write(b, ycc, b16, data.ctypes.data, w, h) == 1
    want, got = open(a, "rb").read(), open(b, "rb").read()
    assert len(want) > data.size and want == got, (len(want), len(got), want[:64].hex(), got[:64].hex())
    ref.close()


@pytest.mark.gpu
@needs_ref
def test_exported_tiffs_match_the_reference(built, tmp_path):
    from jpegsnoop_b200 import CimgDecode, BatchDecoder
    cases = JC.small_cases()
    ref = Oracle("ref_fixed")
    ref.lib.ref_export_tiff.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    dec = CimgDecode()
    bd = BatchDecoder(); bd.set_batch([j for _, j in cases]); bd.decode(); bd.sync()
    for i, (name, j) in enumerate(cases):
        want = ref.decode(j); got = dec.decode(j)
        assert not JC.compare(want, got), name
        three = got.pix_cb is not None
        for mode in (0, 1, 2):
            a, b = str(tmp_path / f"ref{mode}.tif"), str(tmp_path / f"new{mode}.tif")
            for p in (a, b):
                if os.path.exists(p):
                    os.remove(p)
            ok_ref = ref.lib.ref_export_tiff(ref.ctx, a.encode(), mode)
            ok_new = dec.ExportTiff(b, mode)
            assert bool(ok_ref) == ok_new == (three or mode != 2), (name, mode, ok_ref, ok_new)
            if not ok_new:
                continue
            wb, gb = open(a, "rb").read(), open(b, "rb").read()
            assert wb == gb, (name, mode, len(wb), len(gb))
            # the batch entry point hands out the same sample array
            arr = bd.export(i, mode)
            assert arr.tobytes() == wb[len(wb) - arr.size:], (name, mode)
    # the export follows the preview the DIB shows (the reference reads m_pDibTemp, whatever mode painted it)
    name, j = cases[0]
    ref.decode(j); dec.decode(j)
    ref.set_preview_mode(6); dec.SetPreviewMode(6)
    a, b = str(tmp_path / "ref_y.tif"), str(tmp_path / "new_y.tif")
    ref.lib.ref_export_tiff(ref.ctx, a.encode(), 0); assert dec.ExportTiff(b, 0)
    assert open(a, "rb").read() == open(b, "rb").read()
    ref.set_preview_mode(1); ref.close()
