"""SURVEY.md §8f N1 — the reference's marker parser driving the new CimgDecode.

oracle/Makefile's `n1` target compiles the reference's CjfifDecode (source/JfifDecode.cpp, unmodified, in place) twice:
with the reference's CimgDecode (oracle/_ref/libn1_ref.so) and with this repository's host class built against the same
DocLog / WindowBuf / SnoopConfig headers (libn1_new.so, which calls libjsgpu.so).  The same files go through
CjfifDecode::ProcessFile (JfifDecode.cpp:7297) in both; the whole report — EXIF, DQT/DHT/SOF/SOS dumps, the scan-decode
statistics — and every pixel buffer must be identical."""
import ctypes as C
import io
import os
import numpy as np
import pytest

import jpeg_cases as JC

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class N1:
    def __init__(self, which):
        path = os.path.join(ROOT, "oracle", "_ref", f"libn1_{which}.so")
        if not os.path.exists(path):
            pytest.skip(f"{path} not built (needs /root/reference at build time)")
        L = self.L = C.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_DEEPBIND)      # both builds define the same C++ symbols: each binds to its own
        L.n1_create.restype = C.c_void_p
        L.n1_process.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.n1_num_lines.argtypes = [C.c_void_p]; L.n1_num_err_lines.argtypes = [C.c_void_p]
        L.n1_line.restype = C.c_char_p; L.n1_line.argtypes = [C.c_void_p, C.c_int]
        L.n1_image_size.argtypes = [C.c_void_p, C.c_void_p]
        L.n1_pix.restype = C.POINTER(C.c_int16); L.n1_pix.argtypes = [C.c_void_p, C.c_int]
        L.n1_dib.restype = C.POINTER(C.c_uint8); L.n1_dib.argtypes = [C.c_void_p]
        L.n1_preview_ready.argtypes = [C.c_void_p]
        L.n1_file_pos_mcu.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_void_p]
        L.n1_blk_ycc.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_void_p]
        self.c = C.c_void_p(L.n1_create())

    def process(self, jpeg):
        buf = np.frombuffer(jpeg, np.uint8).copy()
        self.L.n1_process(self.c, buf.ctypes.data, buf.size)
        lines = [self.L.n1_line(self.c, i).decode("latin-1") for i in range(self.L.n1_num_lines(self.c))]
        xy = np.zeros(2, np.uint32); self.L.n1_image_size(self.c, xy.ctypes.data)
        W, H = int(xy[0]), int(xy[1])
        out = {"lines": lines, "size": (W, H), "ready": bool(self.L.n1_preview_ready(self.c))}
        if out["ready"] and W and H:
            for k, name in enumerate(("y", "cb", "cr")):
                p = self.L.n1_pix(self.c, k)
                out[name] = np.ctypeslib.as_array(p, shape=(H, W)).copy() if p else None
            p = self.L.n1_dib(self.c)
            out["dib"] = np.ctypeslib.as_array(p, shape=(H, W, 4)).copy() if p else None
            bb = np.zeros(2, np.uint32); self.L.n1_file_pos_mcu(self.c, 1, 1, bb.ctypes.data); out["mcu11"] = tuple(int(v) for v in bb)
            ycc = np.zeros(3, np.int32); self.L.n1_blk_ycc(self.c, 1, 1, ycc.ctypes.data); out["blk11"] = tuple(int(v) for v in ycc)
        return out


def _files():
    from PIL import Image
    def exif_jpeg(img, **kw):
        im = Image.fromarray(img)
        ex = im.getexif(); ex[0x010F] = "TestMake"; ex[0x0110] = "TestModel"; ex[0x0131] = "softw 1.0"
        b = io.BytesIO(); im.save(b, "JPEG", exif=ex.tobytes(), comment=b"hello comment", **kw); return b.getvalue()
    return [("exif_420_q85", exif_jpeg(JC.synth_rgb(320, 240, 77), quality=85, subsampling=2)),
            ("exif_444_dri_rows_opt", exif_jpeg(JC.synth_rgb(200, 120, 78), quality=92, subsampling=0, restart_marker_rows=1, optimize=True)),
            ("gray_dri3", JC.enc(JC.synth_rgb(160, 96, 79)[:, :, 0], quality=70, restart_marker_blocks=3)),
            ("422_norst_800x600", JC.enc(JC.synth_rgb(800, 600, 80), quality=80, subsampling=1))]


def test_reference_parser_runs_on_the_new_class_up_to_the_scan(built):
    """CPU: everything CjfifDecode does through the setters (DQT / SOF / DHT / SOS handling, EXIF ...) is identical with the
    new class; without a GPU its DecodeScanImg refuses loudly (no CPU fallback) where the reference starts decoding."""
    ref, new = N1("ref"), N1("new")
    for name, j in _files():
        a, b = ref.process(j), new.process(j)
        ia = a["lines"].index("*** Decoding SCAN Data ***")
        assert b["lines"][:ia] == a["lines"][:ia], name
        assert a["ready"]
        try:
            import torch
            has_gpu = torch.cuda.is_available()
        except Exception:
            has_gpu = False
        if not has_gpu:
            assert not b["ready"] and any("GPU scan decoder unavailable" in l for l in b["lines"]), name


@pytest.mark.gpu
def test_reference_parser_with_the_new_class_matches_the_all_reference_build(built):
    """GPU: the whole report and every buffer CjfifDecode's clients read (GetPixMapPtrs, GetBitmapPtr, LookupFilePosMcu,
    LookupBlkYCC), all-reference build vs reference parser + new class.  Both builds are the float-IDCT default."""
    ref, new = N1("ref"), N1("new")
    for name, j in _files():
        a, b = ref.process(j), new.process(j)
        assert b["ready"] and a["ready"], name
        diff = [(i, x, y) for i, (x, y) in enumerate(zip(a["lines"], b["lines"])) if x != y]
        assert not diff and len(a["lines"]) == len(b["lines"]), (name, len(a["lines"]), len(b["lines"]), diff[:4])
        assert a["size"] == b["size"] and a["mcu11"] == b["mcu11"] and a["blk11"] == b["blk11"], name
        for k in ("y", "cb", "cr", "dib"):
            assert (a[k] is None) == (b[k] is None), (name, k)
            if a[k] is not None:
                assert np.array_equal(a[k], b[k]), (name, k)
