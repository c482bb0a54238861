"""Seeded synthetic baseline-JPEG generator (ctypes face of libjssynth.so, csrc/tools/synth_jpeg.cpp).
Input tool for tests and bench: it makes the files, it is not on the decode path."""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_L = None

SUBS = {"444": 0, "422": 1, "420": 2, "gray": 3}


def _lib():
    global _L
    if _L is None:
        p = os.path.join(HERE, "libjssynth.so")
        if not os.path.exists(p):
            raise RuntimeError(f"{p} missing: run __graft_entry__.build()")
        _L = C.CDLL(p)
        _L.jssynth_encode.restype = C.c_longlong
        _L.jssynth_encode.argtypes = [C.c_int] * 6 + [C.c_ulonglong, C.c_void_p, C.c_ulonglong]
        _L.jssynth_encode_batch.restype = C.c_longlong
        _L.jssynth_encode_batch.argtypes = [C.c_int] + [C.c_void_p] * 7 + [C.c_int, C.c_void_p, C.c_ulonglong, C.c_void_p]
    return _L


def encode(width, height, subsampling="420", quality=85, restart_interval=0, optimize=False, seed=0):
    L = _lib()
    cap = width * height * 3 + 4096
    buf = np.empty(cap, np.uint8)
    n = L.jssynth_encode(width, height, SUBS[subsampling], quality, restart_interval, int(optimize), seed, buf.ctypes.data, cap)
    if n <= 0:
        raise RuntimeError(f"jssynth_encode failed ({n})")
    return buf[:n].tobytes()


def encode_batch(specs, threads=None):
    """specs: list of dicts(width,height,subsampling,quality,restart_interval,optimize,seed).
    Returns (uint8 array with all files back to back, offsets uint64[n+1])."""
    L = _lib()
    n = len(specs)
    if threads is None:
        threads = min(os.cpu_count() or 1, 64)
    arr = lambda k, conv=int: np.array([conv(s[k]) for s in specs], np.int32)
    w, h, q, ri = arr("width"), arr("height"), arr("quality"), arr("restart_interval")
    ss = np.array([SUBS[s["subsampling"]] for s in specs], np.int32)
    op = np.array([int(s.get("optimize", False)) for s in specs], np.int32)
    seed = np.array([s["seed"] for s in specs], np.uint64)
    cap = int(sum(int(a) * int(b) for a, b in zip(w, h)) * 1.2) + 4096 * n
    out = np.empty(cap, np.uint8); offs = np.zeros(n + 1, np.uint64)
    r = L.jssynth_encode_batch(n, w.ctypes.data, h.ctypes.data, ss.ctypes.data, q.ctypes.data, ri.ctypes.data, op.ctypes.data,
                               seed.ctypes.data, threads, out.ctypes.data, cap, offs.ctypes.data)
    if r < 0:
        cap = -r; out = np.empty(cap, np.uint8)
        r = L.jssynth_encode_batch(n, w.ctypes.data, h.ctypes.data, ss.ctypes.data, q.ctypes.data, ri.ctypes.data, op.ctypes.data,
                                   seed.ctypes.data, threads, out.ctypes.data, cap, offs.ctypes.data)
    if r <= 0:
        raise RuntimeError(f"jssynth_encode_batch failed ({r})")
    return out[:r], offs
