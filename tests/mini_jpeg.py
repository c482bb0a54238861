"""TEST INFRASTRUCTURE: a minimal baseline JPEG writer (numpy), independent of the reference and of
the CUDA path, for inputs Pillow cannot produce: hand-chosen Huffman code-length distributions
(BITS/HUFFVAL), arbitrary sampling factors (e.g. 4:1:1, 4:4:0) and restart intervals.  Not fast, not
pretty: float FDCT, T.81 Annex-K quantisation tables scaled libjpeg-style."""
import numpy as np

ZZ = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
               35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63])
QL = np.array([16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
               18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99])
QC = np.array([17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99] + [99] * 32)


def scaled_q(base, quality):
    s = 5000 // quality if quality < 50 else 200 - 2 * quality
    return np.clip((base * s + 50) // 100, 1, 255).astype(np.int64)


def all_ac_symbols():
    return [0x00, 0xF0] + [(r << 4) | s for r in range(16) for s in range(1, 11)]


def bits_from_lengths(lengths):
    """BITS[1..16] from a list of code lengths (one per symbol, in HUFFVAL order, non-decreasing)."""
    assert all(1 <= l <= 16 for l in lengths) and list(lengths) == sorted(lengths)
    kraft = sum(2.0 ** -l for l in lengths)
    assert kraft <= 1.0 - 2.0 ** -16 + 1e-12, f"Kraft sum {kraft} leaves no room for the reserved all-ones code"
    return [sum(1 for l in lengths if l == i) for i in range(1, 17)]


def canonical_codes(bits, vals):
    """symbol -> (code, length), T.81 Annex C."""
    out, code, k = {}, 0, 0
    for length in range(1, 17):
        for _ in range(bits[length - 1]):
            out[vals[k]] = (code, length); code += 1; k += 1
        code <<= 1
    return out


def long_code_table(symbols, short=(2, 2, 3, 4, 4, 5, 5, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7, 8, 8, 8, 8), n11=24):
    """A legal but unusual length distribution: a few short codes, then n11 codes of 11 bits spread
    over many 10-bit prefixes, the rest at 16 bits.  Forces a large second-level look-up table."""
    n = len(symbols)
    lengths = list(short[:min(len(short), n)])
    lengths += [11] * min(n11, n - len(lengths))
    lengths += [16] * (n - len(lengths))
    lengths = sorted(lengths)
    return bits_from_lengths(lengths), list(symbols)


class BitWriter:
    def __init__(self):
        self.out = bytearray(); self.acc = 0; self.n = 0

    def put(self, code, length):
        self.acc = (self.acc << length) | (code & ((1 << length) - 1)); self.n += length
        while self.n >= 8:
            b = (self.acc >> (self.n - 8)) & 0xFF
            self.out.append(b)
            if b == 0xFF:
                self.out.append(0)
            self.n -= 8
        self.acc &= (1 << self.n) - 1 if self.n else 0

    def flush(self):
        if self.n:
            self.put((1 << (8 - self.n)) - 1, 8 - self.n)


def _fdct_blocks(plane):
    """plane: (H8, W8) float, level-shifted; returns (H8/8, W8/8, 8, 8) DCT-II coefficients."""
    n = np.arange(8)
    C = np.cos((2 * n[None, :] + 1) * n[:, None] * np.pi / 16) * np.where(n[:, None] == 0, np.sqrt(0.5), 1.0) * 0.5   # [u][x]
    H, W = plane.shape
    blk = plane.reshape(H // 8, 8, W // 8, 8).transpose(0, 2, 1, 3)
    return np.einsum("vy,abyx,ux->abvu", C, blk, C)


def _seg(marker, payload):
    return bytes([0xFF, marker]) + (len(payload) + 2).to_bytes(2, "big") + bytes(payload)


def encode(rgb_or_gray, quality=85, samp=((1, 1), (1, 1), (1, 1)), dri=0, dc_tabs=None, ac_tabs=None):
    """Baseline JPEG.  samp: (H,V) per component.  dc_tabs / ac_tabs: [(bits, vals)] for table ids 0 (luma)
    and 1 (chroma); default = canonical tables with a generic length profile."""
    img = np.asarray(rgb_or_gray)
    gray = img.ndim == 2
    H, W = img.shape[:2]
    if gray:
        comps = [img.astype(np.float64)]; samp = ((1, 1),)
    else:
        r, g, b = [img[..., i].astype(np.float64) for i in range(3)]
        comps = [0.299 * r + 0.587 * g + 0.114 * b, -0.168736 * r - 0.331264 * g + 0.5 * b + 128, 0.5 * r - 0.418688 * g - 0.081312 * b + 128]
    hmax = max(h for h, v in samp); vmax = max(v for h, v in samp)
    mw, mh = 8 * hmax, 8 * vmax
    mx, my = (W + mw - 1) // mw, (H + mh - 1) // mh
    qt = [scaled_q(QL, quality), scaled_q(QC, quality)]
    dc_syms = list(range(12)); ac_syms = all_ac_symbols()
    if dc_tabs is None:
        dc_tabs = [(bits_from_lengths([2, 3, 3, 3, 3, 3, 4, 5, 6, 7, 8, 9]), dc_syms)] * 2
    if ac_tabs is None:
        ac_tabs = [long_code_table(ac_syms, n11=0)] * 2
    dcc = [canonical_codes(*t) for t in dc_tabs]; acc = [canonical_codes(*t) for t in ac_tabs]
    coefs = []
    for ci, p in enumerate(comps):
        h, v = samp[ci]
        Wp, Hp = mx * mw, my * mh
        pad = np.pad(p, ((0, Hp - H), (0, Wp - W)), mode="edge")
        fx, fy = hmax // h, vmax // v
        if fx > 1 or fy > 1:
            pad = pad.reshape(Hp // fy, fy, Wp // fx, fx).mean(axis=(1, 3))
        d = _fdct_blocks(pad - 128.0)
        q = qt[0 if ci == 0 else 1].reshape(8, 8)
        coefs.append(np.rint(d / q).astype(np.int64))
    bw = BitWriter(); pred = [0] * len(comps); out = bytearray(); rst = 0

    def put_block(ci, blk):
        t = 0 if ci == 0 else 1
        z = blk.reshape(64)[ZZ]
        diff = int(z[0]) - pred[ci]; pred[ci] = int(z[0])
        s = 0 if diff == 0 else int(abs(diff)).bit_length()
        bw.put(*dcc[t][s])
        if s:
            bw.put(diff if diff > 0 else diff + (1 << s) - 1, s)
        run = 0
        last = int(np.max(np.nonzero(z)[0])) if np.any(z[1:]) else 0
        for k in range(1, last + 1):
            a = int(z[k])
            if a == 0:
                run += 1; continue
            while run > 15:
                bw.put(*acc[t][0xF0]); run -= 16
            s = abs(a).bit_length()
            bw.put(*acc[t][(run << 4) | s]); bw.put(a if a > 0 else a + (1 << s) - 1, s); run = 0
        if last < 63:
            bw.put(*acc[t][0x00])
    n = 0
    for yy in range(my):
        for xx in range(mx):
            if dri and n and n % dri == 0:
                bw.flush(); out += bw.out; out += bytes([0xFF, 0xD0 + (rst & 7)]); rst += 1
                bw = BitWriter(); pred = [0] * len(comps)
            for ci in range(len(comps)):
                h, v = samp[ci]
                for by in range(v):
                    for bx in range(h):
                        put_block(ci, coefs[ci][yy * v + by, xx * h + bx])
            n += 1
    bw.flush(); out += bw.out
    f = bytearray(b"\xFF\xD8")
    for i in range(1 if gray else 2):
        f += _seg(0xDB, bytes([i]) + bytes(int(x) for x in qt[i][ZZ]))
    f += _seg(0xC0, bytes([8]) + H.to_bytes(2, "big") + W.to_bytes(2, "big") + bytes([len(comps)]) +
              b"".join(bytes([ci + 1, (samp[ci][0] << 4) | samp[ci][1], 0 if ci == 0 else 1]) for ci in range(len(comps))))
    for i in range(1 if gray else 2):
        f += _seg(0xC4, bytes([0x00 | i]) + bytes(dc_tabs[i][0]) + bytes(dc_tabs[i][1]))
        f += _seg(0xC4, bytes([0x10 | i]) + bytes(ac_tabs[i][0]) + bytes(ac_tabs[i][1]))
    if dri:
        f += _seg(0xDD, dri.to_bytes(2, "big"))
    f += _seg(0xDA, bytes([len(comps)]) + b"".join(bytes([ci + 1, 0x00 if ci == 0 else 0x11]) for ci in range(len(comps))) + bytes([0, 63, 0]))
    return bytes(f + out + b"\xFF\xD9")
