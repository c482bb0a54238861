"""Round-2 groundwork (NOT product code): how fast does a baseline-JPEG Huffman decoder that starts at an arbitrary bit
offset of a scan WITHOUT restart markers fall into step with the true decode?  "In step" = same bit position at a symbol
start, same block-in-MCU index and same zig-zag position — the state a self-synchronising parallel decoder
has to agree on at sub-sequence boundaries.  Prints the distribution of bits/symbols needed, per start assumption."""
import sys, os, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from jpegsnoop_b200 import synth
from jpegsnoop_b200.host import parse_jpeg


def tables_from(t, cls, tid):
    n = t.dht_size[cls][tid]
    tab = {}
    for i in range(n):
        l = t.dht_len[cls][tid][i]
        if 1 <= l <= 16:
            tab[(l, t.dht_bits[cls][tid][i] >> (32 - l))] = t.dht_code[cls][tid][i]
    return tab


def main(w=1920, h=1080, nstarts=300, seed=1):
    j = synth.encode(w, h, "420", 85, 0, False, seed=seed)          # no DRI
    t, d, start = parse_jpeg(j)
    raw = bytes(j)[start:]
    out = bytearray(); i = 0
    while i < len(raw):                                           # unstuff up to the first real marker
        b = raw[i]
        if b == 0xFF:
            if i + 1 < len(raw) and raw[i + 1] == 0x00: out.append(0xFF); i += 2; continue
            if i + 1 < len(raw) and raw[i + 1] != 0xFF: break
        out.append(b); i += 1
    bits = int.from_bytes(bytes(out), "big"); nbits = len(out) * 8
    def peek(pos, n): return (bits >> (nbits - pos - n)) & ((1 << n) - 1) if pos + n <= nbits else None
    comps = []                                                    # block-in-MCU -> (dc table, ac table)
    for c in range(d.num_sos_comps):
        for _ in range(d.samp_h[c] * d.samp_v[c]):
            comps.append((tables_from(t, 0, d.dht_dc_sel[c]), tables_from(t, 1, d.dht_ac_sel[c])))
    bpm = len(comps)

    def step(pos, blk, zz):
        """decode one symbol at bit `pos`; returns new (pos, blk, zz) or None when no code matches / data ends"""
        tab = comps[blk][0 if zz == 0 else 1]
        for l in range(1, 17):
            v = peek(pos, l)
            if v is None: return None
            s = tab.get((l, v))
            if s is not None:
                pos += l + (s & 15)
                if zz == 0: zz = 1
                elif s == 0: zz = 64                               # EOB
                else: zz += (s >> 4) + 1
                if zz >= 64: zz = 0; blk = (blk + 1) % bpm
                return pos, blk, zz
        return None

    truth = {}                                                    # bit position of every true symbol start -> (blk, zz)
    pos, blk, zz = 0, 0, 0; nsym = 0
    nmcu = ((w + 15) // 16) * ((h + 15) // 16); blocks_left = nmcu * bpm
    while blocks_left:
        truth[pos] = (blk, zz); r = step(pos, blk, zz); nsym += 1
        if r is None: break
        if r[2] == 0: blocks_left -= 1                            # the symbol closed a block
        pos, blk, zz = r
    end = pos
    print(f"image {w}x{h}: {end} bits, {nsym} symbols, {end / nsym:.2f} bits/symbol, {bpm} blocks/MCU")
    rng = random.Random(7)
    for assume in ("Y-DC", "every block index"):
        need_bits, need_sym, fail = [], [], 0
        for _ in range(nstarts):
            p0 = rng.randrange(0, max(1, end - 200000))
            best = None
            for b0 in ([0] if assume == "Y-DC" else range(bpm)):
                pos, blk, zz = p0, b0, 0; n = 0
                while n < 20000:
                    if truth.get(pos) == (blk, zz): break
                    r = step(pos, blk, zz); n += 1
                    if r is None: n = None; break
                    pos, blk, zz = r
                if n is not None and n < 20000 and (best is None or pos - p0 < best[0]): best = (pos - p0, n)
            if best is None: fail += 1
            else: need_bits.append(best[0]); need_sym.append(best[1])
        a = np.array(need_bits); s = np.array(need_sym)
        print(f"start assumption {assume:18s}: synchronised {len(a)}/{nstarts}; bits to sync: median {np.median(a):.0f}, "
              f"p90 {np.percentile(a, 90):.0f}, p99 {np.percentile(a, 99):.0f}, max {a.max()}; symbols: median {np.median(s):.0f}, p99 {np.percentile(s, 99):.0f}")


if __name__ == "__main__":
    main(*(int(x) for x in sys.argv[1:]))
