"""Round-2 groundwork (NOT product code): a sequential emulation of the self-synchronising parallel Huffman decode
planned for scans without restart markers, checked against a plain sequential decode of the same scan.

  pass A   every sub-sequence i (S bits) is decoded from its first bit with the guess "block 0 of an MCU, DC";
           it runs on past its end until it reaches a symbol start at or after bit (i+1)*S and records that exit
           state  X_i = (bit position, block-in-MCU, zig-zag index).
  pass B   repeated until nothing changes: sub-sequence i is decoded again FROM X_{i-1} (sub-sequence 0 from the true
           start); if its new exit equals the stored X_i it is settled.  Because a decoder locks onto the true symbol
           grid within a few hundred bits, almost every X_i is right after pass A already and the fix-up wave dies fast.
  pass C   with true entry states: count the blocks each sub-sequence completes -> exclusive prefix sum = index of its
           first block -> every sub-sequence writes its (block, zig-zag, value) triples independently; the DC
           predictors are prefix sums of the DC differences per component.
The emulation reports the number of B rounds and verifies the coefficient blocks bit for bit."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from jpegsnoop_b200 import synth
from jpegsnoop_b200.host import parse_jpeg
from selfsync_probe import tables_from


def main(w=640, h=480, S=4096, seed=3):
    j = synth.encode(w, h, "420", 85, 0, False, seed=seed)
    t, d, start = parse_jpeg(j)
    raw = bytes(j)[start:]
    out = bytearray(); i = 0
    while i < len(raw):
        b = raw[i]
        if b == 0xFF:
            if i + 1 < len(raw) and raw[i + 1] == 0x00: out.append(0xFF); i += 2; continue
            if i + 1 < len(raw) and raw[i + 1] != 0xFF: break
        out.append(b); i += 1
    bits = int.from_bytes(bytes(out), "big"); nbits = len(out) * 8
    def peek(pos, n): return (bits >> (nbits - pos - n)) & ((1 << n) - 1) if pos + n <= nbits else None
    comps = []
    for c in range(d.num_sos_comps):
        for _ in range(d.samp_h[c] * d.samp_v[c]):
            comps.append((tables_from(t, 0, d.dht_dc_sel[c]), tables_from(t, 1, d.dht_ac_sel[c]), c))
    bpm = len(comps)
    nblocks = ((w + 15) // 16) * ((h + 15) // 16) * bpm

    def step(pos, blk, zz):
        """one symbol: returns (pos', blk', zz', zig-zag slot written or None, value, block closed?) or None"""
        tab = comps[blk][0 if zz == 0 else 1]
        for l in range(1, 17):
            v = peek(pos, l)
            if v is None: return None
            s = tab.get((l, v))
            if s is None: continue
            size = s & 15; pos += l
            val = 0
            if size:
                x = peek(pos, size)
                if x is None: return None
                val = x if x >> (size - 1) else x - (1 << size) + 1
                pos += size
            slot = None
            if zz == 0: slot, zz = 0, 1
            elif s == 0: zz = 64
            else:
                zz += s >> 4
                if size: slot = zz
                zz += 1
            closed = zz >= 64
            if closed: zz = 0; blk = (blk + 1) % bpm
            return pos, blk, zz, slot, val, closed
        return None

    # ---- ground truth: plain sequential decode --------------------------------------------------------------
    truth = np.zeros((nblocks, 64), np.int32); pos, blk, zz, nb = 0, 0, 0, 0
    while nb < nblocks:
        r = step(pos, blk, zz)
        if r[3] is not None and r[3] < 64: truth[nb, r[3]] = r[4]
        pos, blk, zz = r[0], r[1], r[2]
        if r[5]: nb += 1
    end = pos
    nsub = (end + S - 1) // S

    def run(i, state, emit=None):
        """decode sub-sequence i from `state` to the first symbol start at/after (i+1)*S (or the end of data)"""
        pos, blk, zz = state; lim = min((i + 1) * S, end); closed = 0
        while pos < lim:
            r = step(pos, blk, zz)
            if r is None: return None, closed
            if emit is not None: emit(r[3], r[4], r[5])
            pos, blk, zz = r[0], r[1], r[2]; closed += r[5]
        return (pos, blk, zz), closed

    # ---- pass A ------------------------------------------------------------------------------------------------
    X = [run(i, (i * S, 0, 0))[0] for i in range(nsub)]
    # ---- pass B ------------------------------------------------------------------------------------------------
    rounds = 0; changed = True; redo = 0
    while changed:
        changed = False; rounds += 1; newX = list(X)
        for i in range(nsub):                                       # "in parallel": every i reads the OLD X[i-1]
            entry = (0, 0, 0) if i == 0 else X[i - 1]
            if entry is None: continue
            x, _ = run(i, entry); redo += 1
            if x != X[i]: newX[i] = x; changed = True
        X = newX
    # ---- pass C ------------------------------------------------------------------------------------------------
    counts = []
    for i in range(nsub):
        entry = (0, 0, 0) if i == 0 else X[i - 1]
        counts.append(run(i, entry)[1])
    first = np.concatenate([[0], np.cumsum(counts)[:-1]])
    got = np.zeros((nblocks + 8, 64), np.int32)
    for i in range(nsub):
        entry = (0, 0, 0) if i == 0 else X[i - 1]
        cur = [int(first[i])]
        def emit(slot, val, closed, cur=cur):
            if slot is not None and slot < 64 and cur[0] < nblocks: got[cur[0], slot] = val
            if closed: cur[0] += 1
        run(i, entry, emit)
    ok = np.array_equal(got[:nblocks], truth)
    print(f"{w}x{h} no-DRI: {end} bits, {nsub} sub-sequences of {S} bits; pass B rounds until stable: {rounds} "
          f"({redo} sub-sequence decodes in total, {redo / nsub:.2f} per sub-sequence); blocks {sum(counts)}/{nblocks}; "
          f"coefficients identical to the sequential decode: {ok}")
    return ok


if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:]]
    sys.exit(0 if main(*a) else 1)
