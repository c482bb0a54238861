"""CPU: host models of the two stage-A kernels' arithmetic, independent of a GPU.
* k_unstuff_lane (jsgpu_huff.cu): per-lane stuffed-zero nibble, big-endian PRMT packing, 64-bit shift register, 0xFF pad and
  stuffed-byte list against a byte-by-byte unstuffing (tools/models/unstuff_lane_model.py restates the kernel's expressions).
* k_marker_scan2 (jsgpu_kernels.cu): the look-back combining rule "(count, end-of-scan seen)" is associative and the 32-lane tree
  reduction keeps file order, so any window split gives the sequential answer."""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools", "models"))


def test_unstuff_lane_arithmetic_equals_bytewise_unstuffing():
    import unstuff_lane_model as M
    rnd = random.Random(7)
    for _ in range(6000):
        n = rnd.randint(80, 240)
        buf = [rnd.choice([0, 0xFF, 0xFF, 0, rnd.randint(0, 255), rnd.randint(0, 255)]) for _ in range(n)]
        s0 = rnd.randint(0, 40); length = rnd.choice([0, 1, 2, 3, 4, 5, rnd.randint(0, n - s0)])
        assert M.lane_unstuff(buf, s0, length) == M.ref_unstuff(buf, s0, length), (s0, length)


TERM = 1 << 61
M32 = 0xFFFFFFFF


def _combine(left, right):                     # mc_combine
    if left & TERM:
        return left & (TERM | M32)
    return ((left + right) & M32) | (right & TERM)


def test_marker_lookback_rule_is_associative_and_order_preserving():
    rnd = random.Random(3)
    word = lambda: rnd.randint(0, 9) | (TERM if rnd.random() < 0.15 else 0)
    for _ in range(20000):
        a, b, c = word(), word(), word()
        assert _combine(_combine(a, b), c) == _combine(a, _combine(b, c))
    for _ in range(3000):                       # the shuffle tree of the look-back window: lane `stop` is the leftmost chunk
        ws = [word() for _ in range(32)]; stop = rnd.randint(0, 31)
        acc = [ws[l] if l <= stop else 0 for l in range(32)]
        d = 1
        while d < 32:
            acc = [_combine(acc[l + d], acc[l]) if l + d <= stop else acc[l] for l in range(32)]
            d *= 2
        want = ws[stop]
        for l in range(stop - 1, -1, -1):
            want = _combine(want, ws[l])
        assert acc[0] == want
    # what the rule means: RST markers behind the first terminating marker do not count
    chunks = [(3, False), (2, True), (5, False)]
    tot = 0
    for cnt, term in chunks:
        tot = _combine(tot, cnt | (TERM if term else 0))
    assert tot & M32 == 5 and tot & TERM
