#!/usr/bin/env python
"""Evidence for DESIGN.md §4 (K2): could the IDCT kernel skip coefficients that are zero?  K2 gives every LANE one block,
so a multiply-add can only be skipped when the coefficient is zero in all 32 blocks of the warp (32 horizontally adjacent
blocks of one component row).  This script quantises bench-like content (tests/jpeg_cases.synth_rgb, IJG q85 tables, 4:2:0)
and prints, per coefficient position, the fraction of 32-block groups in which it is non-zero somewhere."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from scipy.fft import dctn
import jpeg_cases as JC

img = JC.synth_rgb(1920, 1088, 3).astype(np.float64)
R, G, B = img[..., 0], img[..., 1], img[..., 2]
Y = 0.299 * R + 0.587 * G + 0.114 * B - 128; Cb = -0.168736 * R - 0.331264 * G + 0.5 * B; Cr = 0.5 * R - 0.418688 * G - 0.081312 * B
sub = lambda c: c.reshape(c.shape[0] // 2, 2, c.shape[1] // 2, 2).mean(axis=(1, 3))
Cb, Cr = sub(Cb), sub(Cr)
ql = np.array([16,11,10,16,24,40,51,61,12,12,14,19,26,58,60,55,14,13,16,24,40,57,69,56,14,17,22,29,51,87,80,62,18,22,37,56,68,109,103,77,24,35,55,64,81,104,113,92,49,64,78,87,103,121,120,101,72,92,95,98,112,100,103,99]).reshape(8, 8)
qc = np.array([17,18,24,47,99,99,99,99,18,21,26,66,99,99,99,99,24,26,56,99,99,99,99,99,47,66,99,99,99,99,99,99] + [99] * 32).reshape(8, 8)
scale = lambda q, Q=85: np.clip((q * (200 - 2 * Q) + 50) // 100, 1, 255)
ql, qc = scale(ql), scale(qc)
def coefs(p, q):
    h, w = p.shape; b = p.reshape(h // 8, 8, w // 8, 8).transpose(0, 2, 1, 3)
    return np.round(dctn(b, axes=(2, 3), norm="ortho") / q).astype(int)
for name, c in (("Y", coefs(Y, ql)), ("Cb", coefs(Cb, qc)), ("Cr", coefs(Cr, qc))):
    by, bx = c.shape[:2]; nz = c.reshape(by, bx, 64) != 0
    ng = bx // 32; g = nz[:, :ng * 32].reshape(by, ng, 32, 64).any(axis=2).reshape(-1, 64).mean(0)
    print(f"{name}: non-zero AC per block {nz[..., 1:].sum(-1).mean():.1f}; coefficient positions non-zero somewhere in a 32-block group: {g[1:].mean():.3f} of 63")
    print(np.round(g.reshape(8, 8), 2))
