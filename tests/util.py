"""Shared helpers for the test-suite (test infrastructure; may import oracle/)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle import bindings as ob  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def lcg_plane(seed, h, w, lo, hi):
    """32-bit LCG s = s*1664525 + 1013904223, vectorised by jumping ahead."""
    n = h * w
    out = np.empty(n, dtype=np.uint32)
    s = np.uint64(seed & 0xFFFFFFFF)
    a, c, m = 1664525, 1013904223, 1 << 32
    # sequential generation in chunks (python ints) is slow for big planes: use the closed form via numpy powers
    idx = np.arange(1, n + 1, dtype=np.uint64)
    # compute a^k and (a^k - 1)/(a - 1) * c mod 2^32 with repeated squaring over bits
    ak = np.ones(n, dtype=np.uint64)
    ck = np.zeros(n, dtype=np.uint64)
    base_a, base_c = np.uint64(a), np.uint64(c)
    mask = np.uint64(m - 1)
    k = idx.copy()
    while k.any():
        bit = (k & np.uint64(1)).astype(bool)
        ck = np.where(bit, (ck * base_a + base_c) & mask, ck)
        ak = np.where(bit, (ak * base_a) & mask, ak)
        base_c = (base_c * (base_a + np.uint64(1))) & mask
        base_a = (base_a * base_a) & mask
        k >>= np.uint64(1)
    out = ((ak * s + ck) & mask).astype(np.uint32)
    vals = lo + ((out >> np.uint32(8)) % np.uint32(hi - lo + 1))
    return vals.reshape(h, w).astype(np.uint16)


def random_ycbcr(seed, w, h, chroma, bpp, alpha=False, full_range_values=True):
    sh = 1 if chroma in (1, 2) else 0
    sv = 1 if chroma == 1 else 0
    maxv = (1 << bpp) - 1
    y = lcg_plane(seed, h, w, 0, maxv)
    cb = cr = None
    if chroma:
        cw, ch = (w + sh) >> sh, (h + sv) >> sv
        cb = lcg_plane(seed + 1, ch, cw, 0, maxv)
        cr = lcg_plane(seed + 2, ch, cw, 0, maxv)
    a = lcg_plane(seed + 3, h, w, 0, maxv) if alpha else None
    return y, cb, cr, a


from oracle.bindings import oracle_postprocess, ref_plugin, ref_postprocess  # noqa: E402,F401
