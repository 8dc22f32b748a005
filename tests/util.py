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


_plugin = None


def ref_plugin():
    """liboracle_plugin.so (+ libheif_ref.so): the unmodified reference. None if not built (e.g. reference absent)."""
    global _plugin
    if _plugin is None:
        p = os.path.join(ob.REF, "liboracle_plugin.so")
        if not os.path.exists(p) or not os.path.exists(os.path.join(ob.REF, "libheif_ref.so")):
            return None
        ob.lib()
        # RTLD_LOCAL on purpose: libheif_ref.so exports thousands of C++ symbols that must not interpose on torch
        _plugin = C.CDLL(p)
    return _plugin


def _p16(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.uint16).ctypes.data_as(C.POINTER(C.c_uint16))


def ref_postprocess(y, cb, cr, a, chroma, bpp, nclx, ops, out_chroma, only_preferred=0, upsampling=2, hdr_to_8bit=0):
    """Run the UNMODIFIED reference: HeifPixelImage transforms + convert_colorspace. nclx = (cp, tc, mc, full) or None."""
    pl = ref_plugin()
    h, w = y.shape
    dt = np.uint8 if bpp == 8 else np.uint16
    arrs = [None if p is None else np.ascontiguousarray(p.astype(dt)) for p in (y, cb, cr, a)]
    ptr = [None if p is None else p.ctypes.data_as(C.c_void_p) for p in arrs]
    ops_a = (C.c_int * (5 * max(1, len(ops))))(*[v for o in ops for v in (list(o) + [0] * 5)[:5]])
    cap = (max(w, h) + 64) ** 2 * 8 * 2
    out = np.empty(cap, dtype=np.uint8)
    ow, oh, rb, npl = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    has = 0 if nclx is None else 1
    cp, tc, mc, fr = nclx if nclx else (2, 2, 2, 0)
    colorspace = 1  # heif_colorspace_RGB
    rc = pl.ref_postprocess(ptr[0], ptr[1], ptr[2], ptr[3], w, h, chroma, bpp, has, cp, tc, mc, int(fr), ops_a, len(ops),
                            colorspace, out_chroma, only_preferred, upsampling, hdr_to_8bit,
                            out.ctypes.data_as(C.c_void_p), C.c_size_t(cap), C.byref(ow), C.byref(oh), C.byref(rb), C.byref(npl))
    if rc != 0:
        raise RuntimeError(f"ref_postprocess rc={rc}")
    n = rb.value * oh.value * npl.value
    return out[:n].copy(), ow.value, oh.value, npl.value


def oracle_postprocess(y, cb, cr, a, chroma, bpp, nclx, ops, out_chroma):
    """C restatement (oracle/color_oracle.c)."""
    l = ob.lib()
    h, w = y.shape
    cp, tc, mc, fr = nclx if nclx else (2, 2, 2, 1)   # image without nclx: defaults + full range (yuv2rgb.cc:203-215)
    ops_a = (C.c_int * (5 * max(1, len(ops))))(*[v for o in ops for v in (list(o) + [0] * 5)[:5]])
    cap = (max(w, h) + 64) ** 2 * 8
    out = np.empty(cap, dtype=np.uint8)
    ow, oh = C.c_int(), C.c_int()
    l.co_postprocess.restype = C.c_long
    keep = [np.ascontiguousarray(p, dtype=np.uint16) if p is not None else None for p in (y, cb, cr, a)]
    n = l.co_postprocess(*[None if k is None else k.ctypes.data_as(C.c_void_p) for k in keep], w, h, chroma, bpp, cp, mc, int(fr),
                         ops_a, len(ops), out_chroma, out.ctypes.data_as(C.c_void_p), C.byref(ow), C.byref(oh))
    if n < 0:
        raise RuntimeError("co_postprocess failed")
    return out[:n].copy(), ow.value, oh.value
