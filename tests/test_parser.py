"""CPU: the product's host HEVC front-end (CABAC + syntax -> command stream, libheif_b200/csrc/b200_hevc_parse.cc)
against the C restatement's syntax-level state (oracle/hevc_oracle.c: hevc_oracle_debug_maps), which is itself pinned
on FFmpeg's pixels by test_oracle_hevc.py.  Compared: QpY and filterEdgeFlag maps (8x8), luma/chroma intra modes (4x4),
TU count and an order-independent digest of every coefficient level with its position."""
import ctypes as C

import numpy as np
import pytest

from hevc_cases import all_streams, cpu_extra_streams
from libheif_b200 import _lib
from oracle import bindings as ob


def product_digest(au):
    l = _lib.lib()
    n8 = 2048 * 2048 // 64
    qp8 = np.zeros(n8, np.int8); edge8 = np.zeros(n8, np.uint8); lm = np.zeros(n8 * 4, np.uint8); cm = np.zeros(n8 * 4, np.uint8)
    out5 = (C.c_ulonglong * 5)()
    l.b200_debug_parse.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_ulonglong)]
    _lib.check(l.b200_debug_parse(au, len(au), qp8.ctypes.data, edge8.ctypes.data, lm.ctypes.data, cm.ctypes.data, out5))
    w, h = out5[3], out5[4]
    return dict(hash=out5[0], ncoef=out5[1], ntu=out5[2], w=w, h=h, qp8=qp8[:w * h // 64].copy(), edge8=edge8[:w * h // 64].copy(),
                lm=lm[:w * h // 16].copy(), cm=cm[:w * h // 16].copy())


def oracle_digest(au):
    l = ob.lib()
    n8 = 2048 * 2048 // 64
    qp8 = np.zeros(n8, np.int8); edge8 = np.zeros(n8, np.uint8); lm = np.zeros(n8 * 4, np.uint8); cm = np.zeros(n8 * 4, np.uint8)
    out3 = (C.c_ulonglong * 3)(); dims = (C.c_int * 2)()
    l.hevc_oracle_debug_maps.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_ulonglong), C.POINTER(C.c_int)]
    rc = l.hevc_oracle_debug_maps(au, len(au), qp8.ctypes.data, edge8.ctypes.data, lm.ctypes.data, cm.ctypes.data, out3, dims)
    assert rc == 0
    w, h = dims[0], dims[1]
    return dict(hash=out3[0], ncoef=out3[1], ntu=out3[2], w=w, h=h, qp8=qp8[:w * h // 64].copy(), edge8=edge8[:w * h // 64].copy(),
                lm=lm[:w * h // 16].copy(), cm=cm[:w * h // 16].copy())


@pytest.mark.parametrize("name,au", all_streams() + cpu_extra_streams(), ids=[s[0] for s in all_streams() + cpu_extra_streams()])
def test_front_end_matches_restatement(name, au):
    p, o = product_digest(au), oracle_digest(au)
    assert (p["w"], p["h"]) == (o["w"], o["h"])
    assert p["ntu"] == o["ntu"]
    assert p["ncoef"] == o["ncoef"]
    assert np.array_equal(p["lm"], o["lm"]), "luma intra modes differ"
    assert np.array_equal(p["cm"], o["cm"]), "chroma intra modes differ"
    assert np.array_equal(p["qp8"], o["qp8"]), "QpY map differs"
    assert np.array_equal(p["edge8"], o["edge8"]), "deblocking edge flags differ"
    assert p["hash"] == o["hash"], "coefficient digest differs"


def test_rejects_garbage():
    l = _lib.lib()
    out5 = (C.c_ulonglong * 5)()
    buf = np.zeros(1 << 16, np.uint8)
    rc = l.b200_debug_parse(b"\x00\x00\x00\x05\x40\x01\x0c\x01\xff", 9, buf.ctypes.data, buf.ctypes.data, buf.ctypes.data, buf.ctypes.data, out5)
    assert rc != 0


def test_mutated_streams_never_crash_the_front_end():
    """Bit flips / byte substitutions / deletions in every fixture: the host front-end (the syntax decoder the GPU kernel
    shares) must either decode or fail with an error code.  (The same loop ran 20 000 mutations under ASan + UBSan while
    this round was developed; two findings in the header parser -- ue(v) with 32 leading zeros, an unchecked
    log2_sao_offset_scale -- were fixed.)"""
    import random
    from hevc_cases import cpu_extra_streams
    l = _lib.lib()
    n8 = 2048 * 2048 // 64
    qp8 = np.zeros(n8, np.int8); edge8 = np.zeros(n8, np.uint8); lm = np.zeros(n8 * 4, np.uint8); cm = np.zeros(n8 * 4, np.uint8)
    out5 = (C.c_ulonglong * 5)()
    l.b200_debug_parse.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_ulonglong)]
    streams = [a for _, a in all_streams() + cpu_extra_streams() if len(a) < 100000]
    rng = random.Random(0xB200)
    decoded = 0
    for _ in range(400):
        b = bytearray(rng.choice(streams))
        mode = rng.random()
        for _ in range(rng.choice([1, 1, 2, 4, 16])):
            p = rng.randrange(len(b))
            if mode < 0.7:
                b[p] ^= 1 << rng.randrange(8)
            elif mode < 0.85:
                b[p] = rng.randrange(256)
            else:
                del b[p:p + rng.randrange(1, 8)]
        rc = l.b200_debug_parse(bytes(b), len(b), qp8.ctypes.data, edge8.ctypes.data, lm.ctypes.data, cm.ctypes.data, out5)
        decoded += rc == 0
    assert decoded < 400          # most mutations must be rejected (sanity of the test itself)
