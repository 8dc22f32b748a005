"""Host HEVC-intra encoder wrapper (b200_hevc_encode_intra) and the synthetic source images of SURVEY.md 8(d)."""
import ctypes as C

import numpy as np

from . import _lib


class EncParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "width", "height", "bit_depth", "chroma_format_idc", "log2_ctb_size", "qp", "init_qp",
        "max_transform_hierarchy_depth_intra", "sao", "sign_data_hiding", "transform_skip", "strong_intra_smoothing",
        "cu_qp_delta", "diff_cu_qp_delta_depth", "dqp_range", "cb_qp_offset", "cr_qp_offset", "slice_chroma_qp_offsets",
        "slice_cb_qp_offset", "slice_cr_qp_offset", "wpp", "slice_ctb_rows", "dependent_slice_segments",
        "loop_filter_across_slices", "slice_loop_filter_across_slices", "deblocking_disabled", "beta_offset_div2",
        "tc_offset_div2", "slice_deblocking_override", "slice_deblocking_disabled", "slice_beta_offset_div2",
        "slice_tc_offset_div2", "mode_decision", "split_threshold", "still_picture", "vui_present",
        "colour_description_present", "colour_primaries", "transfer_characteristics", "matrix_coefficients", "full_range")] + \
        [("seed", C.c_uint32), ("scaling_lists", C.c_int), ("pcm", C.c_int), ("transquant_bypass", C.c_int), ("tile_cols", C.c_int), ("tile_rows", C.c_int), ("tiles_uniform", C.c_int),
         ("loop_filter_across_tiles", C.c_int), ("slice_per_tile", C.c_int)]


def default_params(**kw) -> EncParams:
    l = _lib.lib()
    p = EncParams()
    l.b200_hevc_enc_params_default(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, int(v))
    return p


def encode_intra(y, cb=None, cr=None, **kw) -> bytes:
    """Encode one picture; returns the access unit as length-prefixed NALs (what libheif pushes to a decoder plugin)."""
    l = _lib.lib()
    h, w = y.shape
    bd = kw.get("bit_depth", 8)
    dt = np.uint8 if bd == 8 else np.uint16
    y = np.ascontiguousarray(y, dtype=dt)
    chroma = cb is not None
    if chroma:
        cb = np.ascontiguousarray(cb, dtype=dt)
        cr = np.ascontiguousarray(cr, dtype=dt)
    cfmt = 0
    if chroma:                                             # chroma format from the plane shapes (4:2:0 / 4:2:2 / 4:4:4)
        cfmt = 3 if cb.shape == y.shape else (2 if cb.shape[0] == h else 1)
    p = default_params(width=w, height=h, chroma_format_idc=cfmt, **kw)
    out = C.POINTER(C.c_uint8)()
    n = C.c_size_t()
    l.b200_hevc_encode_intra.argtypes = [C.POINTER(EncParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t,
                                         C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.c_size_t)]
    l.b200_free.argtypes = [C.c_void_p]
    _lib.check(l.b200_hevc_encode_intra(C.byref(p), y.ctypes.data, cb.ctypes.data if chroma else None,
                                        cr.ctypes.data if chroma else None, y.strides[0], cb.strides[0] if chroma else 0,
                                        C.byref(out), C.byref(n)))
    data = bytes(C.cast(out, C.POINTER(C.c_uint8 * n.value)).contents)
    l.b200_free(out)
    return data


def synthetic_image(seed: int, width: int, height: int, bit_depth: int = 8, chroma=True):
    """Source picture of SURVEY.md 8(d): smooth gradient + 8 random-oriented sinusoid gratings + 1/16-amplitude noise,
    all driven by the 32-bit LCG s = s*1664525 + 1013904223 seeded with `seed`."""
    s = seed & 0xFFFFFFFF

    def nxt():
        nonlocal s
        s = (s * 1664525 + 1013904223) & 0xFFFFFFFF
        return s >> 8

    maxv = (1 << bit_depth) - 1
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float32)
    planes = []
    # chroma: False / 0 = 4:0:0, True / 1 = 4:2:0, 2 = 4:2:2, 3 = 4:4:4 (chroma_format_idc)
    cfmt = int(chroma)
    for c in range(3 if cfmt else 1):
        subx = 2 if (c and cfmt in (1, 2)) else 1
        suby = 2 if (c and cfmt == 1) else 1
        h, w = (height + suby - 1) // suby, (width + subx - 1) // subx
        X, Y = xx[:h, :w] * subx, yy[:h, :w] * suby
        gx, gy = (nxt() % 200 - 100) / 100.0, (nxt() % 200 - 100) / 100.0
        img = 0.5 + 0.25 * (gx * (X / max(width, 1) - 0.5) + gy * (Y / max(height, 1) - 0.5))
        for _ in range(8):
            ang = (nxt() % 3600) / 3600.0 * np.pi
            freq = 2 * np.pi / (4 + nxt() % 120)
            ph = (nxt() % 1000) / 1000.0 * 2 * np.pi
            amp = (0.02 + (nxt() % 100) / 1500.0) * (0.5 if c else 1.0)
            img = img + amp * np.sin(freq * (np.cos(ang) * X + np.sin(ang) * Y) + ph)
        rs = np.random.RandomState(nxt() & 0x7FFFFFFF)       # noise: bulk generator seeded from the LCG stream
        img = img + (rs.rand(h, w).astype(np.float32) - 0.5) / 16.0
        planes.append(np.clip(np.rint(img * maxv), 0, maxv).astype(np.uint16 if bit_depth > 8 else np.uint8))
    return planes if chroma else [planes[0], None, None]
