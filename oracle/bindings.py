"""
oracle/bindings.py -- TEST INFRASTRUCTURE ONLY (ctypes access to oracle/_ref/*.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
  * liboracle.so        : C restatements (hevc_oracle.c, color_oracle.c) + FFmpeg wrapper (ffhevc.c)
  * libheif_ref.so      : the UNMODIFIED reference libheif core, compiled from /root/reference by oracle/Makefile
  * liboracle_plugin.so : CPU decoder plugin (FFmpeg or restatement) registered into libheif_ref.so
"""
import ctypes as C
import glob
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def avcodec_dir():
    """Directory of the FFmpeg libraries bundled in the opencv-python-headless wheel (no import of cv2 needed)."""
    import importlib.util
    spec = importlib.util.find_spec("cv2")
    if spec is None:
        return None
    site = os.path.dirname(os.path.dirname(spec.origin))
    d = os.path.join(site, "opencv_python_headless.libs")
    return d if glob.glob(os.path.join(d, "libavcodec-*")) else None


class _FFPic(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("cw", C.c_int), ("ch", C.c_int), ("bit_depth", C.c_int),
                ("chroma", C.c_int), ("full_range_name", C.c_int), ("plane", C.POINTER(C.c_uint16) * 3)]


class _HOPic(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("cw", C.c_int), ("ch", C.c_int), ("bit_depth", C.c_int),
                ("chroma_format", C.c_int), ("vui_colour_present", C.c_int), ("colour_primaries", C.c_int),
                ("transfer_characteristics", C.c_int), ("matrix_coeffs", C.c_int), ("full_range", C.c_int),
                ("video_signal_present", C.c_int), ("plane", C.POINTER(C.c_uint16) * 3)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(REF, "liboracle.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/_ref/liboracle.so missing: run `make -C oracle` (or __graft_entry__.build())")
        _lib = C.CDLL(path)
        _lib.ffhevc_init.argtypes = [C.c_char_p]
        _lib.ffhevc_decode.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.POINTER(_FFPic)]
        _lib.hevc_oracle_decode.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.POINTER(_HOPic)]
    return _lib


def _planes(pic, chroma):
    out = []
    for c in range(3 if chroma else 1):
        w, h = (pic.cw, pic.ch) if c else (pic.width, pic.height)
        a = np.ctypeslib.as_array(pic.plane[c], shape=(h, w)).copy()
        out.append(a)
    return out


_ff_ready = False


def ffmpeg_decode(au: bytes, threads: int = 1):
    """Decode one length-prefixed HEVC access unit with FFmpeg. Returns (planes[list of uint16 HxW], bit_depth, chroma)."""
    global _ff_ready
    l = lib()
    if not _ff_ready:
        d = avcodec_dir()
        if d is None or l.ffhevc_init(d.encode()) != 0:
            raise RuntimeError("FFmpeg (cv2 wheel) not available")
        _ff_ready = True
    pic = _FFPic()
    rc = l.ffhevc_decode(au, len(au), threads, C.byref(pic))
    if rc != 0:
        raise RuntimeError(f"ffhevc_decode rc={rc}")
    pl = _planes(pic, pic.chroma)
    l.ffhevc_free_picture(C.byref(pic))
    return pl, pic.bit_depth, pic.chroma


def restatement_decode(au: bytes, stage: int = 0):
    """Decode with the C restatement. stage 0 final, 1 before deblocking, 2 after deblocking before SAO."""
    l = lib()
    pic = _HOPic()
    rc = l.hevc_oracle_decode(au, len(au), stage, C.byref(pic))
    if rc != 0:
        raise RuntimeError(f"hevc_oracle_decode rc={rc}")
    pl = _planes(pic, pic.chroma_format)
    info = dict(bit_depth=pic.bit_depth, chroma=pic.chroma_format, cp=pic.colour_primaries,
                tc=pic.transfer_characteristics, mc=pic.matrix_coeffs, full_range=pic.full_range)
    l.hevc_oracle_free_picture(C.byref(pic))
    return pl, info


# ---------------------------------------------------------------------------------- post-stage wrappers
_plugin = None


def ref_plugin():
    """liboracle_plugin.so (+ libheif_ref.so): the unmodified reference. None if not built (e.g. reference absent)."""
    global _plugin
    if _plugin is None:
        p = os.path.join(REF, "liboracle_plugin.so")
        if not os.path.exists(p) or not os.path.exists(os.path.join(REF, "libheif_ref.so")):
            return None
        lib()
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


def oracle_postprocess(y, cb, cr, a, chroma, bpp, nclx, ops, out_chroma, bilinear=0):
    """C restatement (oracle/color_oracle.c)."""
    l = lib()
    h, w = y.shape
    cp, tc, mc, fr = nclx if nclx else (2, 2, 2, 1)   # image without nclx: defaults + full range (yuv2rgb.cc:203-215)
    ops_a = (C.c_int * (5 * max(1, len(ops))))(*[v for o in ops for v in (list(o) + [0] * 5)[:5]])
    cap = (max(w, h) + 64) ** 2 * 8
    out = np.empty(cap, dtype=np.uint8)
    ow, oh = C.c_int(), C.c_int()
    l.co_postprocess2.restype = C.c_long
    keep = [np.ascontiguousarray(p, dtype=np.uint16) if p is not None else None for p in (y, cb, cr, a)]
    n = l.co_postprocess2(*[None if k is None else k.ctypes.data_as(C.c_void_p) for k in keep], w, h, chroma, bpp, cp, mc, int(fr),
                          ops_a, len(ops), out_chroma, int(bilinear), out.ctypes.data_as(C.c_void_p), C.byref(ow), C.byref(oh))
    if n < 0:
        raise RuntimeError("co_postprocess failed")
    return out[:n].copy(), ow.value, oh.value


# ---- a11 / a12: overlay and nearest-neighbour scaling -----------------------------------------------------------------
def ref_overlay(cw, ch, bkg, children):
    """UNMODIFIED reference: fill_RGB_16bit + HeifPixelImage::overlay per child. children = [(rgb(3,h,w) u8, alpha(h,w) u8 or None, dx, dy)]."""
    pl = ref_plugin()
    n = len(children)
    packed = []
    for rgb, al, _, _ in children:
        parts = [np.ascontiguousarray(rgb, dtype=np.uint8).reshape(-1)]
        if al is not None:
            parts.append(np.ascontiguousarray(al, dtype=np.uint8).reshape(-1))
        packed.append(np.concatenate(parts))
    ptrs = (C.c_void_p * max(1, n))(*[p.ctypes.data_as(C.c_void_p) for p in packed])
    arr = lambda vals: (C.c_int * max(1, n))(*vals)
    out = np.empty(3 * cw * ch, dtype=np.uint8)
    b = (C.c_uint16 * 4)(*bkg)
    rc = pl.ref_overlay(cw, ch, b, n, ptrs, arr([c[0].shape[2] for c in children]), arr([c[0].shape[1] for c in children]),
                        arr([0 if c[1] is None else 1 for c in children]), arr([c[2] for c in children]), arr([c[3] for c in children]),
                        out.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError(f"ref_overlay rc={rc}")
    return out.reshape(3, ch, cw)


def oracle_overlay(cw, ch, bkg, children):
    """C restatement (oracle/color_oracle.c: co_fill_rgb16 + co_overlay)."""
    l = lib()
    canvas = np.empty(3 * cw * ch, dtype=np.uint8)
    l.co_fill_rgb16(canvas.ctypes.data_as(C.c_void_p), cw, ch, (C.c_uint16 * 4)(*bkg))
    for rgb, al, dx, dy in children:
        parts = [np.ascontiguousarray(rgb, dtype=np.uint8).reshape(-1)]
        if al is not None:
            parts.append(np.ascontiguousarray(al, dtype=np.uint8).reshape(-1))
        ov = np.concatenate(parts)
        l.co_overlay(canvas.ctypes.data_as(C.c_void_p), cw, ch, ov.ctypes.data_as(C.c_void_p), rgb.shape[2], rgb.shape[1], 0 if al is None else 1,
                     C.c_int32(dx), C.c_int32(dy))
    return canvas.reshape(3, ch, cw)


def ref_scale_nn(colorspace, chroma, bpp, planes, w, h, ow, oh, has_alpha=False):
    """UNMODIFIED reference scale_nearest_neighbor. planes: list of 2-D arrays in channel order (+ alpha last). Returns the packed output bytes."""
    pl = ref_plugin()
    dt = np.uint8 if bpp <= 8 else np.uint16
    packed = np.concatenate([np.ascontiguousarray(p, dtype=dt).reshape(-1).view(np.uint8) for p in planes])
    cap = (ow + 8) * (oh + 8) * 8 * (len(planes) + 1)
    out = np.empty(cap, dtype=np.uint8)
    n = pl.ref_scale_nn(colorspace, chroma, bpp, int(has_alpha), w, h, ow, oh, packed.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.c_size_t(cap))
    if n < 0:
        raise RuntimeError(f"ref_scale_nn rc={n}")
    return out[:n].copy()


def oracle_scale_plane(plane, out_w, out_h, image_in, image_out, comps=1):
    """C restatement of one plane of scale_nearest_neighbor (co_scale_nearest_plane). plane: (h, w*comps) u8/u16."""
    l = lib()
    p = np.ascontiguousarray(plane)
    bps = p.dtype.itemsize
    out = np.empty((out_h, out_w * comps), dtype=p.dtype)
    l.co_scale_nearest_plane(p.ctypes.data_as(C.c_void_p), C.c_size_t(p.shape[1] * bps), out.ctypes.data_as(C.c_void_p), C.c_size_t(out_w * comps * bps),
                             C.c_uint32(out_w), C.c_uint32(out_h), C.c_uint32(image_in[0]), C.c_uint32(image_in[1]), C.c_uint32(image_out[0]), C.c_uint32(image_out[1]),
                             comps, bps)
    return out


# ---- encoder side (N3): interleaved 8-bit RGB(A) -> YCbCr --------------------------------------------------------------------
def _ycc_planes(w, h, out_chroma, alpha):
    sh, sv = (2 if out_chroma in (1, 2) else 1), (2 if out_chroma == 1 else 1)
    cw, ch = (w + sh - 1) // sh, (h + sv - 1) // sv
    return np.empty((h, w), np.uint8), np.empty((ch, cw), np.uint8), np.empty((ch, cw), np.uint8), (np.empty((h, w), np.uint8) if alpha else None)


def ref_rgb_to_ycbcr(rgb, has_alpha, out_chroma, nclx):
    """UNMODIFIED reference: convert_colorspace(interleaved RGB(A) 8 bit -> YCbCr out_chroma, target nclx = (cp, tc, mc, full_range))."""
    pl = ref_plugin()
    h, wb = rgb.shape
    w = wb // (4 if has_alpha else 3)
    y, cb, cr, a = _ycc_planes(w, h, out_chroma, has_alpha)
    src = np.ascontiguousarray(rgb)
    rc = pl.ref_rgb_to_ycbcr(src.ctypes.data_as(C.c_void_p), w, h, int(has_alpha), out_chroma, nclx[0], nclx[1], nclx[2], int(nclx[3]),
                             y.ctypes.data_as(C.c_void_p), cb.ctypes.data_as(C.c_void_p), cr.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p) if a is not None else None)
    if rc != 0:
        raise RuntimeError(f"ref_rgb_to_ycbcr rc={rc}")
    return y, cb, cr, a


def oracle_rgb_to_ycbcr(rgb, has_alpha, out_chroma, nclx):
    """C restatement (oracle/color_oracle.c: co_rgb_to_ycbcr)."""
    l = lib()
    h, wb = rgb.shape
    w = wb // (4 if has_alpha else 3)
    y, cb, cr, a = _ycc_planes(w, h, out_chroma, has_alpha)
    src = np.ascontiguousarray(rgb)
    l.co_rgb_to_ycbcr.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    rc = l.co_rgb_to_ycbcr(src.ctypes.data, src.strides[0], w, h, int(has_alpha), out_chroma, nclx[2], nclx[0], int(nclx[3]),
                           y.ctypes.data, cb.ctypes.data, cr.ctypes.data, a.ctypes.data if a is not None else None)
    if rc != 0:
        raise RuntimeError("co_rgb_to_ycbcr failed")
    return y, cb, cr, a
