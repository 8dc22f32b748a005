"""Colour post-stage: host mirror of libheif's convert_colorspace()/rotate_ccw()/mirror_inplace()/crop().

Reference interfaces mirrored (argument meaning and order of application):
  convert_colorspace(img, colorspace, chroma, ...)        libheif/color-conversion/colorconversion.cc:490-623
  HeifPixelImage::rotate_ccw / mirror_inplace / crop      libheif/image/pixelimage.cc:1175-1546
  ImageItem::decode_image transform loop                  libheif/image-items/image_item.cc:947-1020
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _lib

CHROMA_MONO, CHROMA_420, CHROMA_422, CHROMA_444 = 0, 1, 2, 3
CHROMA_INTERLEAVED_RGB, CHROMA_INTERLEAVED_RGBA = 10, 11
CHROMA_INTERLEAVED_RRGGBB_BE, CHROMA_INTERLEAVED_RRGGBBAA_BE = 12, 13
CHROMA_INTERLEAVED_RRGGBB_LE, CHROMA_INTERLEAVED_RRGGBBAA_LE = 14, 15

_BYTES_PER_PIXEL = {10: 3, 11: 4, 12: 6, 13: 8, 14: 6, 15: 8}


class Geometry:
    """Chain of irot / imir / clap transforms, composed in the order libheif applies them."""

    def __init__(self, width: int, height: int, chroma: int = 1):
        """chroma: B200_CHROMA_* of the picture the chain applies to (decides where the reference converts to 4:4:4 first)."""
        self.g = _lib.Geometry()
        _lib.lib().b200_geometry_init(width, height, chroma, C.byref(self.g))

    def rotate_ccw(self, degrees: int) -> "Geometry":
        _lib.check(_lib.lib().b200_geometry_rotate_ccw(C.byref(self.g), degrees))
        return self

    def mirror(self, direction: int) -> "Geometry":
        """direction: heif_transform_mirror_direction (0 = vertical: top<->bottom, 1 = horizontal: left<->right)."""
        _lib.check(_lib.lib().b200_geometry_mirror(C.byref(self.g), direction))
        return self

    def crop(self, left: int, right: int, top: int, bottom: int) -> "Geometry":
        _lib.check(_lib.lib().b200_geometry_crop(C.byref(self.g), left, right, top, bottom))
        return self

    @property
    def size(self):
        return self.g.out_w, self.g.out_h


@dataclass
class YCbCrImage:
    """Decoded picture as the decoder plugin hands it over (decoder_libde265.cc:97-171): planes + nclx."""
    y: object
    cb: Optional[object] = None
    cr: Optional[object] = None
    alpha: Optional[object] = None
    chroma: int = CHROMA_420
    bit_depth: int = 8
    colour_primaries: int = 2
    transfer_characteristics: int = 2
    matrix_coefficients: int = 2
    full_range: bool = False
    _keep: list = field(default_factory=list, repr=False)


def _fill_planes(img: YCbCrImage, ptr, stride):
    p = _lib.Planes()
    h, w = img.y.shape
    p.y = ptr(img.y); p.y_stride = stride(img.y)
    if img.chroma != CHROMA_MONO:
        p.cb = ptr(img.cb); p.cr = ptr(img.cr); p.c_stride = stride(img.cb)
        assert stride(img.cb) == stride(img.cr)
    if img.alpha is not None:
        p.alpha = ptr(img.alpha); p.alpha_stride = stride(img.alpha)
    p.width, p.height, p.chroma, p.bit_depth = w, h, img.chroma, img.bit_depth
    p.colour_primaries, p.transfer_characteristics = img.colour_primaries, img.transfer_characteristics
    p.matrix_coefficients, p.full_range = img.matrix_coefficients, int(bool(img.full_range))
    return p


def _out_shape(out_chroma, w, h, bit_depth):
    if out_chroma == CHROMA_444:
        return (3, h, w), (np.uint16 if bit_depth > 8 else np.uint8)
    return (h, w * _BYTES_PER_PIXEL[out_chroma]), np.uint8


def convert_colorspace(img: YCbCrImage, out_chroma: int, geometry: Optional[Geometry] = None, out=None, stream=None, bilinear: bool = False):
    """Device -> device. `img` planes are CUDA torch tensors (uint8, or int16/uint16 for >8 bit).

    Returns a CUDA uint8 tensor [H, W*bytes_per_pixel] (interleaved) or [3, H, W] (planar RGB 4:4:4)."""
    import torch
    l = _lib.lib()
    h, w = img.y.shape
    geom = geometry or Geometry(w, h)
    ow, oh = geom.size
    shape, dt = _out_shape(out_chroma, ow, oh, img.bit_depth)
    tdt = torch.uint8 if dt == np.uint8 else torch.int16
    if out is None:
        out = torch.empty(shape, dtype=tdt, device=img.y.device)
    planes = _fill_planes(img, lambda t: t.data_ptr(), lambda t: t.stride(0) * t.element_size())
    opt = _lib.ColorOptions(out_chroma, 0, 1 if bilinear else 0)
    s = stream if stream is not None else torch.cuda.current_stream(img.y.device)
    pipe = C.c_int(0)
    if out_chroma == CHROMA_444:
        o, og, ob = out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr()
        ostride = out.stride(1) * out.element_size()
    else:
        o, og, ob = out.data_ptr(), None, None
        ostride = out.stride(0) * out.element_size()
    with torch.cuda.device(img.y.device):
        _lib.check(l.b200_color_convert_device(C.byref(planes), C.byref(geom.g), C.byref(opt), o, og, ob, ostride,
                                               C.c_void_p(s.cuda_stream), C.byref(pipe)))
    return out


def convert_colorspace_host(img: YCbCrImage, out_chroma: int, geometry: Optional[Geometry] = None):
    """Host -> host through the C ABI (H2D + kernel + D2H inside the call). Planes are numpy arrays."""
    l = _lib.lib()
    h, w = img.y.shape
    geom = geometry or Geometry(w, h)
    ow, oh = geom.size
    shape, dt = _out_shape(out_chroma, ow, oh, img.bit_depth)
    out = np.empty(shape, dtype=dt)
    planes = _fill_planes(img, lambda a: a.ctypes.data, lambda a: a.strides[0])
    opt = _lib.ColorOptions(out_chroma, 0, 0)
    pipe = C.c_int(0)
    if out_chroma == CHROMA_444:
        o, og, ob, ostride = out[0].ctypes.data, out[1].ctypes.data, out[2].ctypes.data, out.strides[1]
    else:
        o, og, ob, ostride = out.ctypes.data, None, None, out.strides[0]
    _lib.check(l.b200_color_convert_host(C.byref(planes), C.byref(geom.g), C.byref(opt), o, og, ob, ostride, C.byref(pipe)))
    return out, pipe.value


def _ycc_out_planes(w, h, out_chroma, want_alpha, alloc):
    sh = 0 if out_chroma == CHROMA_444 else 1
    sv = 1 if out_chroma == CHROMA_420 else 0
    y = alloc((h, w))
    cb = alloc(((h + sv) >> sv, (w + sh) >> sh))
    cr = alloc(((h + sv) >> sv, (w + sh) >> sh))
    a = alloc((h, w)) if want_alpha else None
    return y, cb, cr, a


def rgb_to_ycbcr(rgb, out_chroma: int = CHROMA_420, matrix_coefficients: int = 6, colour_primaries: int = 1, full_range: bool = True,
                 want_alpha: Optional[bool] = None, stream=None) -> YCbCrImage:
    """Encoder-side direction, device -> device: interleaved RGB / RGBA (CUDA uint8 tensor [H, W, 3 or 4]) -> YCbCrImage of
    CUDA uint8 planes, as Op_RGB24_32_to_YCbCr does (libheif/color-conversion/rgb2yuv.cc:575-808).
    want_alpha: None = an alpha plane iff the source has one (what the reference's planner targets for has_alpha)."""
    import torch
    assert rgb.dtype == torch.uint8 and rgb.dim() == 3 and rgb.shape[2] in (3, 4) and rgb.stride(2) == 1 and rgb.stride(1) == rgb.shape[2]
    h, w, bpp = rgb.shape
    if want_alpha is None:
        want_alpha = bpp == 4
    y, cb, cr, a = _ycc_out_planes(w, h, out_chroma, want_alpha, lambda s: torch.empty(s, dtype=torch.uint8, device=rgb.device))
    img = YCbCrImage(y, cb, cr, a, chroma=out_chroma, bit_depth=8, colour_primaries=colour_primaries,
                     matrix_coefficients=matrix_coefficients, full_range=full_range)
    planes = _fill_planes(img, lambda t: t.data_ptr(), lambda t: t.stride(0) * t.element_size())
    s = stream if stream is not None else torch.cuda.current_stream(rgb.device)
    with torch.cuda.device(rgb.device):
        _lib.check(_lib.lib().b200_rgb_to_ycbcr_device(C.c_void_p(rgb.data_ptr()), C.c_size_t(rgb.stride(0)), int(bpp == 4), C.byref(planes),
                                                       C.c_void_p(s.cuda_stream)))
    return img


def rgb_to_ycbcr_host(rgb: np.ndarray, out_chroma: int = CHROMA_420, matrix_coefficients: int = 6, colour_primaries: int = 1,
                      full_range: bool = True, want_alpha: Optional[bool] = None) -> YCbCrImage:
    """Host -> host through the C ABI (H2D + kernel + D2H inside the call). rgb: uint8 [H, W, 3 or 4], rows may be strided."""
    assert rgb.dtype == np.uint8 and rgb.ndim == 3 and rgb.shape[2] in (3, 4) and rgb.strides[2] == 1 and rgb.strides[1] == rgb.shape[2]
    h, w, bpp = rgb.shape
    if want_alpha is None:
        want_alpha = bpp == 4
    y, cb, cr, a = _ycc_out_planes(w, h, out_chroma, want_alpha, lambda s: np.empty(s, np.uint8))
    img = YCbCrImage(y, cb, cr, a, chroma=out_chroma, bit_depth=8, colour_primaries=colour_primaries,
                     matrix_coefficients=matrix_coefficients, full_range=full_range)
    planes = _fill_planes(img, lambda x: x.ctypes.data, lambda x: x.strides[0])
    _lib.check(_lib.lib().b200_rgb_to_ycbcr_host(C.c_void_p(rgb.ctypes.data), C.c_size_t(rgb.strides[0]), int(bpp == 4), C.byref(planes)))
    return img
