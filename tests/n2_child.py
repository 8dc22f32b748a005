"""Child of tests/test_n2_patch.py: heif_decode_image of a set of files through the reference library named by B200_REF_LIB
(unmodified libheif_ref.so, or libheif_ref_b200.so = the same sources + the GPU colour operation of SURVEY 8f N2), always with
the oracle's CPU decoder plugin, so that only the colour stage differs.  Prints the md5 of every decoded picture."""
import hashlib
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import heic_writer as hw  # noqa: E402
from oracle import refheif as rh  # noqa: E402
from libheif_b200 import hevc_enc  # noqa: E402

rh.load()
rh.register_cpu_decoder()
tmp = tempfile.mkdtemp()
res = {}


def au_file(name):
    return open(os.path.join(ROOT, "tests", "golden", "streams", name), "rb").read()


cases = []
# C1: the access unit of examples/example.heic (1280x854, VUI-less -> limited range, default coefficients)
cases.append(("c1_example_rgb", [au_file("example_primary_1280x854.au")], 1, 1, rh.CHROMA_INTERLEAVED_RGB))
cases.append(("c1_example_rgba", [au_file("example_primary_1280x854.au")], 1, 1, rh.CHROMA_INTERLEAVED_RGBA))
# C2-like: one 8-bit tile, BT.601 limited and full range
for fr in (0, 1):
    y, cb, cr = hevc_enc.synthetic_image(0xB200 + fr, 512, 384, 8, True)
    au = hevc_enc.encode_intra(y, cb, cr, bit_depth=8, log2_ctb_size=5, qp=27, wpp=1, vui_present=1, colour_description_present=1, colour_primaries=1,
                               transfer_characteristics=13, matrix_coefficients=6, full_range=fr)
    cases.append((f"c2_601_full{fr}", [au], 1, 1, rh.CHROMA_INTERLEAVED_RGB))
# C3-like: a 3x2 grid of 8-bit tiles
tiles = []
for k in range(6):
    y, cb, cr = hevc_enc.synthetic_image(700 + k, 256, 256, 8, True)
    tiles.append(hevc_enc.encode_intra(y, cb, cr, bit_depth=8, log2_ctb_size=5, qp=27, wpp=1, seed=0xB200, vui_present=1, colour_description_present=1,
                                       colour_primaries=1, transfer_characteristics=13, matrix_coefficients=6, full_range=0))
cases.append(("c3_grid", tiles, 3, 2, rh.CHROMA_INTERLEAVED_RGB))
# C4-like: 10-bit BT.2020 limited -> RRGGBB_LE, and 12-bit -> RRGGBBAA_BE
for bd, outc in ((10, 14), (12, 13)):
    y, cb, cr = hevc_enc.synthetic_image(900 + bd, 320, 192, bd, True)
    au = hevc_enc.encode_intra(y, cb, cr, bit_depth=bd, log2_ctb_size=5, qp=24, wpp=1, vui_present=1, colour_description_present=1, colour_primaries=9,
                               transfer_characteristics=16, matrix_coefficients=9, full_range=0)
    cases.append((f"c4_{bd}bit_{outc}", [au], 1, 1, outc))
for name, aus, cols, rows, outc in cases:
    path = os.path.join(tmp, name + ".heic")
    hw.write_heic(path, aus, cols=cols, rows=rows)
    out = rh.decode_file(path, chroma=outc, decoder_id="b200-oracle", threads=4)
    res[name] = [list(out.shape), hashlib.md5(out.tobytes()).hexdigest()]
print("RESULT " + json.dumps(res))
