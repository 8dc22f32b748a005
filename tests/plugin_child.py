"""Child process of tests/test_plugin*.py (never imports torch: libheif_ref.so is loaded RTLD_GLOBAL here).
usage: plugin_child.py encode-cpu | roundtrip-gpu"""
import ctypes as C
import hashlib
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from oracle import bindings as ob  # noqa: E402
from oracle import refheif as rh  # noqa: E402

mode = sys.argv[1]
h = rh.load()
b200 = C.CDLL(os.path.join(ROOT, "libheif_b200", "libb200heif.so"))
b200.b200_get_decoder_plugin.restype = C.c_void_p
b200.b200_get_encoder_plugin.restype = C.c_void_p
assert b200.b200_plugin_bind_libheif(None) == 0, "plugin could not resolve the libheif C API"
rh.check(h.heif_register_encoder_plugin(b200.b200_get_encoder_plugin()), "register encoder plugin")
rh.register_cpu_decoder()

sys.path.insert(0, os.path.join(ROOT, "tests"))
from libheif_b200.hevc_enc import synthetic_image  # noqa: E402  (pure numpy helper)

res = {}
tmp = tempfile.mkdtemp()
nclx = (1, 13, 6, 1)
# single image, odd size -> conformance window + (possibly) clap written by libheif
y, cb, cr = synthetic_image(1, 200, 136, 8, True)
img = rh.make_ycbcr_image(y, cb, cr, 8, nclx)
rh.encode_file(os.path.join(tmp, "single.heic"), [img], quality=70)
cpu_single = rh.decode_file(os.path.join(tmp, "single.heic"), decoder_id="b200-oracle")
res["single_shape"] = list(cpu_single.shape)
ref_y = np.repeat(y[:, :, None], 3, 2)
res["single_psnr_luma_vs_green"] = float(10 * np.log10(255 ** 2 / max(1e-9, np.mean((cpu_single.reshape(136, 200, 3)[:, :, 1].astype(float) - y.astype(float)) ** 2))))
# 3x2 grid of 128x128 tiles through heif_context_encode_grid (same encoder instance for every tile, grid.cc:886-906)
tiles = []
for k in range(6):
    ty, tcb, tcr = synthetic_image(100 + k, 128, 128, 8, True)
    tiles.append(rh.make_ycbcr_image(ty, tcb, tcr, 8, nclx))
rh.encode_file(os.path.join(tmp, "grid.heic"), tiles, columns=3, rows=2, quality=60, params={"log2-ctb-size": 5})
cpu_grid = rh.decode_file(os.path.join(tmp, "grid.heic"), decoder_id="b200-oracle", threads=4)
res["grid_shape"] = list(cpu_grid.shape)
res["grid_md5_cpu"] = hashlib.md5(cpu_grid.tobytes()).hexdigest()
res["single_md5_cpu"] = hashlib.md5(cpu_single.tobytes()).hexdigest()
# oracle/heic_writer.py: the same tiles wrapped by our own ISOBMFF writer must decode to the same picture as the file the
# reference's writer produced (same encoder parameters -> byte-identical access units)
from oracle import heic_writer as hw  # noqa: E402
from libheif_b200 import hevc_enc  # noqa: E402
aus = []
for k in range(6):
    ty, tcb, tcr = synthetic_image(100 + k, 128, 128, 8, True)
    aus.append(hevc_enc.encode_intra(ty, tcb, tcr, bit_depth=8, log2_ctb_size=5, qp=51 - (60 * 45 + 50) // 100, wpp=1, seed=0xB200, vui_present=1,
                                     colour_description_present=1, colour_primaries=1, transfer_characteristics=13, matrix_coefficients=6, full_range=1))
hw.write_heic(os.path.join(tmp, "grid_own.heic"), aus, cols=3, rows=2)
own = rh.decode_file(os.path.join(tmp, "grid_own.heic"), decoder_id="b200-oracle", threads=4)
res["grid_md5_own_writer"] = hashlib.md5(own.tobytes()).hexdigest()
hw.write_heic(os.path.join(tmp, "single_own.heic"), aus[:1])
res["single_own_shape"] = list(rh.decode_file(os.path.join(tmp, "single_own.heic"), decoder_id="b200-oracle").shape)
if mode == "roundtrip-gpu":
    rh.check(h.heif_register_decoder_plugin(b200.b200_get_decoder_plugin()), "register decoder plugin")
    for name in ("single", "grid"):
        a = rh.decode_file(os.path.join(tmp, name + ".heic"), decoder_id="b200", threads=8)      # explicit selection (decoder.cc:330-338)
        res[name + "_md5_gpu"] = hashlib.md5(a.tobytes()).hexdigest()
        bdef = rh.decode_file(os.path.join(tmp, name + ".heic"), threads=8)                        # priority selection: 200 > 500? (oracle reports 500)
        res[name + "_md5_default"] = hashlib.md5(bdef.tobytes()).hexdigest()
print("RESULT " + json.dumps(res))
