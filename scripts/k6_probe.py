"""Quick device timing of the K6 colour kernel (development probe; bench.py is the judged harness)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import libheif_b200 as lb

dev = torch.device("cuda:0")
res = {}
for (w, h, bpp, outc, rot) in [(4096, 4096, 8, 10, 0), (16384, 16384, 8, 10, 0), (8192, 8192, 10, 14, 90), (16384, 16384, 8, 10, 90)]:
    dt = torch.uint8 if bpp == 8 else torch.int16
    hi = 255 if bpp == 8 else 1023
    y = torch.randint(0, hi, (h, w), device=dev, dtype=torch.int32).to(dt)
    cb = torch.randint(0, hi, (h // 2, w // 2), device=dev, dtype=torch.int32).to(dt)
    cr = torch.randint(0, hi, (h // 2, w // 2), device=dev, dtype=torch.int32).to(dt)
    img = lb.YCbCrImage(y, cb, cr, chroma=1, bit_depth=bpp, colour_primaries=1, transfer_characteristics=13, matrix_coefficients=6, full_range=False)
    g = lb.Geometry(w, h)
    if rot: g.rotate_ccw(rot)
    out = lb.convert_colorspace(img, outc, g)
    for _ in range(3): lb.convert_colorspace(img, outc, g, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); lb.convert_colorspace(img, outc, g, out=out); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    bpx = 4.5 if bpp == 8 else 9.0
    res[f"{w}x{h}_{bpp}b_out{outc}_rot{rot}"] = dict(ms=ms, mp_s=w * h / ms / 1e3, gb_s=w * h * bpx / ms / 1e6)
print(json.dumps(res, indent=1))
