"""BASELINE.json configs 2-5 as parity cases (config 1 = test_hevc_gpu.py::test_example_heic_rgb_md5, config 3 is also
the bench workload).  Every case runs the whole device path through the C ABI -- HEVC tiles -> canvas planes -> colour
stage -- and is compared bit-exactly with the oracle (C restatement of the HEVC decode pinned on FFmpeg + C restatement
of the reference colour stage pinned on the reference).  Sizes: config 2 at full size; configs 3-5 at sizes the oracle
finishes in seconds, plus size-independent properties at the full sizes where the oracle would take minutes."""
import hashlib

import numpy as np
import pytest

import libheif_b200 as lb
from oracle import bindings as ob
from util import oracle_postprocess

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dec(cuda):
    d = lb.Decoder(host_threads=8)
    yield d
    d.close()


def _encode(seed, w, h, bd, **opts):
    y, cb, cr = lb.hevc_enc.synthetic_image(seed, w, h, bd, True)
    o = dict(log2_ctb_size=5, wpp=1, qp=27, seed=seed)
    o.update(opts)
    return lb.hevc_enc.encode_intra(y, cb, cr, bit_depth=bd, **o)


def _paste(tiles_planes, cols, rows, tw, th):
    W, H = cols * tw, rows * th
    canvas = [np.zeros((H, W), np.uint16), np.zeros((H // 2, W // 2), np.uint16), np.zeros((H // 2, W // 2), np.uint16)]
    for k, pl in enumerate(tiles_planes):
        c0, r0 = k % cols, k // cols
        for c in range(3):
            s = 1 if c == 0 else 2
            canvas[c][r0 * th // s:(r0 + 1) * th // s, c0 * tw // s:(c0 + 1) * tw // s] = pl[c]
    return canvas


def test_config2_4096_single_tile_rgb24(dec):
    """Config 2: one 4096x4096 HEVC-intra tile, 8-bit 4:2:0 -> RGB24 (BT.601 limited, the 2-op float path of the planner)."""
    au = _encode(0xB200, 4096, 4096, 8, vui_present=1, colour_description_present=1, colour_primaries=1, transfer_characteristics=13,
                 matrix_coefficients=6, full_range=0)
    planes, info = ob.restatement_decode(au)
    want, ow, oh = oracle_postprocess(planes[0], planes[1], planes[2], None, 1, 8, (info["cp"], info["tc"], info["mc"], info["full_range"]), [], 10)
    i = dec.decode_grid([au], 1, 1)
    assert (i.width, i.height, i.bit_depth) == (4096, 4096, 8)
    got_planes = dec.planes_host()
    for c in range(3):
        assert np.array_equal(got_planes[c], planes[c]), f"plane {c}"
    rgb = dec.to_rgb_device(lb.CHROMA_INTERLEAVED_RGB)
    assert (oh, ow) == (4096, 4096)
    assert np.array_equal(rgb.cpu().numpy().reshape(-1), want)


def test_config3_grid_slice_and_properties(dec):
    """Config 3: grid of independent 1024x1024 tiles.  A 3x2 slice of the grid is compared with the oracle; the rest of
    the 16x16 grid is covered by the bench and by the property that decoding the same tiles twice / in a different
    grid position gives identical tile pixels (tiles share nothing: libheif/image-items/grid.cc:482-577)."""
    tiles = [_encode(0xB200 + k, 1024, 1024, 8) for k in range(6)]
    ref = [ob.restatement_decode(t)[0] for t in tiles]
    want = _paste(ref, 3, 2, 1024, 1024)
    dec.decode_grid(tiles, cols=3, rows=2)
    got = dec.planes_host()
    for c in range(3):
        assert np.array_equal(got[c], want[c]), f"plane {c}"
    rgb_a = dec.to_rgb_device(lb.CHROMA_INTERLEAVED_RGB).cpu().numpy().reshape(2048, 3072, 3)
    # permute the tiles: every tile's RGB block must move with it
    perm = [5, 3, 1, 0, 2, 4]
    dec.decode_grid([tiles[p] for p in perm], cols=2, rows=3)
    rgb_b = dec.to_rgb_device(lb.CHROMA_INTERLEAVED_RGB).cpu().numpy().reshape(3072, 2048, 3)
    for pos, p in enumerate(perm):
        a = rgb_a[(p // 3) * 1024:(p // 3 + 1) * 1024, (p % 3) * 1024:(p % 3 + 1) * 1024]
        b = rgb_b[(pos // 2) * 1024:(pos // 2 + 1) * 1024, (pos % 2) * 1024:(pos % 2 + 1) * 1024]
        assert np.array_equal(a, b), f"tile {p}"


def test_config4_main10_rotate90_rgb48(dec):
    """Config 4: 10-bit 4:2:0 (nclx 9/16/9 limited), irot = 90 degrees CCW, output interleaved RRGGBB little endian.
    Parity on a 2x2 grid of 512x512 Main10 tiles (1024x1024 canvas) through the whole chain."""
    vui = dict(vui_present=1, colour_description_present=1, colour_primaries=9, transfer_characteristics=16, matrix_coefficients=9, full_range=0)
    tiles = [_encode(0xB200 + 40 + k, 512, 512, 10, **vui) for k in range(4)]
    ref = [ob.restatement_decode(t)[0] for t in tiles]
    canvas = _paste(ref, 2, 2, 512, 512)
    want, ow, oh = oracle_postprocess(canvas[0], canvas[1], canvas[2], None, 1, 10, (9, 16, 9, 0), [(1, 90)], 14)
    i = dec.decode_grid(tiles, cols=2, rows=2)
    assert i.bit_depth == 10
    got = dec.to_rgb_device(lb.CHROMA_INTERLEAVED_RRGGBB_LE, lb.Geometry(1024, 1024).rotate_ccw(90))
    assert (ow, oh) == (1024, 1024)
    assert np.array_equal(got.cpu().numpy().reshape(-1), want)


def test_config4_full_size_post_stage(cuda):
    """Config 4 at full size on the post-stage input (an 8192x8192 picture exceeds every HEVC level, so the
    config is defined on injected planes): 8192x8192 10-bit planes, rotate 90, -> RRGGBB_LE.  Properties: (a) the result
    equals the un-rotated conversion rotated afterwards; (b) a 256x256 window equals the oracle on that window."""
    import torch
    w = h = 8192
    g = torch.Generator(device="cuda"); g.manual_seed(0xB200)
    y = torch.randint(64, 941, (h, w), generator=g, device="cuda", dtype=torch.int16)
    cb = torch.randint(64, 961, (h // 2, w // 2), generator=g, device="cuda", dtype=torch.int16)
    cr = torch.randint(64, 961, (h // 2, w // 2), generator=g, device="cuda", dtype=torch.int16)
    img = lb.YCbCrImage(y, cb, cr, None, chroma=1, bit_depth=10, colour_primaries=9, transfer_characteristics=16, matrix_coefficients=9, full_range=False)
    base = lb.convert_colorspace(img, lb.CHROMA_INTERLEAVED_RRGGBB_LE).view(h, w, 6)
    rot = lb.convert_colorspace(img, lb.CHROMA_INTERLEAVED_RRGGBB_LE, lb.Geometry(w, h).rotate_ccw(90)).view(w, h, 6)
    assert torch.equal(rot, base.transpose(0, 1).flip(0).contiguous())          # rotate_ccw(90): out[v][u] = in[u][w-1-v]
    ty, tx = 4096, 1024
    want, _, _ = oracle_postprocess(y[ty:ty + 256, tx:tx + 256].cpu().numpy().astype(np.uint16), cb[ty // 2:ty // 2 + 128, tx // 2:tx // 2 + 128].cpu().numpy().astype(np.uint16),
                                    cr[ty // 2:ty // 2 + 128, tx // 2:tx // 2 + 128].cpu().numpy().astype(np.uint16), None, 1, 10, (9, 16, 9, 0), [], 14)
    assert np.array_equal(base[ty:ty + 256, tx:tx + 256].contiguous().cpu().numpy().reshape(-1), want)


def test_config5_main12_region_of_interest(dec):
    """Config 5: a tiled 12-bit image of which only some tiles are decoded (region of interest).  The requested tiles are
    drawn by the LCG of SURVEY 8d (seed 0xB2005) from a 32x32 tile grid; only those tiles exist as bitstreams (tiles are
    independent, so the others are never touched) and each is returned individually, here 8 tiles of 256x256."""
    s = 0xB2005
    picks = []
    while len(picks) < 8:
        s = (s * 1664525 + 1013904223) & 0xffffffff
        t = (s >> 8) % 1024
        if t not in picks:
            picks.append(t)
    tiles = [_encode(0xB200 + t, 256, 256, 12, log2_ctb_size=4 + (t % 3)) for t in picks]
    ref = [ob.restatement_decode(t)[0] for t in tiles]
    i = dec.decode_grid(tiles, cols=8, rows=1)                   # a batch of independent tiles, side by side
    assert i.bit_depth == 12
    got = dec.planes_host()
    for k in range(8):
        for c in range(3):
            sdiv = 1 if c == 0 else 2
            assert np.array_equal(got[c][:, k * 256 // sdiv:(k + 1) * 256 // sdiv], ref[k][c]), f"tile {picks[k]} plane {c}"
    # 12-bit -> RRGGBB_LE for the ROI batch (float path), against the colour oracle on the pasted planes
    canvas = _paste(ref, 8, 1, 256, 256)
    want, ow, oh = oracle_postprocess(canvas[0], canvas[1], canvas[2], None, 1, 12, (2, 2, 2, 0), [], 14)   # no VUI: the plugin attaches 2/2/2 limited (decoder_libde265.cc:426-449)
    rgb = dec.to_rgb_device(lb.CHROMA_INTERLEAVED_RRGGBB_LE)
    assert np.array_equal(rgb.cpu().numpy().reshape(-1), want)
