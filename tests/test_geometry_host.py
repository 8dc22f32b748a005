"""Host logic of the geometry composition (no GPU): where the chain reaches the reference's 4:4:4 conversion point
(pixelimage.cc:1187-1215 rotate_ccw, 1370-1381 mirror_inplace, 1458-1467 crop) and how the affine maps compose."""
import numpy as np
import pytest

import libheif_b200 as lb


def _apply(g, u, v):
    m = list(g.g.m)
    return m[0] * u + m[1] * v + m[2], m[3] * u + m[4] * v + m[5]


def _pre(g, u, v):
    m = list(g.g.pre)
    return m[0] * u + m[1] * v + m[2], m[3] * u + m[4] * v + m[5]


@pytest.mark.parametrize("w,h,op,expect", [
    (34, 18, ("rot", 90), 0), (33, 18, ("rot", 90), 1), (34, 17, ("rot", 90), 0),
    (34, 18, ("rot", 180), 0), (33, 18, ("rot", 180), 1), (34, 17, ("rot", 180), 1),
    (34, 18, ("rot", 270), 0), (34, 17, ("rot", 270), 1), (33, 18, ("rot", 270), 0),
    (34, 18, ("mir", 0), 0), (33, 18, ("mir", 0), 1), (34, 17, ("mir", 1), 1),
    (34, 18, ("crop", 2, 31, 4, 15), 0), (34, 18, ("crop", 3, 31, 4, 15), 1), (34, 18, ("crop", 2, 31, 5, 15), 1),
    (34, 18, ("crop", 2, 30, 4, 14), 0),                       # odd right / bottom never need the conversion
])
def test_420_conversion_point_rules(w, h, op, expect):
    g = lb.Geometry(w, h, lb.CHROMA_420)
    if op[0] == "rot":
        g.rotate_ccw(op[1])
    elif op[0] == "mir":
        g.mirror(op[1])
    else:
        g.crop(*op[1:])
    assert g.g.detour == expect
    if expect:                                                # the triggering op itself applies to the 4:4:4 picture
        assert (g.g.pre_w, g.g.pre_h) == (w, h) and list(g.g.pre) == [1, 0, 0, 0, 1, 0]


@pytest.mark.parametrize("chroma,op,expect", [
    (lb.CHROMA_422, ("rot", 90), 1), (lb.CHROMA_422, ("rot", 180), 0), (lb.CHROMA_422, ("mir", 1), 0), (lb.CHROMA_422, ("crop", 3, 31, 4, 15), 1),
    (lb.CHROMA_422, ("crop", 2, 31, 5, 15), 0), (lb.CHROMA_444, ("crop", 3, 31, 5, 15), 0), (lb.CHROMA_444, ("rot", 90), 0), (lb.CHROMA_MONO, ("mir", 0), 0),
])
def test_other_formats(chroma, op, expect):
    g = lb.Geometry(34, 18, chroma)
    if op[0] == "rot":
        g.rotate_ccw(op[1])
    elif op[0] == "mir":
        g.mirror(op[1])
    else:
        g.crop(*op[1:])
    assert g.g.detour == expect


def test_chain_splits_at_the_conversion_point_and_maps_compose():
    """crop (even) -> rotate 90 -> crop (odd): the first two ops stay plane-wise (`pre`), the last one and everything after
    it run on the 4:4:4 picture (`m`).  Check both maps against a brute-force index picture."""
    w, h = 64, 48
    idx = np.arange(w * h).reshape(h, w)
    g = lb.Geometry(w, h, lb.CHROMA_420)
    g.crop(2, 61, 4, 43)              # 60 x 40
    step1 = idx[4:44, 2:62]
    g.rotate_ccw(90)                  # 40 x 60 (width 60 even: no conversion)
    step2 = np.rot90(step1, 1)
    assert g.g.detour == 0
    g.crop(3, 30, 1, 50)              # odd origin -> conversion point here
    assert g.g.detour == 1 and (g.g.pre_w, g.g.pre_h) == (40, 60)
    step3 = step2[1:51, 3:31]
    g.mirror(1)
    step4 = step3[:, ::-1]
    assert g.size == (step4.shape[1], step4.shape[0])
    for (u, v) in [(0, 0), (5, 7), (27, 49), (13, 20)]:
        sx, sy = _pre(g, u, v)                                   # `pre` maps the intermediate picture back to the source
        assert idx[sy, sx] == step2[v, u]
        ix, iy = _apply(g, u, v)                                 # `m` maps the final picture back to the intermediate one
        assert step2[iy, ix] == step4[v, u]


def test_invalid_arguments():
    g = lb.Geometry(32, 24)
    with pytest.raises(lb.B200Error):
        g.rotate_ccw(45)
    with pytest.raises(lb.B200Error):
        g.crop(0, 32, 0, 10)
    with pytest.raises(lb.B200Error):
        g.mirror(2)
