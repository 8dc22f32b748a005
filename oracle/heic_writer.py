"""
oracle/heic_writer.py -- TEST INFRASTRUCTURE ONLY: a minimal ISOBMFF / HEIF writer that wraps already encoded HEVC
access units (length-prefixed NALs, as libheif pushes them to a decoder plugin) into a .heic file the UNMODIFIED
reference libheif reads back: a single 'hvc1' item, or a 'grid' item over cols x rows 'hvc1' tiles
(ISO/IEC 23008-12 6.6.2.3; the reference's reader: libheif/image-items/grid.cc:34-121, hevc_boxes.cc:288-309,
box.cc).  Used by the reference arm of bench.py (heif_decode_image on the same tiles the GPU arm decodes) and by tests.
"""
import struct


def _box(t, payload):
    return struct.pack(">I4s", 8 + len(payload), t) + payload


def _full(t, version, flags, payload):
    return _box(t, struct.pack(">I", (version << 24) | flags) + payload)


def split_nals(au):
    nals, p = [], 0
    while p + 4 <= len(au):
        n = struct.unpack(">I", au[p:p + 4])[0]
        nals.append(au[p + 4:p + 4 + n]); p += 4 + n
    return nals


def _unescape(nal):
    out, z = bytearray(), 0
    for b in nal:
        if z >= 2 and b == 3:
            z = 0; continue
        out.append(b); z = z + 1 if b == 0 else 0
    return bytes(out)


class _Bits:
    def __init__(self, d): self.d, self.p = d, 0
    def u(self, n):
        v = 0
        for _ in range(n):
            v = (v << 1) | ((self.d[self.p >> 3] >> (7 - (self.p & 7))) & 1); self.p += 1
        return v
    def ue(self):
        z = 0
        while self.u(1) == 0: z += 1
        return (1 << z) - 1 + (self.u(z) if z else 0)


def hvcc_from_au(au):
    """HEVCDecoderConfigurationRecord (ISO/IEC 14496-15 8.3.3.1) from the VPS / SPS / PPS of an access unit; returns
    (hvcC payload, item data without the parameter sets, coded width, coded height minus the conformance window)."""
    nals = split_nals(au)
    ps = {32: [], 33: [], 34: []}
    rest = []
    for n in nals:
        t = (n[0] >> 1) & 0x3f
        (ps[t] if t in ps else rest).append(n)
    sps = _unescape(ps[33][0])
    ptl = sps[3:15]                       # general_profile_space .. general_level_idc (12 bytes, max_sub_layers = 0)
    b = _Bits(sps); b.u(16); b.u(4); msl = b.u(3); b.u(1); b.u(96)
    assert msl == 0
    b.ue(); cf = b.ue()
    if cf == 3: b.u(1)
    w, h = b.ue(), b.ue()
    if b.u(1):
        cl, cr, ct, cb = b.ue(), b.ue(), b.ue(), b.ue()
        sx, sy = (2 if cf in (1, 2) else 1), (2 if cf == 1 else 1)
        w -= sx * (cl + cr); h -= sy * (ct + cb)
    bdl, bdc = b.ue(), b.ue()
    rec = bytes([1]) + ptl + struct.pack(">HBBBBHB", 0xF000, 0xFC, 0xFC | cf, 0xF8 | bdl, 0xF8 | bdc, 0, 0x0F) + bytes([3])
    for t in (32, 33, 34):
        rec += bytes([0x80 | t]) + struct.pack(">H", len(ps[t]))
        for n in ps[t]:
            rec += struct.pack(">H", len(n)) + n
    data = b"".join(struct.pack(">I", len(n)) + n for n in rest)
    return rec, data, w, h


def write_heic(path, aus, cols=1, rows=1, out_w=None, out_h=None, nclx=None):
    """aus: row-major list of access units (cols * rows of them; every tile must share the parameter sets of tile 0)."""
    n = cols * rows
    assert len(aus) == n
    recs = [hvcc_from_au(a) for a in aus]
    hvcc, _, tw, th = recs[0]
    grid = n > 1
    out_w = out_w or tw * cols; out_h = out_h or th * rows
    props = [_box(b"hvcC", hvcc), _full(b"ispe", 0, 0, struct.pack(">II", tw, th))]
    if grid:
        props.append(_full(b"ispe", 0, 0, struct.pack(">II", out_w, out_h)))
    if nclx is not None:
        props.append(_box(b"colr", b"nclx" + struct.pack(">HHHB", nclx[0], nclx[1], nclx[2], 0x80 if nclx[3] else 0)))
    colr_idx = len(props) if nclx is not None else 0
    # tiles first, the grid item last: the reference resolves a derived image's tiles while it walks the items in id order
    # (HeifContext::find_first_coded_image_id, context.cc:1373-1420), as its own writer numbers them
    tile_ids = list(range(1, n + 1))
    gid = n + 1
    # item infos
    infes = []
    for i in tile_ids:
        infes.append(_full(b"infe", 2, 1 if grid else 0, struct.pack(">HH4s", i, 0, b"hvc1") + b"\0"))
    if grid:
        infes.append(_full(b"infe", 2, 0, struct.pack(">HH4s", gid, 0, b"grid") + b"\0"))
    iinf = _full(b"iinf", 0, 0, struct.pack(">H", len(infes)) + b"".join(infes))
    hdlr = _full(b"hdlr", 0, 0, struct.pack(">I4s", 0, b"pict") + b"\0" * 12 + b"\0")
    pitm = _full(b"pitm", 0, 0, struct.pack(">H", gid if grid else 1))
    assoc = []
    for i in tile_ids:
        a = [0x80 | 1, 2] + ([colr_idx] if (colr_idx and not grid) else [])
        assoc.append(struct.pack(">HB", i, len(a)) + bytes(a))
    if grid:
        a = [0x80 | 3] + ([colr_idx] if colr_idx else [])
        assoc.append(struct.pack(">HB", gid, len(a)) + bytes(a))
    ipma = _full(b"ipma", 0, 0, struct.pack(">I", len(assoc)) + b"".join(assoc))
    iprp = _box(b"iprp", _box(b"ipco", b"".join(props)) + ipma)
    iref = b""
    idat = b""
    if grid:
        iref = _full(b"iref", 0, 0, _box(b"dimg", struct.pack(">HH", gid, n) + b"".join(struct.pack(">H", i) for i in tile_ids)))
        big = out_w > 65535 or out_h > 65535
        idat = _box(b"idat", struct.pack(">BBBB", 0, 1 if big else 0, rows - 1, cols - 1) + (struct.pack(">II", out_w, out_h) if big else struct.pack(">HH", out_w, out_h)))

    def iloc(offsets):
        ents = []
        for i, off, ln in offsets:
            ents.append(struct.pack(">HHHHII", i, 0, 0, 1, off, ln))
        if grid:
            ents.append(struct.pack(">HHHHII", gid, 1, 0, 1, 0, len(idat) - 8))           # construction_method 1: idat
        return _full(b"iloc", 1, 0, struct.pack(">BBH", 0x44, 0x00, len(ents)) + b"".join(ents))

    ftyp = _box(b"ftyp", b"heic" + struct.pack(">I", 0) + b"mif1heic")
    datas = [r[1] for r in recs]
    dummy = [(i, 0, len(d)) for i, d in zip(tile_ids, datas)]
    meta_len = len(_full(b"meta", 0, 0, hdlr + pitm + iloc(dummy) + iinf + iref + iprp + idat))
    off = len(ftyp) + meta_len + 8
    offs = []
    for i, d in zip(tile_ids, datas):
        offs.append((i, off, len(d))); off += len(d)
    meta = _full(b"meta", 0, 0, hdlr + pitm + iloc(offs) + iinf + iref + iprp + idat)
    assert len(meta) == meta_len
    payload = b"".join(datas)
    if len(payload) + 8 > 0xFFFFFFFF:
        raise ValueError("mdat too large for a 32-bit box")
    with open(path, "wb") as f:
        f.write(ftyp); f.write(meta); f.write(struct.pack(">I4s", 8 + len(payload), b"mdat")); f.write(payload)
    return path
