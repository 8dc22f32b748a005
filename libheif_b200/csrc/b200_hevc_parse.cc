// b200_hevc_parse.cc -- host HEVC front-end: NAL units -> headers -> (optionally) command stream.
//
// Replaces the serial half of what libde265 does behind libheif/plugins/decoder_libde265.cc:322-457
// (de265_push_NAL / de265_decode).  Two stages:
//   parse_headers()      emulation-prevention removal, VPS/SPS/PPS/slice-segment headers (H.265 7.3.1-7.3.6), security
//                        limit, and the split of every slice segment into CABAC sub-streams (entry points, 7.3.6.1).
//   parse_access_unit()  = parse_headers() + slice data decoded on the host with the shared syntax decoder
//                        (b200_hevc_syntax.h), sequentially; the device front-end (b200_hevc_entropy.cu) runs the very
//                        same code with one warp per sub-stream instead.
// Input framing is libheif's: [uint32 BE length][NAL]... (libheif/codecs/decoder.cc:275-308).
#define B200_SYNTAX_HOST_ONLY 1   // this translation unit uses the shared syntax decoder on the host only
#include "b200_hevc.h"
#include <algorithm>

namespace b200 {
namespace {

struct BitRd {
  const uint8_t* d; size_t n; size_t pos;
  // reads past the end return zeros; overrun() reports them, and every bitstream-controlled loop below is bounded by
  // the specification's range for its count, so a short NAL can neither hang nor over-read
  bool overrun() const { return pos > 8 * n; }
  unsigned bit() { unsigned v = (pos >> 3) < n ? (d[pos >> 3] >> (7 - (pos & 7))) & 1 : 0; pos++; return v; }
  unsigned bits(int k) { unsigned v = 0; while (k-- > 0) v = (v << 1) | bit(); return v; }
  // ue(v), 9.2: more than 31 leading zero bits cannot be a valid code; saturate (callers range-check the value)
  unsigned ue() { int z = 0; while (bit() == 0 && z < 32) z++; if (z >= 32) return 0xffffffffu; return z ? ((1u << z) - 1 + bits(z)) : 0; }
  int se() { unsigned k = ue(); return (k & 1) ? (int)((k + 1) >> 1) : -(int)(k >> 1); }
};

struct Sps {
  bool valid = false; int separate_colour_plane = 0; int chroma_format_idc = 1, width = 0, height = 0, conf_l = 0, conf_r = 0, conf_t = 0, conf_b = 0;
  int bit_depth = 8, log2_max_poc_lsb = 4, log2_min_cb = 3, log2_ctb = 4, log2_min_tb = 2, log2_max_tb = 5, max_th_depth_intra = 0;
  int sao = 0, strong_intra = 0, num_st_rps = 0, long_term = 0, num_lt_sps = 0, temporal_mvp = 0;
  int st_num_delta[65] = {0};
  int vui_signal = 0, vui_full_range = 0, vui_colour = 0, vui_cp = 2, vui_tc = 2, vui_mc = 2;
  int scaling_enabled = 0, sl_present = 0; sl::Lists lists;
  int pcm = 0, pcm_bd_y = 8, pcm_bd_c = 8, log2_min_pcm = 3, log2_max_pcm = 3, pcm_lf_disabled = 0;
};
struct Pps {
  bool valid = false; int sps_id = 0, dependent_slices = 0, output_flag_present = 0, num_extra_bits = 0, sign_hiding = 0;
  int init_qp = 26, transform_skip = 0, cu_qp_delta = 0, diff_cu_qp_delta_depth = 0, cb_qp_offset = 0, cr_qp_offset = 0;
  int slice_chroma_qp_offsets = 0, wpp = 0, lf_across_slices = 0, deblock_override_enabled = 0, deblock_disabled = 0;
  int beta_offset = 0, tc_offset = 0, slice_ext_present = 0, log2_sao_scale_luma = 0, log2_sao_scale_chroma = 0;
  int sl_present = 0; sl::Lists lists;
  int tq_bypass = 0;
  int tiles = 0, num_tile_cols = 1, num_tile_rows = 1, uniform_spacing = 1, lf_across_tiles = 1; int col_width[20] = {0}, row_height[22] = {0};
};

inline int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
int ceil_log2(unsigned v) { int r = 0; while ((1u << r) < v) r++; return r; }

// 7.4.2: strip emulation_prevention_three_byte; records the NAL offsets of the removed bytes
// (A 0x03 is an emulation prevention byte iff the two NAL bytes before it are zero: the byte-serial rule "two zeros seen since
// the last removal" says the same, because the byte before a candidate can only be zero if it is not a removed 0x03.  Slice data
// holds one 0x03 per ~256 bytes and hardly any emulation prevention: memchr + block copies run at memory speed, the byte loop
// this replaces took 2/3 of the header stage.)
size_t unescape(const uint8_t* in, size_t n, uint8_t* out, std::vector<uint32_t>* epb) {
  size_t o = 0, copied = 0, from = 2;
  while (from < n) {
    const uint8_t* p = static_cast<const uint8_t*>(memchr(in + from, 3, n - from));
    if (!p) break;
    const size_t k = (size_t)(p - in);
    if (in[k - 1] == 0 && in[k - 2] == 0) {
      memcpy(out + o, in + copied, k - copied); o += k - copied; copied = k + 1;
      if (epb) epb->push_back((uint32_t)k);
    }
    from = k + 1;
  }
  memcpy(out + o, in + copied, n - copied); o += n - copied;
  return o;
}

class HeaderParser {
 public:
  HeaderParser(PictureHeaders& out, const ParseLimits& lim) : P(out), L(lim) {}
  int run(const uint8_t* data, size_t size) {
    P.slices.clear(); P.subs.clear(); P.rbsp.clear(); P.ctu_slice.clear();
    P.rbsp.reserve(size + 64);
    // per-thread scratch for one NAL's RBSP: grow-only and uninitialised on purpose (a fresh 200 KB vector per tile cost a
    // memset plus an mmap / munmap pair)
    struct Scratch { uint8_t* p = nullptr; size_t cap = 0; ~Scratch() { free(p); }
                     uint8_t* data() const { return p; }
                     bool fit(size_t n) { if (n <= cap) return true; uint8_t* q = static_cast<uint8_t*>(realloc(p, n + n / 4)); if (!q) return false; p = q; cap = n + n / 4; return true; } };
    static thread_local Scratch rbsp;
    if (!rbsp.fit(size + 16)) return set_error(B200_E_INVALID, "out of memory");
    size_t p = 0; int rc = B200_OK;
    while (p + 4 <= size && rc == B200_OK) {
      uint32_t n = ((uint32_t)data[p] << 24) | (data[p + 1] << 16) | (data[p + 2] << 8) | data[p + 3];
      p += 4;
      if (n > size - p) return set_error(B200_E_BITSTREAM, "NAL length %u exceeds buffer", n);
      if (n >= 2) {
        int type = (data[p] >> 1) & 0x3f;
        if (type == 33 || type == 34 || (type >= 16 && type <= 21) || type <= 9) {
          epb.clear();
          size_t rn = unescape(data + p, n, rbsp.data(), &epb);
          memset(rbsp.data() + rn, 0, 8);
          if (type == 33) rc = parse_sps(rbsp.data(), rn);
          else if (type == 34) rc = parse_pps(rbsp.data(), rn);
          else if (type <= 9) rc = set_error(B200_E_UNSUPPORTED, "non-IRAP picture (NAL type %d): inter prediction is not supported", type);
          else rc = slice_segment(rbsp.data(), rn, type);
        }
      }
      p += n;
    }
    if (rc != B200_OK) return rc;
    if (!started) return set_error(B200_E_BITSTREAM, "no picture in access unit");
    return finish();
  }

 private:
  PictureHeaders& P; const ParseLimits& L;
  Sps sps_tab[16]; Pps pps_tab[64];
  const Sps* S = nullptr; const Pps* PP = nullptr;
  bool started = false;
  std::vector<uint32_t> epb;
  int slice_idx = -1, slice_addr_rs = 0, slice_qp = 26, sao_luma = 0, sao_chroma = 0, last_seg_sub = -1;
  SliceInfo cur_slice{}; int n_slices = 0;                 // parameters of the current slice; regions (slice x tile) are made from it
  std::vector<int> col_bd, row_bd, ts2rs, rs2ts; std::vector<uint16_t> tile_of;   // 6.5.1: tile boundaries (CTBs), CtbAddrTsToRs / RsToTs, TileId by raster address
  int total = 0;
  struct Seg { int addr; int first_sub; int nsubs; };
  std::vector<Seg> segs;

  static void skip_ptl(BitRd& b, int msl) {
    b.bits(8); b.bits(32); b.bits(4); b.bits(32); b.bits(11); b.bit(); b.bits(8);
    int pp[8], lp[8];
    for (int i = 0; i < msl; i++) { pp[i] = b.bit(); lp[i] = b.bit(); }
    if (msl > 0) for (int i = msl; i < 8; i++) b.bits(2);
    for (int i = 0; i < msl; i++) { if (pp[i]) { b.bits(32); b.bits(32); b.bits(24); } if (lp[i]) b.bits(8); }
  }
  // 7.3.7 st_ref_pic_set; false = malformed (delta_idx_minus1 > idx - 1, or more than 16 pictures: sps_max_dec_pic_buffering <= 16)
  static bool st_rps(BitRd& b, Sps& s, int idx, int num) {
    int inter = idx ? b.bit() : 0;
    if (inter) {
      unsigned di = 1; if (idx == num) { const unsigned v = b.ue(); if (v >= (unsigned)idx) return false; di = v + 1; }
      b.bit(); b.ue();
      const int ref = idx - (int)di;                   // 0 <= ref < idx <= 64
      int cnt = 0;
      for (int j = 0; j <= s.st_num_delta[ref]; j++) { int used = b.bit(), ud = 1; if (!used) ud = b.bit(); if (used || ud) cnt++; }
      if (cnt > 16) return false;
      s.st_num_delta[idx] = cnt;
    } else {
      const unsigned nn = b.ue(), np = b.ue();
      if (nn > 16 || np > 16 || nn + np > 16) return false;
      for (unsigned i = 0; i < nn + np; i++) { b.ue(); b.bit(); }
      s.st_num_delta[idx] = (int)(nn + np);
    }
    return !b.overrun();
  }
  // 7.3.4 scaling_list_data; false = malformed
  static bool scaling_list_data(BitRd& b, sl::Lists& L) {
    sl::set_all_default(L);
    for (int s = 0; s < 4; s++) for (int m = 0; m < 6; m += (s == 3 ? 3 : 1)) {
      if (!b.bit()) {                                             // scaling_list_pred_mode_flag = 0: default list or a copy
        const unsigned delta = b.ue();
        if (delta > (unsigned)(s == 3 ? m / 3 : m)) return false;
        if (delta == 0) sl::set_default(L, s, m);
        else { const int ref = m - (int)delta * (s == 3 ? 3 : 1); memcpy(L.list[s][m], L.list[s][ref], 64); L.dc[s][m] = L.dc[s][ref]; }
      } else {
        int next = 8; const int num = s == 0 ? 16 : 64;
        if (s > 1) { const int dc = b.se(); if (dc < -7 || dc > 247) return false; next = dc + 8; L.dc[s][m] = (uint8_t)next; }
        for (int i = 0; i < num; i++) { const int d = b.se(); if (d < -128 || d > 127) return false; next = (next + d + 256) % 256; L.list[s][m][i] = (uint8_t)next; }
      }
    }
    return !b.overrun();
  }
  static bool skip_hrd(BitRd& b, int msl) {
    int nal = b.bit(), vcl = b.bit(), sub = 0;
    if (nal || vcl) { sub = b.bit(); if (sub) { b.bits(8); b.bits(5); b.bit(); b.bits(5); } b.bits(4); b.bits(4); if (sub) b.bits(4); b.bits(5); b.bits(5); b.bits(5); }
    for (int i = 0; i <= msl; i++) {
      int gen = b.bit(), within = 1, low = 0, cnt = 0;
      if (!gen) within = b.bit();
      if (within) b.ue(); else low = b.bit();
      if (!low) { const unsigned v = b.ue(); if (v > 31) return false; cnt = (int)v; }          // cpb_cnt_minus1: 0..31 (E.3.2)
      for (int k = 0; k < nal + vcl; k++) for (int c = 0; c <= cnt; c++) { b.ue(); b.ue(); if (sub) { b.ue(); b.ue(); } b.bit(); }
      if (b.overrun()) return false;
    }
    return true;
  }
  int parse_sps(const uint8_t* r, size_t n) {                                  // 7.3.2.2
    BitRd b{r, n, 16}; Sps s;
    b.bits(4); int msl = b.bits(3); b.bit();
    skip_ptl(b, msl);
    unsigned id = b.ue(); if (id > 15) return set_error(B200_E_BITSTREAM, "sps id");
    // every ue(v) is range-checked BEFORE it is used in arithmetic (a saturated code is 0xffffffff): 7.4.3.2.1 ranges
    { const unsigned v = b.ue(); if (v > 3) return set_error(B200_E_BITSTREAM, "chroma_format_idc"); s.chroma_format_idc = (int)v; }
    if (s.chroma_format_idc == 3) s.separate_colour_plane = b.bit();
    { const unsigned w = b.ue(), h = b.ue(); if (w == 0 || h == 0 || w > 16384 || h > 16384) return set_error(B200_E_BITSTREAM, "picture size %ux%u", w, h); s.width = (int)w; s.height = (int)h; }
    if (b.bit()) {
      const unsigned cl = b.ue(), cr = b.ue(), ct = b.ue(), cbm = b.ue();
      const uint64_t sub = s.chroma_format_idc == 1 || s.chroma_format_idc == 2 ? 2 : 1, subh = s.chroma_format_idc == 1 ? 2 : 1;
      if (sub * ((uint64_t)cl + cr) >= (uint64_t)s.width || subh * ((uint64_t)ct + cbm) >= (uint64_t)s.height) return set_error(B200_E_BITSTREAM, "conformance window larger than the picture");
      s.conf_l = (int)cl; s.conf_r = (int)cr; s.conf_t = (int)ct; s.conf_b = (int)cbm;
    }
    int bdc;
    { const unsigned bl = b.ue(), bc = b.ue(); if (bl > 8 || bc > 8) return set_error(B200_E_BITSTREAM, "bit depth"); s.bit_depth = 8 + (int)bl; bdc = 8 + (int)bc; }
    { const unsigned v = b.ue(); if (v > 12) return set_error(B200_E_BITSTREAM, "log2_max_pic_order_cnt_lsb"); s.log2_max_poc_lsb = 4 + (int)v; }
    int sub = b.bit();
    for (int i = sub ? 0 : msl; i <= msl; i++) { b.ue(); b.ue(); b.ue(); }
    {
      const unsigned a = b.ue(), d1 = b.ue(), t = b.ue(), d2 = b.ue();
      if (a > 3 || d1 > 3 || a + d1 > 3 || 3 + a + d1 < 4) return set_error(B200_E_BITSTREAM, "coding block size configuration");       // MinCb 8..64, CTB 16..64
      s.log2_min_cb = 3 + (int)a; s.log2_ctb = s.log2_min_cb + (int)d1;
      if (t > 3 || 2 + (int)t >= s.log2_min_cb) return set_error(B200_E_BITSTREAM, "log2_min_luma_transform_block_size");                 // MinTb < MinCb
      s.log2_min_tb = 2 + (int)t;
      if (d2 > 3 || s.log2_min_tb + (int)d2 > std::min(5, s.log2_ctb)) return set_error(B200_E_BITSTREAM, "log2_diff_max_min_luma_transform_block_size");
      s.log2_max_tb = s.log2_min_tb + (int)d2;
    }
    { const unsigned inter = b.ue(), intra = b.ue(); const unsigned mx = (unsigned)(s.log2_ctb - s.log2_min_tb);
      if (inter > mx || intra > mx) return set_error(B200_E_BITSTREAM, "max_transform_hierarchy_depth");
      s.max_th_depth_intra = (int)intra; }
    s.scaling_enabled = b.bit();
    if (s.scaling_enabled) { s.sl_present = b.bit(); if (s.sl_present && !scaling_list_data(b, s.lists)) return set_error(B200_E_BITSTREAM, "sps scaling_list_data"); }
    b.bit(); s.sao = b.bit();
    s.pcm = b.bit();
    if (s.pcm) {                                                               // 7.4.3.2.1: PCM bit depths and coding block sizes
      s.pcm_bd_y = 1 + (int)b.bits(4); s.pcm_bd_c = 1 + (int)b.bits(4);
      const unsigned lo = b.ue(), df = b.ue();
      if (lo > 2 || df > 2) return set_error(B200_E_BITSTREAM, "pcm coding block size");
      s.log2_min_pcm = 3 + (int)lo; s.log2_max_pcm = s.log2_min_pcm + (int)df;
      s.pcm_lf_disabled = b.bit();
      if (s.pcm_bd_y > s.bit_depth || s.pcm_bd_c > bdc) return set_error(B200_E_BITSTREAM, "pcm sample bit depth");
      if (s.log2_min_pcm < std::min(s.log2_min_cb, 5) || s.log2_min_pcm > std::min(s.log2_ctb, 5) || s.log2_max_pcm > std::min(s.log2_ctb, 5))
        return set_error(B200_E_BITSTREAM, "pcm coding block size");
    }
    { const unsigned v = b.ue(); if (v > 64) return set_error(B200_E_BITSTREAM, "num_short_term_ref_pic_sets"); s.num_st_rps = (int)v; }
    for (int i = 0; i < s.num_st_rps; i++) if (!st_rps(b, s, i, s.num_st_rps)) return set_error(B200_E_BITSTREAM, "short-term reference picture set %d", i);
    s.long_term = b.bit();
    if (s.long_term) { const unsigned v = b.ue(); if (v > 32) return set_error(B200_E_BITSTREAM, "num_long_term_ref_pics_sps"); s.num_lt_sps = (int)v; for (int i = 0; i < s.num_lt_sps; i++) { b.bits(s.log2_max_poc_lsb); b.bit(); } }
    s.temporal_mvp = b.bit(); s.strong_intra = b.bit();
    if (b.bit()) {   // VUI (E.2.1)
      if (b.bit()) { if (b.bits(8) == 255) { b.bits(16); b.bits(16); } }
      if (b.bit()) b.bit();
      s.vui_signal = b.bit();
      if (s.vui_signal) { b.bits(3); s.vui_full_range = b.bit(); s.vui_colour = b.bit(); if (s.vui_colour) { s.vui_cp = b.bits(8); s.vui_tc = b.bits(8); s.vui_mc = b.bits(8); } }
      if (b.bit()) { b.ue(); b.ue(); }
      b.bit(); b.bit(); b.bit();
      if (b.bit()) { b.ue(); b.ue(); b.ue(); b.ue(); }
      if (b.bit()) { b.bits(32); b.bits(32); if (b.bit()) b.ue(); if (b.bit() && !skip_hrd(b, msl)) return set_error(B200_E_BITSTREAM, "hrd_parameters"); }
      if (b.bit()) { b.bits(3); b.ue(); b.ue(); b.ue(); b.ue(); b.ue(); }
    }
    if (b.bit()) {   // sps_extension
      int range = b.bit(); b.bits(7);
      if (range) { int f[9]; for (int i = 0; i < 9; i++) f[i] = b.bit(); if (f[0] || f[1] || f[2] || f[4] || f[5] || f[7] || f[8]) return set_error(B200_E_UNSUPPORTED, "range-extension coding tools are not supported"); }
    }
    if (s.chroma_format_idc == 3 && s.separate_colour_plane) return set_error(B200_E_UNSUPPORTED, "separate_colour_plane_flag");
    if (s.bit_depth != bdc || s.bit_depth > 12) return set_error(B200_E_UNSUPPORTED, "bit depth luma %d chroma %d", s.bit_depth, bdc);
    if (b.overrun()) return set_error(B200_E_BITSTREAM, "sequence parameter set is truncated");
    if (s.log2_ctb > 6 || s.log2_ctb < 4 || s.log2_max_tb > 5 || s.log2_min_cb > s.log2_ctb || s.log2_max_tb > s.log2_ctb) return set_error(B200_E_BITSTREAM, "block size configuration");
    if (s.width <= 0 || s.height <= 0 || s.width > 16384 || s.height > 16384 || (s.width & ((1 << s.log2_min_cb) - 1)) || (s.height & ((1 << s.log2_min_cb) - 1)))
      return set_error(B200_E_BITSTREAM, "picture size %dx%d", s.width, s.height);
    s.valid = true; sps_tab[id] = s;
    return B200_OK;
  }
  int parse_pps(const uint8_t* r, size_t n) {                                  // 7.3.2.3
    BitRd b{r, n, 16}; Pps p;
    unsigned id = b.ue(); if (id > 63) return set_error(B200_E_BITSTREAM, "pps id");
    { const unsigned v = b.ue(); if (v > 15) return set_error(B200_E_BITSTREAM, "pps sps id"); p.sps_id = (int)v; }
    p.dependent_slices = b.bit(); p.output_flag_present = b.bit(); p.num_extra_bits = b.bits(3);
    p.sign_hiding = b.bit(); b.bit(); b.ue(); b.ue();
    { const int v = b.se(); if (v < -26 - 24 || v > 25) return set_error(B200_E_BITSTREAM, "init_qp_minus26"); p.init_qp = 26 + v; }      // -(26 + QpBdOffset) .. 25; re-checked against the SPS bit depth per slice
    b.bit();
    p.transform_skip = b.bit(); p.cu_qp_delta = b.bit();
    if (p.cu_qp_delta) { const unsigned v = b.ue(); if (v > 3) return set_error(B200_E_BITSTREAM, "diff_cu_qp_delta_depth"); p.diff_cu_qp_delta_depth = (int)v; }
    p.cb_qp_offset = b.se(); p.cr_qp_offset = b.se(); p.slice_chroma_qp_offsets = b.bit();
    if (p.cb_qp_offset < -12 || p.cb_qp_offset > 12 || p.cr_qp_offset < -12 || p.cr_qp_offset > 12) return set_error(B200_E_BITSTREAM, "pps chroma qp offset");
    b.bit(); b.bit();
    p.tq_bypass = b.bit();
    p.tiles = b.bit();
    p.wpp = b.bit();
    if (p.tiles) {                                                            // 7.3.2.3 / 7.4.3.3
      if (p.wpp) return set_error(B200_E_UNSUPPORTED, "HEVC tiles together with wavefront parallel processing are not supported");
      const unsigned nc = b.ue(), nr = b.ue();
      if (nc > 19 || nr > 21) return set_error(B200_E_BITSTREAM, "number of tile columns / rows");
      p.num_tile_cols = (int)nc + 1; p.num_tile_rows = (int)nr + 1;
      p.uniform_spacing = b.bit();
      if (!p.uniform_spacing) {
        for (int i = 0; i + 1 < p.num_tile_cols; i++) { const unsigned v = b.ue(); if (v > 4096) return set_error(B200_E_BITSTREAM, "tile column width"); p.col_width[i] = (int)v + 1; }
        for (int i = 0; i + 1 < p.num_tile_rows; i++) { const unsigned v = b.ue(); if (v > 4096) return set_error(B200_E_BITSTREAM, "tile row height"); p.row_height[i] = (int)v + 1; }
      }
      p.lf_across_tiles = b.bit();
    }
    p.lf_across_slices = b.bit();
    if (b.bit()) { p.deblock_override_enabled = b.bit(); p.deblock_disabled = b.bit(); if (!p.deblock_disabled) { const int be = b.se(), tc = b.se(); if (be < -6 || be > 6 || tc < -6 || tc > 6) return set_error(B200_E_BITSTREAM, "pps deblocking offsets"); p.beta_offset = 2 * be; p.tc_offset = 2 * tc; } }
    p.sl_present = b.bit();
    if (p.sl_present && !scaling_list_data(b, p.lists)) return set_error(B200_E_BITSTREAM, "pps scaling_list_data");
    b.bit(); b.ue(); p.slice_ext_present = b.bit();
    if (b.bit()) {
      int range = b.bit(); b.bits(7);
      if (range) {
        if (p.transform_skip && b.ue() != 0) return set_error(B200_E_UNSUPPORTED, "transform skip larger than 4x4");
        if (b.bit()) return set_error(B200_E_UNSUPPORTED, "cross-component prediction");
        if (b.bit()) return set_error(B200_E_UNSUPPORTED, "chroma QP offset lists");
        const unsigned sl = b.ue(), sc = b.ue();               // log2_sao_offset_scale_*: 0 .. max(0, BitDepth - 10) (7.4.3.3.2)
        if (sl > 6 || sc > 6) return set_error(B200_E_BITSTREAM, "log2_sao_offset_scale out of range");
        p.log2_sao_scale_luma = (int)sl; p.log2_sao_scale_chroma = (int)sc;
      }
    }
    if (b.overrun()) return set_error(B200_E_BITSTREAM, "picture parameter set is truncated");
    p.valid = true; pps_tab[id] = p;
    return B200_OK;
  }

  int start_picture() {
    const int W = S->width, H = S->height, log2ctb = S->log2_ctb, ctb = 1 << log2ctb;
    if (L.max_image_size_pixels && (uint64_t)W * H > L.max_image_size_pixels)
      return set_error(B200_E_LIMIT, "coded picture %dx%d exceeds the security limit of %llu pixels", W, H, (unsigned long long)L.max_image_size_pixels);
    PicDesc& d = P.desc; memset(&d, 0, sizeof d);
    d.width = W; d.height = H; d.log2_ctb = log2ctb; d.wctb = (W + ctb - 1) >> log2ctb; d.hctb = (H + ctb - 1) >> log2ctb;
    d.bit_depth = S->bit_depth; d.chroma = S->chroma_format_idc;
    const int sub = (d.chroma == 1 || d.chroma == 2) ? 2 : 1, subh = d.chroma == 1 ? 2 : 1;      // SubWidthC, SubHeightC: conformance window units
    d.crop_x = S->conf_l * sub; d.crop_y = S->conf_t * subh;
    d.out_w = W - (S->conf_l + S->conf_r) * sub; d.out_h = H - (S->conf_t + S->conf_b) * subh;
    if (d.out_w <= 0 || d.out_h <= 0) return set_error(B200_E_BITSTREAM, "conformance window");
    d.strong_intra = S->strong_intra; d.sao_enabled = S->sao; d.w8 = W >> 3; d.h8 = H >> 3;
    d.scaling_idx = -1;
    P.scaling_enabled = S->scaling_enabled != 0;
    if (P.scaling_enabled) {                                      // 7.4.5: PPS lists, else SPS lists, else the default lists
      if (PP->sl_present) sl::derive(PP->lists, P.scaling);
      else if (S->sl_present) sl::derive(S->lists, P.scaling);
      else { sl::Lists L; sl::set_all_default(L); sl::derive(L, P.scaling); }
    }
    total = d.wctb * d.hctb;
    P.ctu_slice.assign((size_t)total, 0xffff);
    { // 6.5.1: tile boundaries and the conversion between raster and tile scan (identity without tiles)
      const int ntc = PP->tiles ? PP->num_tile_cols : 1, ntr = PP->tiles ? PP->num_tile_rows : 1;
      if (ntc > d.wctb || ntr > d.hctb) return set_error(B200_E_BITSTREAM, "more tile columns / rows than CTBs");
      col_bd.assign(1, 0); row_bd.assign(1, 0);
      for (int i = 0; i < ntc; i++) { const int w = (!PP->tiles || PP->uniform_spacing) ? ((i + 1) * d.wctb) / ntc - (i * d.wctb) / ntc : (i + 1 < ntc ? PP->col_width[i] : d.wctb - col_bd[(size_t)i]);
        if (w <= 0 || col_bd[(size_t)i] + w > d.wctb) return set_error(B200_E_BITSTREAM, "tile column widths"); col_bd.push_back(col_bd[(size_t)i] + w); }
      for (int i = 0; i < ntr; i++) { const int h = (!PP->tiles || PP->uniform_spacing) ? ((i + 1) * d.hctb) / ntr - (i * d.hctb) / ntr : (i + 1 < ntr ? PP->row_height[i] : d.hctb - row_bd[(size_t)i]);
        if (h <= 0 || row_bd[(size_t)i] + h > d.hctb) return set_error(B200_E_BITSTREAM, "tile row heights"); row_bd.push_back(row_bd[(size_t)i] + h); }
      if (col_bd.back() != d.wctb || row_bd.back() != d.hctb) return set_error(B200_E_BITSTREAM, "tiles do not cover the picture");
      ts2rs.assign((size_t)total, 0); rs2ts.assign((size_t)total, 0); tile_of.assign((size_t)total, 0);
      int ts = 0;
      for (int tr = 0; tr < ntr; tr++) for (int tc = 0; tc < ntc; tc++)
        for (int y = row_bd[(size_t)tr]; y < row_bd[(size_t)tr + 1]; y++) for (int x = col_bd[(size_t)tc]; x < col_bd[(size_t)tc + 1]; x++) {
          const int rs = y * d.wctb + x; ts2rs[(size_t)ts] = rs; rs2ts[(size_t)rs] = ts; tile_of[(size_t)rs] = (uint16_t)(tr * ntc + tc); ts++; }
    }
    P.colour_primaries = S->vui_colour ? S->vui_cp : 2; P.transfer_characteristics = S->vui_colour ? S->vui_tc : 2;
    P.matrix_coefficients = S->vui_colour ? S->vui_mc : 2; P.full_range = S->vui_signal ? S->vui_full_range : 0;
    started = true;
    return B200_OK;
  }

  void fill_seq_params() {
    syn::SeqParams& q = P.sp; memset(&q, 0, sizeof q);
    const PicDesc& d = P.desc;
    q.W = d.width; q.H = d.height; q.log2ctb = d.log2_ctb; q.wctb = d.wctb; q.hctb = d.hctb; q.w4 = d.width >> 2; q.w8 = d.w8; q.h8 = d.h8;
    q.chroma = d.chroma; q.bd = d.bit_depth;
    q.log2_min_cb = S->log2_min_cb; q.log2_min_tb = S->log2_min_tb; q.log2_max_tb = S->log2_max_tb; q.max_th_depth_intra = S->max_th_depth_intra;
    q.sao_enabled = S->sao; q.transform_skip = PP->transform_skip; q.cu_qp_delta = PP->cu_qp_delta; q.qg_log2 = d.log2_ctb - PP->diff_cu_qp_delta_depth;
    q.pcm = S->pcm; q.pcm_shift_y = d.bit_depth - S->pcm_bd_y; q.pcm_shift_c = d.bit_depth - S->pcm_bd_c; q.pcm_bd_y = S->pcm_bd_y; q.pcm_bd_c = S->pcm_bd_c;
    q.log2_min_pcm = S->log2_min_pcm; q.log2_max_pcm = S->log2_max_pcm; q.pcm_lf_disabled = S->pcm_lf_disabled; q.tq_bypass = PP->tq_bypass;
    q.tiles = PP->tiles;
    q.sign_hiding = PP->sign_hiding; q.wpp = PP->wpp; q.sao_scale_luma = PP->log2_sao_scale_luma; q.sao_scale_chroma = PP->log2_sao_scale_chroma;
    const int ctb = 1 << d.log2_ctb;
    q.tu_slots = (ctb / 4) * (ctb / 4) * (d.chroma >= 2 ? 3 : 1);          // 4:2:2 / 4:4:4: every chroma block is a command of its own
    q.coef_slots = d.chroma == 3 ? ctb * ctb * 3 : (d.chroma == 2 ? ctb * ctb * 2 : ctb * ctb * (d.chroma ? 3 : 2) / 2);
  }

  // 7.3.6.1 slice_segment_header; then the split of the segment data into sub-streams
  int slice_segment(const uint8_t* r, size_t n, int nal_type) {
    BitRd b{r, n, 16};
    int first = b.bit();
    if (nal_type >= 16 && nal_type <= 23) b.bit();
    unsigned pid = b.ue();
    if (pid > 63 || !pps_tab[pid].valid || !sps_tab[pps_tab[pid].sps_id].valid) return set_error(B200_E_BITSTREAM, "slice refers to missing parameter sets");
    const Pps* p = &pps_tab[pid]; const Sps* s = &sps_tab[p->sps_id];
    if (first) {
      if (started) return set_error(B200_E_UNSUPPORTED, "more than one picture in the access unit");
      S = s; PP = p; int rc = start_picture(); if (rc) return rc;
      fill_seq_params();
      if (P.sp.qg_log2 < 3) return set_error(B200_E_BITSTREAM, "diff_cu_qp_delta_depth");
      { const int mx = P.sp.bd > 10 ? P.sp.bd - 10 : 0;
        if (P.sp.sao_scale_luma > mx || P.sp.sao_scale_chroma > mx) return set_error(B200_E_BITSTREAM, "log2_sao_offset_scale exceeds BitDepth - 10"); }
    } else if (!started) return set_error(B200_E_BITSTREAM, "slice segment before the first one of the picture");
    else if (p != PP) return set_error(B200_E_UNSUPPORTED, "slice segments of one picture use different PPS");
    P.desc.pps_cb_qp_offset = p->cb_qp_offset; P.desc.pps_cr_qp_offset = p->cr_qp_offset;
    P.desc.log2_sao_scale_luma = p->log2_sao_scale_luma; P.desc.log2_sao_scale_chroma = p->log2_sao_scale_chroma;
    const int wctb = P.desc.wctb;
    int dependent = 0, seg_addr = 0;
    if (!first) { if (p->dependent_slices) dependent = b.bit(); seg_addr = b.bits(ceil_log2((unsigned)total)); if (seg_addr >= total) return set_error(B200_E_BITSTREAM, "slice_segment_address"); }
    const int seg_ts = rs2ts[(size_t)seg_addr];                   // segments are consecutive in TILE scan (6.5.1); all bookkeeping below is in that scan
    if (!segs.empty() && seg_ts <= segs.back().addr) return set_error(B200_E_BITSTREAM, "slice segments out of order");
    if (!dependent) {
      b.bits(p->num_extra_bits);
      if (b.ue() != 2) return set_error(B200_E_UNSUPPORTED, "P/B slices are not supported (intra-only decoder)");
      if (p->output_flag_present) b.bit();
      if (nal_type != 19 && nal_type != 20) {
        b.bits(S->log2_max_poc_lsb);
        if (!b.bit()) { Sps tmp = *S; if (!st_rps(b, tmp, S->num_st_rps, S->num_st_rps)) return set_error(B200_E_BITSTREAM, "slice short-term reference picture set"); }
        else if (S->num_st_rps > 1) b.bits(ceil_log2((unsigned)S->num_st_rps));
        if (S->long_term) {
          unsigned nsps = 0; if (S->num_lt_sps > 0) nsps = b.ue();
          const unsigned npics = b.ue();
          if (nsps > 32 || npics > 32) return set_error(B200_E_BITSTREAM, "long-term reference picture count");
          for (unsigned i = 0; i < nsps + npics; i++) {
            if (i < nsps) { if (S->num_lt_sps > 1) b.bits(ceil_log2((unsigned)S->num_lt_sps)); } else { b.bits(S->log2_max_poc_lsb); b.bit(); }
            if (b.bit()) b.ue();
          }
        }
        if (S->temporal_mvp) b.bit();
      }
      sao_luma = sao_chroma = 0;
      if (S->sao) { sao_luma = b.bit(); if (P.desc.chroma) sao_chroma = b.bit(); }
      { const int dq = b.se(), qbd = 6 * (S->bit_depth - 8);
        if (dq < -128 || dq > 128 || p->init_qp + dq < -qbd || p->init_qp + dq > 51) return set_error(B200_E_BITSTREAM, "SliceQpY %d outside [%d, 51]", p->init_qp + dq, -qbd);   // 7.4.7.1
        slice_qp = p->init_qp + dq; }
      int cb_off = 0, cr_off = 0;
      if (p->slice_chroma_qp_offsets) { cb_off = b.se(); cr_off = b.se(); if (cb_off < -12 || cb_off > 12 || cr_off < -12 || cr_off > 12) return set_error(B200_E_BITSTREAM, "slice chroma qp offset"); }
      int dis = p->deblock_disabled, beta = p->beta_offset, tc = p->tc_offset, ovr = 0;
      if (p->deblock_override_enabled) ovr = b.bit();
      if (ovr) { dis = b.bit(); if (!dis) { const int be = b.se(), t2 = b.se(); if (be < -6 || be > 6 || t2 < -6 || t2 > 6) return set_error(B200_E_BITSTREAM, "slice deblocking offsets"); beta = 2 * be; tc = 2 * t2; } }
      int across = p->lf_across_slices;
      if (p->lf_across_slices && (sao_luma || sao_chroma || !dis)) across = b.bit();
      if (P.slices.size() >= 65000) return set_error(B200_E_UNSUPPORTED, "too many slices");
      SliceInfo& si = cur_slice; si = SliceInfo{}; si.cb_qp_offset = (int8_t)clip3(-24, 24, p->cb_qp_offset + cb_off); si.cr_qp_offset = (int8_t)clip3(-24, 24, p->cr_qp_offset + cr_off);
      si.beta_offset = (int8_t)clip3(-12, 12, beta); si.tc_offset = (int8_t)clip3(-12, 12, tc);
      si.deblocking_disabled = (uint8_t)dis; si.lf_across_slices = (uint8_t)across; si.first_ctb_rs = (uint32_t)seg_addr;
      si.slice_id = (uint16_t)n_slices++; si.lf_across_tiles = (uint8_t)p->lf_across_tiles;
      slice_idx = 0; slice_addr_rs = seg_addr;
    } else if (slice_idx < 0) return set_error(B200_E_BITSTREAM, "dependent slice segment without a slice");
    std::vector<uint32_t> entry;
    if (p->wpp || p->tiles) { const unsigned ne = b.ue(); if (ne > (unsigned)total) return set_error(B200_E_BITSTREAM, "num_entry_point_offsets"); if (ne > 0) { const unsigned lm1 = b.ue(); if (lm1 > 31) return set_error(B200_E_BITSTREAM, "offset_len_minus1"); const int len = (int)lm1 + 1; for (unsigned i = 0; i < ne; i++) { const uint64_t v = (uint64_t)b.bits(len) + 1; if (v > n) return set_error(B200_E_BITSTREAM, "entry point offset beyond the NAL"); entry.push_back((uint32_t)v); } } }
    if (p->slice_ext_present) { const unsigned len = b.ue(); if (len > 256) return set_error(B200_E_BITSTREAM, "slice_segment_header_extension_length"); for (unsigned i = 0; i < len; i++) b.bits(8); }
    b.bit(); b.pos = (b.pos + 7) & ~(size_t)7;
    if (b.overrun()) return set_error(B200_E_BITSTREAM, "slice segment header is truncated");
    const size_t hdr_rbsp = b.pos >> 3;
    if (hdr_rbsp > n) return set_error(B200_E_BITSTREAM, "slice header runs past the NAL");
    // ---- append the segment's data to the picture's RBSP buffer (4-byte aligned start)
    while (P.rbsp.size() & 3) P.rbsp.push_back(0);
    const uint32_t base = (uint32_t)P.rbsp.size();
    P.rbsp.insert(P.rbsp.end(), r + hdr_rbsp, r + n);
    const uint32_t data_len = (uint32_t)(n - hdr_rbsp);
    // ---- sub-streams: with WPP one per CTB row of the segment, located by the entry points (NAL offsets include the
    // emulation prevention bytes, 7.4.7.1: convert to RBSP offsets)
    size_t hdr_epb = 0;
    { size_t nal_pos = 0, k = 0; // number of removed bytes inside the header: NAL position of RBSP byte hdr_rbsp
      for (nal_pos = hdr_rbsp; k < epb.size() && epb[k] < nal_pos + 1; k++) nal_pos++;
      hdr_epb = k; }
    const size_t hdr_nal = hdr_rbsp + hdr_epb;
    Seg sg; sg.addr = seg_ts; sg.first_sub = (int)P.subs.size(); sg.nsubs = 0;
    auto add_sub = [&](uint32_t cb, uint32_t ce, uint32_t byte_begin, bool first_sub) {
      syn::Substream ss{}; ss.pic = 0; ss.byte_begin = base + byte_begin; ss.byte_end = base + data_len; ss.ctb_begin = cb; ss.ctb_end = ce;
      // region = this slice inside the tile the sub-stream starts in (a sub-stream never leaves its tile)
      { const uint16_t tid = tile_of[(size_t)ts2rs[(size_t)cb]];
        if (P.slices.empty() || P.slices.back().slice_id != cur_slice.slice_id || P.slices.back().tile_id != tid) { SliceInfo r = cur_slice; r.tile_id = tid; P.slices.push_back(r); }
        slice_idx = (int)P.slices.size() - 1; }
      ss.slice_addr_rs = (uint32_t)slice_addr_rs; ss.slice_idx = slice_idx; ss.slice_qp = slice_qp; ss.sao_luma = (uint8_t)sao_luma; ss.sao_chroma = (uint8_t)sao_chroma;
      ss.init_contexts = (uint8_t)(first_sub && !dependent); ss.last_of_segment = 0;
      ss.prev = (first_sub && dependent) ? last_seg_sub : -1;
      ss.wake_ctb2 = ss.wake_end = -1; ss.deps = 0;
      P.subs.push_back(ss); sg.nsubs++;
    };
    if (p->wpp) {
      // rows covered by this segment are only known once the next segment's address is: create row sub-streams lazily
      // from the entry points (row k of the segment <-> entry k-1)
      uint32_t nal_off = 0;
      uint32_t cb = (uint32_t)seg_addr;
      for (size_t k = 0; k <= entry.size(); k++) {
        const uint32_t row_end = (cb / (uint32_t)wctb + 1) * (uint32_t)wctb;
        size_t abs_nal = hdr_nal + nal_off, cnt = 0;
        while (cnt < epb.size() && epb[cnt] < abs_nal) cnt++;
        const uint32_t rb = (uint32_t)(abs_nal - cnt - hdr_rbsp);
        if (rb > data_len) return set_error(B200_E_BITSTREAM, "entry point beyond the slice segment data");
        add_sub(cb, row_end, rb, k == 0);
        if (k < entry.size()) nal_off += entry[k];
        cb = row_end;
        if (cb >= (uint32_t)total) break;
      }
    } else if (p->tiles) {
      // one sub-stream per tile the segment touches: sub-stream k + 1 starts at the k-th tile boundary after the segment's start
      uint32_t nal_off = 0, cb = (uint32_t)seg_ts;
      for (size_t k = 0; k <= entry.size(); k++) {
        uint32_t te = cb + 1;
        while (te < (uint32_t)total && tile_of[(size_t)ts2rs[te]] == tile_of[(size_t)ts2rs[cb]]) te++;       // end of this tile in tile scan
        size_t abs_nal = hdr_nal + nal_off, cnt = 0;
        while (cnt < epb.size() && epb[cnt] < abs_nal) cnt++;
        const uint32_t rb = (uint32_t)(abs_nal - cnt - hdr_rbsp);
        if (rb > data_len) return set_error(B200_E_BITSTREAM, "entry point beyond the slice segment data");
        add_sub(cb, te, rb, k == 0);
        if (k > 0) { syn::Substream& ss = P.subs.back(); ss.init_contexts = 1; ss.prev = -1; }        // 9.3.1: the first CTB of a tile initialises the context variables
        if (k < entry.size()) nal_off += entry[k];
        cb = te;
        if (cb >= (uint32_t)total) break;
      }
      // (a segment that begins at a tile's first CTB initialises its contexts there, too -- also a dependent one)
      { syn::Substream& f = P.subs[(size_t)sg.first_sub]; const int rs0 = ts2rs[(size_t)seg_ts];
        if (rs0 % wctb == col_bd[(size_t)(tile_of[(size_t)rs0] % (uint16_t)(col_bd.size() - 1))] && (seg_ts == 0 || tile_of[(size_t)ts2rs[(size_t)seg_ts - 1]] != tile_of[(size_t)rs0])) { f.init_contexts = 1; f.prev = -1; } }
    } else add_sub((uint32_t)seg_ts, (uint32_t)total, 0, true);
    segs.push_back(sg);
    last_seg_sub = (int)P.subs.size() - 1;
    return B200_OK;
  }

  // Close the segments (each ends where the next begins), fix the CTB ranges of the sub-streams, fill ctu_slice.
  int finish() {
    for (size_t si = 0; si < segs.size(); si++) {
      const uint32_t seg_end = si + 1 < segs.size() ? (uint32_t)segs[si + 1].addr : (uint32_t)total;
      Seg& sg = segs[si];
      int keep = 0;
      for (int k = 0; k < sg.nsubs; k++) {
        syn::Substream& ss = P.subs[(size_t)sg.first_sub + k];
        if (ss.ctb_begin >= seg_end) break;
        if (ss.ctb_end > seg_end) ss.ctb_end = seg_end;
        keep++;
      }
      if (keep == 0) return set_error(B200_E_BITSTREAM, "empty slice segment");
      if (PP->wpp) {
        // every CTB row of the segment needs its own entry point
        const syn::Substream& lastk = P.subs[(size_t)sg.first_sub + keep - 1];
        if (lastk.ctb_end != seg_end) return set_error(B200_E_BITSTREAM, "missing entry points for WPP rows");
      }
      for (int k = keep; k < sg.nsubs; k++) P.subs[(size_t)sg.first_sub + k].ctb_begin = P.subs[(size_t)sg.first_sub + k].ctb_end = 0;   // unused
      P.subs[(size_t)sg.first_sub + keep - 1].last_of_segment = 1;
      sg.nsubs = keep;
      for (int k = 0; k < keep; k++) {
        const syn::Substream& ss = P.subs[(size_t)sg.first_sub + k];
        for (uint32_t a = ss.ctb_begin; a < ss.ctb_end; a++) P.ctu_slice[(size_t)ts2rs[a]] = (uint16_t)ss.slice_idx;
      }
    }
    // drop unused sub-streams (entry points past the segment end) while keeping `prev` links valid
    std::vector<int> remap(P.subs.size(), -1); std::vector<syn::Substream> out;
    for (size_t i = 0; i < P.subs.size(); i++) if (P.subs[i].ctb_end > P.subs[i].ctb_begin) { remap[i] = (int)out.size(); out.push_back(P.subs[i]); }
    for (auto& ss : out) if (ss.prev >= 0) { int pr = ss.prev; while (pr >= 0 && remap[(size_t)pr] < 0) pr--; ss.prev = pr >= 0 ? remap[(size_t)pr] : -1; }
    P.subs.swap(out);
    // tile scan -> what the syntax decoder walks: raster address of the first CTB, CTB count, CTB columns of its tile
    for (auto& ss : P.subs) {
      const uint32_t cnt = ss.ctb_end - ss.ctb_begin; const int rs0 = ts2rs[ss.ctb_begin];
      const int tc = (int)(tile_of[(size_t)rs0] % (uint16_t)(col_bd.size() - 1));
      ss.tile_x0 = (uint16_t)col_bd[(size_t)tc]; ss.tile_x1 = (uint16_t)col_bd[(size_t)tc + 1];
      ss.ctb_begin = (uint32_t)rs0; ss.ctb_end = (uint32_t)rs0 + cnt;
    }
    if (segs.empty() || segs[0].addr != 0) return set_error(B200_E_BITSTREAM, "first slice segment missing");
    for (int a = 0; a < total; a++) if (P.ctu_slice[(size_t)a] == 0xffff) return set_error(B200_E_BITSTREAM, "picture incomplete (missing slice segments)");
    P.desc.nslices = (int)P.slices.size();
    while (P.rbsp.size() & 3) P.rbsp.push_back(0);
    for (int k = 0; k < 16; k++) P.rbsp.push_back(0);              // zero tail the CABAC refill may read (even size)
    return B200_OK;
  }
};

struct HostSync {      // sequential execution: every dependency is already satisfied
  uint32_t dense_tu = 0, dense_coef = 0, dense_tu_cap = 0, dense_coef_cap = 0;
  uint64_t end_bit_position = 0;
  int err = 0;
  B200_HD void wait_row(int, int) {}
  B200_HD void publish_row(int, int) {}
  B200_HD void wait_substream(int) {}
  B200_HD void finish_substream(int, int e) { if (e && !err) err = e; }
  B200_HD void notify(int) {}
};

}  // namespace

int parse_headers(const uint8_t* data, size_t size, const ParseLimits& limits, PictureHeaders& out) {
  if (!data || size < 6) return set_error(B200_E_BITSTREAM, "empty access unit");
  HeaderParser p(out, limits);
  return p.run(data, size);
}

int parse_access_unit(const uint8_t* data, size_t size, const ParseLimits& limits, ParsedPicture& out) {
  int rc = parse_headers(data, size, limits, out.hdr);
  if (rc) return rc;
  const PictureHeaders& H = out.hdr;
  out.desc = H.desc;
  syn::SeqParams sp = H.sp; sp.dense = 1;
  const size_t nctb = (size_t)H.desc.wctb * H.desc.hctb, n8 = (size_t)H.desc.w8 * H.desc.h8, n4 = n8 * 4;
  const size_t tu_cap = (size_t)H.desc.width * H.desc.height / 16 * (H.desc.chroma >= 2 ? 3 : 1), coef_cap = (size_t)H.desc.width * H.desc.height * (H.desc.chroma == 3 ? 6 : (H.desc.chroma == 2 ? 4 : (H.desc.chroma ? 3 : 2))) / 2;
  if (out.ctus.size() < nctb) out.ctus.resize(nctb);
  if (out.tus.size() < tu_cap) out.tus.resize(tu_cap);
  if (out.coefs.size() < coef_cap) out.coefs.resize(coef_cap);
  if (out.qp8.size() < n8) out.qp8.resize(n8);
  if (out.edge8.size() < n8) out.edge8.resize(n8);
  if (out.ipm4.size() < n4) out.ipm4.resize(n4);
  if (out.cd8.size() < n8) out.cd8.resize(n8);
  out.wpp_ctx.resize((size_t)H.desc.hctb * syn::CTX_STRIDE);
  out.end_state.resize(H.subs.size() * syn::CTX_STRIDE);
  out.slices = H.slices;
  syn::PicBuffers pb{};
  pb.rbsp = H.rbsp.data(); pb.rbsp_size = (uint32_t)H.rbsp.size();
  pb.tus = out.tus.data(); pb.coefs = out.coefs.data(); pb.ctus = out.ctus.data(); pb.slices = out.slices.data(); pb.ctu_slice = H.ctu_slice.data();
  pb.qp8 = out.qp8.data(); pb.edge8 = out.edge8.data(); pb.ipm4 = out.ipm4.data(); pb.cd8 = out.cd8.data();
  pb.wpp_ctx = out.wpp_ctx.data(); pb.end_state = out.end_state.data();
  for (size_t a = 0; a < nctb; a++) out.ctus[a].slice_idx = H.ctu_slice[a];
  HostSync sync; sync.dense_tu_cap = (uint32_t)tu_cap; sync.dense_coef_cap = (uint32_t)coef_cap;
  syn::U2 ctx[syn::CTX_COUNT];
  for (size_t i = 0; i < H.subs.size(); i++) {
    syn::Decoder dec;
    int e = syn::run_substream<syn::CfgRuntime>(dec, sp, pb, H.subs.data(), (int)i, ctx, sync);
    if (e == syn::SYN_E_OVERFLOW) return set_error(B200_E_BITSTREAM, "slice data produces more transform units / coefficients than the picture can hold");
    if (e) return set_error(B200_E_BITSTREAM, "corrupt slice data (sub-stream %zu, CTB %u..%u)", i, H.subs[i].ctb_begin, H.subs[i].ctb_end);
    // cross-check of the entry points: the next sub-stream of the same segment starts where this one ended
    if (!H.subs[i].last_of_segment && i + 1 < H.subs.size()) {
      const uint64_t next = ((sync.end_bit_position + 7) >> 3);
      if (next != H.subs[i + 1].byte_begin) return set_error(B200_E_BITSTREAM, "entry point offset does not match the end of the previous sub-stream");
    }
  }
  out.n_tus = sync.dense_tu; out.n_coefs = sync.dense_coef;
  return B200_OK;
}

// test hook (tests/test_parser.py): the emulation-prevention removal of the header stage on a bare byte string; returns the RBSP
// length, writes at most epb_cap removed-byte offsets and their total count
int unescape_for_tests(const uint8_t* in, size_t n, uint8_t* out, uint32_t* epb_out, size_t epb_cap, size_t* epb_count) {
  std::vector<uint32_t> epb;
  const size_t o = unescape(in, n, out, &epb);
  for (size_t i = 0; i < epb.size() && i < epb_cap; i++) epb_out[i] = epb[i];
  if (epb_count) *epb_count = epb.size();
  return (int)o;
}

}  // namespace b200

extern "C" int b200_debug_unescape(const uint8_t* in, size_t n, uint8_t* out, uint32_t* epb_out, size_t epb_cap, size_t* epb_count) {
  if (!in || !out) return -1;
  return b200::unescape_for_tests(in, n, out, epb_out, epb_cap, epb_count);
}
