// b200_hevc_parse.cc -- host HEVC front-end: NAL units -> command stream for the sm_100a reconstruction kernels.
//
// Replaces the serial half of what libde265 does behind libheif/plugins/decoder_libde265.cc:322-457
// (de265_push_NAL / de265_decode): emulation-prevention removal, VPS/SPS/PPS/slice-header parsing (H.265 7.3.1-7.3.6),
// CABAC (9.3) and the coding-quadtree syntax (7.3.8), intra-mode derivation (8.4.2) and QP derivation (8.6.1).
// Input framing is libheif's: [uint32 BE length][NAL]... (libheif/codecs/decoder.cc:275-308).
// No pixel is touched here; see b200_hevc_types.h for the division of labour.
#include "b200_hevc.h"
#include <algorithm>

namespace b200 {
namespace {

// ---------------------------------------------------------------------------------------------- bit reader
struct BitRd {
  const uint8_t* d; size_t n; size_t pos;
  unsigned bit() { unsigned v = (pos >> 3) < n ? (d[pos >> 3] >> (7 - (pos & 7))) & 1 : 0; pos++; return v; }
  unsigned bits(int k) { unsigned v = 0; while (k-- > 0) v = (v << 1) | bit(); return v; }
  unsigned ue() { int z = 0; while (bit() == 0 && z < 32) z++; return z ? ((1u << z) - 1 + bits(z)) : 0; }
  int se() { unsigned k = ue(); return (k & 1) ? (int)((k + 1) >> 1) : -(int)(k >> 1); }
};

struct Sps {
  bool valid = false; int chroma_format_idc = 1, width = 0, height = 0, conf_l = 0, conf_r = 0, conf_t = 0, conf_b = 0;
  int bit_depth = 8, log2_max_poc_lsb = 4, log2_min_cb = 3, log2_ctb = 4, log2_min_tb = 2, log2_max_tb = 5, max_th_depth_intra = 0;
  int sao = 0, strong_intra = 0, num_st_rps = 0, long_term = 0, num_lt_sps = 0, temporal_mvp = 0;
  int st_num_delta[65] = {0};
  int vui_signal = 0, vui_full_range = 0, vui_colour = 0, vui_cp = 2, vui_tc = 2, vui_mc = 2;
};
struct Pps {
  bool valid = false; int sps_id = 0, dependent_slices = 0, output_flag_present = 0, num_extra_bits = 0, sign_hiding = 0;
  int init_qp = 26, transform_skip = 0, cu_qp_delta = 0, diff_cu_qp_delta_depth = 0, cb_qp_offset = 0, cr_qp_offset = 0;
  int slice_chroma_qp_offsets = 0, wpp = 0, lf_across_slices = 0, deblock_override_enabled = 0, deblock_disabled = 0;
  int beta_offset = 0, tc_offset = 0, slice_ext_present = 0, log2_sao_scale_luma = 0, log2_sao_scale_chroma = 0;
};

enum { CTX_SAO_MERGE = 0, CTX_SAO_TYPE = 1, CTX_SPLIT_CU = 2, CTX_PART_MODE = 5, CTX_PREV_INTRA = 6,
       CTX_CHROMA_PRED = 7, CTX_SPLIT_TR = 8, CTX_CBF_LUMA = 11, CTX_CBF_CHROMA = 13, CTX_QP_DELTA = 18,
       CTX_TSKIP = 20, CTX_LAST_X = 22, CTX_LAST_Y = 40, CTX_CSBF = 58, CTX_SIG = 62, CTX_GT1 = 104,
       CTX_GT2 = 128, CTX_COUNT = 134 };

const uint8_t kInitI[CTX_COUNT] = {
  153, 200, 139, 141, 157, 184, 184, 63, 153, 138, 138, 111, 141, 94, 138, 182, 154, 154, 154, 154, 139, 139,
  110, 110, 124, 125, 140, 153, 125, 127, 140, 109, 111, 143, 127, 111, 79, 108, 123, 63,
  110, 110, 124, 125, 140, 153, 125, 127, 140, 109, 111, 143, 127, 111, 79, 108, 123, 63,
  91, 171, 134, 141,
  111, 111, 125, 110, 110, 94, 124, 108, 124, 107, 125, 141, 179, 153, 125, 107, 125, 141, 179, 153, 125,
  107, 125, 141, 179, 153, 125, 140, 139, 182, 182, 152, 136, 152, 136, 153, 136, 139, 111, 136, 139, 111,
  140, 92, 137, 138, 140, 152, 138, 139, 153, 74, 149, 92, 139, 107, 122, 152, 140, 179, 166, 182, 140, 227, 122, 197,
  138, 153, 136, 167, 152, 152};
const uint8_t kLps[64][4] = {
  {128,176,208,240},{128,167,197,227},{128,158,187,216},{123,150,178,205},{116,142,169,195},{111,135,160,185},
  {105,128,152,175},{100,122,144,166},{95,116,137,158},{90,110,130,150},{85,104,123,142},{81,99,117,135},
  {77,94,111,128},{73,89,105,122},{69,85,100,116},{66,80,95,110},{62,76,90,104},{59,72,86,99},{56,69,81,94},
  {53,65,77,89},{51,62,73,85},{48,59,69,80},{46,56,66,76},{43,53,63,72},{41,50,59,69},{39,48,56,65},
  {37,45,54,62},{35,43,51,59},{33,41,48,56},{32,39,46,53},{30,37,43,50},{29,35,41,48},{27,33,39,45},
  {26,31,37,43},{24,30,35,41},{23,28,33,39},{22,27,32,37},{21,26,30,35},{20,24,29,33},{19,23,27,31},
  {18,22,26,30},{17,21,25,28},{16,20,23,27},{15,19,22,25},{14,18,21,24},{14,17,20,23},{13,16,19,22},
  {12,15,18,21},{12,14,17,20},{11,14,16,19},{11,13,15,18},{10,12,15,17},{10,12,14,16},{9,11,13,15},
  {9,11,12,14},{8,10,12,14},{8,9,11,13},{7,9,11,12},{7,9,10,12},{7,8,10,11},{6,8,9,11},{6,7,9,10},
  {6,7,8,9},{2,2,2,2}};
const uint8_t kTransLps[64] = {0,0,1,2,2,4,4,5,6,7,8,9,9,11,11,12,13,13,15,15,16,16,18,18,19,19,21,21,22,22,23,24,
  24,25,26,26,27,27,28,29,29,30,30,30,31,32,32,33,33,33,34,34,35,35,35,36,36,36,37,37,37,38,38,63};
const uint8_t kSigMap4[16] = {0, 1, 4, 5, 2, 3, 4, 5, 6, 6, 8, 8, 7, 7, 8, 8};

uint8_t g_sx[4][3][64], g_sy[4][3][64];
bool g_scan_ready = false;
void init_scans() {
  if (g_scan_ready) return;
  for (int l = 0; l <= 3; l++) {
    int n = 1 << l, i = 0, x = 0, y = 0; bool stop = false;
    while (!stop) {
      while (y >= 0) { if (x < n && y < n) { g_sx[l][0][i] = (uint8_t)x; g_sy[l][0][i] = (uint8_t)y; i++; } y--; x++; }
      y = x; x = 0; if (i >= n * n) stop = true;
    }
    i = 0; for (y = 0; y < n; y++) for (x = 0; x < n; x++) { g_sx[l][1][i] = (uint8_t)x; g_sy[l][1][i] = (uint8_t)y; i++; }
    i = 0; for (x = 0; x < n; x++) for (y = 0; y < n; y++) { g_sx[l][2][i] = (uint8_t)x; g_sy[l][2][i] = (uint8_t)y; i++; }
  }
  g_scan_ready = true;
}
struct ScanInit { ScanInit() { init_scans(); } } g_scan_init;

inline int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
int ceil_log2(unsigned v) { int r = 0; while ((1u << r) < v) r++; return r; }

size_t unescape(const uint8_t* in, size_t n, uint8_t* out) {
  size_t o = 0; int zeros = 0;
  for (size_t i = 0; i < n; i++) {
    if (zeros >= 2 && in[i] == 3) { zeros = 0; continue; }
    out[o++] = in[i];
    zeros = in[i] == 0 ? zeros + 1 : 0;
  }
  return o;
}

// ---------------------------------------------------------------------------------------------- CABAC (9.3.4.3)
// Literal 9-bit-offset arithmetic decoder with a 64-bit bit reservoir, so that the bit position after a terminating
// bin is the one the specification defines (needed to find the next WPP sub-stream without trusting entry points).
struct Cabac {
  const uint8_t* d; size_t n; size_t byte_pos; uint64_t res; int avail; unsigned range, offset;
  void start(const uint8_t* data, size_t size, size_t start_byte) { d = data; n = size; byte_pos = start_byte; res = 0; avail = 0; range = 510; offset = take(9); }
  inline void refill() { while (avail <= 56) { uint64_t b = byte_pos < n ? d[byte_pos] : 0; byte_pos++; res |= b << (56 - avail); avail += 8; } }
  inline unsigned take(int k) { if (avail < k) refill(); unsigned v = (unsigned)(res >> (64 - k)); res <<= k; avail -= k; return v; }
  size_t bit_position() const { return byte_pos * 8 - (size_t)avail; }
  inline int bin(uint8_t& c) {
    unsigned state = c >> 1, mps = c & 1;
    unsigned lps = kLps[state][(range >> 6) & 3];
    range -= lps;
    int b;
    if (offset >= range) {
      b = !mps; offset -= range; range = lps;
      if (state == 0) mps ^= 1;
      c = (uint8_t)((kTransLps[state] << 1) | mps);
      int sh = __builtin_clz(range) - 23;
      range <<= sh; offset = (offset << sh) | take(sh);
    } else {
      b = (int)mps;
      if (state < 62) c = (uint8_t)(c + 2);
      if (range < 256) { range <<= 1; offset = (offset << 1) | take(1); }
    }
    return b;
  }
  inline int bypass() { offset = (offset << 1) | take(1); if (offset >= range) { offset -= range; return 1; } return 0; }
  inline unsigned bypass_bits(int k) { unsigned v = 0; while (k-- > 0) v = (v << 1) | (unsigned)bypass(); return v; }
  inline int terminate() {
    range -= 2;
    if (offset >= range) return 1;
    if (range < 256) { range <<= 1; offset = (offset << 1) | take(1); }
    return 0;
  }
  // after a terminating bin == 1 every bit written by the encoder's flush has been consumed: next sub-stream starts
  // at the next byte boundary
  void restart_aligned() { size_t p = (bit_position() + 7) >> 3; start(d, n, p); }
};

void init_ctx(uint8_t* ctx, int slice_qp) {
  int qp = clip3(0, 51, slice_qp);
  for (int i = 0; i < CTX_COUNT; i++) {
    int iv = kInitI[i], m = (iv >> 4) * 5 - 45, nn = ((iv & 15) << 3) - 16;
    int pre = clip3(1, 126, ((m * qp) >> 4) + nn);
    int mps = pre > 63, st = mps ? pre - 64 : 63 - pre;
    ctx[i] = (uint8_t)((st << 1) | mps);
  }
}

struct SaoRaw { int type[3], band[3], eo[3], off[3][4]; };

// ---------------------------------------------------------------------------------------------- parser
class Parser {
 public:
  Parser(ParsedPicture& out, const ParseLimits& lim) : P(out), L(lim) {}
  int run(const uint8_t* data, size_t size) {
    std::vector<uint8_t> rbsp(size + 16);
    size_t p = 0; int rc = B200_OK;
    while (p + 4 <= size && rc == B200_OK) {
      uint32_t n = ((uint32_t)data[p] << 24) | (data[p + 1] << 16) | (data[p + 2] << 8) | data[p + 3];
      p += 4;
      if (n > size - p) return set_error(B200_E_BITSTREAM, "NAL length %u exceeds buffer", n);
      if (n >= 2) {
        int type = (data[p] >> 1) & 0x3f;
        if (type == 33 || type == 34 || (type >= 16 && type <= 21) || type <= 9) {
          size_t rn = unescape(data + p, n, rbsp.data());
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
    for (size_t i = 0; i < slice_of4.size(); i++) if (!slice_of4[i]) return set_error(B200_E_BITSTREAM, "picture incomplete (missing slice segments)");
    finalize();
    return B200_OK;
  }

 private:
  ParsedPicture& P; const ParseLimits& L;
  Sps sps_tab[16]; Pps pps_tab[64];
  const Sps* S = nullptr; const Pps* PP = nullptr;
  bool started = false;
  int W = 0, H = 0, log2ctb = 0, ctb = 0, wctb = 0, hctb = 0, w4 = 0, h4 = 0, w8 = 0, h8 = 0, chroma = 0, bd = 8;
  std::vector<uint16_t> slice_of4; std::vector<uint8_t> ipm4, cd4, edge4; std::vector<int8_t> qp4;
  std::vector<SaoRaw> sao_raw;
  Cabac cabac; uint8_t ctx[CTX_COUNT], ctx_wpp[CTX_COUNT];
  int slice_qp = 26, sao_luma = 0, sao_chroma = 0, slice_idx = -1, slice_addr_rs = 0, cur_cb_off = 0, cur_cr_off = 0;
  int qg_log2 = 3, is_dqp_coded = 0, dqp_val = 0, qpy_prev_qg = 0, last_cu_qpy = 0, first_qg = 1, cur_qpy = 26;
  int err = B200_OK;

  struct Cu { int x0, y0, log2cb, nxn, lmode[4], cmode; };

  bool avail(int x, int y) const {
    if (x < 0 || y < 0 || x >= W || y >= H) return false;
    unsigned s = slice_of4[(size_t)(y >> 2) * w4 + (x >> 2)];
    return s != 0 && s == (unsigned)(slice_idx + 1);
  }

  // -------- parameter sets (7.3.2.2 / 7.3.2.3)
  static void skip_ptl(BitRd& b, int msl) {
    b.bits(8); b.bits(32); b.bits(4); b.bits(32); b.bits(11); b.bit(); b.bits(8);
    int pp[8], lp[8];
    for (int i = 0; i < msl; i++) { pp[i] = b.bit(); lp[i] = b.bit(); }
    if (msl > 0) for (int i = msl; i < 8; i++) b.bits(2);
    for (int i = 0; i < msl; i++) { if (pp[i]) { b.bits(32); b.bits(32); b.bits(24); } if (lp[i]) b.bits(8); }
  }
  static void st_rps(BitRd& b, Sps& s, int idx, int num) {
    int inter = idx ? b.bit() : 0;
    if (inter) {
      int di = 1; if (idx == num) di = b.ue() + 1;
      b.bit(); b.ue();
      int ref = std::max(0, idx - di), cnt = 0;
      for (int j = 0; j <= s.st_num_delta[ref]; j++) { int used = b.bit(), ud = 1; if (!used) ud = b.bit(); if (used || ud) cnt++; }
      s.st_num_delta[idx] = cnt;
    } else { int nn = b.ue(), np = b.ue(); for (int i = 0; i < nn + np; i++) { b.ue(); b.bit(); } s.st_num_delta[idx] = nn + np; }
  }
  static void skip_hrd(BitRd& b, int msl) {
    int nal = b.bit(), vcl = b.bit(), sub = 0;
    if (nal || vcl) { sub = b.bit(); if (sub) { b.bits(8); b.bits(5); b.bit(); b.bits(5); } b.bits(4); b.bits(4); if (sub) b.bits(4); b.bits(5); b.bits(5); b.bits(5); }
    for (int i = 0; i <= msl; i++) {
      int gen = b.bit(), within = 1, low = 0, cnt = 0;
      if (!gen) within = b.bit();
      if (within) b.ue(); else low = b.bit();
      if (!low) cnt = b.ue();
      for (int k = 0; k < nal + vcl; k++) for (int c = 0; c <= cnt; c++) { b.ue(); b.ue(); if (sub) { b.ue(); b.ue(); } b.bit(); }
    }
  }
  int parse_sps(const uint8_t* r, size_t n) {
    BitRd b{r, n, 16}; Sps s;
    b.bits(4); int msl = b.bits(3); b.bit();
    skip_ptl(b, msl);
    unsigned id = b.ue(); if (id > 15) return set_error(B200_E_BITSTREAM, "sps id");
    s.chroma_format_idc = b.ue();
    if (s.chroma_format_idc == 3) b.bit();
    s.width = b.ue(); s.height = b.ue();
    if (b.bit()) { s.conf_l = b.ue(); s.conf_r = b.ue(); s.conf_t = b.ue(); s.conf_b = b.ue(); }
    s.bit_depth = 8 + b.ue(); int bdc = 8 + b.ue();
    s.log2_max_poc_lsb = 4 + b.ue();
    int sub = b.bit();
    for (int i = sub ? 0 : msl; i <= msl; i++) { b.ue(); b.ue(); b.ue(); }
    s.log2_min_cb = 3 + b.ue(); s.log2_ctb = s.log2_min_cb + b.ue();
    s.log2_min_tb = 2 + b.ue(); s.log2_max_tb = s.log2_min_tb + b.ue();
    b.ue(); s.max_th_depth_intra = b.ue();
    if (b.bit()) return set_error(B200_E_UNSUPPORTED, "scaling lists are not supported");
    b.bit(); s.sao = b.bit();
    if (b.bit()) return set_error(B200_E_UNSUPPORTED, "PCM is not supported");
    s.num_st_rps = b.ue(); if (s.num_st_rps > 64) return set_error(B200_E_BITSTREAM, "num_short_term_ref_pic_sets");
    for (int i = 0; i < s.num_st_rps; i++) st_rps(b, s, i, s.num_st_rps);
    s.long_term = b.bit();
    if (s.long_term) { s.num_lt_sps = b.ue(); for (int i = 0; i < s.num_lt_sps; i++) { b.bits(s.log2_max_poc_lsb); b.bit(); } }
    s.temporal_mvp = b.bit(); s.strong_intra = b.bit();
    if (b.bit()) {   // VUI (E.2.1)
      if (b.bit()) { if (b.bits(8) == 255) { b.bits(16); b.bits(16); } }
      if (b.bit()) b.bit();
      s.vui_signal = b.bit();
      if (s.vui_signal) { b.bits(3); s.vui_full_range = b.bit(); s.vui_colour = b.bit(); if (s.vui_colour) { s.vui_cp = b.bits(8); s.vui_tc = b.bits(8); s.vui_mc = b.bits(8); } }
      if (b.bit()) { b.ue(); b.ue(); }
      b.bit(); b.bit(); b.bit();
      if (b.bit()) { b.ue(); b.ue(); b.ue(); b.ue(); }
      if (b.bit()) { b.bits(32); b.bits(32); if (b.bit()) b.ue(); if (b.bit()) skip_hrd(b, msl); }
      if (b.bit()) { b.bits(3); b.ue(); b.ue(); b.ue(); b.ue(); b.ue(); }
    }
    if (b.bit()) {   // sps_extension
      int range = b.bit(); b.bits(7);
      if (range) { int f[9]; for (int i = 0; i < 9; i++) f[i] = b.bit(); if (f[0] || f[1] || f[2] || f[4] || f[5] || f[7] || f[8]) return set_error(B200_E_UNSUPPORTED, "range-extension coding tools are not supported"); }
    }
    if (s.chroma_format_idc > 1) return set_error(B200_E_UNSUPPORTED, "chroma_format_idc %d (only 4:2:0 and 4:0:0)", s.chroma_format_idc);
    if (s.bit_depth != bdc || s.bit_depth > 12) return set_error(B200_E_UNSUPPORTED, "bit depth luma %d chroma %d", s.bit_depth, bdc);
    if (s.log2_ctb > 6 || s.log2_ctb < 4 || s.log2_max_tb > 5 || s.log2_min_cb > s.log2_ctb) return set_error(B200_E_BITSTREAM, "block size configuration");
    if (s.width <= 0 || s.height <= 0 || s.width > 16384 || s.height > 16384 || (s.width & ((1 << s.log2_min_cb) - 1)) || (s.height & ((1 << s.log2_min_cb) - 1)))
      return set_error(B200_E_BITSTREAM, "picture size %dx%d", s.width, s.height);
    s.valid = true; sps_tab[id] = s;
    return B200_OK;
  }
  int parse_pps(const uint8_t* r, size_t n) {
    BitRd b{r, n, 16}; Pps p;
    unsigned id = b.ue(); if (id > 63) return set_error(B200_E_BITSTREAM, "pps id");
    p.sps_id = b.ue(); if (p.sps_id > 15) return set_error(B200_E_BITSTREAM, "pps sps id");
    p.dependent_slices = b.bit(); p.output_flag_present = b.bit(); p.num_extra_bits = b.bits(3);
    p.sign_hiding = b.bit(); b.bit(); b.ue(); b.ue();
    p.init_qp = 26 + b.se(); b.bit();
    p.transform_skip = b.bit(); p.cu_qp_delta = b.bit();
    if (p.cu_qp_delta) p.diff_cu_qp_delta_depth = b.ue();
    p.cb_qp_offset = b.se(); p.cr_qp_offset = b.se(); p.slice_chroma_qp_offsets = b.bit();
    b.bit(); b.bit();
    if (b.bit()) return set_error(B200_E_UNSUPPORTED, "transquant bypass is not supported");
    if (b.bit()) return set_error(B200_E_UNSUPPORTED, "HEVC tiles are not supported");
    p.wpp = b.bit();
    p.lf_across_slices = b.bit();
    if (b.bit()) { p.deblock_override_enabled = b.bit(); p.deblock_disabled = b.bit(); if (!p.deblock_disabled) { p.beta_offset = 2 * b.se(); p.tc_offset = 2 * b.se(); } }
    if (b.bit()) return set_error(B200_E_UNSUPPORTED, "scaling lists are not supported");
    b.bit(); b.ue(); p.slice_ext_present = b.bit();
    if (b.bit()) {
      int range = b.bit(); b.bits(7);
      if (range) {
        if (p.transform_skip && b.ue() != 0) return set_error(B200_E_UNSUPPORTED, "transform skip larger than 4x4");
        if (b.bit()) return set_error(B200_E_UNSUPPORTED, "cross-component prediction");
        if (b.bit()) return set_error(B200_E_UNSUPPORTED, "chroma QP offset lists");
        p.log2_sao_scale_luma = b.ue(); p.log2_sao_scale_chroma = b.ue();
      }
    }
    p.valid = true; pps_tab[id] = p;
    return B200_OK;
  }

  int start_picture() {
    W = S->width; H = S->height; log2ctb = S->log2_ctb; ctb = 1 << log2ctb; chroma = S->chroma_format_idc; bd = S->bit_depth;
    if (L.max_image_size_pixels && (uint64_t)W * H > L.max_image_size_pixels)
      return set_error(B200_E_LIMIT, "coded picture %dx%d exceeds the security limit of %llu pixels", W, H, (unsigned long long)L.max_image_size_pixels);
    wctb = (W + ctb - 1) >> log2ctb; hctb = (H + ctb - 1) >> log2ctb; w4 = W >> 2; h4 = H >> 2; w8 = W >> 3; h8 = H >> 3;
    size_t n4 = (size_t)w4 * h4;
    slice_of4.assign(n4, 0); ipm4.assign(n4, 1); cd4.assign(n4, 0); edge4.assign(n4, 0); qp4.assign(n4, 0);
    sao_raw.assign((size_t)wctb * hctb, SaoRaw{});
    P.ctus.assign((size_t)wctb * hctb, CtuInfo{});
    P.tus.clear(); P.coefs.clear(); P.slices.clear();
    P.tus.reserve((size_t)W * H / 96); P.coefs.reserve((size_t)W * H / 6);
    PicDesc& d = P.desc; memset(&d, 0, sizeof d);
    d.width = W; d.height = H; d.log2_ctb = log2ctb; d.wctb = wctb; d.hctb = hctb; d.bit_depth = bd; d.chroma = chroma;
    int sub = chroma ? 2 : 1;
    d.crop_x = S->conf_l * sub; d.crop_y = S->conf_t * sub;
    d.out_w = W - (S->conf_l + S->conf_r) * sub; d.out_h = H - (S->conf_t + S->conf_b) * sub;
    if (d.out_w <= 0 || d.out_h <= 0) return set_error(B200_E_BITSTREAM, "conformance window");
    d.strong_intra = S->strong_intra; d.sao_enabled = S->sao; d.w8 = w8; d.h8 = h8;
    P.colour_primaries = S->vui_colour ? S->vui_cp : 2; P.transfer_characteristics = S->vui_colour ? S->vui_tc : 2;
    P.matrix_coefficients = S->vui_colour ? S->vui_mc : 2; P.full_range = S->vui_signal ? S->vui_full_range : 0;
    started = true;
    return B200_OK;
  }

  // -------- slice segment (7.3.6.1, 7.3.8.1)
  int slice_segment(const uint8_t* r, size_t n, int nal_type) {
    BitRd b{r, n, 16};
    int first = b.bit();
    if (nal_type >= 16 && nal_type <= 23) b.bit();
    unsigned pid = b.ue();
    if (pid > 63 || !pps_tab[pid].valid || !sps_tab[pps_tab[pid].sps_id].valid) return set_error(B200_E_BITSTREAM, "slice refers to missing parameter sets");
    const Pps* p = &pps_tab[pid]; const Sps* s = &sps_tab[p->sps_id];
    if (first) { if (started) return set_error(B200_E_UNSUPPORTED, "more than one picture in the access unit"); S = s; PP = p; int rc = start_picture(); if (rc) return rc; }
    else if (!started) return set_error(B200_E_BITSTREAM, "slice segment before the first one of the picture");
    PP = p;
    P.desc.pps_cb_qp_offset = p->cb_qp_offset; P.desc.pps_cr_qp_offset = p->cr_qp_offset;
    P.desc.log2_sao_scale_luma = p->log2_sao_scale_luma; P.desc.log2_sao_scale_chroma = p->log2_sao_scale_chroma;
    int dependent = 0, seg_addr = 0, total = wctb * hctb;
    if (!first) { if (p->dependent_slices) dependent = b.bit(); seg_addr = b.bits(ceil_log2((unsigned)total)); if (seg_addr >= total) return set_error(B200_E_BITSTREAM, "slice_segment_address"); }
    if (!dependent) {
      b.bits(p->num_extra_bits);
      if (b.ue() != 2) return set_error(B200_E_UNSUPPORTED, "P/B slices are not supported (intra-only decoder)");
      if (p->output_flag_present) b.bit();
      if (nal_type != 19 && nal_type != 20) {
        b.bits(S->log2_max_poc_lsb);
        if (!b.bit()) { Sps tmp = *S; st_rps(b, tmp, S->num_st_rps, S->num_st_rps); }
        else if (S->num_st_rps > 1) b.bits(ceil_log2((unsigned)S->num_st_rps));
        if (S->long_term) {
          int nsps = 0; if (S->num_lt_sps > 0) nsps = b.ue();
          int npics = b.ue();
          for (int i = 0; i < nsps + npics; i++) {
            if (i < nsps) { if (S->num_lt_sps > 1) b.bits(ceil_log2((unsigned)S->num_lt_sps)); } else { b.bits(S->log2_max_poc_lsb); b.bit(); }
            if (b.bit()) b.ue();
          }
        }
        if (S->temporal_mvp) b.bit();
      }
      sao_luma = sao_chroma = 0;
      if (S->sao) { sao_luma = b.bit(); if (chroma) sao_chroma = b.bit(); }
      slice_qp = p->init_qp + b.se();
      cur_cb_off = cur_cr_off = 0;
      if (p->slice_chroma_qp_offsets) { cur_cb_off = b.se(); cur_cr_off = b.se(); }
      int dis = p->deblock_disabled, beta = p->beta_offset, tc = p->tc_offset, ovr = 0;
      if (p->deblock_override_enabled) ovr = b.bit();
      if (ovr) { dis = b.bit(); if (!dis) { beta = 2 * b.se(); tc = 2 * b.se(); } }
      int across = p->lf_across_slices;
      if (p->lf_across_slices && (sao_luma || sao_chroma || !dis)) across = b.bit();
      if (P.slices.size() >= 65000) return set_error(B200_E_UNSUPPORTED, "too many slices");
      SliceInfo si{}; si.cb_qp_offset = (int8_t)clip3(-24, 24, p->cb_qp_offset + cur_cb_off); si.cr_qp_offset = (int8_t)clip3(-24, 24, p->cr_qp_offset + cur_cr_off);
      si.beta_offset = (int8_t)clip3(-12, 12, beta); si.tc_offset = (int8_t)clip3(-12, 12, tc);
      si.deblocking_disabled = (uint8_t)dis; si.lf_across_slices = (uint8_t)across; si.first_ctb_rs = (uint32_t)seg_addr;
      P.slices.push_back(si);
      slice_idx = (int)P.slices.size() - 1; slice_addr_rs = seg_addr;
    } else if (slice_idx < 0) return set_error(B200_E_BITSTREAM, "dependent slice segment without a slice");
    if (p->wpp) { int ne = b.ue(); if (ne > 0) { int len = b.ue() + 1; for (int i = 0; i < ne; i++) b.bits(len); } }
    if (p->slice_ext_present) { int len = b.ue(); for (int i = 0; i < len; i++) b.bits(8); }
    b.bit(); b.pos = (b.pos + 7) & ~(size_t)7;
    qg_log2 = log2ctb - p->diff_cu_qp_delta_depth;
    if (qg_log2 < 3) return set_error(B200_E_BITSTREAM, "diff_cu_qp_delta_depth");
    if (!dependent) { init_ctx(ctx, slice_qp); last_cu_qpy = slice_qp; first_qg = 1; }
    cabac.start(r, n, b.pos >> 3);
    int a = seg_addr;
    for (;;) {
      int rx = a % wctb, ry = a / wctb;
      if (p->wpp && rx == 0 && (a != seg_addr || (dependent && ry > 0))) {
        if (avail(ctb, (ry - 1) << log2ctb)) memcpy(ctx, ctx_wpp, sizeof ctx);
        else if (a != seg_addr) init_ctx(ctx, slice_qp);
        first_qg = 1;
      }
      CtuInfo& ci = P.ctus[a];
      ci.tu_start = (uint32_t)P.tus.size(); ci.slice_idx = (uint16_t)slice_idx;
      if (S->sao) parse_sao(rx, ry);
      coding_quadtree(rx << log2ctb, ry << log2ctb, log2ctb, 0);
      if (err) return err;
      size_t cnt = P.tus.size() - ci.tu_start;
      ci.tu_count = (uint16_t)cnt;
      if (p->wpp && rx == 1) memcpy(ctx_wpp, ctx, sizeof ctx);
      int end = cabac.terminate();
      a++;
      if (end) break;
      if (a >= total) return set_error(B200_E_BITSTREAM, "slice data runs past the picture");
      if (p->wpp && a % wctb == 0) { if (!cabac.terminate()) return set_error(B200_E_BITSTREAM, "end_of_subset_one_bit"); cabac.restart_aligned(); }
      if (cabac.byte_pos > n + 16) return set_error(B200_E_BITSTREAM, "slice data truncated");
    }
    return B200_OK;
  }

  // -------- SAO (7.3.8.3)
  void parse_sao(int rx, int ry) {
    int addr = ry * wctb + rx;
    SaoRaw& sp = sao_raw[addr]; sp = SaoRaw{};
    if (!sao_luma && !sao_chroma) return;
    int ml = 0, mu = 0;
    if (rx > 0 && addr - 1 >= slice_addr_rs) ml = cabac.bin(ctx[CTX_SAO_MERGE]);
    if (ry > 0 && !ml && addr - wctb >= slice_addr_rs) mu = cabac.bin(ctx[CTX_SAO_MERGE]);
    if (ml) { sp = sao_raw[addr - 1]; return; }
    if (mu) { sp = sao_raw[addr - wctb]; return; }
    for (int c = 0; c < (chroma ? 3 : 1); c++) {
      if ((c == 0 && !sao_luma) || (c > 0 && !sao_chroma)) continue;
      if (c < 2) { int t = 0; if (cabac.bin(ctx[CTX_SAO_TYPE])) t = cabac.bypass() ? 2 : 1; sp.type[c] = t; } else sp.type[2] = sp.type[1];
      if (!sp.type[c]) continue;
      int cmax = (1 << (std::min(bd, 10) - 5)) - 1, av[4];
      for (int i = 0; i < 4; i++) { int v = 0; while (v < cmax && cabac.bypass()) v++; av[i] = v; }
      int sc = c == 0 ? PP->log2_sao_scale_luma : PP->log2_sao_scale_chroma;
      if (sp.type[c] == 1) {
        for (int i = 0; i < 4; i++) if (av[i] && cabac.bypass()) av[i] = -av[i];
        sp.band[c] = (int)cabac.bypass_bits(5);
        for (int i = 0; i < 4; i++) sp.off[c][i] = av[i] * (1 << sc);
      } else {
        if (c == 0) sp.eo[0] = (int)cabac.bypass_bits(2); else if (c == 1) sp.eo[1] = (int)cabac.bypass_bits(2); else sp.eo[2] = sp.eo[1];
        sp.off[c][0] = av[0] << sc; sp.off[c][1] = av[1] << sc; sp.off[c][2] = -(av[2] << sc); sp.off[c][3] = -(av[3] << sc);
      }
    }
  }

  // -------- QP (8.6.1)
  void derive_qpy(int xcb, int ycb) {
    int mask = (1 << qg_log2) - 1, xqg = xcb & ~mask, yqg = ycb & ~mask, cm = ~(ctb - 1);
    int qa = qpy_prev_qg, qb = qpy_prev_qg;
    if (avail(xqg - 1, yqg) && ((xqg - 1) & cm) == (xqg & cm)) qa = qp4[(size_t)(yqg >> 2) * w4 + ((xqg - 1) >> 2)];
    if (avail(xqg, yqg - 1) && ((yqg - 1) & cm) == (yqg & cm)) qb = qp4[(size_t)((yqg - 1) >> 2) * w4 + (xqg >> 2)];
    int pred = (qa + qb + 1) >> 1, qbd = 6 * (bd - 8);
    cur_qpy = ((pred + dqp_val + 52 + 2 * qbd) % (52 + qbd)) - qbd;
  }

  // -------- residual_coding (7.3.8.11): emits sparse (pos, level) entries; returns the number of coefficients
  int residual(int log2n, int c, int mode, int& tskip) {
    const int n = 1 << log2n;
    tskip = 0;
    if (PP->transform_skip && log2n == 2) tskip = cabac.bin(ctx[CTX_TSKIP + (c ? 1 : 0)]);
    int cmax = (log2n << 1) - 1, off, shift;
    if (c == 0) { off = 3 * (log2n - 2) + ((log2n - 1) >> 2); shift = (log2n + 1) >> 2; } else { off = 15; shift = log2n - 2; }
    int lx = 0, ly = 0;
    while (lx < cmax && cabac.bin(ctx[CTX_LAST_X + off + (lx >> shift)])) lx++;
    while (ly < cmax && cabac.bin(ctx[CTX_LAST_Y + off + (ly >> shift)])) ly++;
    if (lx > 3) { int nb = (lx >> 1) - 1; lx = (1 << nb) * (2 + (lx & 1)) + (int)cabac.bypass_bits(nb); }
    if (ly > 3) { int nb = (ly >> 1) - 1; ly = (1 << nb) * (2 + (ly & 1)) + (int)cabac.bypass_bits(nb); }
    int scan = 0;
    if (log2n == 2 || (log2n == 3 && c == 0)) { if (mode >= 6 && mode <= 14) scan = 2; else if (mode >= 22 && mode <= 30) scan = 1; }
    if (scan == 2) std::swap(lx, ly);
    if (lx >= n || ly >= n) { err = set_error(B200_E_BITSTREAM, "last significant coefficient outside the block"); return 0; }
    const int l2sb = log2n - 2;
    const uint8_t *sbx = g_sx[l2sb][scan], *sby = g_sy[l2sb][scan], *px = g_sx[2][scan], *py = g_sy[2][scan];
    int last_sb = 0, last_pos = 0;
    { int xs = lx >> 2, ys = ly >> 2, xp = lx & 3, yp = ly & 3, nsb = 1 << (2 * l2sb);
      for (int i = 0; i < nsb; i++) if (sbx[i] == xs && sby[i] == ys) { last_sb = i; break; }
      for (int k = 0; k < 16; k++) if (px[k] == xp && py[k] == yp) { last_pos = k; break; } }
    uint8_t csbf[8][8]; memset(csbf, 0, sizeof csbf);
    int carry = 1, count = 0; bool first_done = false;
    for (int i = last_sb; i >= 0; i--) {
      int xs = sbx[i], ys = sby[i], infer_dc = 0, coded;
      if (i < last_sb && i > 0) {
        int cs = 0;
        if (xs + 1 < (1 << l2sb)) cs |= csbf[ys][xs + 1];
        if (ys + 1 < (1 << l2sb)) cs |= csbf[ys + 1][xs];
        coded = cabac.bin(ctx[CTX_CSBF + (cs ? 1 : 0) + (c ? 2 : 0)]);
        infer_dc = 1;
      } else coded = 1;
      csbf[ys][xs] = (uint8_t)coded;
      if (!coded) continue;
      uint8_t sig[16] = {0};
      int prev = 0;
      if (xs + 1 < (1 << l2sb)) prev |= csbf[ys][xs + 1];
      if (ys + 1 < (1 << l2sb)) prev |= csbf[ys + 1][xs] << 1;
      int start = i == last_sb ? last_pos - 1 : 15;
      if (i == last_sb) sig[last_pos] = 1;
      for (int k = start; k >= 0; k--) {
        if (k > 0 || !infer_dc) {
          int xc = (xs << 2) + px[k], yc = (ys << 2) + py[k], sc;
          if (log2n == 2) sc = kSigMap4[(yc << 2) + xc];
          else if (xc + yc == 0) sc = 0;
          else {
            int xp = xc & 3, yp = yc & 3;
            if (prev == 0) sc = (xp + yp == 0) ? 2 : (xp + yp < 3) ? 1 : 0;
            else if (prev == 1) sc = yp == 0 ? 2 : (yp == 1 ? 1 : 0);
            else if (prev == 2) sc = xp == 0 ? 2 : (xp == 1 ? 1 : 0);
            else sc = 2;
            if (c == 0) { if (xs || ys) sc += 3; sc += log2n == 3 ? (scan == 0 ? 9 : 15) : 21; } else sc += log2n == 3 ? 9 : 12;
          }
          sig[k] = (uint8_t)cabac.bin(ctx[CTX_SIG + (c == 0 ? sc : 27 + sc)]);
          if (sig[k]) infer_dc = 0;
        } else sig[k] = 1;
      }
      uint8_t g1[16] = {0};
      int first_sig = 16, last_sig = -1, ng1 = 0, last_g1 = -1, g1ctx = 1, g2 = 0;
      int ctx_set = (i == 0 || c > 0) ? 0 : 2;
      if (first_done && carry == 0) ctx_set++;
      first_done = true;
      bool any = false;
      for (int k = 15; k >= 0; k--) if (sig[k]) {
        any = true;
        if (ng1 < 8) {
          g1[k] = (uint8_t)cabac.bin(ctx[CTX_GT1 + ctx_set * 4 + std::min(3, g1ctx) + (c ? 16 : 0)]);
          ng1++;
          if (g1[k]) { g1ctx = 0; if (last_g1 < 0) last_g1 = k; } else if (g1ctx > 0) g1ctx++;
        }
        if (last_sig < 0) last_sig = k;
        first_sig = k;
      }
      if (any) carry = g1ctx;
      bool hidden = PP->sign_hiding && (last_sig - first_sig > 3);
      if (last_g1 >= 0) g2 = cabac.bin(ctx[CTX_GT2 + ctx_set + (c ? 4 : 0)]);
      unsigned signs = 0; int nsign = 0;
      for (int k = 15; k >= 0; k--) if (sig[k] && (!hidden || k != first_sig)) nsign++;
      signs = cabac.bypass_bits(nsign);
      int nsig = 0, sum = 0, rice = 0, sidx = nsign;
      for (int k = 15; k >= 0; k--) if (sig[k]) {
        int base = 1 + g1[k] + (k == last_g1 ? g2 : 0), a = base;
        if (base == ((nsig < 8) ? ((k == last_g1) ? 3 : 2) : 1)) {
          int pre = 0; while (pre < 32 && cabac.bypass()) pre++;
          int rem = pre <= 3 ? (pre << rice) + (int)cabac.bypass_bits(rice) : (((1 << (pre - 3)) + 3 - 1) << rice) + (int)cabac.bypass_bits(pre - 3 + rice);
          a = base + rem;
          if (a > 3 * (1 << rice)) rice = std::min(rice + 1, 4);
        }
        int neg = 0;
        if (!hidden || k != first_sig) { sidx--; neg = (signs >> sidx) & 1; }
        int v = neg ? -a : a;
        if (hidden) { sum += a; if (k == first_sig && (sum & 1)) v = -v; }
        CoefEntry e; e.pos = (uint16_t)((((ys << 2) + py[k]) << log2n) + (xs << 2) + px[k]); e.level = (int16_t)clip3(-32768, 32767, v);
        P.coefs.push_back(e); count++;
        nsig++;
      }
    }
    return count;
  }

  // -------- transform tree / unit (7.3.8.8, 7.3.8.10)
  void mark_tu(int x0, int y0, int log2n) {
    int n4 = 1 << (log2n - 2), bx = x0 >> 2, by = y0 >> 2;
    for (int y = 0; y < n4; y++) for (int x = 0; x < n4; x++) {
      size_t i = (size_t)(by + y) * w4 + bx + x;
      slice_of4[i] = (uint16_t)(slice_idx + 1); qp4[i] = (int8_t)cur_qpy;
      if (x == 0) edge4[i] |= 1;
      if (y == 0) edge4[i] |= 2;
    }
  }

  void transform_unit(const Cu& cu, int x0, int y0, int log2n, int blk, int cbf_l, int cbf_cb, int cbf_cr, int pcb, int pcr) {
    int cbf_c = chroma ? (log2n > 2 ? (cbf_cb | cbf_cr) : (pcb | pcr)) : 0;
    if ((cbf_l || cbf_c) && PP->cu_qp_delta && !is_dqp_coded) {
      int v = 0;
      while (v < 5 && cabac.bin(ctx[CTX_QP_DELTA + (v ? 1 : 0)])) v++;
      if (v == 5) { int k = 0; while (k < 16 && cabac.bypass()) { v += 1 << k; k++; } v += (int)cabac.bypass_bits(k); }
      if (v && cabac.bypass()) v = -v;
      is_dqp_coded = 1; dqp_val = v;
      derive_qpy(cu.x0, cu.y0);
    }
    int pu = cu.nxn ? ((y0 >= cu.y0 + (1 << (cu.log2cb - 1))) ? 2 : 0) + ((x0 >= cu.x0 + (1 << (cu.log2cb - 1))) ? 1 : 0) : 0;
    int lmode = cu.lmode[pu];
    TuCmd t{};
    size_t coef0 = P.coefs.size();
    int ts_l = 0, ts_cb = 0, ts_cr = 0, nl = 0, ncb = 0, ncr = 0;
    if (cbf_l) nl = residual(log2n, 0, lmode, ts_l);
    int chroma_here = 0, ccb = 0, ccr = 0;
    if (chroma) {
      if (log2n > 2) { chroma_here = 1; ccb = cbf_cb; ccr = cbf_cr; if (ccb) ncb = residual(log2n - 1, 1, cu.cmode, ts_cb); if (ccr) ncr = residual(log2n - 1, 2, cu.cmode, ts_cr); }
      else if (blk == 3) { chroma_here = 1; ccb = pcb; ccr = pcr; if (ccb) ncb = residual(2, 1, cu.cmode, ts_cb); if (ccr) ncr = residual(2, 2, cu.cmode, ts_cr); }
    }
    mark_tu(x0, y0, log2n);
    t.w0 = (uint32_t)(x0 >> 2) | ((uint32_t)(y0 >> 2) << 12) | ((uint32_t)(log2n - 2) << 24) | ((uint32_t)cbf_l << 26) | ((uint32_t)ccb << 27) |
           ((uint32_t)ccr << 28) | ((uint32_t)chroma_here << 29) | ((uint32_t)ts_l << 30) | ((uint32_t)ts_cb << 31);
    t.w1 = (uint32_t)lmode | ((uint32_t)cu.cmode << 6) | ((uint32_t)(cur_qpy + 64) << 12) | ((uint32_t)ts_cr << 20);
    t.w2 = (uint32_t)coef0;
    t.w3 = (uint32_t)nl | ((uint32_t)ncb << 11) | ((uint32_t)ncr << 21);
    P.tus.push_back(t);
  }

  void transform_tree(const Cu& cu, int x0, int y0, int log2n, int depth, int blk, int pcb, int pcr, int max_depth) {
    if (err) return;
    int split;
    if (log2n <= S->log2_max_tb && log2n > S->log2_min_tb && depth < max_depth && !(cu.nxn && depth == 0)) split = cabac.bin(ctx[CTX_SPLIT_TR + 5 - log2n]);
    else split = (log2n > S->log2_max_tb || (cu.nxn && depth == 0)) ? 1 : 0;
    if (split && log2n <= 2) { err = set_error(B200_E_BITSTREAM, "transform split below 4x4"); return; }
    int cb = 0, cr = 0;
    if (chroma) {
      if (log2n > 2) { if (depth == 0 || pcb) cb = cabac.bin(ctx[CTX_CBF_CHROMA + depth]); if (depth == 0 || pcr) cr = cabac.bin(ctx[CTX_CBF_CHROMA + depth]); }
      else { cb = pcb; cr = pcr; }
    }
    if (split) {
      int h = 1 << (log2n - 1);
      for (int k = 0; k < 4; k++) transform_tree(cu, x0 + (k & 1) * h, y0 + (k >> 1) * h, log2n - 1, depth + 1, k, cb, cr, max_depth);
    } else {
      int cl = cabac.bin(ctx[CTX_CBF_LUMA + (depth == 0 ? 1 : 0)]);
      if (log2n > 2) transform_unit(cu, x0, y0, log2n, blk, cl, cb, cr, 0, 0);
      else transform_unit(cu, x0, y0, log2n, blk, cl, 0, 0, pcb, pcr);
    }
  }

  int luma_mode(int x, int y, int prev, int mpm_idx, int rem) const {
    int ca = 1, cb = 1;
    if (avail(x - 1, y)) ca = ipm4[(size_t)(y >> 2) * w4 + ((x - 1) >> 2)];
    if (avail(x, y - 1) && (y - 1) >= ((y >> log2ctb) << log2ctb)) cb = ipm4[(size_t)((y - 1) >> 2) * w4 + (x >> 2)];
    int cand[3];
    if (ca == cb) { if (ca < 2) { cand[0] = 0; cand[1] = 1; cand[2] = 26; } else { cand[0] = ca; cand[1] = 2 + ((ca + 29) % 32); cand[2] = 2 + ((ca - 2 + 1) % 32); } }
    else { cand[0] = ca; cand[1] = cb; if (ca != 0 && cb != 0) cand[2] = 0; else if (ca != 1 && cb != 1) cand[2] = 1; else cand[2] = 26; }
    if (prev) return cand[mpm_idx];
    if (cand[0] > cand[1]) std::swap(cand[0], cand[1]);
    if (cand[0] > cand[2]) std::swap(cand[0], cand[2]);
    if (cand[1] > cand[2]) std::swap(cand[1], cand[2]);
    int m = rem;
    for (int i = 0; i < 3; i++) if (m >= cand[i]) m++;
    return m;
  }

  void coding_unit(int x0, int y0, int log2cb, int depth) {
    Cu cu{}; cu.x0 = x0; cu.y0 = y0; cu.log2cb = log2cb;
    int n = 1 << log2cb;
    if (log2cb == S->log2_min_cb) cu.nxn = !cabac.bin(ctx[CTX_PART_MODE]);
    if (cu.nxn && log2cb == 3 && S->log2_min_tb > 2) { err = set_error(B200_E_BITSTREAM, "NxN partition with 8x8 minimum transform"); return; }
    int np = cu.nxn ? 4 : 1, pb = cu.nxn ? n / 2 : n, prev[4], mi[4] = {0}, rem[4] = {0};
    for (int i = 0; i < np; i++) prev[i] = cabac.bin(ctx[CTX_PREV_INTRA]);
    for (int i = 0; i < np; i++) { if (prev[i]) { mi[i] = cabac.bypass(); if (mi[i]) mi[i] += cabac.bypass(); } else rem[i] = (int)cabac.bypass_bits(5); }
    for (int i = 0; i < np; i++) {
      int px = x0 + (i & 1) * pb, py = y0 + (i >> 1) * pb;
      int m = luma_mode(px, py, prev[i], mi[i], rem[i]);
      cu.lmode[i] = m;
      for (int yy = 0; yy < pb; yy += 4) for (int xx = 0; xx < pb; xx += 4) {
        size_t idx = (size_t)((py + yy) >> 2) * w4 + ((px + xx) >> 2);
        ipm4[idx] = (uint8_t)m; slice_of4[idx] = (uint16_t)(slice_idx + 1);   // earlier PUs of this CU are available (6.4.2)
      }
    }
    if (chroma) {
      int v = 4; if (cabac.bin(ctx[CTX_CHROMA_PRED])) v = (int)cabac.bypass_bits(2);
      static const uint8_t tab[4] = {0, 26, 10, 1};
      if (v == 4) cu.cmode = cu.lmode[0]; else { cu.cmode = tab[v]; if (cu.cmode == cu.lmode[0]) cu.cmode = 34; }
    }
    for (int yy = 0; yy < n; yy += 4) for (int xx = 0; xx < n; xx += 4) {
      size_t idx = (size_t)((y0 + yy) >> 2) * w4 + ((x0 + xx) >> 2);
      slice_of4[idx] = 0; cd4[idx] = (uint8_t)depth;
    }
    if (!PP->cu_qp_delta) cur_qpy = slice_qp; else derive_qpy(x0, y0);
    transform_tree(cu, x0, y0, log2cb, 0, 0, 0, 0, S->max_th_depth_intra + cu.nxn);
    for (int yy = 0; yy < n; yy += 4) for (int xx = 0; xx < n; xx += 4) qp4[(size_t)((y0 + yy) >> 2) * w4 + ((x0 + xx) >> 2)] = (int8_t)cur_qpy;
    last_cu_qpy = cur_qpy;
  }

  void coding_quadtree(int x0, int y0, int log2cb, int depth) {
    if (err) return;
    int n = 1 << log2cb, split;
    if (x0 + n <= W && y0 + n <= H && log2cb > S->log2_min_cb) {
      int inc = 0;
      if (avail(x0 - 1, y0) && cd4[(size_t)(y0 >> 2) * w4 + ((x0 - 1) >> 2)] > depth) inc++;
      if (avail(x0, y0 - 1) && cd4[(size_t)((y0 - 1) >> 2) * w4 + (x0 >> 2)] > depth) inc++;
      split = cabac.bin(ctx[CTX_SPLIT_CU + inc]);
    } else split = log2cb > S->log2_min_cb;
    if (PP->cu_qp_delta && log2cb >= qg_log2) {
      is_dqp_coded = 0; dqp_val = 0;
      if (!split || log2cb == qg_log2) { if (first_qg) { qpy_prev_qg = slice_qp; first_qg = 0; } else qpy_prev_qg = last_cu_qpy; }
    }
    if (split) {
      int h = n >> 1;
      for (int k = 0; k < 4; k++) { int x1 = x0 + (k & 1) * h, y1 = y0 + (k >> 1) * h; if (x1 < W && y1 < H) coding_quadtree(x1, y1, log2cb - 1, depth + 1); }
    } else coding_unit(x0, y0, log2cb, depth);
  }

  // -------- per-picture maps for the in-loop filters
  void finalize() {
    P.desc.nslices = (int)P.slices.size();
    P.qp8.resize((size_t)w8 * h8); P.edge8.resize((size_t)w8 * h8);
    for (int by = 0; by < h8; by++) for (int bx = 0; bx < w8; bx++) {
      size_t i4 = (size_t)(by * 2) * w4 + bx * 2;
      P.qp8[(size_t)by * w8 + bx] = qp4[i4];
      uint8_t e = 0;
      int sq = slice_of4[i4] - 1;
      const SliceInfo& sl = P.slices[sq];
      if (!sl.deblocking_disabled) {                       // 8.7.2.3: filterEdgeFlag
        if ((edge4[i4] & 1) && bx > 0) { int sp = slice_of4[i4 - 1] - 1; if (sp == sq || sl.lf_across_slices) e |= 1; }
        if ((edge4[i4] & 2) && by > 0) { int sp = slice_of4[i4 - (size_t)w4] - 1; if (sp == sq || sl.lf_across_slices) e |= 2; }
      }
      P.edge8[(size_t)by * w8 + bx] = e;
    }
    for (int a = 0; a < wctb * hctb; a++) {
      const SaoRaw& r = sao_raw[a]; CtuInfo& ci = P.ctus[a];
      for (int c = 0; c < 3; c++) {
        ci.sao[c].type = (uint8_t)r.type[c];
        ci.sao[c].band_or_class = (uint8_t)(r.type[c] == 1 ? r.band[c] : r.eo[c]);
        for (int k = 0; k < 4; k++) ci.sao[c].offset[k] = (int8_t)clip3(-128, 127, r.off[c][k]);
      }
    }
  }
};

}  // namespace

int parse_access_unit(const uint8_t* data, size_t size, const ParseLimits& limits, ParsedPicture& out) {
  if (!data || size < 6) return set_error(B200_E_BITSTREAM, "empty access unit");
  Parser p(out, limits);
  return p.run(data, size);
}

}  // namespace b200
