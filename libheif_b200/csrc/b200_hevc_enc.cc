// b200_hevc_enc.cc -- host-side HEVC intra-only encoder (fixture generator / heif_encoder_plugin back end).
//
// Role in the reference: the x265 plugin (libheif/plugins/encoder_x265.cc:752-1051 encode_image,
// :1186-1244 get_compressed_data) as driven by Encoder_HEVC::encode (libheif/codecs/hevc_enc.cc:33-115):
// one call per image/tile, output = VPS, SPS, PPS and slice NAL units without start codes.
// x265 is absent from this image, so synthetic inputs for the decoder (BASELINE configs 2-5) are
// produced by this closed-loop encoder.  It is deliberately simple (no RDO) but emits every syntax element
// the decoder supports: CTB 16/32/64, CU quadtree, 2Nx2N / NxN, all 35 intra modes, TU trees with TB 4..32,
// DST 4x4, transform skip, sign-data hiding, cu_qp_delta, chroma QP offsets, SAO band/edge with merges,
// deblocking overrides, WPP entry points, multiple slices and dependent slice segments, 8..12 bit,
// 4:2:0 and 4:0:0.  Syntax follows ITU-T H.265 7.3 / 9.3; the reconstruction loop follows 8.4 / 8.6.
#include "b200_internal.h"
#include "b200_hevc_scaling.h"
#include <algorithm>
#include <vector>

namespace b200 {
namespace enc {

// ------------------------------------------------------------------------------------------ tables
enum { CTX_SAO_MERGE = 0, CTX_SAO_TYPE = 1, CTX_SPLIT_CU = 2, CTX_PART_MODE = 5, CTX_PREV_INTRA = 6,
       CTX_CHROMA_PRED = 7, CTX_SPLIT_TR = 8, CTX_CBF_LUMA = 11, CTX_CBF_CHROMA = 13, CTX_QP_DELTA = 18,
       CTX_TSKIP = 20, CTX_LAST_X = 22, CTX_LAST_Y = 40, CTX_CSBF = 58, CTX_SIG = 62, CTX_GT1 = 104,
       CTX_GT2 = 128, CTX_TQ_BYPASS = 134, CTX_COUNT = 135 };

static const uint8_t kCtxInitI[CTX_COUNT] = {
  153, 200, 139, 141, 157, 184, 184, 63, 153, 138, 138, 111, 141, 94, 138, 182, 154, 154, 154, 154, 139, 139,
  110, 110, 124, 125, 140, 153, 125, 127, 140, 109, 111, 143, 127, 111, 79, 108, 123, 63,
  110, 110, 124, 125, 140, 153, 125, 127, 140, 109, 111, 143, 127, 111, 79, 108, 123, 63,
  91, 171, 134, 141,
  111, 111, 125, 110, 110, 94, 124, 108, 124, 107, 125, 141, 179, 153, 125, 107, 125, 141, 179, 153, 125,
  107, 125, 141, 179, 153, 125, 140, 139, 182, 182, 152, 136, 152, 136, 153, 136, 139, 111, 136, 139, 111,
  140, 92, 137, 138, 140, 152, 138, 139, 153, 74, 149, 92, 139, 107, 122, 152, 140, 179, 166, 182, 140, 227, 122, 197,
  138, 153, 136, 167, 152, 152,
  154};                                        // cu_transquant_bypass_flag (Table 9-8)

static const uint8_t kRangeLps[64][4] = {
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
static const uint8_t kTransLps[64] = {0,0,1,2,2,4,4,5,6,7,8,9,9,11,11,12,13,13,15,15,16,16,18,18,19,19,21,21,22,22,23,24,
  24,25,26,26,27,27,28,29,29,30,30,30,31,32,32,33,33,33,34,34,35,35,35,36,36,36,37,37,37,38,38,63};

static const int8_t kDctT[32] = {64, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67,
                                 64, 61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9, 4};
static const int8_t kDst4[4][4] = {{29, 55, 74, 84}, {74, 74, 0, -74}, {84, -29, -74, 55}, {55, -84, 74, -29}};
static const int8_t kAngle[35] = {0, 0, 32, 26, 21, 17, 13, 9, 5, 2, 0, -2, -5, -9, -13, -17, -21, -26, -32,
                                  -26, -21, -17, -13, -9, -5, -2, 0, 2, 5, 9, 13, 17, 21, 26, 32};
static const int16_t kInvAngle[35] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, -4096, -1638, -910, -630, -482, -390, -315, -256,
                                      -315, -390, -482, -630, -910, -1638, -4096, 0, 0, 0, 0, 0, 0, 0, 0, 0};
static const uint8_t kSigCtxMap4[16] = {0, 1, 4, 5, 2, 3, 4, 5, 6, 6, 8, 8, 7, 7, 8, 8};
static const uint8_t kQpcTab[14] = {29, 30, 31, 32, 33, 33, 34, 34, 35, 35, 36, 36, 37, 37};
static const uint8_t kLevelScale[6] = {40, 45, 51, 57, 64, 72};
static const int kQuantScale[6] = {26214, 23302, 20560, 18396, 16384, 14564};

static int16_t g_mat[4][32][32];       // [log2n-2][k][n] DCT matrices
static uint8_t g_scan_x[4][3][64], g_scan_y[4][3][64];
static bool g_tables_ready = false;

static void init_tables() {
  if (g_tables_ready) return;
  for (int l = 2; l <= 5; l++) {
    int n = 1 << l;
    for (int k = 0; k < n; k++) for (int x = 0; x < n; x++) {
      int v;
      if (k == 0) v = 64;
      else {
        int j = ((k << (5 - l)) * (2 * x + 1)) & 127, sgn = 1;
        if (j > 64) j = 128 - j;
        if (j > 32) { j = 64 - j; sgn = -1; }
        v = sgn * kDctT[j];
      }
      g_mat[l - 2][k][x] = (int16_t)v;
    }
  }
  for (int l = 0; l <= 3; l++) {
    int n = 1 << l, i = 0, x = 0, y = 0;
    bool stop = false;
    while (!stop) {
      while (y >= 0) { if (x < n && y < n) { g_scan_x[l][0][i] = (uint8_t)x; g_scan_y[l][0][i] = (uint8_t)y; i++; } y--; x++; }
      y = x; x = 0;
      if (i >= n * n) stop = true;
    }
    i = 0; for (y = 0; y < n; y++) for (x = 0; x < n; x++) { g_scan_x[l][1][i] = (uint8_t)x; g_scan_y[l][1][i] = (uint8_t)y; i++; }
    i = 0; for (x = 0; x < n; x++) for (y = 0; y < n; y++) { g_scan_x[l][2][i] = (uint8_t)x; g_scan_y[l][2][i] = (uint8_t)y; i++; }
  }
  g_tables_ready = true;
}

static inline int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }

// ------------------------------------------------------------------------------------------ bit I/O
struct BitWriter {
  std::vector<uint8_t> buf; int nbits = 0; uint8_t cur = 0;
  void put(unsigned v, int n) { for (int i = n - 1; i >= 0; i--) { cur = (uint8_t)((cur << 1) | ((v >> i) & 1)); if (++nbits == 8) { buf.push_back(cur); cur = 0; nbits = 0; } } }
  void ue(unsigned v) { unsigned x = v + 1; int len = 0; while ((x >> len) > 1) len++; put(0, len); put(x, len + 1); }
  void se(int v) { ue(v > 0 ? 2 * v - 1 : -2 * v); }
  void trailing() { put(1, 1); while (nbits) put(0, 1); }
  void align_zero() { while (nbits) put(0, 1); }
};

static void append_nal(std::vector<uint8_t>& out, int type, const std::vector<uint8_t>& rbsp) {
  std::vector<uint8_t> nal;
  nal.push_back((uint8_t)(type << 1)); nal.push_back(1);
  int zeros = 0;
  for (uint8_t b : rbsp) {
    if (zeros >= 2 && b <= 3) { nal.push_back(3); zeros = 0; }
    nal.push_back(b);
    zeros = b == 0 ? zeros + 1 : 0;
  }
  uint32_t n = (uint32_t)nal.size();
  out.push_back((uint8_t)(n >> 24)); out.push_back((uint8_t)(n >> 16)); out.push_back((uint8_t)(n >> 8)); out.push_back((uint8_t)n);
  out.insert(out.end(), nal.begin(), nal.end());
}
static size_t escaped_size(const std::vector<uint8_t>& d) {
  size_t n = 0; int zeros = 0;
  for (uint8_t b : d) { if (zeros >= 2 && b <= 3) { n++; zeros = 0; } n++; zeros = b == 0 ? zeros + 1 : 0; }
  return n;
}

// ------------------------------------------------------------------------------------------ CABAC encoder (9.3.4.5)
struct Ctx { uint8_t state, mps; };
struct Cabac {
  BitWriter bw; unsigned low = 0, range = 510; int outstanding = 0; bool first = true;
  void reset() { bw = BitWriter(); low = 0; range = 510; outstanding = 0; first = true; }
  void put_bit(unsigned b) {
    if (first) first = false; else bw.put(b, 1);
    while (outstanding > 0) { bw.put(1 - b, 1); outstanding--; }
  }
  void renorm() {
    while (range < 256) {
      if (low < 256) put_bit(0);
      else if (low >= 512) { low -= 512; put_bit(1); }
      else { low -= 256; outstanding++; }
      range <<= 1; low <<= 1;
    }
  }
  void bin(Ctx& c, int b) {
    unsigned lps = kRangeLps[c.state][(range >> 6) & 3];
    range -= lps;
    if (b != c.mps) { low += range; range = lps; if (c.state == 0) c.mps = 1 - c.mps; c.state = kTransLps[c.state]; }
    else if (c.state < 62) c.state++;
    renorm();
  }
  void bypass(int b) {
    low <<= 1;
    if (b) low += range;
    if (low >= 1024) { put_bit(1); low -= 1024; }
    else if (low < 512) put_bit(0);
    else { low -= 512; outstanding++; }
  }
  void bypass_bits(unsigned v, int n) { for (int i = n - 1; i >= 0; i--) bypass((v >> i) & 1); }
  void restart() { low = 0; range = 510; outstanding = 0; first = true; }     // 9.3.2.5 after pcm_sample(): same bit writer, fresh interval
  void terminate(int b) {
    range -= 2;
    if (b) { low += range; range = 2; renorm(); put_bit((low >> 9) & 1); bw.put(((low >> 7) & 3) | 1, 2); bw.align_zero(); }
    else renorm();
  }
};

static void init_contexts(Ctx* ctx, int slice_qp) {
  int qp = clip3(0, 51, slice_qp);
  for (int i = 0; i < CTX_COUNT; i++) {
    int iv = kCtxInitI[i], m = (iv >> 4) * 5 - 45, n = ((iv & 15) << 3) - 16;
    int pre = clip3(1, 126, ((m * qp) >> 4) + n);
    ctx[i].mps = pre > 63; ctx[i].state = (uint8_t)(ctx[i].mps ? pre - 64 : 63 - pre);
  }
}

struct Lcg { uint32_t s; uint32_t next() { s = s * 1664525u + 1013904223u; return s >> 8; } int range(int n) { return (int)(next() % (uint32_t)n); } };

struct SaoParams { int type[3], band_pos[3], eo_class[3], abs[3][4], sign[3][4]; int merge_left, merge_up; };

// ------------------------------------------------------------------------------------------ encoder
class Encoder {
 public:
  Encoder(const b200_hevc_enc_params& p, const uint16_t* const src[3], const int stride[3]) : P(p) {
    init_tables();
    W = (p.width + 7) & ~7; H = (p.height + 7) & ~7;       // multiples of MinCbSizeY (8); conformance window crops
    cfmt = p.chroma_format_idc; chroma = cfmt ? 1 : 0;
    sx = (cfmt == 1 || cfmt == 2) ? 1 : 0; sy = cfmt == 1 ? 1 : 0;       // SubWidthC = 1 << sx, SubHeightC = 1 << sy (Table 6-1)
    Wc = chroma ? W >> sx : 0; Hc = chroma ? H >> sy : 0;
    log2ctb = p.log2_ctb_size; ctb = 1 << log2ctb;
    wctb = (W + ctb - 1) >> log2ctb; hctb = (H + ctb - 1) >> log2ctb;
    w4 = W / 4; h4 = H / 4;
    bd = p.bit_depth;
    for (int c = 0; c < (chroma ? 3 : 1); c++) {
      int pw = c ? Wc : W, ph = c ? Hc : H, sw = c ? (p.width + (1 << sx) - 1) >> sx : p.width, sh = c ? (p.height + (1 << sy) - 1) >> sy : p.height;
      org[c].assign((size_t)pw * ph, 0); rec[c].assign((size_t)pw * ph, 0);
      for (int y = 0; y < ph; y++) for (int x = 0; x < pw; x++)
        org[c][(size_t)y * pw + x] = src[c][(size_t)std::min(y, sh - 1) * stride[c] + std::min(x, sw - 1)];   // edge padding
    }
    slice_of4.assign((size_t)w4 * h4, 0); ipm4.assign((size_t)w4 * h4, 1); qp4.assign((size_t)w4 * h4, 0); cd4.assign((size_t)w4 * h4, 0);
    rng.s = p.seed ? p.seed : 0xB200u;
    log2_min_tb = 2; log2_max_tb = std::min(5, log2ctb);
    max_th_depth = clip3(0, 4, p.max_transform_hierarchy_depth_intra);
    qg_log2 = log2ctb - clip3(0, log2ctb - 3, p.diff_cu_qp_delta_depth);
    // scaling lists (7.3.4): what gets coded (or defaulted) and the factors the closed-loop reconstruction uses
    sl::set_all_default(sl_lists);
    if (p.scaling_lists >= 2) {
      for (int s = 0; s < 4; s++) for (int m = 0; m < 6; m += (s == 3 ? 3 : 1)) {
        const int kind = (int)rng.range(4);                       // 0 default, 1 copy of the previous matrix, 2/3 explicit
        sl_kind[s][m] = (uint8_t)((kind == 1 && m == 0) ? 0 : (kind >= 2 ? 2 : kind));
        if (sl_kind[s][m] == 0) sl::set_default(sl_lists, s, m);
        else if (sl_kind[s][m] == 1) { const int ref = m - (s == 3 ? 3 : 1); memcpy(sl_lists.list[s][m], sl_lists.list[s][ref], 64); sl_lists.dc[s][m] = sl_lists.dc[s][ref]; }
        else {
          const int num = s == 0 ? 16 : 64; int v = 8 + (int)rng.range(16);
          if (s > 1) sl_lists.dc[s][m] = (uint8_t)(8 + rng.range(40));
          for (int i = 0; i < num; i++) { v = clip3(1, 255, v + (int)rng.range(7) - 2); sl_lists.list[s][m][i] = (uint8_t)v; }
        }
      }
    }
    {
      const int tc = std::max(1, std::min(p.tile_cols, wctb)), tr = std::max(1, std::min(p.tile_rows, hctb));
      tiles = (tc > 1 || tr > 1) && !p.wpp;
      if (tiles) {
        auto bounds = [&](int n, int size) {
          std::vector<int> b(1, 0);
          if (p.tiles_uniform) for (int i = 1; i <= n; i++) b.push_back((i * size) / n);               // 6.5.1 (6-3)
          else { std::vector<int> cuts; while ((int)cuts.size() < n - 1) { int c = 1 + (int)rng.range(size - 1); if (std::find(cuts.begin(), cuts.end(), c) == cuts.end()) cuts.push_back(c); }
                 std::sort(cuts.begin(), cuts.end()); for (int c : cuts) b.push_back(c); b.push_back(size); }
          return b;
        };
        col_bd = bounds(tc, wctb); row_bd = bounds(tr, hctb);
      }
    }
    pcm_bd_y = p.pcm == 2 ? bd - 1 : bd; pcm_bd_c = p.pcm == 2 ? bd - 2 : bd;
    sl_on = p.scaling_lists != 0;
    if (sl_on) sl::derive(sl_lists, sl_f);
  }

  void encode(std::vector<uint8_t>& out) {
    write_vps(out); write_sps(out); write_pps(out);
    int rows_per_slice = P.slice_ctb_rows > 0 ? P.slice_ctb_rows : hctb;
    slice_idx = 0; region_idx = 0;
    ctb_region.assign((size_t)wctb * hctb, -1);
    if (tiles) {
      // CTBs in tile scan (6.5.1): tile after tile in raster order of the tiles, raster inside each tile
      std::vector<int> order, tile_start;
      for (size_t tr = 0; tr + 1 < row_bd.size(); tr++) for (size_t tc = 0; tc + 1 < col_bd.size(); tc++) {
        tile_start.push_back((int)order.size());
        for (int y = row_bd[tr]; y < row_bd[tr + 1]; y++) for (int x = col_bd[tc]; x < col_bd[tc + 1]; x++) order.push_back(y * wctb + x);
      }
      tile_start.push_back((int)order.size());
      if (P.slice_per_tile) {
        for (size_t t = 0; t + 1 < tile_start.size(); t++) {
          std::vector<int> o(order.begin() + tile_start[t], order.begin() + tile_start[t + 1]);
          encode_slice_segment_list(out, o, {0, (int)o.size()}, false, o[0]);
          slice_idx++;
        }
      } else encode_slice_segment_list(out, order, tile_start, false, 0);
      return;
    }
    for (int r0 = 0; r0 < hctb; r0 += rows_per_slice) {
      int r1 = std::min(hctb, r0 + rows_per_slice);
      if (P.dependent_slice_segments && P.wpp == 0 && r1 - r0 > 1) {
        // split the slice into an independent segment and dependent segments, one CTB row each
        for (int r = r0; r < r1; r++) encode_slice_segment(out, r * wctb, (r + 1) * wctb, r != r0, r0 * wctb);
      } else encode_slice_segment(out, r0 * wctb, r1 * wctb, false, r0 * wctb);
      slice_idx++;
    }
  }

  const std::vector<uint16_t>& recon(int c) const { return rec[c]; }
  int coded_w() const { return W; } int coded_h() const { return H; }

 private:
  b200_hevc_enc_params P;
  int cfmt = 1, sx = 1, sy = 1;
  int W, H, Wc, Hc, chroma, log2ctb, ctb, wctb, hctb, w4, h4, bd, log2_min_tb, log2_max_tb, max_th_depth, qg_log2;
  std::vector<uint16_t> org[3], rec[3];
  std::vector<uint16_t> slice_of4; std::vector<uint8_t> ipm4, cd4; std::vector<int8_t> qp4;
  Lcg rng;
  Cabac cabac; Ctx ctx[CTX_COUNT], ctx_wpp[CTX_COUNT];
  std::vector<SaoParams> sao;
  int slice_idx = 0, slice_addr_rs = 0, slice_qp = 26;
  int region_idx = 0;                                  // counts (slice, tile) regions: availability = same region
  bool tiles = false; std::vector<int> col_bd, row_bd;   // tile column / row boundaries in CTBs (6.5.1)
  std::vector<int> ctb_region;                         // region of every coded CTB (SAO merge candidates)
  int is_qp_delta_coded = 0, cu_qp_delta_val = 0, qpy_prev_qg = 0, last_cu_qpy = 0, first_qg = 1, cur_qpy = 0, qg_target_qp = 0;
  int cu_x0 = 0, cu_y0 = 0;
  sl::Lists sl_lists; sl::Factors sl_f; uint8_t sl_kind[4][6] = {}; bool sl_on = false;
  int pcm_bd_y = 8, pcm_bd_c = 8; bool cu_bypass = false;

  void write_scaling_list_data(BitWriter& b) {                    // 7.3.4
    for (int s = 0; s < 4; s++) for (int m = 0; m < 6; m += (s == 3 ? 3 : 1)) {
      if (sl_kind[s][m] < 2) { b.put(0, 1); b.ue(sl_kind[s][m]); continue; }       // pred_mode_flag = 0: delta 0 = default, 1 = previous matrix
      b.put(1, 1);
      int next = 8; const int num = s == 0 ? 16 : 64;
      if (s > 1) { b.se((int)sl_lists.dc[s][m] - 8); next = sl_lists.dc[s][m]; }
      for (int i = 0; i < num; i++) { int d = (int)sl_lists.list[s][m][i] - next; if (d > 127) d -= 256; if (d < -128) d += 256; b.se(d); next = sl_lists.list[s][m][i]; }
    }
  }
  int scaling_factor(int c, int log2n, int pos) const {          // m[x][y] of 8.6.4.2 for raster position pos of an n x n block
    if (!sl_on) return 16;
    const int n = 1 << log2n, x = pos & (n - 1), y = pos >> log2n, sid = log2n - 2;
    if (sid == 0) return sl_f.m[c][0][y * 4 + x];
    if (pos == 0 && sid >= 2) return sl_f.dc[c][sid];
    return sl_f.m[c][sid][((y >> (log2n - 3)) << 3) + (x >> (log2n - 3))];
  }

  int stride_of(int c) const { return c ? Wc : W; }
  bool avail(int x, int y) const {
    if (x < 0 || y < 0 || x >= W || y >= H) return false;
    unsigned s = slice_of4[(size_t)(y >> 2) * w4 + (x >> 2)];      // 1 + region (slice x tile) of a decoded block: 6.4.1 wants same slice AND same tile
    return s != 0 && s == (unsigned)(region_idx + 1);
  }

  // ---------------------------------------------------------------------------- parameter sets
  void profile_tier_level(BitWriter& b) {
    int profile = cfmt >= 2 ? 4 : (bd == 8 ? (P.still_picture ? 3 : 1) : (bd == 10 && chroma ? 2 : 4));
    b.put(0, 2); b.put(0, 1); b.put(profile, 5);
    uint32_t compat = 0;
    if (profile == 1) compat = (1u << 30) | (1u << 29);       // Main => also Main 10 compatible
    else if (profile == 2) compat = 1u << 29;
    else if (profile == 3) compat = (1u << 28) | (1u << 30) | (1u << 29);
    else compat = 1u << 27;
    b.put(compat, 32);
    b.put(1, 1); b.put(0, 1); b.put(0, 1); b.put(1, 1);        // progressive, !interlaced, !non_packed, frame_only
    if (profile == 4) {                                         // RExt constraint flags: Main 12 / Monochrome 12 family
      b.put(1, 1);                                              // max_12bit_constraint
      b.put(bd <= 10, 1); b.put(bd <= 8, 1);                    // max_10bit, max_8bit
      b.put(cfmt <= 2, 1); b.put(cfmt <= 1, 1); b.put(chroma == 0, 1);   // max_422chroma, max_420chroma, max_monochrome
      b.put(1, 1); b.put(1, 1); b.put(1, 1);                    // intra, one_picture_only, lower_bit_rate
      b.put(0, 32); b.put(0, 2);                                // reserved 34 bits
    } else { b.put(0, 32); b.put(0, 11); }
    b.put(0, 1);                                                // general_inbld / reserved
    long px = (long)W * H;
    int level = px <= 36864 ? 30 : px <= 122880 ? 60 : px <= 245760 ? 63 : px <= 552960 ? 90 : px <= 983040 ? 93 :
                px <= 2228224 ? 120 : px <= 8912896 ? 150 : 180;
    b.put(level, 8);
  }
  void write_vps(std::vector<uint8_t>& out) {
    BitWriter b;
    b.put(0, 4); b.put(1, 1); b.put(1, 1); b.put(0, 6); b.put(0, 3); b.put(1, 1); b.put(0xffff, 16);
    profile_tier_level(b);
    b.put(1, 1);                      // sub_layer_ordering_info_present
    b.ue(0); b.ue(0); b.ue(0);        // max_dec_pic_buffering_minus1, num_reorder, max_latency
    b.put(0, 6); b.ue(0);             // max_layer_id, num_layer_sets_minus1
    b.put(0, 1);                      // timing_info_present
    b.put(0, 1);                      // extension
    b.trailing();
    append_nal(out, 32, b.buf);
  }
  void write_sps(std::vector<uint8_t>& out) {
    BitWriter b;
    b.put(0, 4); b.put(0, 3); b.put(1, 1);
    profile_tier_level(b);
    b.ue(0);
    b.ue(cfmt);
    if (cfmt == 3) b.put(0, 1);                                   // separate_colour_plane_flag
    b.ue(W); b.ue(H);
    int cr = (W - P.width) >> (chroma ? sx : 0), cbm = (H - P.height) >> (chroma ? sy : 0);     // conformance window in chroma units
    if (cr || cbm) { b.put(1, 1); b.ue(0); b.ue(cr); b.ue(0); b.ue(cbm); } else b.put(0, 1);
    b.ue(bd - 8); b.ue(bd - 8);
    b.ue(4);                          // log2_max_pic_order_cnt_lsb_minus4
    b.put(1, 1); b.ue(0); b.ue(0); b.ue(0);
    b.ue(0);                          // log2_min_luma_coding_block_size_minus3 (8)
    b.ue(log2ctb - 3);
    b.ue(log2_min_tb - 2); b.ue(log2_max_tb - log2_min_tb);
    b.ue(0); b.ue(max_th_depth);
    b.put(sl_on ? 1 : 0, 1);          // scaling_list_enabled
    if (sl_on) { b.put(P.scaling_lists == 2 ? 1 : 0, 1); if (P.scaling_lists == 2) write_scaling_list_data(b); }
    b.put(0, 1);                      // amp
    b.put(P.sao ? 1 : 0, 1);
    b.put(P.pcm ? 1 : 0, 1);          // pcm_enabled
    if (P.pcm) {
      b.put(pcm_bd_y - 1, 4); b.put(pcm_bd_c - 1, 4);
      b.ue(0); b.ue(std::min(5, log2ctb) - 3);                      // Log2MinIpcmCbSizeY = 3, Log2MaxIpcmCbSizeY = min(CtbLog2SizeY, 5)
      b.put(P.pcm == 2 ? 1 : 0, 1);                                 // pcm_loop_filter_disabled_flag
    }
    b.ue(0);                          // num_short_term_ref_pic_sets
    b.put(0, 1);                      // long_term_ref_pics_present
    b.put(0, 1);                      // temporal_mvp
    b.put(P.strong_intra_smoothing ? 1 : 0, 1);
    if (P.vui_present) {
      b.put(1, 1);
      b.put(0, 1); b.put(0, 1);       // aspect_ratio_info, overscan_info
      b.put(1, 1);                    // video_signal_type_present
      b.put(5, 3); b.put(P.full_range ? 1 : 0, 1);
      if (P.colour_description_present) { b.put(1, 1); b.put(P.colour_primaries, 8); b.put(P.transfer_characteristics, 8); b.put(P.matrix_coefficients, 8); }
      else b.put(0, 1);
      b.put(0, 1); b.put(0, 1); b.put(0, 1); b.put(0, 1);   // chroma_loc, neutral_chroma, field_seq, frame_field_info
      b.put(0, 1); b.put(0, 1); b.put(0, 1);                // default_display_window, timing_info, bitstream_restriction
    } else b.put(0, 1);
    b.put(0, 1);                      // sps_extension_present
    b.trailing();
    append_nal(out, 33, b.buf);
  }
  void write_pps(std::vector<uint8_t>& out) {
    BitWriter b;
    b.ue(0); b.ue(0);
    b.put(P.dependent_slice_segments ? 1 : 0, 1);
    b.put(0, 1); b.put(0, 3);
    b.put(P.sign_data_hiding ? 1 : 0, 1);
    b.put(0, 1);
    b.ue(0); b.ue(0);
    b.se(P.init_qp - 26);
    b.put(0, 1);                      // constrained_intra_pred
    b.put(P.transform_skip ? 1 : 0, 1);
    b.put(P.cu_qp_delta ? 1 : 0, 1);
    if (P.cu_qp_delta) b.ue(log2ctb - qg_log2);
    b.se(P.cb_qp_offset); b.se(P.cr_qp_offset);
    b.put(P.slice_chroma_qp_offsets ? 1 : 0, 1);
    b.put(0, 1); b.put(0, 1);
    b.put(P.transquant_bypass ? 1 : 0, 1);   // transquant_bypass_enabled
    b.put(tiles ? 1 : 0, 1);          // tiles_enabled_flag
    b.put(P.wpp ? 1 : 0, 1);
    if (tiles) {
      b.ue((unsigned)col_bd.size() - 2); b.ue((unsigned)row_bd.size() - 2);
      b.put(P.tiles_uniform ? 1 : 0, 1);
      if (!P.tiles_uniform) {
        for (size_t i = 0; i + 2 < col_bd.size(); i++) b.ue((unsigned)(col_bd[i + 1] - col_bd[i] - 1));
        for (size_t i = 0; i + 2 < row_bd.size(); i++) b.ue((unsigned)(row_bd[i + 1] - row_bd[i] - 1));
      }
      b.put(P.loop_filter_across_tiles ? 1 : 0, 1);
    }
    b.put(P.loop_filter_across_slices ? 1 : 0, 1);
    b.put(1, 1);                      // deblocking_filter_control_present
    b.put(1, 1);                      // deblocking_filter_override_enabled
    b.put(P.deblocking_disabled ? 1 : 0, 1);
    if (!P.deblocking_disabled) { b.se(P.beta_offset_div2); b.se(P.tc_offset_div2); }
    b.put(P.scaling_lists == 3 ? 1 : 0, 1);   // pps_scaling_list_data_present
    if (P.scaling_lists == 3) write_scaling_list_data(b);
    b.put(0, 1);                      // lists_modification_present
    b.ue(0);                          // log2_parallel_merge_level_minus2
    b.put(0, 1);                      // slice_segment_header_extension_present
    b.put(0, 1);                      // pps_extension_present
    b.trailing();
    append_nal(out, 34, b.buf);
  }

  // ---------------------------------------------------------------------------- slice segment
  void encode_slice_segment(std::vector<uint8_t>& out, int addr0, int addr1, bool dependent, int slice_addr) {
    std::vector<int> order; for (int a = addr0; a < addr1; a++) order.push_back(a);
    encode_slice_segment_list(out, order, {0, (int)order.size()}, dependent, slice_addr);
  }
  // order: raster addresses of the segment's CTBs in coding order; starts: indices into `order` where a tile begins (+ the end)
  void encode_slice_segment_list(std::vector<uint8_t>& out, const std::vector<int>& order, const std::vector<int>& starts, bool dependent, int slice_addr) {
    slice_addr_rs = slice_addr;
    const int addr0 = order.front();
    int total = wctb * hctb;
    if (!dependent) {
      slice_qp = clip3(0, 51, P.qp + (slice_idx ? (int)(rng.range(5)) - 2 : 0));
      init_contexts(ctx, slice_qp);
      last_cu_qpy = slice_qp; first_qg = 1;
      region_idx++;
    }
    if (sao.empty()) sao.resize((size_t)total);
    // slice data: one CABAC sub-stream per CTB row when WPP is on, one per tile when tiles are on
    std::vector<std::vector<uint8_t>> substreams;
    cabac.reset();
    size_t next_start = 1;
    for (size_t k = 0; k < order.size(); k++) {
      const int a = order[k];
      int rx = a % wctb, ry = a / wctb;
      if (tiles && next_start + 1 < starts.size() && (int)k == starts[next_start]) {      // first CTB of the next tile: 9.3.1 initialisation
        next_start++;
        init_contexts(ctx, slice_qp); first_qg = 1; region_idx++;
      }
      if (P.wpp && rx == 0 && a != addr0) {
        if (avail((rx + 1) << log2ctb, (ry - 1) << log2ctb)) memcpy(ctx, ctx_wpp, sizeof ctx); else init_contexts(ctx, slice_qp);
        first_qg = 1;
      }
      ctb_region[(size_t)a] = region_idx;
      if (P.sao) { choose_sao(rx, ry); write_sao(rx, ry); }
      coding_quadtree(rx << log2ctb, ry << log2ctb, log2ctb, 0);
      if (P.wpp && rx == 1) memcpy(ctx_wpp, ctx, sizeof ctx);
      bool end = k + 1 == order.size();
      cabac.terminate(end ? 1 : 0);                        // end_of_slice_segment_flag
      const bool tile_end = tiles && next_start < starts.size() && (int)k + 1 == starts[next_start];
      if (!end && ((P.wpp && (a + 1) % wctb == 0) || tile_end)) {
        cabac.terminate(1);                                // end_of_subset_one_bit (+ byte_alignment)
        substreams.push_back(cabac.bw.buf); cabac.reset();
      }
    }
    substreams.push_back(cabac.bw.buf);
    // header
    BitWriter b;
    bool first = addr0 == 0;
    b.put(first ? 1 : 0, 1);
    b.put(0, 1);                                           // no_output_of_prior_pics (IRAP)
    b.ue(0);
    if (!first) {
      if (P.dependent_slice_segments) b.put(dependent ? 1 : 0, 1);
      int bits = 0; while ((1 << bits) < total) bits++;
      b.put(addr0, bits);
    }
    if (!dependent) {
      b.ue(2);                                             // slice_type I
      if (P.sao) { b.put(1, 1); if (chroma) b.put(1, 1); }
      b.se(slice_qp - P.init_qp);
      if (P.slice_chroma_qp_offsets) { b.se(P.slice_cb_qp_offset); b.se(P.slice_cr_qp_offset); }
      bool override = P.slice_deblocking_override != 0;
      b.put(override ? 1 : 0, 1);
      bool dis = P.deblocking_disabled;
      if (override) {
        dis = P.slice_deblocking_disabled != 0;
        b.put(dis ? 1 : 0, 1);
        if (!dis) { b.se(P.slice_beta_offset_div2); b.se(P.slice_tc_offset_div2); }
      }
      if (P.loop_filter_across_slices && (P.sao || !dis)) b.put(P.slice_loop_filter_across_slices ? 1 : 0, 1);
    }
    if (P.wpp || tiles) {
      int ne = (int)substreams.size() - 1;
      b.ue(ne);
      if (ne > 0) { b.ue(31); for (int i = 0; i < ne; i++) b.put((unsigned)(escaped_size(substreams[i]) - 1), 32); }
    }
    b.trailing();                                          // byte_alignment()
    std::vector<uint8_t> rbsp = b.buf;
    for (auto& s : substreams) rbsp.insert(rbsp.end(), s.begin(), s.end());
    append_nal(out, 19 /* IDR_W_RADL */, rbsp);
  }

  // ---------------------------------------------------------------------------- SAO (7.3.8.3)
  void choose_sao(int rx, int ry) {
    int addr = ry * wctb + rx;
    SaoParams& s = sao[addr];
    memset(&s, 0, sizeof s);
    bool left_ok = rx > 0 && ctb_region[(size_t)addr - 1] == region_idx, up_ok = ry > 0 && ctb_region[(size_t)(addr - wctb)] == region_idx;
    int r = rng.range(16);
    if (left_ok && r < 3) { s = sao[addr - 1]; s.merge_left = 1; s.merge_up = 0; return; }
    if (up_ok && r < 6) { s = sao[addr - wctb]; s.merge_up = 1; s.merge_left = 0; return; }
    int cmax = (1 << (std::min(bd, 10) - 5)) - 1;
    for (int c = 0; c < (chroma ? 2 : 1); c++) {
      int t = rng.range(8);
      s.type[c] = t < 3 ? 0 : (t < 5 ? 1 : 2);
      for (int i = 0; i < 4; i++) { s.abs[c][i] = rng.range(4) == 0 ? rng.range(cmax + 1) : rng.range(std::min(cmax, 2) + 1); s.sign[c][i] = rng.range(2); }
      s.band_pos[c] = rng.range(32); s.eo_class[c] = rng.range(4);
    }
    if (chroma) {
      s.type[2] = s.type[1]; s.eo_class[2] = s.eo_class[1];
      for (int i = 0; i < 4; i++) { s.abs[2][i] = rng.range(std::min(cmax, 2) + 1); s.sign[2][i] = rng.range(2); }
      s.band_pos[2] = rng.range(32);
    }
  }
  void write_sao(int rx, int ry) {
    int addr = ry * wctb + rx;
    const SaoParams& s = sao[addr];
    if (rx > 0 && ctb_region[(size_t)addr - 1] == region_idx) cabac.bin(ctx[CTX_SAO_MERGE], s.merge_left);
    if (s.merge_left) return;
    if (ry > 0 && ctb_region[(size_t)(addr - wctb)] == region_idx) cabac.bin(ctx[CTX_SAO_MERGE], s.merge_up);
    if (s.merge_up) return;
    int cmax = (1 << (std::min(bd, 10) - 5)) - 1;
    for (int c = 0; c < (chroma ? 3 : 1); c++) {
      if (c < 2) {
        cabac.bin(ctx[CTX_SAO_TYPE], s.type[c] != 0);
        if (s.type[c]) cabac.bypass(s.type[c] == 2);
      }
      if (!s.type[c]) continue;
      for (int i = 0; i < 4; i++) { int v = s.abs[c][i]; for (int k = 0; k < v; k++) cabac.bypass(1); if (v < cmax) cabac.bypass(0); }
      if (s.type[c] == 1) {
        for (int i = 0; i < 4; i++) if (s.abs[c][i]) cabac.bypass(s.sign[c][i]);
        cabac.bypass_bits(s.band_pos[c], 5);
      } else if (c < 2) cabac.bypass_bits(s.eo_class[c], 2);
    }
  }

  // ---------------------------------------------------------------------------- intra prediction (8.4.4.2)
  void predict(int c, int x0, int y0, int log2n, int mode, uint16_t* dst /* n*n */) const {
    const int n = 1 << log2n, shx = c ? sx : 0, shy = c ? sy : 0, st = stride_of(c);
    const uint16_t* pl = rec[c].data();
    int refbuf[129], fbuf[129]; uint8_t av[129];
    bool any = false;
    for (int i = 0; i <= 4 * n; i++) {
      int px, py;
      if (i < 2 * n) { px = x0 - 1; py = y0 + 2 * n - 1 - i; } else if (i == 2 * n) { px = x0 - 1; py = y0 - 1; } else { px = x0 + (i - 2 * n - 1); py = y0 - 1; }
      av[i] = avail(px << shx, py << shy);
      if (av[i]) { refbuf[i] = pl[(size_t)py * st + px]; any = true; }
    }
    if (!any) for (int i = 0; i <= 4 * n; i++) refbuf[i] = 1 << (bd - 1);
    else {
      int first = 0; while (!av[first]) first++;
      for (int i = 0; i < first; i++) refbuf[i] = refbuf[first];
      for (int i = first + 1; i <= 4 * n; i++) if (!av[i]) refbuf[i] = refbuf[i - 1];
    }
    int* ref = refbuf;
    if ((c == 0 || cfmt == 3) && mode != 1 && n != 4) {            // 8.4.4.2.3: filtering of the neighbours for luma, and for chroma in 4:4:4
      int dist = std::min(std::abs(mode - 26), std::abs(mode - 10));
      int thr = n == 8 ? 7 : (n == 16 ? 1 : 0);
      if (dist > thr) {
        int corner = ref[2 * n], bl = ref[0], tr = ref[4 * n];
        if (P.strong_intra_smoothing && c == 0 && n == 32 && std::abs(corner + tr - 2 * ref[3 * n]) < (1 << (bd - 5)) && std::abs(corner + bl - 2 * ref[n]) < (1 << (bd - 5))) {
          fbuf[2 * n] = corner; fbuf[0] = bl; fbuf[4 * n] = tr;
          for (int y = 0; y < 63; y++) fbuf[2 * n - 1 - y] = ((63 - y) * corner + (y + 1) * bl + 32) >> 6;
          for (int x = 0; x < 63; x++) fbuf[2 * n + 1 + x] = ((63 - x) * corner + (x + 1) * tr + 32) >> 6;
        } else {
          fbuf[0] = ref[0]; fbuf[4 * n] = ref[4 * n];
          for (int i = 1; i < 4 * n; i++) fbuf[i] = (ref[i - 1] + 2 * ref[i] + ref[i + 1] + 2) >> 2;
        }
        ref = fbuf;
      }
    }
    auto LEFT = [&](int y) { return ref[2 * n - 1 - y]; };
    auto TOP = [&](int x) { return ref[2 * n + 1 + x]; };
    const int maxv = (1 << bd) - 1;
    if (mode == 0) {
      for (int y = 0; y < n; y++) for (int x = 0; x < n; x++)
        dst[y * n + x] = (uint16_t)(((n - 1 - x) * LEFT(y) + (x + 1) * TOP(n) + (n - 1 - y) * TOP(x) + (y + 1) * LEFT(n) + n) >> (log2n + 1));
    } else if (mode == 1) {
      int sum = n; for (int i = 0; i < n; i++) sum += LEFT(i) + TOP(i);
      int dc = sum >> (log2n + 1);
      for (int i = 0; i < n * n; i++) dst[i] = (uint16_t)dc;
      if (c == 0 && n < 32) {
        dst[0] = (uint16_t)((LEFT(0) + 2 * dc + TOP(0) + 2) >> 2);
        for (int x = 1; x < n; x++) dst[x] = (uint16_t)((TOP(x) + 3 * dc + 2) >> 2);
        for (int y = 1; y < n; y++) dst[y * n] = (uint16_t)((LEFT(y) + 3 * dc + 2) >> 2);
      }
    } else {
      int ang = kAngle[mode], ia = kInvAngle[mode];
      int rbuf[98]; int* r = rbuf + 32;
      if (mode >= 18) {
        for (int x = 0; x <= n; x++) r[x] = TOP(x - 1);
        if (ang < 0) { int last = (n * ang) >> 5; if (last < -1) for (int x = last; x <= -1; x++) r[x] = LEFT(-1 + ((x * ia + 128) >> 8)); }
        else for (int x = n + 1; x <= 2 * n; x++) r[x] = TOP(x - 1);
        for (int y = 0; y < n; y++) {
          int idx = ((y + 1) * ang) >> 5, f = ((y + 1) * ang) & 31;
          for (int x = 0; x < n; x++) dst[y * n + x] = (uint16_t)(f ? ((32 - f) * r[x + idx + 1] + f * r[x + idx + 2] + 16) >> 5 : r[x + idx + 1]);
        }
        if (mode == 26 && c == 0 && n < 32) for (int y = 0; y < n; y++) dst[y * n] = (uint16_t)clip3(0, maxv, TOP(0) + ((LEFT(y) - LEFT(-1)) >> 1));
      } else {
        for (int x = 0; x <= n; x++) r[x] = LEFT(x - 1);
        if (ang < 0) { int last = (n * ang) >> 5; if (last < -1) for (int x = last; x <= -1; x++) r[x] = TOP(-1 + ((x * ia + 128) >> 8)); }
        else for (int x = n + 1; x <= 2 * n; x++) r[x] = LEFT(x - 1);
        for (int x = 0; x < n; x++) {
          int idx = ((x + 1) * ang) >> 5, f = ((x + 1) * ang) & 31;
          for (int y = 0; y < n; y++) dst[y * n + x] = (uint16_t)(f ? ((32 - f) * r[y + idx + 1] + f * r[y + idx + 2] + 16) >> 5 : r[y + idx + 1]);
        }
        if (mode == 10 && c == 0 && n < 32) for (int x = 0; x < n; x++) dst[x] = (uint16_t)clip3(0, maxv, LEFT(0) + ((TOP(x) - TOP(-1)) >> 1));
      }
    }
  }

  // ---------------------------------------------------------------------------- transforms
  void forward(const int* res, int* coef, int log2n, bool dst4, bool tskip) const {
    int n = 1 << log2n;
    if (tskip) { int s = 15 - bd - log2n; for (int i = 0; i < n * n; i++) coef[i] = res[i] << s; return; }
    int tmp[1024];
    int s1 = log2n + bd - 9, s2 = log2n + 6;
    for (int k = 0; k < n; k++) for (int x = 0; x < n; x++) {        // columns: tmp[k][x] = sum_y M[k][y] res[y][x]
      long e = 0;
      for (int y = 0; y < n; y++) e += (dst4 ? kDst4[k][y] : g_mat[log2n - 2][k][y]) * res[y * n + x];
      tmp[k * n + x] = (int)((e + (s1 > 0 ? (1 << (s1 - 1)) : 0)) >> s1);
    }
    for (int k = 0; k < n; k++) for (int y = 0; y < n; y++) {        // rows
      long e = 0;
      for (int x = 0; x < n; x++) e += (dst4 ? kDst4[k][x] : g_mat[log2n - 2][k][x]) * tmp[y * n + x];
      coef[y * n + k] = (int)((e + (1 << (s2 - 1))) >> s2);
    }
  }
  void inverse(const int16_t* d, int* res, int log2n, bool dst4, bool tskip) const {     // 8.6.4.2
    int n = 1 << log2n, bs = 20 - bd;
    if (tskip) { for (int i = 0; i < n * n; i++) res[i] = (((int)d[i] << 7) + (1 << (bs - 1))) >> bs; return; }
    int tmp[1024];
    for (int x = 0; x < n; x++) for (int y = 0; y < n; y++) {
      int e = 0;
      for (int k = 0; k < n; k++) e += d[k * n + x] * (dst4 ? kDst4[k][y] : g_mat[log2n - 2][k][y]);
      tmp[y * n + x] = clip3(-32768, 32767, (e + 64) >> 7);
    }
    for (int y = 0; y < n; y++) for (int x = 0; x < n; x++) {
      int e = 0;
      for (int k = 0; k < n; k++) e += tmp[y * n + k] * (dst4 ? kDst4[k][x] : g_mat[log2n - 2][k][x]);
      res[y * n + x] = (e + (1 << (bs - 1))) >> bs;
    }
  }
  int chroma_qp(int qpy, int off) const {
    int qbd = 6 * (bd - 8), qpi = clip3(-qbd, 57, qpy + off);
    int qpc = cfmt != 1 ? std::min(qpi, 51) : (qpi < 30 ? qpi : (qpi >= 43 ? qpi - 6 : kQpcTab[qpi - 30]));      // Table 8-10 only for ChromaArrayType == 1
    return qpc + qbd;
  }

  // ---------------------------------------------------------------------------- QP (8.6.1)
  int predict_qpy(int xcb, int ycb) const {
    int mask = (1 << qg_log2) - 1, xqg = xcb & ~mask, yqg = ycb & ~mask, cm = ~(ctb - 1);
    int qa = qpy_prev_qg, qb = qpy_prev_qg;
    if (avail(xqg - 1, yqg) && ((xqg - 1) & cm) == (xqg & cm)) qa = qp4[(size_t)(yqg >> 2) * w4 + ((xqg - 1) >> 2)];
    if (avail(xqg, yqg - 1) && ((yqg - 1) & cm) == (yqg & cm)) qb = qp4[(size_t)((yqg - 1) >> 2) * w4 + (xqg >> 2)];
    return (qa + qb + 1) >> 1;
  }

  // ---------------------------------------------------------------------------- residual (7.3.8.11)
  // quantise + (optionally) sign-hide; returns cbf. levels are in raster order [y][x].
  bool quantise(const int* coef, int16_t* lev, int log2n, int qp, int scan) const {
    int n = 1 << log2n, ts = 15 - bd - log2n, qbits = 14 + qp / 6 + ts;
    long add = 171L << (qbits - 9);
    bool any = false;
    for (int i = 0; i < n * n; i++) {
      long a = std::labs((long)coef[i]);
      int l = (int)((a * kQuantScale[qp % 6] + add) >> qbits);
      l = std::min(l, 32767);
      lev[i] = (int16_t)(coef[i] < 0 ? -l : l);
      any |= l != 0;
    }
    if (any && P.sign_data_hiding && !cu_bypass) {
      int l2sb = log2n - 2;
      for (int i = 0; i < (1 << (2 * l2sb)); i++) {
        int xs = g_scan_x[l2sb][scan][i], ys = g_scan_y[l2sb][scan][i];
        int first = 16, last = -1, sum = 0;
        for (int k = 0; k < 16; k++) {
          int v = lev[((ys << 2) + g_scan_y[2][scan][k]) * n + (xs << 2) + g_scan_x[2][scan][k]];
          if (v) { if (first == 16) first = k; last = k; sum += std::abs(v); }
        }
        if (last - first > 3) {
          int16_t& f = lev[((ys << 2) + g_scan_y[2][scan][first]) * n + (xs << 2) + g_scan_x[2][scan][first]];
          if ((sum & 1) != (f < 0 ? 1 : 0)) {               // parity must equal the sign of the first coefficient
            int16_t& t = lev[((ys << 2) + g_scan_y[2][scan][last]) * n + (xs << 2) + g_scan_x[2][scan][last]];
            t = (int16_t)(t < 0 ? t - 1 : t + 1);
          }
        }
      }
    }
    return any;
  }

  void write_residual(const int16_t* lev, int log2n, int c, int scan, bool tskip) {
    const int n = 1 << log2n, l2sb = log2n - 2;
    if (P.transform_skip && log2n == 2 && !cu_bypass) cabac.bin(ctx[CTX_TSKIP + (c ? 1 : 0)], tskip);
    const uint8_t *sbx = g_scan_x[l2sb][scan], *sby = g_scan_y[l2sb][scan], *px = g_scan_x[2][scan], *py = g_scan_y[2][scan];
    int last_sb = -1, last_pos = -1;
    for (int i = (1 << (2 * l2sb)) - 1; i >= 0 && last_sb < 0; i--) for (int k = 15; k >= 0; k--)
      if (lev[((sby[i] << 2) + py[k]) * n + (sbx[i] << 2) + px[k]]) { last_sb = i; last_pos = k; break; }
    int lx = (sbx[last_sb] << 2) + px[last_pos], ly = (sby[last_sb] << 2) + py[last_pos];
    if (scan == 2) std::swap(lx, ly);
    static const uint8_t group[32] = {0, 1, 2, 3, 4, 4, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7, 8, 8, 8, 8, 8, 8, 8, 8, 9, 9, 9, 9, 9, 9, 9, 9};
    static const uint8_t min_in_group[10] = {0, 1, 2, 3, 4, 6, 8, 12, 16, 24};
    int cmax = (log2n << 1) - 1, off, shift;
    if (c == 0) { off = 3 * (log2n - 2) + ((log2n - 1) >> 2); shift = (log2n + 1) >> 2; } else { off = 15; shift = log2n - 2; }
    int gx = group[lx], gy = group[ly];
    for (int k = 0; k < gx; k++) cabac.bin(ctx[CTX_LAST_X + off + (k >> shift)], 1);
    if (gx < cmax) cabac.bin(ctx[CTX_LAST_X + off + (gx >> shift)], 0);
    for (int k = 0; k < gy; k++) cabac.bin(ctx[CTX_LAST_Y + off + (k >> shift)], 1);
    if (gy < cmax) cabac.bin(ctx[CTX_LAST_Y + off + (gy >> shift)], 0);
    if (gx > 3) cabac.bypass_bits(lx - min_in_group[gx], (gx >> 1) - 1);
    if (gy > 3) cabac.bypass_bits(ly - min_in_group[gy], (gy >> 1) - 1);
    uint8_t csbf[8][8]; memset(csbf, 0, sizeof csbf);
    int carry = 1; bool first_done = false;
    for (int i = last_sb; i >= 0; i--) {
      int xs = sbx[i], ys = sby[i];
      int16_t v[16]; bool coded = false;
      for (int k = 0; k < 16; k++) { v[k] = lev[((ys << 2) + py[k]) * n + (xs << 2) + px[k]]; coded |= v[k] != 0; }
      bool infer_dc = false;
      if (i < last_sb && i > 0) {
        int cs = 0;
        if (xs + 1 < (1 << l2sb)) cs |= csbf[ys][xs + 1];
        if (ys + 1 < (1 << l2sb)) cs |= csbf[ys + 1][xs];
        cabac.bin(ctx[CTX_CSBF + (cs ? 1 : 0) + (c ? 2 : 0)], coded);
        infer_dc = true;
      } else coded = true;
      csbf[ys][xs] = coded;
      if (!coded) continue;
      int prev = 0;
      if (xs + 1 < (1 << l2sb)) prev |= csbf[ys][xs + 1];
      if (ys + 1 < (1 << l2sb)) prev |= csbf[ys + 1][xs] << 1;
      int start = i == last_sb ? last_pos - 1 : 15;
      for (int k = start; k >= 0; k--) {
        int xc = (xs << 2) + px[k], yc = (ys << 2) + py[k];
        if (k > 0 || !infer_dc) {
          int sc;
          if (log2n == 2) sc = kSigCtxMap4[(yc << 2) + xc];
          else if (xc + yc == 0) sc = 0;
          else {
            int xp = xc & 3, yp = yc & 3;
            if (prev == 0) sc = (xp + yp == 0) ? 2 : (xp + yp < 3) ? 1 : 0;
            else if (prev == 1) sc = yp == 0 ? 2 : (yp == 1 ? 1 : 0);
            else if (prev == 2) sc = xp == 0 ? 2 : (xp == 1 ? 1 : 0);
            else sc = 2;
            if (c == 0) { if (xs || ys) sc += 3; sc += log2n == 3 ? (scan == 0 ? 9 : 15) : 21; }
            else sc += log2n == 3 ? 9 : 12;
          }
          cabac.bin(ctx[CTX_SIG + (c == 0 ? sc : 27 + sc)], v[k] != 0);
          if (v[k]) infer_dc = false;
        }
      }
      int first_sig = 16, last_sig = -1, ng1 = 0, last_g1 = -1, g1ctx = 1;
      int ctx_set = (i == 0 || c > 0) ? 0 : 2;
      if (first_done && carry == 0) ctx_set++;
      first_done = true;
      bool any = false;
      for (int k = 15; k >= 0; k--) if (v[k]) {
        any = true;
        if (ng1 < 8) {
          int g = std::abs(v[k]) > 1;
          cabac.bin(ctx[CTX_GT1 + ctx_set * 4 + std::min(3, g1ctx) + (c ? 16 : 0)], g);
          ng1++;
          if (g) { g1ctx = 0; if (last_g1 < 0) last_g1 = k; } else if (g1ctx > 0) g1ctx++;
        }
        if (last_sig < 0) last_sig = k;
        first_sig = k;
      }
      if (any) carry = g1ctx;
      bool hidden = P.sign_data_hiding && !cu_bypass && (last_sig - first_sig > 3);
      if (last_g1 >= 0) cabac.bin(ctx[CTX_GT2 + ctx_set + (c ? 4 : 0)], std::abs(v[last_g1]) > 2);
      for (int k = 15; k >= 0; k--) if (v[k] && (!hidden || k != first_sig)) cabac.bypass(v[k] < 0);
      int nsig = 0, rice = 0, cnt1 = 0;
      for (int k = 15; k >= 0; k--) if (v[k]) {
        int a = std::abs(v[k]);
        int g1 = cnt1 < 8 ? (a > 1) : 0; if (cnt1 < 8) cnt1++;
        int g2 = (k == last_g1) ? (a > 2) : 0;
        int base = 1 + g1 + g2;
        if (base == ((nsig < 8) ? ((k == last_g1) ? 3 : 2) : 1)) {
          int rem = a - base;
          if ((rem >> rice) <= 3) { int pre = rem >> rice; for (int t = 0; t < pre; t++) cabac.bypass(1); cabac.bypass(0); cabac.bypass_bits(rem & ((1 << rice) - 1), rice); }
          else {
            int q = (rem >> rice) - 2, kk = 0; while ((q >> (kk + 1)) > 0) kk++;
            int pre = kk + 3;
            for (int t = 0; t < pre; t++) cabac.bypass(1); cabac.bypass(0);
            cabac.bypass_bits(rem - (((1 << kk) + 2) << rice), kk + rice);
          }
          if (a > 3 * (1 << rice)) rice = std::min(rice + 1, 4);
        }
        nsig++;
      }
    }
  }

  // one transform block: predict, transform, quantise, (write), reconstruct. Returns cbf.
  struct TbResult { bool cbf; int16_t lev[1024]; int scan; bool tskip; };
  void code_tb(int c, int x0, int y0, int log2n, int mode, TbResult& r) {
    const int n = 1 << log2n, st = stride_of(c);
    uint16_t pred[1024]; int res[1024], coef[1024];
    predict(c, x0, y0, log2n, mode, pred);
    for (int y = 0; y < n; y++) for (int x = 0; x < n; x++) res[y * n + x] = (int)org[c][(size_t)(y0 + y) * st + x0 + x] - pred[y * n + x];
    bool dst4 = c == 0 && log2n == 2;
    if (cu_bypass) {                                  // cu_transquant_bypass_flag: the residual is coded as is (8.6.2), lossless
      r.tskip = false; r.scan = 0;
      if (log2n == 2 || (log2n == 3 && (c == 0 || cfmt == 3))) { if (mode >= 6 && mode <= 14) r.scan = 2; else if (mode >= 22 && mode <= 30) r.scan = 1; }
      r.cbf = false;
      for (int i = 0; i < n * n; i++) { r.lev[i] = (int16_t)res[i]; r.cbf |= res[i] != 0; }
      uint16_t* rq = rec[c].data();
      for (int y = 0; y < n; y++) for (int x = 0; x < n; x++) rq[(size_t)(y0 + y) * st + x0 + x] = org[c][(size_t)(y0 + y) * st + x0 + x];
      return;
    }
    r.tskip = P.transform_skip && log2n == 2 && rng.range(4) == 0;
    r.scan = 0;
    if (log2n == 2 || (log2n == 3 && (c == 0 || cfmt == 3))) { if (mode >= 6 && mode <= 14) r.scan = 2; else if (mode >= 22 && mode <= 30) r.scan = 1; }
    forward(res, coef, log2n, dst4, r.tskip);
    int qp = c == 0 ? qg_target_qp + 6 * (bd - 8) : chroma_qp(qg_target_qp, (c == 1 ? P.cb_qp_offset : P.cr_qp_offset) + (P.slice_chroma_qp_offsets ? (c == 1 ? P.slice_cb_qp_offset : P.slice_cr_qp_offset) : 0));
    r.cbf = quantise(coef, r.lev, log2n, qp, r.scan);
    // NOTE: the QP used here (qg_target_qp) is only valid if a cu_qp_delta can still be sent (or already
    // was); the caller re-runs with the predicted QP when neither holds.
    uint16_t* rp = rec[c].data();
    if (r.cbf) {
      int16_t d[1024]; int bs = bd + log2n - 5, scale = kLevelScale[qp % 6] << (qp / 6);
      for (int i = 0; i < n * n; i++) { long t = ((long)r.lev[i] * scaling_factor(c, log2n, i) * scale + (1L << (bs - 1))) >> bs; d[i] = (int16_t)(t < -32768 ? -32768 : (t > 32767 ? 32767 : t)); }
      inverse(d, res, log2n, dst4, r.tskip);
      int maxv = (1 << bd) - 1;
      for (int y = 0; y < n; y++) for (int x = 0; x < n; x++) rp[(size_t)(y0 + y) * st + x0 + x] = (uint16_t)clip3(0, maxv, pred[y * n + x] + res[y * n + x]);
    } else for (int y = 0; y < n; y++) for (int x = 0; x < n; x++) rp[(size_t)(y0 + y) * st + x0 + x] = pred[y * n + x];
  }

  void mark_tu(int x0, int y0, int log2n) {
    int n4 = 1 << (log2n - 2);
    for (int y = 0; y < n4; y++) for (int x = 0; x < n4; x++) { size_t i = (size_t)((y0 >> 2) + y) * w4 + (x0 >> 2) + x; slice_of4[i] = (uint16_t)(region_idx + 1); qp4[i] = (int8_t)cur_qpy; }
  }

  // ---------------------------------------------------------------------------- transform tree (7.3.8.8)
  struct Cu { int x0, y0, log2cb, nxn, lmode[4], cmode[4]; };      // cmode: IntraPredModeC per prediction unit (one per PU only in 4:4:4)

  // Decide the TU split structure first (so that cbf_cb/cbf_cr of inner nodes are known before they are
  // written) by coding leaves depth-first into a node list, then emit the syntax in a second walk.
  // chroma: [t] = the upper / lower square block of a 4:2:2 transform unit (t = 0 only otherwise)
  struct Node { int x0, y0, log2n, depth, blk, split, child[4]; bool cbf_l, cbf_cb[2], cbf_cr[2]; TbResult *l, *cb[2], *cr[2]; };
  std::vector<Node> nodes; std::vector<TbResult*> pool;
  TbResult* new_tb() { TbResult* t = new TbResult; pool.push_back(t); return t; }

  int build_tree(const Cu& cu, int x0, int y0, int log2n, int depth, int blk, int max_depth, int parent) {
    Node nd{}; nd.x0 = x0; nd.y0 = y0; nd.log2n = log2n; nd.depth = depth; nd.blk = blk;
    bool can_split = log2n <= log2_max_tb && log2n > log2_min_tb && depth < max_depth && !(cu.nxn && depth == 0);
    if (can_split) nd.split = rng.range(log2n >= 5 ? 2 : 3) == 0;
    else nd.split = (log2n > log2_max_tb || (cu.nxn && depth == 0)) ? 1 : 0;
    int me = (int)nodes.size(); nodes.push_back(nd);
    if (nd.split) {
      int h = 1 << (log2n - 1);
      bool cb = false, cr = false;
      for (int k = 0; k < 4; k++) {
        int ch = build_tree(cu, x0 + (k & 1) * h, y0 + (k >> 1) * h, log2n - 1, depth + 1, k, max_depth, me);
        nodes[me].child[k] = ch; cb |= nodes[ch].cbf_cb[0] || nodes[ch].cbf_cb[1]; cr |= nodes[ch].cbf_cr[0] || nodes[ch].cbf_cr[1];
      }
      nodes[me].cbf_cb[0] = cb; nodes[me].cbf_cr[0] = cr; nodes[me].cbf_cb[1] = nodes[me].cbf_cr[1] = false;
      if (log2n == 3 && chroma && cfmt != 3) {   // 4x4 luma children: the chroma 4x4 blocks are coded with child 3 at this node's origin
        for (int t = 0; t < 2; t++) {
          nodes[me].cbf_cb[t] = nodes[nodes[me].child[3]].cb[t] ? nodes[nodes[me].child[3]].cb[t]->cbf : false;
          nodes[me].cbf_cr[t] = nodes[nodes[me].child[3]].cr[t] ? nodes[nodes[me].child[3]].cr[t]->cbf : false;
        }
      }
    } else {
      int pu = cu.nxn ? ((y0 >= cu.y0 + (1 << (cu.log2cb - 1))) ? 2 : 0) + ((x0 >= cu.x0 + (1 << (cu.log2cb - 1))) ? 1 : 0) : 0;
      Node& m = nodes[me];
      m.l = new_tb(); code_tb(0, x0, y0, log2n, cu.lmode[pu], *m.l); m.cbf_l = m.l->cbf;
      mark_tu(x0, y0, log2n);
      if (chroma) {
        const int nb = cfmt == 2 ? 2 : 1;                       // 4:2:2: two square blocks, one above the other (7.3.8.10)
        if (log2n > 2 || cfmt == 3) {
          const int lc = cfmt == 3 ? log2n : log2n - 1, cm = cu.cmode[cfmt == 3 ? pu : 0];
          for (int t = 0; t < nb; t++) { m.cb[t] = new_tb(); code_tb(1, x0 >> sx, (y0 >> sy) + (t << lc), lc, cm, *m.cb[t]); m.cbf_cb[t] = m.cb[t]->cbf; }
          for (int t = 0; t < nb; t++) { m.cr[t] = new_tb(); code_tb(2, x0 >> sx, (y0 >> sy) + (t << lc), lc, cm, *m.cr[t]); m.cbf_cr[t] = m.cr[t]->cbf; }
        } else if (blk == 3) {
          const Node& par = nodes[parent];
          for (int t = 0; t < nb; t++) { m.cb[t] = new_tb(); code_tb(1, par.x0 >> sx, (par.y0 >> sy) + (t << 2), 2, cu.cmode[0], *m.cb[t]); }
          for (int t = 0; t < nb; t++) { m.cr[t] = new_tb(); code_tb(2, par.x0 >> sx, (par.y0 >> sy) + (t << 2), 2, cu.cmode[0], *m.cr[t]); }
        }
      }
    }
    return me;
  }

  // parent_cb / parent_cr: the cbf_cb / cbf_cr flags of the parent node ([1]: lower 4:2:2 block)
  void write_tree(const Cu& cu, int me, const bool parent_cb[2], const bool parent_cr[2], int max_depth) {
    const Node& nd = nodes[me];
    bool can_split = nd.log2n <= log2_max_tb && nd.log2n > log2_min_tb && nd.depth < max_depth && !(cu.nxn && nd.depth == 0);
    if (can_split) cabac.bin(ctx[CTX_SPLIT_TR + 5 - nd.log2n], nd.split);
    bool cb[2] = {false, false}, cr[2] = {false, false};
    if (chroma) {
      if (nd.log2n > 2 || cfmt == 3) {
        const bool two = cfmt == 2 && (!nd.split || nd.log2n == 3);
        if (nd.depth == 0 || parent_cb[0]) { cb[0] = nd.cbf_cb[0]; cabac.bin(ctx[CTX_CBF_CHROMA + nd.depth], cb[0]); if (two) { cb[1] = nd.cbf_cb[1]; cabac.bin(ctx[CTX_CBF_CHROMA + nd.depth], cb[1]); } }
        if (nd.depth == 0 || parent_cr[0]) { cr[0] = nd.cbf_cr[0]; cabac.bin(ctx[CTX_CBF_CHROMA + nd.depth], cr[0]); if (two) { cr[1] = nd.cbf_cr[1]; cabac.bin(ctx[CTX_CBF_CHROMA + nd.depth], cr[1]); } }
      } else { cb[0] = parent_cb[0]; cb[1] = parent_cb[1]; cr[0] = parent_cr[0]; cr[1] = parent_cr[1]; }
    }
    if (nd.split) { for (int k = 0; k < 4; k++) write_tree(cu, nd.child[k], cb, cr, max_depth); return; }
    cabac.bin(ctx[CTX_CBF_LUMA + (nd.depth == 0 ? 1 : 0)], nd.cbf_l);
    bool cbf_chroma = chroma && (cb[0] || cb[1] || cr[0] || cr[1]);
    if ((nd.cbf_l || cbf_chroma) && P.cu_qp_delta && !is_qp_delta_coded) {
      int v = cu_qp_delta_val, a = std::abs(v);
      for (int k = 0; k < std::min(a, 5); k++) cabac.bin(ctx[CTX_QP_DELTA + (k ? 1 : 0)], 1);
      if (a < 5) cabac.bin(ctx[CTX_QP_DELTA + (a ? 1 : 0)], 0);
      else { int rem = a - 5, k = 0; while (rem >= (1 << k)) { cabac.bypass(1); rem -= 1 << k; k++; } cabac.bypass(0); cabac.bypass_bits(rem, k); }
      if (a) cabac.bypass(v < 0);
      is_qp_delta_coded = 1;
    }
    if (nd.cbf_l) write_residual(nd.l->lev, nd.log2n, 0, nd.l->scan, nd.l->tskip);
    if (chroma) {
      const int nb = cfmt == 2 ? 2 : 1;
      if (nd.log2n > 2 || cfmt == 3) {
        const int lc = cfmt == 3 ? nd.log2n : nd.log2n - 1;
        for (int t = 0; t < nb; t++) if (cb[t]) write_residual(nd.cb[t]->lev, lc, 1, nd.cb[t]->scan, nd.cb[t]->tskip);
        for (int t = 0; t < nb; t++) if (cr[t]) write_residual(nd.cr[t]->lev, lc, 2, nd.cr[t]->scan, nd.cr[t]->tskip);
      } else if (nd.blk == 3) {
        for (int t = 0; t < nb; t++) if (parent_cb[t]) write_residual(nd.cb[t]->lev, 2, 1, nd.cb[t]->scan, nd.cb[t]->tskip);
        for (int t = 0; t < nb; t++) if (parent_cr[t]) write_residual(nd.cr[t]->lev, 2, 2, nd.cr[t]->scan, nd.cr[t]->tskip);
      }
    }
  }

  // ---------------------------------------------------------------------------- coding unit (7.3.8.5)
  void mpm(int x, int y, int cand[3]) const {
    int ca = 1, cb = 1;
    if (avail(x - 1, y)) ca = ipm4[(size_t)(y >> 2) * w4 + ((x - 1) >> 2)];
    if (avail(x, y - 1) && (y - 1) >= ((y >> log2ctb) << log2ctb)) cb = ipm4[(size_t)((y - 1) >> 2) * w4 + (x >> 2)];
    if (ca == cb) {
      if (ca < 2) { cand[0] = 0; cand[1] = 1; cand[2] = 26; }
      else { cand[0] = ca; cand[1] = 2 + ((ca + 29) % 32); cand[2] = 2 + ((ca - 2 + 1) % 32); }
    } else {
      cand[0] = ca; cand[1] = cb;
      if (ca != 0 && cb != 0) cand[2] = 0; else if (ca != 1 && cb != 1) cand[2] = 1; else cand[2] = 26;
    }
  }

  int choose_mode(int x0, int y0, int log2n, const int cand[3]) {
    if (P.mode_decision == 0) return rng.range(35);
    const int n = 1 << log2n;
    int tries[8] = {0, 1, 10, 26, cand[0], 2 + rng.range(33), 2 + rng.range(33), 2 + rng.range(33)};
    int best = 0; long best_cost = -1;
    uint16_t pred[1024];
    int lg = std::min(log2n, 5);                       // evaluate on (at most) 32x32 at the CU origin
    int m = 1 << lg;
    for (int t = 0; t < 8; t++) {
      predict(0, x0, y0, lg, tries[t], pred);
      long sad = 0;
      for (int y = 0; y < m; y++) for (int x = 0; x < m; x++) sad += std::abs((int)org[0][(size_t)(y0 + y) * W + x0 + x] - pred[y * m + x]);
      if (best_cost < 0 || sad < best_cost) { best_cost = sad; best = tries[t]; }
    }
    (void)n;
    return best;
  }

  void coding_unit(int x0, int y0, int log2cb, int depth) {
    Cu cu{}; cu.x0 = x0; cu.y0 = y0; cu.log2cb = log2cb;
    const int n = 1 << log2cb;
    cu_bypass = false;
    if (P.transquant_bypass) { cu_bypass = P.transquant_bypass == 2 || rng.range(4) == 0; cabac.bin(ctx[CTX_TQ_BYPASS], cu_bypass); }
    if (log2cb == 3) { cu.nxn = rng.range(3) == 0; cabac.bin(ctx[CTX_PART_MODE], !cu.nxn); }
    if (P.pcm && !cu.nxn && log2cb <= std::min(5, log2ctb)) {
      const bool pcm = rng.range(6) == 0;
      cabac.terminate(pcm ? 1 : 0);                   // pcm_flag (terminate bin); value 1: flush, stop bit, pcm_alignment_zero_bits
      if (pcm) {
        for (int c = 0; c < (chroma ? 3 : 1); c++) {
          const int shx = c ? sx : 0, shy = c ? sy : 0, pbd = c ? pcm_bd_c : pcm_bd_y, st = stride_of(c);
          uint16_t* rq = rec[c].data();
          for (int y = 0; y < (n >> shy); y++) for (int x = 0; x < (n >> shx); x++) {
            const size_t idx = (size_t)((y0 >> shy) + y) * st + (x0 >> shx) + x;
            const unsigned v = org[c][idx] >> (bd - pbd);
            cabac.bw.put(v, pbd);                     // pcm_sample_luma / pcm_sample_chroma
            rq[idx] = (uint16_t)(v << (bd - pbd));
          }
        }
        cabac.restart();
        for (int yy = 0; yy < n; yy += 4) for (int xx = 0; xx < n; xx += 4) {
          const size_t idx = (size_t)((y0 + yy) >> 2) * w4 + ((x0 + xx) >> 2);
          ipm4[idx] = 1; slice_of4[idx] = (uint16_t)(region_idx + 1); cd4[idx] = (uint8_t)depth;       // a PCM unit counts as INTRA_DC for its neighbours (8.4.2)
        }
        // no transform tree, no cu_qp_delta: QpY = predicted QP (+ the delta already coded in this quantization group, 8.6.1)
        const int pred_qp = P.cu_qp_delta ? predict_qpy(x0, y0) : slice_qp;
        cur_qpy = P.cu_qp_delta ? pred_qp + (is_qp_delta_coded ? cu_qp_delta_val : 0) : slice_qp;
        for (int yy = 0; yy < n; yy += 4) for (int xx = 0; xx < n; xx += 4) qp4[(size_t)((y0 + yy) >> 2) * w4 + ((x0 + xx) >> 2)] = (int8_t)cur_qpy;
        last_cu_qpy = cur_qpy;
        return;
      }
    }
    int np = cu.nxn ? 4 : 1, pb = cu.nxn ? n / 2 : n;
    int prev[4], mpm_idx[4], rem[4];
    for (int i = 0; i < np; i++) {
      int px = x0 + (i & 1) * pb, py = y0 + (i >> 1) * pb, cand[3];
      mpm(px, py, cand);
      int mode = choose_mode(px, py, cu.nxn ? 2 : log2cb, cand);
      cu.lmode[i] = mode;
      prev[i] = 0; mpm_idx[i] = 0; rem[i] = 0;
      for (int k = 0; k < 3; k++) if (cand[k] == mode) { prev[i] = 1; mpm_idx[i] = k; }
      if (!prev[i]) {
        int s[3] = {cand[0], cand[1], cand[2]}; std::sort(s, s + 3);
        int r = mode; for (int k = 2; k >= 0; k--) if (r > s[k]) r--;
        rem[i] = r;
      }
      for (int yy = 0; yy < pb; yy += 4) for (int xx = 0; xx < pb; xx += 4) {
        size_t idx = (size_t)((py + yy) >> 2) * w4 + ((px + xx) >> 2);
        ipm4[idx] = (uint8_t)mode; slice_of4[idx] = (uint16_t)(region_idx + 1);
      }
    }
    for (int i = 0; i < np; i++) cabac.bin(ctx[CTX_PREV_INTRA], prev[i]);
    for (int i = 0; i < np; i++) {
      if (prev[i]) { cabac.bypass(mpm_idx[i] > 0); if (mpm_idx[i] > 0) cabac.bypass(mpm_idx[i] > 1); }
      else cabac.bypass_bits(rem[i], 5);
    }
    if (chroma) {
      // intra_chroma_pred_mode: one per prediction unit in 4:4:4, else one per coding unit (7.3.8.5); 8.4.3 + Table 8-3 for 4:2:2
      static const uint8_t tab[4] = {0, 26, 10, 1};
      static const uint8_t k422[35] = {0, 1, 2, 2, 2, 2, 3, 5, 7, 8, 10, 12, 13, 15, 17, 18, 19, 20, 21, 22, 23, 23, 24, 24, 25, 25, 26, 27, 27, 28, 28, 29, 29, 30, 31};
      for (int i = 0; i < (cfmt == 3 ? np : 1); i++) {
        int v = rng.range(8); if (v > 4) v = 4;
        int m = (v < 4 && tab[v] == cu.lmode[i]) ? 34 : (v == 4 ? cu.lmode[i] : tab[v]);
        if (cfmt == 2) m = k422[m];
        cu.cmode[i] = m;
        cabac.bin(ctx[CTX_CHROMA_PRED], v != 4);
        if (v != 4) cabac.bypass_bits(v, 2);
      }
    }
    for (int yy = 0; yy < n; yy += 4) for (int xx = 0; xx < n; xx += 4) {
      size_t idx = (size_t)((y0 + yy) >> 2) * w4 + ((x0 + xx) >> 2);
      slice_of4[idx] = 0; cd4[idx] = (uint8_t)depth;
    }
    // QP of this CU: target QP if the delta of this quantization group was (or can still be) sent, else predicted
    int pred_qp = P.cu_qp_delta ? predict_qpy(x0, y0) : slice_qp;
    int qbd = 6 * (bd - 8);
    if (P.cu_qp_delta && !is_qp_delta_coded) {
      cu_qp_delta_val = clip3(-(26 + qbd / 2), 25 + qbd / 2, qg_target_qp - pred_qp);
      qg_target_qp = pred_qp + cu_qp_delta_val;
    } else if (!P.cu_qp_delta) qg_target_qp = slice_qp;
    cur_qpy = qg_target_qp;
    nodes.clear();
    int max_depth = max_th_depth + cu.nxn;
    int root = build_tree(cu, x0, y0, log2cb, 0, 0, max_depth, -1);
    bool any_cbf = false;
    for (const Node& nd : nodes) if (!nd.split) for (int t = 0; t < 2; t++) any_cbf |= nd.cbf_l || (nd.cb[t] && nd.cb[t]->cbf) || (nd.cr[t] && nd.cr[t]->cbf);
    if (P.cu_qp_delta && !is_qp_delta_coded && !any_cbf) {
      // nothing coded: the decoder will use the predicted QP (CuQpDeltaVal stays 0); pixels are pure prediction so
      // the reconstruction above is already what the decoder produces.
      cur_qpy = pred_qp;
    }
    { const bool none[2] = {false, false}; write_tree(cu, root, none, none, max_depth); }
    for (TbResult* t : pool) delete t;
    pool.clear();
    for (int yy = 0; yy < n; yy += 4) for (int xx = 0; xx < n; xx += 4) qp4[(size_t)((y0 + yy) >> 2) * w4 + ((x0 + xx) >> 2)] = (int8_t)cur_qpy;
    last_cu_qpy = cur_qpy;
  }

  long block_activity(int x0, int y0, int n) const {
    long s = 0, s2 = 0;
    for (int y = 0; y < n; y += 2) for (int x = 0; x < n; x += 2) { int v = org[0][(size_t)(y0 + y) * W + x0 + x]; s += v; s2 += (long)v * v; }
    long cnt = (long)(n / 2) * (n / 2);
    return (s2 - s * s / cnt) / cnt >> (2 * (bd - 8));
  }

  void coding_quadtree(int x0, int y0, int log2cb, int depth) {
    int n = 1 << log2cb;
    bool split;
    if (x0 + n <= W && y0 + n <= H && log2cb > 3) {
      long act = block_activity(x0, y0, n);
      int r = rng.range(8);
      split = P.mode_decision == 0 ? r < 4 : (act > (long)P.split_threshold * (log2cb - 2) ? r != 0 : r == 0);
      int inc = 0;
      if (avail(x0 - 1, y0) && cd4[(size_t)(y0 >> 2) * w4 + ((x0 - 1) >> 2)] > depth) inc++;
      if (avail(x0, y0 - 1) && cd4[(size_t)((y0 - 1) >> 2) * w4 + (x0 >> 2)] > depth) inc++;
      cabac.bin(ctx[CTX_SPLIT_CU + inc], split);
    } else split = log2cb > 3;
    if (P.cu_qp_delta && log2cb >= qg_log2) {
      is_qp_delta_coded = 0; cu_qp_delta_val = 0;
      if (!split || log2cb == qg_log2) {
        if (first_qg) { qpy_prev_qg = slice_qp; first_qg = 0; } else qpy_prev_qg = last_cu_qpy;
        qg_target_qp = clip3(0, 51, slice_qp + (P.dqp_range ? rng.range(2 * P.dqp_range + 1) - P.dqp_range : 0));
      }
    }
    if (split) {
      int h = n >> 1;
      for (int k = 0; k < 4; k++) { int x1 = x0 + (k & 1) * h, y1 = y0 + (k >> 1) * h; if (x1 < W && y1 < H) coding_quadtree(x1, y1, log2cb - 1, depth + 1); }
    } else coding_unit(x0, y0, log2cb, depth);
  }
};

}  // namespace enc
}  // namespace b200

extern "C" {

void b200_hevc_enc_params_default(b200_hevc_enc_params* p) {
  memset(p, 0, sizeof *p);
  p->tile_cols = p->tile_rows = 1; p->tiles_uniform = 1; p->loop_filter_across_tiles = 1; p->slice_per_tile = 0;
  p->bit_depth = 8; p->chroma_format_idc = 1; p->log2_ctb_size = 5; p->qp = 27; p->init_qp = 26;
  p->max_transform_hierarchy_depth_intra = 1; p->sao = 1; p->sign_data_hiding = 1; p->cu_qp_delta = 1;
  p->diff_cu_qp_delta_depth = 1; p->dqp_range = 3; p->strong_intra_smoothing = 1; p->loop_filter_across_slices = 1;
  p->slice_loop_filter_across_slices = 1; p->mode_decision = 1; p->split_threshold = 40; p->seed = 0xB200;
  p->still_picture = 1; p->vui_present = 0; p->colour_primaries = 2; p->transfer_characteristics = 2; p->matrix_coefficients = 2;
}

int b200_hevc_encode_intra(const b200_hevc_enc_params* p, const void* y, const void* cb, const void* cr, size_t y_stride,
                           size_t c_stride, uint8_t** out_data, size_t* out_size) {
  using namespace b200;
  if (!p || !y || !out_data || !out_size) return set_error(B200_E_INVALID, "null argument");
  if (p->width < 8 || p->height < 8 || p->width > 16384 || p->height > 16384) return set_error(B200_E_INVALID, "size %dx%d", p->width, p->height);
  if (p->bit_depth < 8 || p->bit_depth > 12) return set_error(B200_E_UNSUPPORTED, "bit depth %d", p->bit_depth);
  if (p->log2_ctb_size < 4 || p->log2_ctb_size > 6) return set_error(B200_E_INVALID, "log2_ctb_size %d", p->log2_ctb_size);
  if (p->chroma_format_idc && (!cb || !cr)) return set_error(B200_E_INVALID, "missing chroma planes");
  const int bps = p->bit_depth > 8 ? 2 : 1;
  std::vector<uint16_t> tmp[3];
  const uint16_t* src[3] = {nullptr, nullptr, nullptr}; int stride[3] = {0, 0, 0};
  const void* in[3] = {y, cb, cr};
  for (int c = 0; c < (p->chroma_format_idc ? 3 : 1); c++) {
    const int csx = (p->chroma_format_idc == 1 || p->chroma_format_idc == 2) ? 1 : 0, csy = p->chroma_format_idc == 1 ? 1 : 0;
    int w = c ? (p->width + (1 << csx) - 1) >> csx : p->width, h = c ? (p->height + (1 << csy) - 1) >> csy : p->height;
    size_t st = c ? c_stride : y_stride;
    tmp[c].resize((size_t)w * h);
    for (int yy = 0; yy < h; yy++) for (int xx = 0; xx < w; xx++)
      tmp[c][(size_t)yy * w + xx] = bps == 1 ? ((const uint8_t*)in[c])[yy * st + xx] : ((const uint16_t*)((const uint8_t*)in[c] + yy * st))[xx];
    src[c] = tmp[c].data(); stride[c] = w;
  }
  enc::Encoder e(*p, src, stride);
  std::vector<uint8_t> out;
  e.encode(out);
  *out_data = (uint8_t*)malloc(out.size() ? out.size() : 1);
  if (!*out_data) return set_error(B200_E_INVALID, "out of memory");
  memcpy(*out_data, out.data(), out.size());
  *out_size = out.size();
  return B200_OK;
}

void b200_free(void* p) { free(p); }

}  // extern "C"
