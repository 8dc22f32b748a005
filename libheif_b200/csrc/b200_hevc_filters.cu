// b200_hevc_filters.cu -- K3 deblocking (H.265 8.7.2) and K4 SAO (8.7.3) + K5 conformance crop / tile paste.
//
// The in-loop filters libde265 applies before libheif/plugins/decoder_libde265.cc:97-171 copies the planes out.
// Intra pictures: bS = 2 on every transform edge of the 8x8 grid (host supplies the filterEdgeFlag map and the
// QpY map, one byte per 8x8 block each).  Deblocking runs in place on the reconstruction planes: one pass over all
// vertical edges of all pictures of the batch, then one over all horizontal edges (8.7.2 orders them picture-wide;
// edges of one direction are independent because they are 8 samples apart and touch at most 3 samples per side).
// SAO reads the deblocked planes and writes the final samples straight into the destination (conformance window
// applied, destination pointer already offset to the tile's paste position: K5 folded into K4's store).
#include <cstdlib>
#include "b200_hevc.h"

namespace b200 {

__constant__ uint8_t c_tc[54] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,1,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,5,5,6,6,7,8,9,10,11,13,14,16,18,20,22,24};
__constant__ uint8_t c_beta[52] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,6,7,8,9,10,11,12,13,14,15,16,17,18,20,22,24,26,28,30,32,34,36,38,40,42,44,46,48,50,52,54,56,58,60,62,64};
__constant__ uint8_t c_qpc2[14] = {29, 30, 31, 32, 33, 33, 34, 34, 35, 35, 36, 36, 37, 37};

__device__ __forceinline__ int clip3d(int lo, int hi, int v) { return min(max(v, lo), hi); }

// one thread = one 4-sample luma edge segment (+ the 2-sample chroma segments lying on it)
template <typename T, int VERT>
__global__ void __launch_bounds__(256) deblock_kernel(const DeviceBatch b, int max_w8, int max_seg) {
  const PicDesc& pic = b.pics[blockIdx.y];
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  // VERT: edges at x = 8*e, segments along y (4 rows each); else edges at y = 8*e, segments along x
  const int nedge = VERT ? pic.w8 : pic.h8, nseg = (VERT ? pic.height : pic.width) >> 2;
  // consecutive threads walk along the contiguous image direction: edges for the vertical pass, segments for the horizontal one
  const int e = VERT ? idx % max_w8 : idx / max_seg, sg = VERT ? idx / max_w8 : idx % max_seg;
  if (e >= nedge || sg >= nseg || e == 0) return;
  const int x = VERT ? e * 8 : sg * 4, y = VERT ? sg * 4 : e * 8;
  const uint8_t* edge8 = b.edge8 + pic.map8_base;
  const int8_t* qp8 = b.qp8 + pic.map8_base;
  const int bi = (y >> 3) * pic.w8 + (x >> 3);
  if (!(edge8[bi] & (VERT ? 1 : 2))) return;
  const int bj = VERT ? bi - 1 : bi - pic.w8;
  const bool keep_q = edge8[bi] & 4, keep_p = edge8[bj] & 4;       // cu_transquant_bypass, pcm + pcm_loop_filter_disabled: nDq / nDp = 0 (8.7.2.5.7)
  const CtuInfo& ci = b.ctus[pic.ctu_base + (y >> pic.log2_ctb) * pic.wctb + (x >> pic.log2_ctb)];
  const SliceInfo sl = b.slices[pic.slice_base + ci.slice_idx];
  const int bd = pic.bit_depth, maxv = (1 << bd) - 1;
  const int qpl = (qp8[bi] + qp8[bj] + 1) >> 1;
  {
    const int beta = c_beta[clip3d(0, 51, qpl + sl.beta_offset)] * (1 << (bd - 8));
    const int tc = c_tc[clip3d(0, 53, qpl + 2 + sl.tc_offset)] * (1 << (bd - 8));
    T* pl = static_cast<T*>(pic.rec[0]);
    const int st = pic.rec_stride[0];
    const int xs = VERT ? 1 : st, ls = VERT ? st : 1;
    T* q = pl + (size_t)y * st + x;
    int s[4][8];                                        // [line][p3 p2 p1 p0 q0 q1 q2 q3]
#pragma unroll
    for (int l = 0; l < 4; l++)
#pragma unroll
      for (int k = 0; k < 8; k++) s[l][k] = q[(ptrdiff_t)(k - 4) * xs + (ptrdiff_t)l * ls];
#define P_(k, l) s[l][3 - (k)]
#define Q_(k, l) s[l][4 + (k)]
    const int dp0 = abs(P_(2, 0) - 2 * P_(1, 0) + P_(0, 0)), dp3 = abs(P_(2, 3) - 2 * P_(1, 3) + P_(0, 3));
    const int dq0 = abs(Q_(2, 0) - 2 * Q_(1, 0) + Q_(0, 0)), dq3 = abs(Q_(2, 3) - 2 * Q_(1, 3) + Q_(0, 3));
    const int dpq0 = dp0 + dq0, dpq3 = dp3 + dq3, dp = dp0 + dp3, dq = dq0 + dq3;
    if (dpq0 + dpq3 < beta) {                            // 8.7.2.5.3
      const bool s0 = 2 * dpq0 < (beta >> 2) && abs(P_(3, 0) - P_(0, 0)) + abs(Q_(0, 0) - Q_(3, 0)) < (beta >> 3) && abs(P_(0, 0) - Q_(0, 0)) < ((5 * tc + 1) >> 1);
      const bool s3 = 2 * dpq3 < (beta >> 2) && abs(P_(3, 3) - P_(0, 3)) + abs(Q_(0, 3) - Q_(3, 3)) < (beta >> 3) && abs(P_(0, 3) - Q_(0, 3)) < ((5 * tc + 1) >> 1);
      const bool strong = s0 && s3;
      const bool dep = dp < ((beta + (beta >> 1)) >> 3), deq = dq < ((beta + (beta >> 1)) >> 3);
#pragma unroll
      for (int l = 0; l < 4; l++) {                       // 8.7.2.5.7
        const int p0 = P_(0, l), p1 = P_(1, l), p2 = P_(2, l), p3 = P_(3, l), q0 = Q_(0, l), q1 = Q_(1, l), q2 = Q_(2, l), q3 = Q_(3, l);
        T* ql = q + (ptrdiff_t)l * ls;
        if (strong) {
          if (!keep_p) {
            ql[-1 * (ptrdiff_t)xs] = (T)clip3d(p0 - 2 * tc, p0 + 2 * tc, (p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
            ql[-2 * (ptrdiff_t)xs] = (T)clip3d(p1 - 2 * tc, p1 + 2 * tc, (p2 + p1 + p0 + q0 + 2) >> 2);
            ql[-3 * (ptrdiff_t)xs] = (T)clip3d(p2 - 2 * tc, p2 + 2 * tc, (2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
          }
          if (!keep_q) {
            ql[0] = (T)clip3d(q0 - 2 * tc, q0 + 2 * tc, (p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
            ql[xs] = (T)clip3d(q1 - 2 * tc, q1 + 2 * tc, (p0 + q0 + q1 + q2 + 2) >> 2);
            ql[2 * (ptrdiff_t)xs] = (T)clip3d(q2 - 2 * tc, q2 + 2 * tc, (p0 + q0 + q1 + 3 * q2 + 2 * q3 + 4) >> 3);
          }
        } else {
          int delta = (9 * (q0 - p0) - 3 * (q1 - p1) + 8) >> 4;
          if (abs(delta) < tc * 10) {
            delta = clip3d(-tc, tc, delta);
            if (!keep_p) ql[-1 * (ptrdiff_t)xs] = (T)clip3d(0, maxv, p0 + delta);
            if (!keep_q) ql[0] = (T)clip3d(0, maxv, q0 - delta);
            if (dep && !keep_p) ql[-2 * (ptrdiff_t)xs] = (T)clip3d(0, maxv, p1 + clip3d(-(tc >> 1), tc >> 1, (((p2 + p0 + 1) >> 1) - p1 + delta) >> 1));
            if (deq && !keep_q) ql[xs] = (T)clip3d(0, maxv, q1 + clip3d(-(tc >> 1), tc >> 1, (((q2 + q0 + 1) >> 1) - q1 - delta) >> 1));
          }
        }
      }
    }
#undef P_
#undef Q_
  }
  // chroma: edges on the 8-sample CHROMA grid (every 16 luma samples where the chroma is sub-sampled across the edge), bS == 2 only
  // (8.7.2.5.5 / 8.7.2.5.8); a 4-luma-sample segment covers 4 >> (sub-sampling along the edge) chroma samples
  const int fsx = (pic.chroma == 1 || pic.chroma == 2) ? 1 : 0, fsy = pic.chroma == 1 ? 1 : 0;
  if (pic.chroma && ((VERT ? x : y) & ((8 << (VERT ? fsx : fsy)) - 1)) == 0) {
    const int len = 4 >> (VERT ? fsy : fsx);
    for (int c = 1; c <= 2; c++) {
      const int qpi = qpl + (c == 1 ? pic.pps_cb_qp_offset : pic.pps_cr_qp_offset);
      const int qpc = pic.chroma != 1 ? min(qpi, 51) : (qpi < 30 ? qpi : (qpi >= 43 ? qpi - 6 : c_qpc2[qpi - 30]));   // Table 8-10 only when ChromaArrayType == 1
      const int tc = c_tc[clip3d(0, 53, qpc + 2 + sl.tc_offset)] * (1 << (bd - 8));
      T* pl = static_cast<T*>(pic.rec[c]);
      const int st = pic.rec_stride[c];
      const int xs = VERT ? 1 : st, ls = VERT ? st : 1;
      T* q = pl + (size_t)(y >> fsy) * st + (x >> fsx);
#pragma unroll
      for (int l = 0; l < 4; l++) {
        if (l >= len) break;
        T* ql = q + (ptrdiff_t)l * ls;
        const int p0 = ql[-(ptrdiff_t)xs], p1 = ql[-2 * (ptrdiff_t)xs], q0 = ql[0], q1 = ql[xs];
        const int delta = clip3d(-tc, tc, ((((q0 - p0) << 2) + p1 - q1 + 4) >> 3));
        if (!keep_p) ql[-(ptrdiff_t)xs] = (T)clip3d(0, maxv, p0 + delta);
        if (!keep_q) ql[0] = (T)clip3d(0, maxv, q0 - delta);
      }
    }
  }
}

// SAO offset of ONE sample with every rule of 8.7.3 (picture bounds, slice boundaries with
// slice_loop_filter_across_slices_enabled_flag): the path for pictures with more than one slice.
template <typename T>
__device__ __noinline__ int sao_sample_generic(const DeviceBatch& b, const PicDesc& pic, const CtuInfo* ctus, const T* src, int st, int c, int x, int y, int w, int h, int lg, int lgy) {   // lg / lgy: log2 of the component's CTB width / height
  int v = src[(size_t)y * st + x];
  const CtuInfo& ci = ctus[(y >> lgy) * pic.wctb + (x >> lg)];
  const SaoComp sp = ci.sao[c];
  if (!pic.sao_enabled || !sp.type) return v;
  const int bd = pic.bit_depth, maxv = (1 << bd) - 1;
  int off = 0;
  if (sp.type == 1) {
    const int k = ((v >> (bd - 5)) - sp.band_or_class) & 31;
    if (k < 4) off = sp.offset[k];
  } else {
    const int e = sp.band_or_class;
    const int hx = e == 1 ? 0 : (e == 3 ? 1 : -1), vy = e == 0 ? 0 : -1;      // first neighbour; second is the opposite
    const int xa = x + hx, ya = y + vy, xb = x - hx, yb = y - vy;
    if (xa >= 0 && xb >= 0 && ya >= 0 && yb >= 0 && xa < w && xb < w && ya < h && yb < h) {
      bool skip = false;
      const SliceInfo* sls = b.slices + pic.slice_base;
      const int cur = ci.slice_idx;
      const int sa = ctus[(ya >> lgy) * pic.wctb + (xa >> lg)].slice_idx, sb = ctus[(yb >> lgy) * pic.wctb + (xb >> lg)].slice_idx;
      // 8.7.3.2: a neighbour in another slice counts only if the later of the two slices filters across its boundary; one in another
      // tile only with loop_filter_across_tiles_enabled_flag (regions = slice x tile, compared through their slice / tile ids)
      const SliceInfo c0 = sls[cur];
      if (sa != cur) { const SliceInfo o = sls[sa]; if ((o.slice_id != c0.slice_id && !(o.slice_id < c0.slice_id ? c0.lf_across_slices : o.lf_across_slices)) || (o.tile_id != c0.tile_id && !c0.lf_across_tiles)) skip = true; }
      if (sb != cur) { const SliceInfo o = sls[sb]; if ((o.slice_id != c0.slice_id && !(o.slice_id < c0.slice_id ? c0.lf_across_slices : o.lf_across_slices)) || (o.tile_id != c0.tile_id && !c0.lf_across_tiles)) skip = true; }
      if (!skip) {
        const int a = src[(size_t)ya * st + xa], bb = src[(size_t)yb * st + xb];
        const int ei = 2 + (v > a) - (v < a) + (v > bb) - (v < bb);
        if (ei != 2) off = sp.offset[ei < 2 ? ei : ei - 1];           // edgeIdx 0,1,3,4 -> SaoOffsetVal[1..4]
      }
    }
  }
  return clip3d(0, maxv, v + off);
}

// One thread = 8 horizontally adjacent samples of one row of one component (always inside one CTB: a chroma CTB is at
// least 8 samples wide); block = 64 segments x 4 rows; grid.z = picture * 3 + component.  The CTB's SAO parameters are
// fetched once per thread, the three source rows with 8- / 16-byte loads, the result leaves with one vector store.
template <typename T>
__global__ void __launch_bounds__(256) sao_kernel(const DeviceBatch b) {
  const int pi = blockIdx.z / 3, c = blockIdx.z % 3;
  const PicDesc& pic = b.pics[pi];
  if (c > 0 && !pic.chroma) return;
  const int sh = (c && (pic.chroma == 1 || pic.chroma == 2)) ? 1 : 0, shy = (c && pic.chroma == 1) ? 1 : 0;     // horizontal / vertical sub-sampling of the component
  const int w = pic.width >> sh, h = pic.height >> shy;
  const int x0 = (blockIdx.x * 64 + threadIdx.x) * 8, y = blockIdx.y * 4 + threadIdx.y;
  if (x0 >= w || y >= h) return;
  const int cx = pic.crop_x >> sh, cy = pic.crop_y >> shy, ow = (pic.out_w + sh) >> sh, oh = (pic.out_h + shy) >> shy;
  const int oy = y - cy;
  if (oy < 0 || oy >= oh) return;                                     // outside the conformance window: never output
  const T* src = static_cast<const T*>(pic.rec[c]);
  const int st = pic.rec_stride[c];
  const int lg = pic.log2_ctb - sh, lgy = pic.log2_ctb - shy;
  const CtuInfo* ctus = b.ctus + pic.ctu_base;
  const int n = min(8, w - x0);
  int res[8];
  if (pic.nslices > 1) {
#pragma unroll
    for (int k = 0; k < 8; k++) res[k] = k < n ? sao_sample_generic<T>(b, pic, ctus, src, st, c, x0 + k, y, w, h, lg, lgy) : 0;
  } else {
    // centre row: samples x0-1 .. x0+8 in cur[0..9]
    int cur[10], up[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, dn[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const T* row = src + (size_t)y * st + x0;
    auto load8 = [&](const T* p, int* d) {                            // 8 samples, vector load (rows and x0 are 8-sample aligned)
      if (sizeof(T) == 1) { const uint2 q = *reinterpret_cast<const uint2*>(p); for (int k = 0; k < 4; k++) { d[k] = (q.x >> (8 * k)) & 0xff; d[4 + k] = (q.y >> (8 * k)) & 0xff; } }
      else { const uint4 q = *reinterpret_cast<const uint4*>(p); const unsigned u[4] = {q.x, q.y, q.z, q.w}; for (int k = 0; k < 4; k++) { d[2 * k] = u[k] & 0xffff; d[2 * k + 1] = u[k] >> 16; } }
    };
    if (n == 8) load8(row, cur + 1);
    else {
#pragma unroll
      for (int k = 0; k < 8; k++) cur[1 + k] = k < n ? (int)row[k] : 0;
    }
    const SaoComp sp = ctus[(y >> lgy) * pic.wctb + (x0 >> lg)].sao[c];
    const int bd = pic.bit_depth, maxv = (1 << bd) - 1;
    const unsigned offs = (unsigned)(uint8_t)sp.offset[0] | ((unsigned)(uint8_t)sp.offset[1] << 8) | ((unsigned)(uint8_t)sp.offset[2] << 16) | ((unsigned)(uint8_t)sp.offset[3] << 24);
    auto offset = [&](int i) { return (int)(int8_t)(offs >> (8 * i)); };     // register-resident SaoOffsetVal[1..4]
    const bool on = pic.sao_enabled && sp.type;
    if (!on) { for (int k = 0; k < 8; k++) res[k] = cur[1 + k]; }
    else if (sp.type == 1) {
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int v = cur[1 + k], kk = ((v >> (bd - 5)) - sp.band_or_class) & 31;
        res[k] = clip3d(0, maxv, v + (kk < 4 ? offset(kk) : 0));
      }
    } else {
      const int e = sp.band_or_class;
      const int hx = e == 1 ? 0 : (e == 3 ? 1 : -1);
      const bool vert = e != 0;                                        // neighbours in the rows above / below
      const bool has_l = x0 > 0, has_r = x0 + 8 < w;
      cur[0] = has_l ? (int)row[-1] : 0; cur[9] = has_r ? (int)row[8] : 0;
      const bool rows_ok = !vert || (y > 0 && y + 1 < h);
      if (vert && rows_ok) {
        const T* ru = row - st; const T* rd = row + st;
        if (n == 8) { load8(ru, up + 1); load8(rd, dn + 1); }
        else {
#pragma unroll
          for (int k = 0; k < 8; k++) { up[1 + k] = k < n ? (int)ru[k] : 0; dn[1 + k] = k < n ? (int)rd[k] : 0; }
        }
        up[0] = has_l ? (int)ru[-1] : 0; up[9] = has_r ? (int)ru[8] : 0;
        dn[0] = has_l ? (int)rd[-1] : 0; dn[9] = has_r ? (int)rd[8] : 0;
      }
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int v = cur[1 + k];
        int off = 0;
        const int xa = x0 + k + hx, xb = x0 + k - hx;
        if (rows_ok && xa >= 0 && xb >= 0 && xa < w && xb < w) {
          // static indices only (hx is uniform per CTB): the arrays stay in registers
          const int a_l = vert ? up[k] : cur[k], a_m = up[1 + k], a_r = vert ? up[2 + k] : cur[2 + k];
          const int b_l = vert ? dn[k] : cur[k], b_m = dn[1 + k], b_r = vert ? dn[2 + k] : cur[2 + k];
          const int a = hx < 0 ? a_l : (hx > 0 ? a_r : a_m), bb = hx < 0 ? b_r : (hx > 0 ? b_l : b_m);
          const int ei = 2 + (v > a) - (v < a) + (v > bb) - (v < bb);
          if (ei != 2) off = offset(ei < 2 ? ei : ei - 1);              // edgeIdx 0,1,3,4 -> SaoOffsetVal[1..4]
        }
        res[k] = clip3d(0, maxv, v + off);
      }
    }
  }
  {
    // cu_transquant_bypass / pcm + pcm_loop_filter_disabled (8.7.3: SaoTypeIdx is treated as 0 there): bit 2 of the 8x8 luma cells
    const uint8_t* cell = b.edge8 + pic.map8_base + ((y << shy) >> 3) * pic.w8 + ((x0 << sh) >> 3);
    const bool k0 = cell[0] & 4, k1 = sh ? (((x0 + 4) << 1) < pic.width && (cell[1] & 4)) : k0;
    if (k0 | k1) {
      const T* row = src + (size_t)y * st + x0;
#pragma unroll
      for (int k = 0; k < 8; k++) if (k < n && (k < 4 ? k0 : k1)) res[k] = (int)row[k];
    }
  }
  // store (conformance window applied; destination already offset to the tile's paste position)
  T* drow = static_cast<T*>(pic.dst[c]) + (size_t)oy * pic.dst_stride[c];
  const int ox0 = x0 - cx;
  if (n == 8 && ox0 >= 0 && ox0 + 8 <= ow && ((reinterpret_cast<uintptr_t>(drow + ox0) & (8 * sizeof(T) - 1)) == 0)) {
    if (sizeof(T) == 1) {
      uint2 q; q.x = res[0] | (res[1] << 8) | (res[2] << 16) | (res[3] << 24); q.y = res[4] | (res[5] << 8) | (res[6] << 16) | (res[7] << 24);
      *reinterpret_cast<uint2*>(drow + ox0) = q;
    } else {
      uint4 q; q.x = res[0] | (res[1] << 16); q.y = res[2] | (res[3] << 16); q.z = res[4] | (res[5] << 16); q.w = res[6] | (res[7] << 16);
      *reinterpret_cast<uint4*>(drow + ox0) = q;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 8; k++) { const int ox = ox0 + k; if (k < n && ox >= 0 && ox < ow) drow[ox] = (T)res[k]; }
  }
}

// ---- K4, second form (default): one thread = 8 horizontally adjacent samples (one SAO unit: always inside one CTB) of FOUR
// consecutive rows (a 4-aligned row group never straddles a CTB either).  All six source rows a thread needs (the row above
// its group .. the row below) are requested up front with 8- / 16-byte loads -- six independent loads in flight per thread
// instead of a dependent chain of three -- and the horizontal neighbours come from the adjacent lanes by shuffle (one warp =
// 256 consecutive samples of a row; only lanes 0 and 31 load their outer neighbour).  The CTB's parameters, the
// bypass / PCM cell and the picture descriptor are fetched once per 32 samples.  Measured on the bench grid: see profiles/README.md.
// (The first form above walks one row per thread: its 393 K blocks of 2 KB were bound by block turnover and load latency; walking
// four rows one after the other in a rolled loop was slower still.)
template <typename T> struct SaoPack;
template <> struct SaoPack<uint8_t> {
  uint2 q;
  __device__ __forceinline__ void zero() { q = make_uint2(0, 0); }
  __device__ __forceinline__ void load_vec(const uint8_t* p) { q = *reinterpret_cast<const uint2*>(p); }
  __device__ __forceinline__ void load_n(const uint8_t* p, int n) { unsigned long long v = 0; for (int k = 0; k < 8; k++) if (k < n) v |= (unsigned long long)p[k] << (8 * k); q.x = (unsigned)v; q.y = (unsigned)(v >> 32); }
  __device__ __forceinline__ int first() const { return (int)(q.x & 0xffu); }
  __device__ __forceinline__ int last() const { return (int)(q.y >> 24); }
  __device__ __forceinline__ void unpack(int* d) const {
#pragma unroll
    for (int k = 0; k < 4; k++) { d[k] = (int)((q.x >> (8 * k)) & 0xffu); d[4 + k] = (int)((q.y >> (8 * k)) & 0xffu); }
  }
  __device__ __forceinline__ static void store_vec(uint8_t* p, const int* r) {
    uint2 o; o.x = (unsigned)r[0] | ((unsigned)r[1] << 8) | ((unsigned)r[2] << 16) | ((unsigned)r[3] << 24); o.y = (unsigned)r[4] | ((unsigned)r[5] << 8) | ((unsigned)r[6] << 16) | ((unsigned)r[7] << 24);
    *reinterpret_cast<uint2*>(p) = o;
  }
};
template <> struct SaoPack<uint16_t> {
  uint4 q;
  __device__ __forceinline__ void zero() { q = make_uint4(0, 0, 0, 0); }
  __device__ __forceinline__ void load_vec(const uint16_t* p) { q = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void load_n(const uint16_t* p, int n) {
    unsigned u[4] = {0, 0, 0, 0};
    for (int k = 0; k < 8; k++) if (k < n) u[k >> 1] |= (unsigned)p[k] << (16 * (k & 1));
    q = make_uint4(u[0], u[1], u[2], u[3]);
  }
  __device__ __forceinline__ int first() const { return (int)(q.x & 0xffffu); }
  __device__ __forceinline__ int last() const { return (int)(q.w >> 16); }
  __device__ __forceinline__ void unpack(int* d) const {
    const unsigned u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; k++) { d[2 * k] = (int)(u[k] & 0xffffu); d[2 * k + 1] = (int)(u[k] >> 16); }
  }
  __device__ __forceinline__ static void store_vec(uint16_t* p, const int* r) {
    uint4 o; o.x = (unsigned)r[0] | ((unsigned)r[1] << 16); o.y = (unsigned)r[2] | ((unsigned)r[3] << 16); o.z = (unsigned)r[4] | ((unsigned)r[5] << 16); o.w = (unsigned)r[6] | ((unsigned)r[7] << 16);
    *reinterpret_cast<uint4*>(p) = o;
  }
};

// 8-bit samples, four per register (SIMD video instructions): one row of one SAO unit = two words.
// tab_pos / tab_neg: magnitudes of the positive / negative offsets by table index (bytes 0..4; index 2 of the edge table and
// index 4 of the band table are zero), looked up for four samples at once with PRMT.
struct SaoTab8 { unsigned pos_lo, pos_hi, neg_lo, neg_hi; };
__device__ __forceinline__ unsigned sao_apply4(unsigned v, unsigned idx, const SaoTab8& t) {
  // idx: one table index (0..4) per byte -> PRMT selector nibbles
  const unsigned u = idx | (idx >> 4), sel = __byte_perm(u, 0, 0x4420);
  const unsigned pos = __byte_perm(t.pos_lo, t.pos_hi, sel), neg = __byte_perm(t.neg_lo, t.neg_hi, sel);
  return __vsubus4(__vaddus4(v, pos), neg);                     // clip to [0, 255]
}
__device__ __forceinline__ unsigned sao_edge_idx4(unsigned v, unsigned a, unsigned bb) {
  // edgeIdx = 2 + sign(v - a) + sign(v - b) per byte (8.7.3.2); the comparison masks are 0xff = -1
  unsigned ei = __vsub4(0x02020202u, __vcmpgtu4(v, a));
  ei = __vadd4(ei, __vcmpltu4(v, a));
  ei = __vsub4(ei, __vcmpgtu4(v, bb));
  return __vadd4(ei, __vcmpltu4(v, bb));
}

constexpr int SAO_ROWS = 4;
template <typename T>
__global__ void __launch_bounds__(256, 3) sao_rows_kernel(const DeviceBatch b, int c_first, int ncomp) {
  // grid.z = picture x component (c_first .. c_first + ncomp - 1): luma and the two chroma planes are launched apart, each with a
  // grid of its own size (one grid sized for luma left half of all blocks empty: 0.2 ms on the bench grid)
  const int pi = blockIdx.z / ncomp, c = c_first + blockIdx.z % ncomp;
  const PicDesc& pic = b.pics[pi];
  if (c > 0 && !pic.chroma) return;                                                     // (uniform per block)
  const int sh = (c && (pic.chroma == 1 || pic.chroma == 2)) ? 1 : 0, shy = (c && pic.chroma == 1) ? 1 : 0;
  const int w = pic.width >> sh, h = pic.height >> shy;
  const int lane = threadIdx.x;                                                           // blockDim.x == 32: one warp per row group
  const int x0 = (blockIdx.x * 32 + lane) * 8, yb = (blockIdx.y * blockDim.y + threadIdx.y) * SAO_ROWS;
  if (blockIdx.x * 256 >= w || yb >= h) return;                                          // (uniform per warp: every lane stays for the shuffles)
  const bool in = x0 < w;
  const int n = in ? min(8, w - x0) : 0;
  const int cx = pic.crop_x >> sh, cy = pic.crop_y >> shy, ow = (pic.out_w + sh) >> sh, oh = (pic.out_h + shy) >> shy;
  const T* src = static_cast<const T*>(pic.rec[c]);
  const int st = pic.rec_stride[c];
  const int lg = pic.log2_ctb - sh, lgy = pic.log2_ctb - shy;
  const CtuInfo* ctus = b.ctus + pic.ctu_base;
  const int bd = pic.bit_depth, maxv = (1 << bd) - 1;
  T* const dplane = static_cast<T*>(pic.dst[c]);
  const int dst_st = pic.dst_stride[c];
  const int ox0 = x0 - cx;
  if (pic.nslices > 1) {
    // several slices: every sample through the rule-complete path (slice-boundary conditions of 8.7.3.2)
    if (!in) return;
    for (int r = 0; r < SAO_ROWS; r++) {
      const int y = yb + r, oy = y - cy;
      if (y >= h || oy < 0 || oy >= oh) continue;
      const uint8_t* cell = b.edge8 + pic.map8_base + ((y << shy) >> 3) * pic.w8 + ((x0 << sh) >> 3);
      const bool k0 = cell[0] & 4, k1 = sh ? (((x0 + 4) << 1) < pic.width && (cell[1] & 4)) : k0;
      T* drow = dplane + (size_t)oy * dst_st;
      for (int k = 0; k < n; k++) {
        const int v = (k < 4 ? k0 : k1) ? (int)src[(size_t)y * st + x0 + k] : sao_sample_generic<T>(b, pic, ctus, src, st, c, x0 + k, y, w, h, lg, lgy);
        const int ox = ox0 + k;
        if (ox >= 0 && ox < ow) drow[ox] = (T)v;
      }
    }
    return;
  }
  // ---- the six source rows (yb - 1 .. yb + 4), all requested before anything is used
  SaoPack<T> row[SAO_ROWS + 2];
  int hl[SAO_ROWS + 2], hr[SAO_ROWS + 2];                                                // outer neighbours (x0 - 1, x0 + 8) of every row
  const bool has_l = in && x0 > 0, has_r = in && x0 + 8 < w;
#pragma unroll
  for (int j = 0; j < SAO_ROWS + 2; j++) {
    const int y = yb - 1 + j;
    row[j].zero(); hl[j] = 0; hr[j] = 0;
    if (in && y >= 0 && y < h) {
      const T* p = src + (size_t)y * st + x0;
      if (n == 8) row[j].load_vec(p); else row[j].load_n(p, n);
      if (lane == 0 && has_l) hl[j] = (int)p[-1];
      if (lane == 31 && has_r) hr[j] = (int)p[8];
    }
  }
  SaoComp sp{}; bool keep0 = false, keep1 = false;
  if (in) {
    sp = ctus[(yb >> lgy) * pic.wctb + (x0 >> lg)].sao[c];
    // cu_transquant_bypass / pcm + pcm_loop_filter_disabled (8.7.3: SaoTypeIdx is treated as 0 there): bit 2 of the 8x8 luma cells
    // (the four rows of the group lie in one row of cells: 4 rows of luma, or 4 chroma rows = 8 luma rows, 4-aligned)
    const uint8_t* cell = b.edge8 + pic.map8_base + ((yb << shy) >> 3) * pic.w8 + ((x0 << sh) >> 3);
    keep0 = cell[0] & 4; keep1 = sh ? (((x0 + 4) << 1) < pic.width && (cell[1] & 4)) : keep0;
  }
#pragma unroll
  for (int j = 0; j < SAO_ROWS + 2; j++) {
    const int fl = __shfl_up_sync(0xffffffffu, row[j].last(), 1), fr = __shfl_down_sync(0xffffffffu, row[j].first(), 1);
    if (lane != 0) hl[j] = fl;
    if (lane != 31) hr[j] = fr;
  }
  if (!in) return;
  const unsigned offs = (unsigned)(uint8_t)sp.offset[0] | ((unsigned)(uint8_t)sp.offset[1] << 8) | ((unsigned)(uint8_t)sp.offset[2] << 16) | ((unsigned)(uint8_t)sp.offset[3] << 24);
  auto offset = [&](int i) { return (int)(int8_t)(offs >> (8 * i)); };                   // register-resident SaoOffsetVal[1..4]
  const int type = (pic.sao_enabled && sp.type) ? (int)sp.type : 0;
  const int e = sp.band_or_class;
  const int hx = e == 1 ? 0 : (e == 3 ? 1 : -1);
  const bool vert = e != 0;
  // packed tables of the 8-bit SIMD path: |offset| of the positive and of the negative offsets by table index
  SaoTab8 edge_tab{0, 0, 0, 0}, band_tab{0, 0, 0, 0};
  if constexpr (sizeof(T) == 1) {
    unsigned pos = 0, neg = 0;                                                          // byte i: max(offset(i), 0) / max(-offset(i), 0)
#pragma unroll
    for (int i = 0; i < 4; i++) { const int o = offset(i); pos |= (unsigned)(o > 0 ? o : 0) << (8 * i); neg |= (unsigned)(o < 0 ? -o : 0) << (8 * i); }
    band_tab.pos_lo = pos; band_tab.neg_lo = neg;                                        // [o0 o1 o2 o3 | 0]
    edge_tab.pos_lo = (pos & 0xffffu) | ((pos & 0xff0000u) << 8); edge_tab.pos_hi = pos >> 24;   // [o0 o1 0 o2 | o3]
    edge_tab.neg_lo = (neg & 0xffffu) | ((neg & 0xff0000u) << 8); edge_tab.neg_hi = neg >> 24;
  }
#pragma unroll
  for (int r = 0; r < SAO_ROWS; r++) {
    const int y = yb + r, oy = y - cy;
    if (y >= h || oy < 0 || oy >= oh) continue;                                           // outside the picture / the conformance window: never output
    if constexpr (sizeof(T) == 1) {
      {
        // 8-bit samples: four per instruction (a partial unit at the right picture edge is zero-filled beyond its n samples)
        const uint2 cv = row[r + 1].q;
        uint2 out = cv;
        if (type == 1) {
          const unsigned bnd = (unsigned)e * 0x01010101u;
          unsigned k0 = __vsub4((cv.x >> 3) & 0x1f1f1f1fu, bnd) & 0x1f1f1f1fu, k1 = __vsub4((cv.y >> 3) & 0x1f1f1f1fu, bnd) & 0x1f1f1f1fu;
          const unsigned m0 = __vcmpltu4(k0, 0x04040404u), m1 = __vcmpltu4(k1, 0x04040404u);
          k0 = (k0 & m0) | (~m0 & 0x04040404u); k1 = (k1 & m1) | (~m1 & 0x04040404u);
          out.x = sao_apply4(cv.x, k0, band_tab); out.y = sao_apply4(cv.y, k1, band_tab);
        } else if (type == 2 && (!vert || (y > 0 && y + 1 < h))) {
          // neighbour vectors: L = the sample to the left at every position, R = the one to the right
          const uint2 uv = row[r].q, dv = row[r + 2].q;
          uint2 a, bb;
          if (e == 0) {
            a.x = (cv.x << 8) | (unsigned)hl[r + 1]; a.y = __funnelshift_l(cv.x, cv.y, 8);
            bb.x = __funnelshift_r(cv.x, cv.y, 8); bb.y = (cv.y >> 8) | ((unsigned)hr[r + 1] << 24);
          } else if (e == 1) { a = uv; bb = dv; }
          else if (e == 2) {
            a.x = (uv.x << 8) | (unsigned)hl[r]; a.y = __funnelshift_l(uv.x, uv.y, 8);
            bb.x = __funnelshift_r(dv.x, dv.y, 8); bb.y = (dv.y >> 8) | ((unsigned)hr[r + 2] << 24);
          } else {
            a.x = __funnelshift_r(uv.x, uv.y, 8); a.y = (uv.y >> 8) | ((unsigned)hr[r] << 24);
            bb.x = (dv.x << 8) | (unsigned)hl[r + 2]; bb.y = __funnelshift_l(dv.x, dv.y, 8);
          }
          unsigned e0 = sao_edge_idx4(cv.x, a.x, bb.x), e1 = sao_edge_idx4(cv.y, a.y, bb.y);
          if (e != 1) {                                                                   // first / last column of the picture: a neighbour is missing
            if (!has_l) e0 = (e0 & 0xffffff00u) | 0x02u;
            if (!has_r) {                                                                 // the unit's last sample (n - 1) is the picture's last column
              const unsigned sft = 8u * (unsigned)((n - 1) & 3), clr = ~(0xffu << sft), two = 0x02u << sft;
              if (n > 4) e1 = (e1 & clr) | two; else e0 = (e0 & clr) | two;
            }
          }
          out.x = sao_apply4(cv.x, e0, edge_tab); out.y = sao_apply4(cv.y, e1, edge_tab);
        }
        if (keep0) out.x = cv.x;
        if (keep1) out.y = cv.y;
        T* drow8 = dplane + (size_t)oy * dst_st;
        if (n == 8 && ox0 >= 0 && ox0 + 8 <= ow && ((reinterpret_cast<uintptr_t>(drow8 + ox0) & 7) == 0)) *reinterpret_cast<uint2*>(drow8 + ox0) = out;
        else {
#pragma unroll
          for (int k = 0; k < 8; k++) { const int ox = ox0 + k; if (k < n && ox >= 0 && ox < ow) drow8[ox] = (T)(((k < 4 ? out.x : out.y) >> (8 * (k & 3))) & 0xffu); }
        }
      }
    } else {
    int cur[10], res[8];
    row[r + 1].unpack(cur + 1); cur[0] = hl[r + 1]; cur[9] = hr[r + 1];
    if (type == 0) {
#pragma unroll
      for (int k = 0; k < 8; k++) res[k] = cur[1 + k];
    } else if (type == 1) {
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int v = cur[1 + k], kk = ((v >> (bd - 5)) - e) & 31;
        res[k] = clip3d(0, maxv, v + (kk < 4 ? offset(kk) : 0));
      }
    } else {
      int up[10], dn[10];
      if (vert) { row[r].unpack(up + 1); up[0] = hl[r]; up[9] = hr[r]; row[r + 2].unpack(dn + 1); dn[0] = hl[r + 2]; dn[9] = hr[r + 2]; }
      else {
#pragma unroll
        for (int k = 0; k < 10; k++) up[k] = dn[k] = 0;
      }
      const bool rows_ok = !vert || (y > 0 && y + 1 < h);
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int v = cur[1 + k];
        int off = 0;
        const int xa = x0 + k + hx, xb = x0 + k - hx;
        if (rows_ok && xa >= 0 && xb >= 0 && xa < w && xb < w) {
          const int a_l = vert ? up[k] : cur[k], a_m = up[1 + k], a_r = vert ? up[2 + k] : cur[2 + k];
          const int b_l = vert ? dn[k] : cur[k], b_m = dn[1 + k], b_r = vert ? dn[2 + k] : cur[2 + k];
          const int a = hx < 0 ? a_l : (hx > 0 ? a_r : a_m), bb = hx < 0 ? b_r : (hx > 0 ? b_l : b_m);
          const int ei = 2 + (v > a) - (v < a) + (v > bb) - (v < bb);
          if (ei != 2) off = offset(ei < 2 ? ei : ei - 1);              // edgeIdx 0,1,3,4 -> SaoOffsetVal[1..4]
        }
        res[k] = clip3d(0, maxv, v + off);
      }
    }
    if (keep0 | keep1) {
#pragma unroll
      for (int k = 0; k < 8; k++) if (k < 4 ? keep0 : keep1) res[k] = cur[1 + k];
    }
    // store (conformance window applied; destination already offset to the tile's paste position)
    T* drow = dplane + (size_t)oy * dst_st;
    if (n == 8 && ox0 >= 0 && ox0 + 8 <= ow && ((reinterpret_cast<uintptr_t>(drow + ox0) & (8 * sizeof(T) - 1)) == 0)) SaoPack<T>::store_vec(drow + ox0, res);
    else {
#pragma unroll
      for (int k = 0; k < 8; k++) { const int ox = ox0 + k; if (k < n && ox >= 0 && ox < ow) drow[ox] = (T)res[k]; }
    }
    }   // (samples wider than 8 bits: scalar arithmetic)
  }
}

int launch_deblock(const DeviceBatch& b, const PicDesc* hp, cudaStream_t s) {
  int max_w8 = 0, max_h8 = 0, max_w = 0, max_h = 0; bool any16 = false, any8 = false;
  for (int i = 0; i < b.npics; i++) {
    max_w8 = max(max_w8, hp[i].w8); max_h8 = max(max_h8, hp[i].h8); max_w = max(max_w, hp[i].width); max_h = max(max_h, hp[i].height);
    if (hp[i].bit_depth > 8) any16 = true; else any8 = true;
  }
  if (any16 && any8) return set_error(B200_E_UNSUPPORTED, "a batch must not mix 8-bit and >8-bit pictures");
  if (!b.npics) return B200_OK;
  {
    const int m = max_w8, total = m * (max_h >> 2);
    dim3 grid((total + 255) / 256, b.npics);
    if (any16) deblock_kernel<uint16_t, 1><<<grid, 256, 0, s>>>(b, m, max_h >> 2); else deblock_kernel<uint8_t, 1><<<grid, 256, 0, s>>>(b, m, max_h >> 2);
  }
  {
    const int total = max_h8 * (max_w >> 2);
    dim3 grid((total + 255) / 256, b.npics);
    if (any16) deblock_kernel<uint16_t, 0><<<grid, 256, 0, s>>>(b, max_w8, max_w >> 2); else deblock_kernel<uint8_t, 0><<<grid, 256, 0, s>>>(b, max_w8, max_w >> 2);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(B200_E_CUDA, "deblock launch: %s", cudaGetErrorString(e));
  return B200_OK;
}

int launch_sao(const DeviceBatch& b, const PicDesc* hp, cudaStream_t s, int* launches) {
  int max_w = 0, max_h = 0, max_cw = 0, max_ch = 0; bool any16 = false;
  for (int i = 0; i < b.npics; i++) {
    max_w = max(max_w, hp[i].width); max_h = max(max_h, hp[i].height); if (hp[i].bit_depth > 8) any16 = true;
    if (hp[i].chroma) { max_cw = max(max_cw, hp[i].width >> (hp[i].chroma != 3 ? 1 : 0)); max_ch = max(max_ch, hp[i].height >> (hp[i].chroma == 1 ? 1 : 0)); }
  }
  if (launches) *launches = 0;
  if (!b.npics) return B200_OK;
  static const bool first_form = getenv("B200_SAO_ROW_PER_THREAD") != nullptr;            // the first form of the kernel (one row per thread), kept for A / B measurements
  if (first_form) {
    const dim3 block(64, 4), grid((max_w / 8 + 63) / 64, (max_h + 3) / 4, b.npics * 3);
    if (any16) sao_kernel<uint16_t><<<grid, block, 0, s>>>(b); else sao_kernel<uint8_t><<<grid, block, 0, s>>>(b);
    if (launches) *launches = 1;
  } else {
    const dim3 block(32, 8);
    const dim3 grid_y((max_w + 255) / 256, (max_h + 8 * SAO_ROWS - 1) / (8 * SAO_ROWS), b.npics);
    if (any16) sao_rows_kernel<uint16_t><<<grid_y, block, 0, s>>>(b, 0, 1); else sao_rows_kernel<uint8_t><<<grid_y, block, 0, s>>>(b, 0, 1);
    if (launches) *launches = 1;
    if (max_cw > 0 && max_ch > 0) {
      const dim3 grid_c((max_cw + 255) / 256, (max_ch + 8 * SAO_ROWS - 1) / (8 * SAO_ROWS), b.npics * 2);
      if (any16) sao_rows_kernel<uint16_t><<<grid_c, block, 0, s>>>(b, 1, 2); else sao_rows_kernel<uint8_t><<<grid_c, block, 0, s>>>(b, 1, 2);
      if (launches) *launches = 2;
    }
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(B200_E_CUDA, "sao launch: %s", cudaGetErrorString(e));
  return B200_OK;
}

}  // namespace b200
