// b200_hevc_filters.cu -- K3 deblocking (H.265 8.7.2) and K4 SAO (8.7.3) + K5 conformance crop / tile paste.
//
// The in-loop filters libde265 applies before libheif/plugins/decoder_libde265.cc:97-171 copies the planes out.
// Intra pictures: bS = 2 on every transform edge of the 8x8 grid (host supplies the filterEdgeFlag map and the
// QpY map, one byte per 8x8 block each).  Deblocking runs in place on the reconstruction planes: one pass over all
// vertical edges of all pictures of the batch, then one over all horizontal edges (8.7.2 orders them picture-wide;
// edges of one direction are independent because they are 8 samples apart and touch at most 3 samples per side).
// SAO reads the deblocked planes and writes the final samples straight into the destination (conformance window
// applied, destination pointer already offset to the tile's paste position: K5 folded into K4's store).
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

int launch_sao(const DeviceBatch& b, const PicDesc* hp, cudaStream_t s) {
  int max_w = 0, max_h = 0; bool any16 = false;
  for (int i = 0; i < b.npics; i++) { max_w = max(max_w, hp[i].width); max_h = max(max_h, hp[i].height); if (hp[i].bit_depth > 8) any16 = true; }
  if (!b.npics) return B200_OK;
  const dim3 block(64, 4), grid((max_w / 8 + 63) / 64, (max_h + 3) / 4, b.npics * 3);
  if (any16) sao_kernel<uint16_t><<<grid, block, 0, s>>>(b); else sao_kernel<uint8_t><<<grid, block, 0, s>>>(b);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(B200_E_CUDA, "sao launch: %s", cudaGetErrorString(e));
  return B200_OK;
}

}  // namespace b200
