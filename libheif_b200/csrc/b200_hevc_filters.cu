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
          ql[-1 * (ptrdiff_t)xs] = (T)clip3d(p0 - 2 * tc, p0 + 2 * tc, (p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
          ql[-2 * (ptrdiff_t)xs] = (T)clip3d(p1 - 2 * tc, p1 + 2 * tc, (p2 + p1 + p0 + q0 + 2) >> 2);
          ql[-3 * (ptrdiff_t)xs] = (T)clip3d(p2 - 2 * tc, p2 + 2 * tc, (2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
          ql[0] = (T)clip3d(q0 - 2 * tc, q0 + 2 * tc, (p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
          ql[xs] = (T)clip3d(q1 - 2 * tc, q1 + 2 * tc, (p0 + q0 + q1 + q2 + 2) >> 2);
          ql[2 * (ptrdiff_t)xs] = (T)clip3d(q2 - 2 * tc, q2 + 2 * tc, (p0 + q0 + q1 + 3 * q2 + 2 * q3 + 4) >> 3);
        } else {
          int delta = (9 * (q0 - p0) - 3 * (q1 - p1) + 8) >> 4;
          if (abs(delta) < tc * 10) {
            delta = clip3d(-tc, tc, delta);
            ql[-1 * (ptrdiff_t)xs] = (T)clip3d(0, maxv, p0 + delta);
            ql[0] = (T)clip3d(0, maxv, q0 - delta);
            if (dep) ql[-2 * (ptrdiff_t)xs] = (T)clip3d(0, maxv, p1 + clip3d(-(tc >> 1), tc >> 1, (((p2 + p0 + 1) >> 1) - p1 + delta) >> 1));
            if (deq) ql[xs] = (T)clip3d(0, maxv, q1 + clip3d(-(tc >> 1), tc >> 1, (((q2 + q0 + 1) >> 1) - q1 - delta) >> 1));
          }
        }
      }
    }
#undef P_
#undef Q_
  }
  // chroma (4:2:0): edges on the 8-sample chroma grid = 16-sample luma grid, bS == 2 only (8.7.2.5.5 / 8.7.2.5.8)
  if (pic.chroma && ((VERT ? x : y) & 15) == 0) {
    for (int c = 1; c <= 2; c++) {
      const int qpi = qpl + (c == 1 ? pic.pps_cb_qp_offset : pic.pps_cr_qp_offset);
      const int qpc = qpi < 30 ? qpi : (qpi >= 43 ? qpi - 6 : c_qpc2[qpi - 30]);
      const int tc = c_tc[clip3d(0, 53, qpc + 2 + sl.tc_offset)] * (1 << (bd - 8));
      T* pl = static_cast<T*>(pic.rec[c]);
      const int st = pic.rec_stride[c];
      const int xs = VERT ? 1 : st, ls = VERT ? st : 1;
      T* q = pl + (size_t)(y >> 1) * st + (x >> 1);
#pragma unroll
      for (int l = 0; l < 2; l++) {
        T* ql = q + (ptrdiff_t)l * ls;
        const int p0 = ql[-(ptrdiff_t)xs], p1 = ql[-2 * (ptrdiff_t)xs], q0 = ql[0], q1 = ql[xs];
        const int delta = clip3d(-tc, tc, ((((q0 - p0) << 2) + p1 - q1 + 4) >> 3));
        ql[-(ptrdiff_t)xs] = (T)clip3d(0, maxv, p0 + delta);
        ql[0] = (T)clip3d(0, maxv, q0 - delta);
      }
    }
  }
}

// one thread = one sample of one component; grid.z = picture * 3 + component
template <typename T>
__global__ void __launch_bounds__(256) sao_kernel(const DeviceBatch b, int ncomp_max) {
  const int pi = blockIdx.z / 3, c = blockIdx.z % 3;
  const PicDesc& pic = b.pics[pi];
  if (c > 0 && !pic.chroma) return;
  const int sh = c ? 1 : 0;
  const int w = pic.width >> sh, h = pic.height >> sh;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w || y >= h) return;
  const int cx = pic.crop_x >> sh, cy = pic.crop_y >> sh, ow = (pic.out_w + sh) >> sh, oh = (pic.out_h + sh) >> sh;
  const int ox = x - cx, oy = y - cy;
  if (ox < 0 || oy < 0 || ox >= ow || oy >= oh) return;              // outside the conformance window: never output
  const T* src = static_cast<const T*>(pic.rec[c]);
  const int st = pic.rec_stride[c];
  int v = src[(size_t)y * st + x];
  const int lg = pic.log2_ctb - sh;
  const CtuInfo* ctus = b.ctus + pic.ctu_base;
  const CtuInfo& ci = ctus[(y >> lg) * pic.wctb + (x >> lg)];
  const SaoComp sp = ci.sao[c];
  if (pic.sao_enabled && sp.type) {
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
        const int sa = ctus[(ya >> lg) * pic.wctb + (xa >> lg)].slice_idx, sb = ctus[(yb >> lg) * pic.wctb + (xb >> lg)].slice_idx;
        if (sa != cur && !(sa < cur ? sls[cur].lf_across_slices : sls[sa].lf_across_slices)) skip = true;
        if (sb != cur && !(sb < cur ? sls[cur].lf_across_slices : sls[sb].lf_across_slices)) skip = true;
        if (!skip) {
          const int a = src[(size_t)ya * st + xa], bb = src[(size_t)yb * st + xb];
          const int ei = 2 + (v > a) - (v < a) + (v > bb) - (v < bb);
          if (ei != 2) off = sp.offset[ei < 2 ? ei : ei - 1];           // edgeIdx 0,1,3,4 -> SaoOffsetVal[1..4]
        }
      }
    }
    v = clip3d(0, maxv, v + off);
  }
  static_cast<T*>(pic.dst[c])[(size_t)oy * pic.dst_stride[c] + ox] = (T)v;
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
  dim3 grid((max_w + 255) / 256, max_h, b.npics * 3);
  if (any16) sao_kernel<uint16_t><<<grid, 256, 0, s>>>(b, 3); else sao_kernel<uint8_t><<<grid, 256, 0, s>>>(b, 3);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(B200_E_CUDA, "sao launch: %s", cudaGetErrorString(e));
  return B200_OK;
}

}  // namespace b200
