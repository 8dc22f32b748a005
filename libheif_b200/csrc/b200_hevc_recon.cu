// b200_hevc_recon.cu -- K1 + K2: scaling, inverse DCT/DST, intra prediction and reconstruction on sm_100a.
//
// Does the per-sample half of what libde265 does inside de265_decode() (reached from
// libheif/plugins/decoder_libde265.cc:386-457): H.265 8.6.2-8.6.4 (scaling, transforms), 8.4.4.2 (intra sample
// prediction incl. reference substitution and smoothing), 8.6.6 (reconstruction), driven by the command stream
// of the entropy stage (b200_hevc_types.h).
//
// Work item = (picture, CTB row, component group): luma, or the Cb + Cr pair.  Chroma prediction never reads luma, so the
// two groups are independent wavefronts; Cb and Cr share position, size, mode and availability, so one warp does both at
// once on its two half-warps.  ONE WARP walks the CTBs of its row left to right; rows advance as a wavefront with a lag
// of two CTBs (above-right dependency) through per-(row, group) progress counters.  A global ticket hands items out in
// "row k of every picture" order, so a dependency always holds a smaller ticket (no co-residency requirement).
//
// Per CTB the warp works in two phases:
//   A. residual phase, lane-parallel over the CTB's transform units (no dependency between them): every lane decodes
//      one TuCmd, derives the neighbour availability of its block (6.4.1, z-scan order) into a packed descriptor, and
//      -- 4x4 blocks being 80 % of all blocks of a typical intra picture -- dequantises and inverse-transforms its own
//      4x4 block entirely in registers (32 blocks per pass).  Larger blocks (8..32) are done by the whole warp, one at a
//      time, with zero-row/column skipping.  Residuals land in shared memory, block-contiguous in z-order.
//   B. prediction phase, block after block (the intra dependency chain): one gather of the 4n+1 neighbours with the
//      substitution process (8.4.4.2.2) folded into the index computation, optional smoothing, prediction, residual
//      add, store into the CTB tile in shared memory.  ~100 warp instructions per block instead of ~900.
// The finished CTB leaves shared memory with one cp.async.bulk (TMA) row copy per lane; the halo (row above incl.
// above-right, column to the left) is kept in shared memory next to the tile.  Integer work: no tensor cores.
#include "b200_hevc.h"

namespace b200 {

#ifndef B200_RECON_MIN_BLOCKS
#define B200_RECON_MIN_BLOCKS 5
#endif
constexpr int WARPS = 4;                       // warps (= work items in flight) per CTA
constexpr int PAD = 16;                        // samples left of the tile / halo row: column PAD - 1 is the left halo
// Transposed DCT matrices MT_n[y][k] = transMatrix_n[k][y] for n = 32, 16, 8 (rows padded by 4 bytes: lanes that read different
// rows hit different banks), shared by the CTA: 32 x 36 + 16 x 20 + 8 x 12 bytes
constexpr int MT32_OFF = 0, MT16_OFF = 32 * 36, MT8_OFF = MT16_OFF + 16 * 20;
constexpr int MAT_BYTES = MT8_OFF + 8 * 12;
constexpr int REF_STRIDE = 136;                // int16 entries per neighbour array (4 * 32 + 1 rounded up)

struct WarpLayout { int tile, top, res, tmp, desc, total; };     // byte offsets inside the warp's shared-memory slice
__host__ __device__ inline WarpLayout warp_layout(int log2ctb, int bps) {
  const int ctb = 1 << log2ctb, tb = ctb < 32 ? ctb : 32;
  WarpLayout L; int o = 0;
  L.tile = o; o += ctb * (ctb + PAD) * bps;                      // luma tile; the Cb + Cr tiles of a chroma item fit inside
  L.top = o; o += (2 * PAD + 2 * ctb + 32) * bps;                // halo row(s)
  o = (o + 15) & ~15;
  L.res = o; o += ctb * ctb * 2;                                 // residuals, int16, z-order block-contiguous
  L.tmp = o; o += (tb * (tb + 2) * 2 > 1024 ? tb * (tb + 2) * 2 : 1024);   // first-stage output (rows padded by 2) / 4x4 scratch / neighbour arrays
  L.desc = o; o += (ctb / 4) * (ctb / 4) * 8;                    // one descriptor per transform block of the CTB
  L.total = (o + 15) & ~15;
  return L;
}

__constant__ int8_t c_dct[32] = {64, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67,
                                 64, 61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9, 4};
__constant__ int8_t c_angle[35] = {0, 0, 32, 26, 21, 17, 13, 9, 5, 2, 0, -2, -5, -9, -13, -17, -21, -26, -32,
                                   -26, -21, -17, -13, -9, -5, -2, 0, 2, 5, 9, 13, 17, 21, 26, 32};
__constant__ int16_t c_inv_angle[35] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, -4096, -1638, -910, -630, -482, -390, -315, -256,
                                        -315, -390, -482, -630, -910, -1638, -4096, 0, 0, 0, 0, 0, 0, 0, 0, 0};
__constant__ uint8_t c_qpc[14] = {29, 30, 31, 32, 33, 33, 34, 34, 35, 35, 36, 36, 37, 37};
__constant__ uint8_t c_level_scale[6] = {40, 45, 51, 57, 64, 72};

__device__ __forceinline__ int clip3i(int lo, int hi, int v) { return min(max(v, lo), hi); }
__device__ __forceinline__ unsigned morton4(unsigned x, unsigned y) {   // z-order index of a 4x4 block inside a CTB
  unsigned sx = (x & 1) | ((x & 2) << 1) | ((x & 4) << 2) | ((x & 8) << 3);
  unsigned sy = (y & 1) | ((y & 2) << 1) | ((y & 4) << 2) | ((y & 8) << 3);
  return sx | (sy << 1);
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// TMA (bulk asynchronous copy) of one finished tile row: shared -> global.  Source and destination 16-byte aligned, size a
// multiple of 16.
__device__ __forceinline__ void bulk_store(void* g, const void* s, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"((unsigned)__cvta_generic_to_shared(s)), "r"(bytes) : "memory");
}

// Command-stream reads.  LIVE = K0 is running concurrently: a neighbouring, not yet written entry may share a cache
// line with one read earlier and L1 is not coherent, so everything goes to L2 (ld.global.cg).  Otherwise the command
// stream is complete and plain loads let consecutive entries hit the L1 line the first one brought in.
template <bool LIVE, class T> __device__ __forceinline__ T ld_cmd(const T* p) { return LIVE ? __ldcg(p) : *p; }
template <bool LIVE> __device__ __forceinline__ TuCmd ld_tu(const TuCmd* p) { const uint4 v = ld_cmd<LIVE>(reinterpret_cast<const uint4*>(p)); return TuCmd{v.x, v.y, v.z, v.w}; }
template <bool LIVE> __device__ __forceinline__ CoefEntry ld_coef(const CoefEntry* p) { const unsigned v = ld_cmd<LIVE>(reinterpret_cast<const unsigned*>(p)); CoefEntry e; e.pos = (uint16_t)(v & 0xffff); e.level = (int16_t)(v >> 16); return e; }
template <bool LIVE> __device__ __forceinline__ CtuInfo ld_ctu(const CtuInfo* p) {
  CtuInfo c; const uint2* s = reinterpret_cast<const uint2*>(p); uint2* d = reinterpret_cast<uint2*>(&c);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(CtuInfo) / 8); i++) d[i] = ld_cmd<LIVE>(s + i);
  return c;
}

__device__ __forceinline__ int chroma_qp(int qpy, int off, int bd, int cfmt = 1) {       // 8.6.1: Table 8-10 when ChromaArrayType == 1, else Min(qPi, 51)
  const int qbd = 6 * (bd - 8);
  const int qpi = clip3i(-qbd, 57, qpy + off);
  const int qpc = cfmt != 1 ? min(qpi, 51) : (qpi < 30 ? qpi : (qpi >= 43 ? qpi - 6 : c_qpc[qpi - 30]));
  return qpc + qbd;
}

// 8.6.3 / 8.6.4.2 scaling: TransCoeffLevel -> d, clipped to 16 bits; m = 16 (flat) or the scaling factor of the position
__device__ __forceinline__ int dequant(int level, int qp, int bd_shift, int m) {
  const long long scale = (long long)(c_level_scale[qp % 6] << (qp / 6)) * m;
  const long long t = ((long long)level * scale + (1LL << (bd_shift - 1))) >> bd_shift;
  return (int)(t < -32768 ? -32768 : (t > 32767 ? 32767 : t));
}

// Descriptor of one transform block, built in phase A, consumed in phase B (uint2):
//  x: bx/4 [0:4) by/4 [4:8) log2n-2 [8:10) mode [10:16) coded [16] (Cr: [17]) availL [18] availCorner [19] availTop [20]
//     available below-left samples / 4 [21:25)  available above-right samples / 4 [25:29)  pcm [29]: the "residuals" are the samples
//  y: offset of the block's residuals inside the component's residual area (samples)
// (tw, th: the component's CTB size; shx: its horizontal sub-sampling when the decoding order inside the CTB has to be judged in
//  LUMA units -- the 4:2:2 chroma planes, whose blocks follow the z-order of the luma quadtree, not of their own coordinates)
__device__ __forceinline__ unsigned make_desc(int bx, int by, int lg, int mode, int coded0, int coded1, int tw, int th, int shx, int cx0, int cy0, int cw, int ch,
                                              bool nbL, bool nbAL, bool nbA, bool nbAR) {
  const int n = 1 << lg;
  const bool fL = bx > 0 || nbL, fT = by > 0 || nbA;
  const bool fC = (bx > 0 && by > 0) ? true : (bx > 0 ? nbA : (by > 0 ? nbL : nbAL));
  const unsigned me = morton4((unsigned)(bx << shx) >> 2, (unsigned)by >> 2);
  bool tr = false, bl = false;
  if (cx0 + bx + n < cw) {
    if (by > 0) { if (bx + n < tw) tr = morton4((unsigned)((bx + n) << shx) >> 2, (unsigned)(by - 1) >> 2) < me; }
    else tr = (bx + n < tw) ? nbA : nbAR;
  }
  if (cy0 + by + n < ch && by + n < th) {
    if (bx > 0) bl = morton4((unsigned)((bx - 1) << shx) >> 2, (unsigned)(by + n) >> 2) < me; else bl = nbL;
  }
  const int trc = tr ? min(n, cw - (cx0 + bx + n)) : 0, blc = bl ? min(n, ch - (cy0 + by + n)) : 0;
  return (unsigned)(bx >> 2) | ((unsigned)(by >> 2) << 4) | ((unsigned)(lg - 2) << 8) | ((unsigned)mode << 10) | ((unsigned)coded0 << 16) | ((unsigned)coded1 << 17) |
         ((unsigned)fL << 18) | ((unsigned)fC << 19) | ((unsigned)fT << 20) | ((unsigned)(blc >> 2) << 21) | ((unsigned)(trc >> 2) << 25);
}

// ---- phase A, 4x4 blocks: the calling lane owns the block.  scr: the warp's [16][32] int16 scratch (column = lane).
template <bool LIVE>
__device__ __forceinline__ void residual4_lane(int16_t* scr, int lane, const CoefEntry* __restrict__ ce, int nnz, int qp, int bd, bool dst, bool tskip, bool raw, int16_t* out,
                                               const uint8_t* __restrict__ sf) {      // raw: cu_transquant_bypass / pcm, the levels ARE the residuals (8.6.2)      // sf: the 16 scaling factors of this component (raster), or nullptr
  int16_t* my = scr + lane;
#pragma unroll
  for (int p = 0; p < 16; p++) my[p * 32] = 0;
  const int bd_shift = bd - 3;                        // bd + log2(4) - 5
#pragma unroll 1
  for (int i = 0; i < nnz; i++) {
    const CoefEntry e = ld_coef<LIVE>(&ce[i]);
    my[(e.pos & 15) * 32] = raw ? e.level : (int16_t)dequant(e.level, qp, bd_shift, sf ? (int)__ldg(sf + (e.pos & 15)) : 16);
  }
  int c[16];
#pragma unroll
  for (int p = 0; p < 16; p++) c[p] = my[p * 32];
  const int bs2 = 20 - bd, rnd = 1 << (bs2 - 1);
  unsigned* o32 = reinterpret_cast<unsigned*>(out);
  if (raw) {
#pragma unroll
    for (int p = 0; p < 16; p += 2) o32[p >> 1] = (unsigned)(c[p] & 0xffff) | ((unsigned)c[p + 1] << 16);
    return;
  }
  if (tskip) {                                        // 8.6.4.2, transform_skip_flag: r = d << 7
#pragma unroll
    for (int p = 0; p < 16; p += 2) {
      const int r0 = ((c[p] << 7) + rnd) >> bs2, r1 = ((c[p + 1] << 7) + rnd) >> bs2;
      o32[p >> 1] = (unsigned)(r0 & 0xffff) | ((unsigned)r1 << 16);
    }
    return;
  }
  // coefficient c[k * 4 + x] (k: vertical frequency).  First stage (columns): t[x][y] = clip16((sum_k c[k][x] * M[k][y] + 64) >> 7)
  int t[16];
#pragma unroll
  for (int x = 0; x < 4; x++) {
    const int c0 = c[x], c1 = c[4 + x], c2 = c[8 + x], c3 = c[12 + x];
    int e0, e1, e2, e3;
    if (dst) {                                        // DST-VII (8.6.4.2): M = {29 55 74 84; 74 74 0 -74; 84 -29 -74 55; 55 -84 74 -29}
      e0 = 29 * c0 + 74 * c1 + 84 * c2 + 55 * c3; e1 = 55 * c0 + 74 * c1 - 29 * c2 - 84 * c3;
      e2 = 74 * c0 - 74 * c2 + 74 * c3;           e3 = 84 * c0 - 74 * c1 + 55 * c2 - 29 * c3;
    } else {                                          // DCT-II: M = {64 64 64 64; 83 36 -36 -83; 64 -64 -64 64; 36 -83 83 -36}
      const int a = 64 * (c0 + c2), b = 64 * (c0 - c2), o0 = 83 * c1 + 36 * c3, o1 = 36 * c1 - 83 * c3;
      e0 = a + o0; e1 = b + o1; e2 = b - o1; e3 = a - o0;
    }
    t[x * 4 + 0] = clip3i(-32768, 32767, (e0 + 64) >> 7); t[x * 4 + 1] = clip3i(-32768, 32767, (e1 + 64) >> 7);
    t[x * 4 + 2] = clip3i(-32768, 32767, (e2 + 64) >> 7); t[x * 4 + 3] = clip3i(-32768, 32767, (e3 + 64) >> 7);
  }
  // second stage (rows): r[y][x] = (sum_k t[k][y] * M[k][x] + rnd) >> bs2
#pragma unroll
  for (int y = 0; y < 4; y++) {
    const int c0 = t[y], c1 = t[4 + y], c2 = t[8 + y], c3 = t[12 + y];
    int e0, e1, e2, e3;
    if (dst) {
      e0 = 29 * c0 + 74 * c1 + 84 * c2 + 55 * c3; e1 = 55 * c0 + 74 * c1 - 29 * c2 - 84 * c3;
      e2 = 74 * c0 - 74 * c2 + 74 * c3;           e3 = 84 * c0 - 74 * c1 + 55 * c2 - 29 * c3;
    } else {
      const int a = 64 * (c0 + c2), b = 64 * (c0 - c2), o0 = 83 * c1 + 36 * c3, o1 = 36 * c1 - 83 * c3;
      e0 = a + o0; e1 = b + o1; e2 = b - o1; e3 = a - o0;
    }
    const int r0 = (e0 + rnd) >> bs2, r1 = (e1 + rnd) >> bs2, r2 = (e2 + rnd) >> bs2, r3 = (e3 + rnd) >> bs2;
    o32[y * 2] = (unsigned)(r0 & 0xffff) | ((unsigned)r1 << 16);
    o32[y * 2 + 1] = (unsigned)(r2 & 0xffff) | ((unsigned)r3 << 16);
  }
}

// ---- phase A, 8x8 .. 32x32 blocks: the whole warp, in place in the block's residual slot (coefficients -> residuals).
template <bool LIVE>
__device__ __noinline__ void residual_big(int16_t* rs, int16_t* tmp, const int8_t* __restrict__ mat, const CoefEntry* __restrict__ ce, int nnz, int lg, int qp, int bd, int lane,
                                          const uint8_t* __restrict__ sf, int sf_dc, bool raw) {   // sf: 8x8 raster scaling factors of (component, size) or nullptr; sf_dc: factor of position (0, 0) for 16x16 / 32x32
  const int n = 1 << lg;
  unsigned* z = reinterpret_cast<unsigned*>(rs);
#pragma unroll 1
  for (int i = lane; i < n * n / 2; i += 32) z[i] = 0;
  __syncwarp();
  const int bd_shift = bd + lg - 5;
  int maxrow = 0, maxcol = 0;
  // coefficients are scattered TRANSPOSED (column x of the block = row x of the buffer): the first transform stage runs down the
  // columns, and with the vertical frequencies k of a column adjacent in memory two of them ride in one register (dp2a)
#pragma unroll 1
  for (int i = lane; i < nnz; i += 32) {
    const CoefEntry e = ld_coef<LIVE>(&ce[i]);
    const int pos = e.pos & (n * n - 1);
    const int x = pos & (n - 1), y = pos >> lg;
    int m = 16;
    if (sf) m = (pos == 0 && lg >= 4) ? sf_dc : (int)__ldg(sf + ((y >> (lg - 3)) << 3) + (x >> (lg - 3)));
    rs[raw ? pos : x * n + y] = raw ? e.level : (int16_t)dequant(e.level, qp, bd_shift, m);
    maxrow = max(maxrow, y); maxcol = max(maxcol, x);
  }
  maxrow = __reduce_max_sync(0xffffffffu, maxrow); maxcol = __reduce_max_sync(0xffffffffu, maxcol);
  __syncwarp();
  if (raw) return;                                    // cu_transquant_bypass / pcm (warp-uniform): no scaling, no transform
  const int8_t* mt = mat + (lg == 5 ? MT32_OFF : (lg == 4 ? MT16_OFF : MT8_OFF));
  const int ms = n + 4, P2 = n + 2;                   // row strides: matrix (bytes), intermediate (int16)
  const int kq1 = (maxrow >> 2) + 1, ncol = ((maxcol >> 2) + 1) << 2;     // groups of 4 vertical frequencies; columns, rounded up to 4 (the extra ones are zero)
  // first stage (columns): u[y][x] = clip16((sum_k coef[k][x] * M[k][y] + 64) >> 7), only columns that hold coefficients
#pragma unroll 1
  for (int i = lane; i < n * ncol; i += 32) {
    const int y = i & (n - 1), x = i >> lg;
    const int2* c = reinterpret_cast<const int2*>(rs + x * n);
    const int* mq = reinterpret_cast<const int*>(mt + y * ms);
    int e = 0;
#pragma unroll 2
    for (int q = 0; q < kq1; q++) { const int2 a = c[q]; const int bq = mq[q]; e = __dp2a_lo(a.x, bq, e); e = __dp2a_hi(a.y, bq, e); }
    tmp[y * P2 + x] = (int16_t)clip3i(-32768, 32767, (e + 64) >> 7);
  }
  __syncwarp();
  // second stage (rows): residual r[y][x] = (sum_k u[y][k] * M[k][x] + rnd) >> bs2
  const int bs2 = 20 - bd, rnd = 1 << (bs2 - 1), kq2 = ncol >> 2;
#pragma unroll 1
  for (int p = lane; p < n * n; p += 32) {
    const int x = p & (n - 1), y = p >> lg;
    const int* u = reinterpret_cast<const int*>(tmp + y * P2);
    const int* mq = reinterpret_cast<const int*>(mt + x * ms);
    int e = 0;
#pragma unroll 2
    for (int q = 0; q < kq2; q++) { const int bq = mq[q]; e = __dp2a_lo(u[2 * q], bq, e); e = __dp2a_hi(u[2 * q + 1], bq, e); }
    rs[p] = (int16_t)((e + rnd) >> bs2);
  }
  __syncwarp();
}

// ---- phase B: one transform block (of one component, or of Cb and Cr on the two half-warps).
//  tl: the lane's component tile (row stride S, sample (x, y) at tl[y * S + PAD + x]), tp: its halo row (sample x at
//  tp[PAD + x]), rs: its residual area, rf: its neighbour array(s), l / lpc: lane index inside / lanes per component,
//  gmask: the lanes working on this component.
template <typename P>
__device__ __forceinline__ void predict_tb(const uint2 d, P* tl, const P* tp, const int16_t* rs, int16_t* rf, int S, int l, int lpc, unsigned gmask, int cidx, bool luma, bool smooth, int bd, int strong_en) {   // luma: boundary filters of DC / horizontal / vertical (cIdx == 0); smooth: 8.4.4.2.3 applies (luma; chroma in 4:4:4)
  const int bx = (int)(d.x & 15) << 2, by = (int)((d.x >> 4) & 15) << 2, lg = 2 + (int)((d.x >> 8) & 3), mode = (int)((d.x >> 10) & 63);
  const int n = 1 << lg, n2 = 2 * n, n4 = 4 * n;
  const bool coded = (d.x >> (16 + cidx)) & 1, pcm = (d.x >> 29) & 1;
  const bool fL = (d.x >> 18) & 1, fC = (d.x >> 19) & 1, fT = (d.x >> 20) & 1;
  const int blc = (int)((d.x >> 21) & 15) << 2, trc = (int)((d.x >> 25) & 15) << 2;
  // ---- neighbour array rf[0 .. 4n]: index 0 = bottom of the below-left column ... 2n = corner ... 4n = end of above-right.
  // Substitution (8.4.4.2.2) = every unavailable index reads the nearest available index below it, or the first available
  // one; the available indices form up to five intervals known per block, so the source index is a handful of min / compare.
  const int first = blc ? n - blc : (fL ? n : (fC ? n2 : (fT ? n2 + 1 : (trc ? 3 * n + 1 : -1))));
  const P* colL = tl + by * S + PAD + bx - 1;                          // left column: sample y at colL[y * S]
  const P* rowA = (by > 0 ? tl + (by - 1) * S : tp) + PAD + bx;        // row above: sample x at rowA[x] (x = -1: corner)
  int dcs = 0;
  const bool small = n4 < lpc;                                         // 4x4 on a full warp: the corner rides in the same pass
  const int jn = small ? n4 + 1 : n4;
#pragma unroll 1
  for (int j = l; j < jn; j += lpc) {
    const int i = j == n4 ? n2 : (j < n2 ? j : j + 1);
    int v = 1 << (bd - 1);
    if (first >= 0) {
      int s = first;
      if (blc && i >= n - blc) s = min(i, n - 1);
      if (fL && i >= n) s = min(i, n2 - 1);
      if (fC && i >= n2) s = n2;
      if (fT && i > n2) s = min(i, 3 * n);
      if (trc && i > 3 * n) s = min(i, 3 * n + trc);
      v = s < n2 ? (int)colL[(n2 - 1 - s) * S] : (int)rowA[s - n2 - 1];
    }
    rf[i] = (int16_t)v;
    if ((i >= n && i < n2) || (i > n2 && i <= 3 * n)) dcs += v;
  }
  if (!small && l == 0) {                                               // corner (index 2n) for blocks that fill every lane
    int v = 1 << (bd - 1);
    if (first >= 0) {
      int s = first;
      if (blc) s = n - 1;
      if (fL) s = n2 - 1;
      if (fC) s = n2;
      v = s < n2 ? (int)colL[(n2 - 1 - s) * S] : (int)rowA[s - n2 - 1];
    }
    rf[n2] = (int16_t)v;
  }
  __syncwarp();
  // ---- smoothing of the neighbours (8.4.4.2.3): luma only in 4:2:0
  const int16_t* ref = rf;
  if (smooth && mode != 1 && n != 4) {
    const int dist = min(abs(mode - 26), abs(mode - 10));
    const int thr = n == 8 ? 7 : (n == 16 ? 1 : 0);
    if (dist > thr) {
      const int corner = rf[n2], bl = rf[0], tr = rf[n4];
      const bool strong = strong_en && n == 32 && abs(corner + tr - 2 * rf[3 * n]) < (1 << (bd - 5)) && abs(corner + bl - 2 * rf[n]) < (1 << (bd - 5));
      int16_t* rb = rf + REF_STRIDE;
#pragma unroll 1
      for (int i = l; i <= n4; i += lpc) {
        int v;
        if (i == 0 || i == n4) v = rf[i];
        else if (strong) {
          if (i == n2) v = corner;
          else if (i < n2) { const int y = n2 - 1 - i; v = ((63 - y) * corner + (y + 1) * bl + 32) >> 6; }
          else { const int x = i - n2 - 1; v = ((63 - x) * corner + (x + 1) * tr + 32) >> 6; }
        } else v = (rf[i - 1] + 2 * rf[i] + rf[i + 1] + 2) >> 2;
        rb[i] = (int16_t)v;
      }
      __syncwarp();
      ref = rb;
    }
  }
  // ---- prediction (8.4.4.2.4 - 8.4.4.2.6) + residual (8.6.6), written straight into the tile
  const int maxv = (1 << bd) - 1;
#define LEFT(y) ((int)ref[n2 - 1 - (y)])
#define TOP(x) ((int)ref[n2 + 1 + (x)])
  int dc = 0;
  if (mode == 1) dc = (__reduce_add_sync(gmask, dcs) + n) >> (lg + 1);
  const int ang = c_angle[mode], ia = c_inv_angle[mode];
  const bool edge = luma && n < 32;
  const int16_t* rsb = rs + d.y;
  P* out = tl + by * S + PAD + bx;
#pragma unroll 1
  for (int e = l; e < n * n; e += lpc) {
    const int x = e & (n - 1), y = e >> lg;
    int v;
    if (pcm) { out[y * S + x] = (P)rsb[e]; continue; }                   // 8.4.4.1: no prediction
    if (mode == 0) v = ((n - 1 - x) * LEFT(y) + (x + 1) * TOP(n) + (n - 1 - y) * TOP(x) + (y + 1) * LEFT(n) + n) >> (lg + 1);
    else if (mode == 1) {
      v = dc;
      if (edge) {
        if (x == 0 && y == 0) v = (LEFT(0) + 2 * dc + TOP(0) + 2) >> 2;
        else if (y == 0) v = (TOP(x) + 3 * dc + 2) >> 2;
        else if (x == 0) v = (LEFT(y) + 3 * dc + 2) >> 2;
      }
    } else if (mode >= 18) {
      const int idx = ((y + 1) * ang) >> 5, f = ((y + 1) * ang) & 31;
      const int k0 = x + idx + 1, k1 = k0 + 1;      // r[k] = p[-1 + k][-1] for k >= 0, projected left column for k < 0
      const int a = k0 >= 0 ? TOP(k0 - 1) : LEFT(-1 + ((k0 * ia + 128) >> 8));
      if (f) { const int b = k1 >= 0 ? TOP(k1 - 1) : LEFT(-1 + ((k1 * ia + 128) >> 8)); v = ((32 - f) * a + f * b + 16) >> 5; } else v = a;
      if (mode == 26 && edge && x == 0) v = clip3i(0, maxv, TOP(0) + ((LEFT(y) - LEFT(-1)) >> 1));
    } else {
      const int idx = ((x + 1) * ang) >> 5, f = ((x + 1) * ang) & 31;
      const int k0 = y + idx + 1, k1 = k0 + 1;
      const int a = k0 >= 0 ? LEFT(k0 - 1) : TOP(-1 + ((k0 * ia + 128) >> 8));
      if (f) { const int b = k1 >= 0 ? LEFT(k1 - 1) : TOP(-1 + ((k1 * ia + 128) >> 8)); v = ((32 - f) * a + f * b + 16) >> 5; } else v = a;
      if (mode == 10 && edge && y == 0) v = clip3i(0, maxv, LEFT(0) + ((TOP(x) - TOP(-1)) >> 1));
    }
    if (coded) v = clip3i(0, maxv, v + (int)rsb[e]);
    out[y * S + x] = (P)v;
  }
#undef LEFT
#undef TOP
  __syncwarp();
}

template <typename P, bool LIVE>
__global__ void __launch_bounds__(WARPS * 32, B200_RECON_MIN_BLOCKS) hevc_recon_kernel(const DeviceBatch b) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int8_t* mat = reinterpret_cast<int8_t*>(smem_raw);                       // the transposed DCT matrices, shared by the CTA
  for (int i = threadIdx.x; i < 1024 + 256 + 64; i += blockDim.x) {
    const int lg = i < 1024 ? 5 : (i < 1280 ? 4 : 3), j0 = i < 1024 ? i : (i < 1280 ? i - 1024 : i - 1280);
    const int n = 1 << lg, y = j0 >> lg, kn = j0 & (n - 1), k = kn << (5 - lg);     // row kn of the n-point matrix = row kn << (5 - log2 n) of the 32-point one
    int v;
    if (k == 0) v = 64;
    else { int j = (k * (2 * y + 1)) & 127, sgn = 1; if (j > 64) j = 128 - j; if (j > 32) { j = 64 - j; sgn = -1; } v = sgn * c_dct[j]; }
    mat[(lg == 5 ? MT32_OFF : (lg == 4 ? MT16_OFF : MT8_OFF)) + y * (n + 4) + kn] = (int8_t)v;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const WarpLayout L = warp_layout(b.max_log2_ctb, (int)sizeof(P));
  unsigned char* wb = smem_raw + MAT_BYTES + (threadIdx.x >> 5) * L.total;
  P* const tile0 = reinterpret_cast<P*>(wb + L.tile);
  P* const top0 = reinterpret_cast<P*>(wb + L.top);
  int16_t* const res0 = reinterpret_cast<int16_t*>(wb + L.res);
  int16_t* const tmp = reinterpret_cast<int16_t*>(wb + L.tmp);
  uint2* const desc = reinterpret_cast<uint2*>(wb + L.desc);
  const unsigned lt_mask = (1u << lane) - 1u;

  for (;;) {
    unsigned t = 0;
    if (lane == 0) t = atomicAdd(b.ticket, 1u);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= (unsigned)b.nrows) break;
    const uint2 pr = b.row_list[t];
    const PicDesc* pic = &b.pics[pr.x];
    // g = 0: luma; 1: Cb + Cr of a 4:2:0 picture on the two half-warps; 2 / 3: the Cb / Cr plane of a 4:2:2 or 4:4:4 picture, handled
    // like luma (full warp, its own commands: b200_hevc_syntax.h transform_unit_x)
    const int g = (int)(pr.y >> 30), ry = (int)(pr.y & 0x3fffffffu);
    const int pl = g >= 2 ? g - 1 : 0;                                      // plane of a luma-like item
    const int cfmt = pic->chroma;
    const int shx = g == 0 ? 0 : (g == 1 ? 1 : (cfmt == 2 ? 1 : 0)), shy = g == 1 ? 1 : 0;
    const bool paired = g == 1;
    const CtuInfo* ctus = b.ctus + pic->ctu_base;
    const SliceInfo* slices = b.slices + pic->slice_base;
    const int log2ctb = pic->log2_ctb, wctb = pic->wctb, bd = pic->bit_depth, strong_en = pic->strong_intra;
    const int tw = (1 << log2ctb) >> shx, th = (1 << log2ctb) >> shy, S = tw + PAD;   // component CTB size, tile row stride
    const int cw = pic->width >> shx, ch = pic->height >> shy;
    const int y0 = ry << log2ctb, cy0 = y0 >> shy;
    const uint8_t* sfac = pic->scaling_idx >= 0 ? b.scaling + (size_t)pic->scaling_idx * 784 : nullptr;      // sl::Factors: m[3][4][64], dc[3][4]
    const TuCmd* tus = b.tus + pic->tu_base;
    const CoefEntry* coefs = b.coefs + pic->coef_base;
    unsigned* prog = b.progress + 3 * pic->progress_base + (g == 3 ? 2 : (g ? 1 : 0));   // counter of (row r, item kind) at prog[3 * r]
    const unsigned* eprog = b.entropy_progress ? b.entropy_progress + pic->progress_base : nullptr;
    // lane roles in phase B
    const int lpc = paired ? 16 : 32, l = lane & (lpc - 1), cidx = paired ? lane >> 4 : 0;
    const unsigned gmask = paired ? (0xffffu << (16 * cidx)) : 0xffffffffu;
    P* const tl = tile0 + cidx * (th * S);
    P* const tp = top0 + cidx * (PAD + 2 * tw + 16);
    int16_t* const rs = res0 + cidx * (tw * th);
    int16_t* const rf = tmp + cidx * REF_STRIDE;
    const int plane = paired ? 1 + cidx : pl;
    P* const recp = static_cast<P*>(pic->rec[plane]);
    const int rst = pic->rec_stride[plane];

    for (int rx = 0; rx < wctb; rx++) {
      const int x0 = rx << log2ctb, cx0 = x0 >> shx;
      // (1) with K0 running concurrently: this CTB's commands must have been published.  Relaxed polling loads (no L1
      // invalidation) with microsecond back-off: waiting rows must not flood L2 with polls.
      if (eprog) {
        int abort = 0;
        if (lane == 0) {
          unsigned spins = 0, ns = 250;
          while (ld_acquire(&eprog[ry]) < (unsigned)(rx + 1)) {
            __nanosleep(ns); if (ns < 8000) ns <<= 1;
            if ((++spins & 31u) == 0 && ld_acquire(b.error_flag)) break;            // K0 failed (corrupt stream): its progress will never come
            if (spins > (1u << 23)) { atomicExch(b.error_flag, 1u); break; }       // ~1 min; turns a would-be hang into an error
          }
          abort = ld_acquire(b.error_flag) != 0u;
        }
        if (__shfl_sync(0xffffffffu, abort, 0)) return;
      }
      const int addr = ry * wctb + rx;
      const CtuInfo ci = ld_ctu<LIVE>(&ctus[addr]);
      const int cur = ci.slice_idx;
      const bool nbL = rx > 0 && (int)ld_cmd<LIVE>(&ctus[addr - 1].slice_idx) == cur;
      const bool nbAL = rx > 0 && ry > 0 && (int)ld_cmd<LIVE>(&ctus[addr - wctb - 1].slice_idx) == cur;
      const bool nbA = ry > 0 && (int)ld_cmd<LIVE>(&ctus[addr - wctb].slice_idx) == cur;
      const bool nbAR = ry > 0 && rx + 1 < wctb && (int)ld_cmd<LIVE>(&ctus[addr - wctb + 1].slice_idx) == cur;
      const SliceInfo sl = slices[cur];
      // ---- phase A: descriptors + residuals, lane-parallel over the CTB's transform units
      int ntb = 0;
#pragma unroll 1
      for (unsigned base = 0; base < ci.tu_count; base += 32) {
        const bool valid = base + lane < ci.tu_count;
        TuCmd cmd{0, 0, 0, 0};
        if (valid) cmd = ld_tu<LIVE>(&tus[ci.tu_start + base + lane]);
        const int log2n = 2 + (int)((cmd.w0 >> 24) & 3);
        const int lx = (int)((cmd.w0 & 0xfff) << 2) - x0, ly = (int)(((cmd.w0 >> 12) & 0xfff) << 2) - y0;   // luma position inside the CTB
        const int qpy = (int)((cmd.w1 >> 12) & 0xff) - 64;
        const bool pcm = (cmd.w1 >> 21) & 1, raw = ((cmd.w1 >> 21) & 3) != 0;
        const int nl = (int)(cmd.w3 & 0x7ff), ncb = (int)((cmd.w3 >> 11) & 0x3ff), ncr = (int)((cmd.w3 >> 21) & 0x3ff);
        const CoefEntry* ce = coefs + cmd.w2;
        bool has; int bx, by, lg, mode, coded0, coded1, qp0, qp1 = 0, ts0, ts1 = 0, n0, n1 = 0; const CoefEntry* ce1 = ce;
        if (!paired) {
          // a luma block, or (4:2:2 / 4:4:4) the block of plane `pl`: the commands of the other planes are skipped
          has = valid && (int)((cmd.w1 >> 23) & 3) == pl;
          bx = lx >> shx; by = ly; lg = log2n; mode = (int)(cmd.w1 & 63); coded0 = (int)((cmd.w0 >> 26) & 1); coded1 = 0;
          qp0 = pl == 0 ? qpy + 6 * (bd - 8) : chroma_qp(qpy, pl == 1 ? sl.cb_qp_offset : sl.cr_qp_offset, bd, cfmt);
          ts0 = (int)((cmd.w0 >> 30) & 1); n0 = nl;
        } else {
          has = valid && ((cmd.w0 >> 29) & 1);
          if (log2n > 2) { bx = lx >> 1; by = ly >> 1; lg = log2n - 1; } else { bx = (lx - 4) >> 1; by = (ly - 4) >> 1; lg = 2; }
          mode = (int)((cmd.w1 >> 6) & 63); coded0 = (int)((cmd.w0 >> 27) & 1); coded1 = (int)((cmd.w0 >> 28) & 1);
          qp0 = chroma_qp(qpy, sl.cb_qp_offset, bd); qp1 = chroma_qp(qpy, sl.cr_qp_offset, bd);
          ts0 = (int)((cmd.w0 >> 31) & 1); ts1 = (int)((cmd.w1 >> 20) & 1);
          ce = ce + nl; n0 = ncb; ce1 = ce + ncb; n1 = ncr;
        }
        if (!has) { coded0 = coded1 = 0; bx = by = 0; lg = 2; }
        coded0 = coded0 && n0 > 0; coded1 = coded1 && n1 > 0;
        const unsigned hb = __ballot_sync(0xffffffffu, has);
        const int idx = ntb + __popc(hb & lt_mask);
        const unsigned roff = morton4((unsigned)bx >> 2, (unsigned)by >> 2) * 16;
        if (has) desc[idx] = make_uint2(make_desc(bx, by, lg, mode, coded0, coded1, tw, th, paired ? 0 : shx, cx0, cy0, cw, ch, nbL, nbAL, nbA, nbAR) | ((unsigned)pcm << 29), roff);
        ntb += __popc(hb);
        // 4x4 blocks: one lane each, in registers
        if (lg == 2) {
#pragma unroll 1
          for (int c2 = 0; c2 < 2; c2++)
            if (c2 ? coded1 : coded0)
              residual4_lane<LIVE>(tmp, lane, c2 ? ce1 : ce, c2 ? n1 : n0, c2 ? qp1 : qp0, bd, g == 0, c2 ? ts1 : ts0, raw, res0 + (c2 ? tw * th : 0) + roff,
                                   sfac ? sfac + (paired ? 1 + c2 : pl) * 256 : nullptr);
        }
        __syncwarp();
        // larger blocks: the whole warp, one block at a time
        unsigned big0 = __ballot_sync(0xffffffffu, lg > 2 && coded0), big1 = __ballot_sync(0xffffffffu, lg > 2 && coded1);
#pragma unroll 1
        for (int c2 = 0; c2 < 2; c2++) {
          unsigned m = c2 ? big1 : big0;
          while (m) {
            const int src = __ffs(m) - 1; m &= m - 1;
            const unsigned long long cp = __shfl_sync(0xffffffffu, (unsigned long long)(c2 ? ce1 : ce), src);
            const int nn = __shfl_sync(0xffffffffu, c2 ? n1 : n0, src), lgg = __shfl_sync(0xffffffffu, lg, src), qq = __shfl_sync(0xffffffffu, c2 ? qp1 : qp0, src);
            const unsigned ro = __shfl_sync(0xffffffffu, roff, src); const bool rw = __shfl_sync(0xffffffffu, (int)raw, src) != 0;
            residual_big<LIVE>(res0 + (c2 ? tw * th : 0) + ro, tmp, mat, reinterpret_cast<const CoefEntry*>(cp), nn, lgg, qq, bd, lane,
                               sfac ? sfac + (paired ? 1 + c2 : pl) * 256 + (lgg - 2) * 64 : nullptr, sfac ? (int)sfac[768 + (paired ? 1 + c2 : pl) * 4 + (lgg - 2)] : 16, rw);
          }
        }
      }
      __syncwarp();
      // (2) wavefront: the above-right CTB of this component group must be reconstructed (lag 2)
      if (ry > 0) {
        int abort = 0;
        if (lane == 0) {
          const unsigned need = (unsigned)min(rx + 2, wctb);
          unsigned spins = 0, ns = 250;
          while (ld_acquire(&prog[3 * (ry - 1)]) < need) {
            __nanosleep(ns); if (ns < 8000) ns <<= 1;
            if ((++spins & 31u) == 0 && ld_acquire(b.error_flag)) break;
            if (spins > (1u << 23)) { atomicExch(b.error_flag, 1u); break; }
          }
          abort = ld_acquire(b.error_flag) != 0u;
        }
        if (__shfl_sync(0xffffffffu, abort, 0)) return;                      // the batch is reported as failed; nothing it produced is used
        // halo row above (corner .. above-right) from HBM/L2: written by another SM during this kernel -> L1-bypassing loads
        const int cnt = 1 + 2 * tw;
        const P* grow = recp + (size_t)(cy0 - 1) * rst;
#pragma unroll 1
        for (int i = l; i < cnt; i += lpc) {
          const int gx = cx0 - 1 + i;
          int v = 0;
          if (gx >= 0 && gx < cw) v = (int)__ldcg(grow + gx);
          tp[PAD - 1 + i] = (P)v;
        }
      }
      __syncwarp();
      // ---- phase B: prediction + reconstruction, block after block
#pragma unroll 1
      for (int k = 0; k < ntb; k++) predict_tb<P>(desc[k], tl, tp, rs, rf, S, l, lpc, gmask, cidx, g == 0, g == 0 || (g >= 2 && cfmt == 3), bd, g == 0 ? strong_en : 0);
      // ---- the finished CTB goes to HBM; its last column becomes the next CTB's left halo
      {
        const int w = min(tw, cw - cx0), h = min(th, ch - cy0);
        const unsigned rowb = (unsigned)(w * (int)sizeof(P));
        P* gdst = recp + (size_t)cy0 * rst + cx0;
        if ((rowb & 15u) == 0) {
          // TMA: one bulk row copy per lane (generic-proxy writes of the tile made visible to the async proxy first)
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
#pragma unroll 1
          for (int y = l; y < h; y += lpc) bulk_store(gdst + (size_t)y * rst, tl + y * S + PAD, rowb);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        } else {
          const int wq = (int)(rowb >> 2);                                     // 4-byte words per row (widths are multiples of 4 samples)
#pragma unroll 1
          for (int i = l; i < h * wq; i += lpc) {
            const int y = i / wq, q = i - y * wq;
            reinterpret_cast<unsigned*>(gdst + (size_t)y * rst)[q] = reinterpret_cast<const unsigned*>(tl + y * S + PAD)[q];
          }
        }
        // last column -> left halo of the next CTB (read before, written after the rows have left the tile)
        P keep0 = 0, keep1 = 0;
        if (l < th) keep0 = tl[l * S + PAD + tw - 1];
        if (l + lpc < th) keep1 = tl[(l + lpc) * S + PAD + tw - 1];
        if ((rowb & 15u) == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // rows are in global memory; the tile may be overwritten
        __syncwarp();
        if (l < th) tl[l * S + PAD - 1] = keep0;
        if (l + lpc < th) tl[(l + lpc) * S + PAD - 1] = keep1;
        __threadfence();
        __syncwarp();                                         // all lanes' stores precede lane 0's release store (cumulativity)
        if (lane == 0) st_release(&prog[3 * ry], (unsigned)(rx + 1));
      }
    }
  }
}

int launch_recon(const DeviceBatch& b, cudaStream_t s) {
  if (b.nrows <= 0) return B200_OK;
  const bool live = b.entropy_progress != nullptr;
  const int bps = b.wide_samples ? 2 : 1;
  const size_t smem = MAT_BYTES + (size_t)warp_layout(b.max_log2_ctb, bps).total * WARPS;
  const void* kern = bps == 2 ? (live ? (const void*)hevc_recon_kernel<uint16_t, true> : (const void*)hevc_recon_kernel<uint16_t, false>)
                              : (live ? (const void*)hevc_recon_kernel<uint8_t, true> : (const void*)hevc_recon_kernel<uint8_t, false>);
  B200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(MAT_BYTES + (size_t)warp_layout(6, 2).total * WARPS)));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int occ = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WARPS * 32, smem);
  if (occ < 1) occ = 1;
  if (b.blocks_per_sm > 0 && b.blocks_per_sm < occ) occ = b.blocks_per_sm;
  const int want = (b.nrows + WARPS - 1) / WARPS;
  const int grid = want < sms * occ ? want : sms * occ;
  void* args[] = {const_cast<DeviceBatch*>(&b)};
  cudaError_t e = cudaLaunchKernel(kern, dim3((unsigned)grid), dim3(WARPS * 32), args, smem, s);
  if (e != cudaSuccess) return set_error(B200_E_CUDA, "recon launch: %s", cudaGetErrorString(e));
  return B200_OK;
}

}  // namespace b200
