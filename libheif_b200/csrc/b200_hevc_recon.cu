// b200_hevc_recon.cu -- K1 + K2: scaling, inverse DCT/DST, intra prediction and reconstruction on sm_100a.
//
// Does the per-sample half of what libde265 does inside de265_decode() (reached from
// libheif/plugins/decoder_libde265.cc:386-457): H.265 8.6.2-8.6.4 (scaling, transforms), 8.4.4.2 (intra sample
// prediction incl. reference substitution and smoothing), 8.6.6 (reconstruction), driven by the command stream
// of the host front-end (b200_hevc_types.h).
//
// Mapping: ONE WARP OWNS ONE CTB ROW of one picture and walks its CTBs left to right (the only order the
// left-neighbour dependency allows); rows of a picture advance as a wavefront with a lag of two CTBs
// (above-right dependency), published through a per-row progress counter (release/acquire in global memory).
// Pictures of a batch (grid tiles) are independent, so a batch exposes (#pictures x CTB rows) warps of work; a
// global ticket hands rows out in picture-major order, which makes the spin-wait deadlock-free without requiring
// co-residency.  The CTB being reconstructed lives in shared memory together with its halo (row above incl.
// above-right, column to the left), all neighbour reads of the intra predictor hit shared memory; the finished
// CTB is written to HBM once, coalesced.  Integer work: no tensor cores.
#include "b200_hevc.h"

namespace b200 {

#ifndef B200_RECON_MIN_BLOCKS
#define B200_RECON_MIN_BLOCKS 5                // register cap: 65536 / (5 * 128) = 102 -> 96 registers per thread
#endif
constexpr int WARPS = 4;                       // warps (= CTB rows in flight) per CTA
constexpr int MAT_BYTES = 1024 + 16;           // 32x32 DCT matrix + 4x4 DST matrix, shared by the CTA

// Per-warp shared-memory working set; sized at launch from the largest CTB of the batch so that small CTBs buy
// occupancy (CTB 64: 17.7 KB per warp, CTB 32: 7.6 KB, CTB 16: 2.5 KB).
struct WarpMem {
  uint16_t* tile_y; uint16_t* tile_c[2];       // current CTB, stride ts / tsc
  int16_t* coef;                               // scaled coefficients [k][x]
  int16_t* tmp;                                // first-stage output, transposed: [x][y]
  uint16_t* top_y; uint16_t* left_y;           // halo: top_y[0] = above-left corner, top_y[1 + x], x < 2 * ctb
  uint16_t* top_c[2]; uint16_t* left_c[2];
  int16_t* ref_r; int16_t* ref_a; int16_t* ref_b;   // neighbour array (index 0 = bottom of left column): available samples only / substituted / filtered
  unsigned* avm;                               // availability bit mask of the neighbour array, 32 entries per word
  int ts, tsc;
};
__host__ __device__ inline size_t warp_mem_bytes(int log2ctb) {
  const int ctb = 1 << log2ctb, tb = ctb < 32 ? ctb : 32;
  size_t n = (size_t)ctb * ctb + 2 * (size_t)(ctb / 2) * (ctb / 2)      // tiles
           + 2 * (size_t)tb * tb                                         // coef + tmp
           + (1 + 2 * ctb + 3) + ctb + 2 * (1 + ctb + 3) + 2 * (ctb / 2) // halos
           + 3 * 136 + 16;                                               // ref_r, ref_a, ref_b, avm
  return (n * 2 + 15) & ~(size_t)15;
}
__device__ inline void warp_mem_init(WarpMem& m, unsigned char* base, int log2ctb) {
  const int ctb = 1 << log2ctb, tb = ctb < 32 ? ctb : 32, cc = ctb / 2;
  uint16_t* p = reinterpret_cast<uint16_t*>(base);
  m.ts = ctb; m.tsc = cc;
  m.tile_y = p; p += ctb * ctb;
  m.tile_c[0] = p; p += cc * cc; m.tile_c[1] = p; p += cc * cc;
  m.coef = reinterpret_cast<int16_t*>(p); p += tb * tb;
  m.tmp = reinterpret_cast<int16_t*>(p); p += tb * tb;
  m.top_y = p; p += 1 + 2 * ctb + 3;
  m.left_y = p; p += ctb;
  m.top_c[0] = p; p += 1 + ctb + 3; m.top_c[1] = p; p += 1 + ctb + 3;
  m.left_c[0] = p; p += cc; m.left_c[1] = p; p += cc;
  m.ref_r = reinterpret_cast<int16_t*>(p); p += 136;
  m.ref_a = reinterpret_cast<int16_t*>(p); p += 136;
  m.ref_b = reinterpret_cast<int16_t*>(p); p += 136;
  m.avm = reinterpret_cast<unsigned*>(p);
}

__constant__ int8_t c_dct[32] = {64, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67,
                                 64, 61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9, 4};
__constant__ int8_t c_dst[16] = {29, 55, 74, 84, 74, 74, 0, -74, 84, -29, -74, 55, 55, -84, 74, -29};
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

struct RowCtx {             // warp-uniform state of the CTB row being decoded
  const PicDesc* pic; const CtuInfo* ctus; const SliceInfo* slices;
  int W, H, log2ctb, ctb, wctb, bd, chroma, strong;
  int rx, ry, x0, y0;       // current CTB
  int cur_slice;
  int nb_slice[4];          // slice of the left, above-left, above, above-right CTB (-1: outside / not decoded)
};

// Availability of the luma location (xn, yn) for a block whose first luma sample is (xc, yc) -- H.265 6.4.1 restated
// for the wavefront schedule: previous CTB rows up to the above-right CTB and the left CTB are complete.
__device__ __forceinline__ bool available(const RowCtx& r, int xn, int yn, int xc, int yc) {
  if (xn < 0 || yn < 0 || xn >= r.W || yn >= r.H) return false;
  const int ncx = xn >> r.log2ctb, ncy = yn >> r.log2ctb;
  if (ncy > r.ry) return false;
  if (ncy == r.ry) {
    if (ncx > r.rx) return false;
    if (ncx == r.rx) {
      const unsigned m = r.ctb - 1;
      return morton4((xn & m) >> 2, (yn & m) >> 2) < morton4((xc & m) >> 2, (yc & m) >> 2);
    }
  }
  // neighbouring CTB: only left, above-left, above and above-right can be referenced (extent <= 2 * nTbS <= CTB size)
  const int k = ncy == r.ry ? 0 : 1 + (ncx - r.rx + 1);
  return r.nb_slice[k] == r.cur_slice;
}

// sample of component c at tile-relative position (tx, ty); tx in [-1, 2*ctb), ty in [-1, ctb)
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

__device__ __forceinline__ int tile_sample(const WarpMem& m, int c, int tx, int ty) {
  if (c == 0) { if (ty < 0) return m.top_y[tx + 1]; if (tx < 0) return m.left_y[ty]; return m.tile_y[ty * m.ts + tx]; }
  if (ty < 0) return m.top_c[c - 1][tx + 1];
  if (tx < 0) return m.left_c[c - 1][ty];
  return m.tile_c[c - 1][ty * m.tsc + tx];
}

// One transform block: 8.4.4.2 prediction into the tile, then (if coded) 8.6.3 scaling + 8.6.4 inverse transform
// + 8.6.6 reconstruction.  (bx, by): position inside the tile in samples of component c.
template <bool LIVE>
__device__ __noinline__ void process_tb(WarpMem& m, const int8_t* __restrict__ mat, const RowCtx& r, int c, int bx, int by, int log2n, int mode,
                           const CoefEntry* __restrict__ ce, int ncoef, int qp, int tskip, int lane) {
  const int n = 1 << log2n, sh = c ? 1 : 0, bd = r.bd;
  const int cx0 = c ? r.x0 >> 1 : r.x0, cy0 = c ? r.y0 >> 1 : r.y0;      // tile origin in component samples
  const int xl = (cx0 + bx) << sh, yl = (cy0 + by) << sh;                 // luma location of the block
  // ---- neighbour array with availability, then substitution (8.4.4.2.2)
  // The left / below-left / corner / above / above-right neighbour regions each lie inside ONE aligned block of the
  // current block's size, so z-scan availability (6.4.1) is decided per region (5 tests per block, done by lanes 0..4)
  // and per sample only the picture bounds remain.
  unsigned regions;
  {
    const int nl = n << sh;
    bool f = false;
    if (lane < 5) {
      const int xn = lane == 4 ? xl + nl : (lane == 3 ? xl : xl - 1);
      const int yn = lane == 0 ? yl : (lane == 1 ? yl + nl : yl - 1);      // 0: left, 1: below-left, 2: corner, 3: above, 4: above-right
      f = available(r, xn, yn, xl, yl);
    }
    regions = __ballot_sync(0xffffffffu, f);
  }
  const int wc = c ? r.W >> 1 : r.W, hc = c ? r.H >> 1 : r.H;
  const int nk = (4 * n + 32) >> 5;                 // 32-entry chunks that hold the 4n+1 neighbours
  // Loops over the chunks stay rolled: this function is the instruction-cache footprint of the kernel (measured: 60 % of
  // all stall cycles were instruction fetch when the three chunk loops were unrolled five-fold).
#pragma unroll 1
  for (int k = 0; k < nk; k++) {
    const int i = lane + 32 * k;
    bool av = false; int v = 0;
    if (i <= 4 * n) {
      int px, py;
      if (i < 2 * n) { const int y = 2 * n - 1 - i; px = bx - 1; py = by + y; av = ((regions >> (y < n ? 0 : 1)) & 1) && (cy0 + py < hc); }
      else if (i == 2 * n) { px = bx - 1; py = by - 1; av = (regions >> 2) & 1; }
      else { const int x = i - 2 * n - 1; px = bx + x; py = by - 1; av = ((regions >> (x < n ? 3 : 4)) & 1) && (cx0 + px < wc); }
      if (av) v = tile_sample(m, c, px, py);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, av);
    if (av) m.ref_r[i] = (int16_t)v;
    if (lane == 0) m.avm[k] = bal;
  }
  __syncwarp();
  {
    int first = -1;
#pragma unroll 1
    for (int k = nk - 1; k >= 0; k--) { const unsigned bal = m.avm[k]; if (bal) first = 32 * k + __ffs(bal) - 1; }
    const int fill = first < 0 ? (1 << (bd - 1)) : (int)m.ref_r[first];
    int carry = -1;                                   // highest available index in earlier chunks
#pragma unroll 1
    for (int k = 0; k < nk; k++) {
      const int i = lane + 32 * k;
      const unsigned bal = m.avm[k];
      const unsigned le = bal & (0xffffffffu >> (31 - lane));
      const int j = le ? 32 * k + 31 - __clz(le) : carry;
      if (i <= 4 * n) m.ref_a[i] = (int16_t)(j < 0 ? fill : (int)m.ref_r[j]);      // j == i when the sample itself is available
      if (bal) carry = 32 * k + 31 - __clz(bal);
    }
    __syncwarp();
  }
  // ---- smoothing of the neighbours (8.4.4.2.3): luma only in 4:2:0
  const int16_t* ref = m.ref_a;
  if (c == 0 && mode != 1 && n != 4) {
    const int dist = min(abs(mode - 26), abs(mode - 10));
    const int thr = n == 8 ? 7 : (n == 16 ? 1 : 0);
    if (dist > thr) {
      const int corner = m.ref_a[2 * n], bl = m.ref_a[0], tr = m.ref_a[4 * n];
      const bool strong = r.strong && n == 32 && abs(corner + tr - 2 * m.ref_a[3 * n]) < (1 << (bd - 5)) &&
                          abs(corner + bl - 2 * m.ref_a[n]) < (1 << (bd - 5));
#pragma unroll 1
      for (int k = 0; k < nk; k++) {
        const int i = lane + 32 * k;
        if (i <= 4 * n) {
          int v;
          if (i == 0 || i == 4 * n) v = m.ref_a[i];
          else if (strong) {
            if (i == 2 * n) v = corner;
            else if (i < 2 * n) { const int y = 2 * n - 1 - i; v = ((63 - y) * corner + (y + 1) * bl + 32) >> 6; }
            else { const int x = i - 2 * n - 1; v = ((63 - x) * corner + (x + 1) * tr + 32) >> 6; }
          } else v = (m.ref_a[i - 1] + 2 * m.ref_a[i] + m.ref_a[i + 1] + 2) >> 2;
          m.ref_b[i] = (int16_t)v;
        }
      }
      __syncwarp();
      ref = m.ref_b;
    }
  }
  // ---- prediction (8.4.4.2.4 - 8.4.4.2.6) written straight into the tile
  uint16_t* tile = c == 0 ? m.tile_y : m.tile_c[c - 1];
  const int ts = c == 0 ? m.ts : m.tsc;
  const int maxv = (1 << bd) - 1;
#define LEFT(y) ((int)ref[2 * n - 1 - (y)])
#define TOP(x) ((int)ref[2 * n + 1 + (x)])
  int dc = 0;
  if (mode == 1) {
    int s = 0;
    for (int i = lane; i < n; i += 32) s += LEFT(i) + TOP(i);
    s = __reduce_add_sync(0xffffffffu, s);
    dc = (s + n) >> (log2n + 1);
  }
  const int ang = c_angle[mode], ia = c_inv_angle[mode];
  const bool edge = c == 0 && n < 32;
  for (int p = lane; p < n * n; p += 32) {
    const int x = p & (n - 1), y = p >> log2n;
    int v;
    if (mode == 0) v = ((n - 1 - x) * LEFT(y) + (x + 1) * TOP(n) + (n - 1 - y) * TOP(x) + (y + 1) * LEFT(n) + n) >> (log2n + 1);
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
    tile[(by + y) * ts + bx + x] = (uint16_t)v;
  }
#undef LEFT
#undef TOP
  __syncwarp();
  if (ncoef == 0) return;
  // ---- scaling (8.6.3, flat scaling list m = 16)
#pragma unroll 1
  for (int i = lane; i < n * n / 2; i += 32) reinterpret_cast<uint32_t*>(m.coef)[i] = 0;
  __syncwarp();
  const int bd_shift = bd + log2n - 5;
  const long long scale = (long long)(c_level_scale[qp % 6] << (qp / 6)) * 16;
  int maxrow = 0, maxcol = 0;
#pragma unroll 1
  for (int i = lane; i < ncoef; i += 32) {
    const CoefEntry e = ld_coef<LIVE>(&ce[i]);
    const long long t = ((long long)e.level * scale + (1LL << (bd_shift - 1))) >> bd_shift;
    m.coef[e.pos] = (int16_t)(t < -32768 ? -32768 : (t > 32767 ? 32767 : t));
    maxrow = max(maxrow, e.pos >> log2n); maxcol = max(maxcol, e.pos & (n - 1));
  }
  maxrow = __reduce_max_sync(0xffffffffu, maxrow); maxcol = __reduce_max_sync(0xffffffffu, maxcol);
  __syncwarp();
  const int bs2 = 20 - bd;
  if (tskip) {                                          // 8.6.4.2, transform_skip_flag: r = d << 7
#pragma unroll 1
    for (int p = lane; p < n * n; p += 32) {
      const int x = p & (n - 1), y = p >> log2n;
      const int res = (((int)m.coef[p] << 7) + (1 << (bs2 - 1))) >> bs2;
      uint16_t* q = &tile[(by + y) * ts + bx + x];
      *q = (uint16_t)clip3i(0, maxv, (int)*q + res);
    }
    __syncwarp();
    return;
  }
  // matrix rows: DST-VII 4x4 for intra luma 4x4 (appended to the DCT matrix in shared memory), else row k of the
  // n-point DCT = row k << (5 - log2n) of the 32-point one
  const bool dst = c == 0 && log2n == 2;
  const int8_t* mrow = dst ? mat + 1024 : mat;
  const int mstride = dst ? 4 : (32 << (5 - log2n));
  // first stage (columns): tmp[x][y] = clip16((sum_k coef[k][x] * M[k][y] + 64) >> 7), only columns that hold coefficients
#pragma unroll 1
  for (int i = lane; i < n * (maxcol + 1); i += 32) {
    const int y = i & (n - 1), x = i >> log2n;
    int e = 0;
#pragma unroll 2
    for (int k = 0; k <= maxrow; k++) e += (int)m.coef[k * n + x] * (int)mrow[k * mstride + y];
    m.tmp[x * n + y] = (int16_t)clip3i(-32768, 32767, (e + 64) >> 7);
  }
  __syncwarp();
  // second stage (rows) + reconstruction (8.6.6)
#pragma unroll 1
  for (int p = lane; p < n * n; p += 32) {
    const int x = p & (n - 1), y = p >> log2n;
    int e = 0;
#pragma unroll 2
    for (int k = 0; k <= maxcol; k++) e += (int)m.tmp[k * n + y] * (int)mrow[k * mstride + x];
    const int res = (e + (1 << (bs2 - 1))) >> bs2;
    uint16_t* q = &tile[(by + y) * ts + bx + x];
    *q = (uint16_t)clip3i(0, maxv, (int)*q + res);
  }
  __syncwarp();
}

__device__ __forceinline__ int chroma_qp(int qpy, int off, int bd) {       // 8.6.1, ChromaArrayType == 1
  const int qbd = 6 * (bd - 8);
  const int qpi = clip3i(-qbd, 57, qpy + off);
  const int qpc = qpi < 30 ? qpi : (qpi >= 43 ? qpi - 6 : c_qpc[qpi - 30]);
  return qpc + qbd;
}

template <bool LIVE>
__global__ void __launch_bounds__(WARPS * 32, B200_RECON_MIN_BLOCKS) hevc_recon_kernel(const DeviceBatch b) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int8_t* mat = reinterpret_cast<int8_t*>(smem_raw);                       // 32x32 DCT matrix, shared by the CTA
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
    const int k = i >> 5, x = i & 31;
    int v;
    if (k == 0) v = 64;
    else { int j = (k * (2 * x + 1)) & 127, sgn = 1; if (j > 64) j = 128 - j; if (j > 32) { j = 64 - j; sgn = -1; } v = sgn * c_dct[j]; }
    mat[i] = (int8_t)v;
  }
  if (threadIdx.x < 16) mat[1024 + threadIdx.x] = c_dst[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  WarpMem m;
  warp_mem_init(m, smem_raw + MAT_BYTES + (threadIdx.x >> 5) * warp_mem_bytes(b.max_log2_ctb), b.max_log2_ctb);
  // process_tb is out of line and takes the warp's descriptors by reference: they live in SHARED memory (as locals they
  // would sit in local memory, one 128-byte line of L1 per word and warp; measured: 35 % of all stall cycles were loads
  // of exactly these fields).  The kernel body keeps its own register copies.
  __shared__ WarpMem s_wm[WARPS];
  __shared__ RowCtx s_row[WARPS];
  WarpMem& wm = s_wm[threadIdx.x >> 5];
  RowCtx& rs = s_row[threadIdx.x >> 5];
  if (lane == 0) wm = m;
  __syncwarp();

  for (;;) {
    unsigned t = 0;
    if (lane == 0) t = atomicAdd(b.ticket, 1u);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= (unsigned)b.nrows) break;
    const uint2 pr = b.row_list[t];
    const PicDesc* pic = &b.pics[pr.x];
    RowCtx r;
    r.pic = pic; r.ctus = b.ctus + pic->ctu_base; r.slices = b.slices + pic->slice_base;
    r.W = pic->width; r.H = pic->height; r.log2ctb = pic->log2_ctb; r.ctb = 1 << r.log2ctb; r.wctb = pic->wctb;
    r.bd = pic->bit_depth; r.chroma = pic->chroma; r.strong = pic->strong_intra;
    r.ry = (int)pr.y; r.y0 = r.ry << r.log2ctb;
    const TuCmd* tus = b.tus + pic->tu_base;
    const CoefEntry* coefs = b.coefs + pic->coef_base;
    unsigned* prog = b.progress + pic->progress_base;
    const unsigned* eprog = b.entropy_progress ? b.entropy_progress + pic->progress_base : nullptr;
    const bool b8 = r.bd == 8;
    const int ctbc = r.ctb >> 1;
    const int nch = r.chroma ? 3 : 1;

    for (r.rx = 0; r.rx < r.wctb; r.rx++) {
      r.x0 = r.rx << r.log2ctb;
      // (1) with K0 running concurrently: this CTB's commands must have been published; (2) wavefront: the above-right
      // CTB must be reconstructed (lag 2).  Relaxed polling loads (no L1 invalidation) with microsecond back-off (a CTB
      // takes ~0.2 ms): waiting rows must not flood L2 with polls.  Everything produced by another SM during this
      // kernel -- commands and the halo row -- is then read with L1-bypassing loads.
      int abort = 0;
      if (lane == 0) {
        unsigned spins = 0, ns = 250;
        if (eprog) while (ld_acquire(&eprog[r.ry]) < (unsigned)(r.rx + 1)) {
          __nanosleep(ns); if (ns < 8000) ns <<= 1;
          if ((++spins & 31u) == 0 && ld_acquire(b.error_flag)) break;            // K0 failed (corrupt stream): its progress will never come
          if (spins > (1u << 23)) { atomicExch(b.error_flag, 1u); break; }       // ~1 min; turns a would-be hang into an error
        }
        if (r.ry > 0) {
          const unsigned need = (unsigned)min(r.rx + 2, r.wctb);
          spins = 0; ns = 250;
          while (ld_acquire(&prog[r.ry - 1]) < need) {
            __nanosleep(ns); if (ns < 8000) ns <<= 1;
            if ((++spins & 31u) == 0 && ld_acquire(b.error_flag)) break;
            if (spins > (1u << 23)) { atomicExch(b.error_flag, 1u); break; }
          }
        }
        abort = ld_acquire(b.error_flag) != 0u;
      }
      abort = __shfl_sync(0xffffffffu, abort, 0);
      if (abort) return;                                     // the batch is reported as failed; nothing it produced is used
      const CtuInfo ci = ld_ctu<LIVE>(&r.ctus[r.ry * r.wctb + r.rx]);
      r.cur_slice = ci.slice_idx;
      r.nb_slice[0] = r.rx > 0 ? (int)ld_cmd<LIVE>(&r.ctus[r.ry * r.wctb + r.rx - 1].slice_idx) : -1;
      r.nb_slice[1] = (r.rx > 0 && r.ry > 0) ? (int)ld_cmd<LIVE>(&r.ctus[(r.ry - 1) * r.wctb + r.rx - 1].slice_idx) : -1;
      r.nb_slice[2] = r.ry > 0 ? (int)ld_cmd<LIVE>(&r.ctus[(r.ry - 1) * r.wctb + r.rx].slice_idx) : -1;
      r.nb_slice[3] = (r.ry > 0 && r.rx + 1 < r.wctb) ? (int)ld_cmd<LIVE>(&r.ctus[(r.ry - 1) * r.wctb + r.rx + 1].slice_idx) : -1;
      TuCmd next_cmd = ci.tu_count ? ld_tu<LIVE>(&tus[ci.tu_start]) : TuCmd{0, 0, 0, 0};
      if (lane == 0) rs = r;                                 // the shared copy process_tb reads (synchronised by the __syncwarp below / in process_tb)
      __syncwarp();
      if (r.ry > 0) {
        // fetch the halo row above from HBM/L2
        __syncwarp();
#pragma unroll 1
        for (int c = 0; c < nch; c++) {
          const int cw = c ? r.W >> 1 : r.W, st = pic->rec_stride[c];
          const int gx0 = (c ? r.x0 >> 1 : r.x0) - 1, gy = (c ? r.y0 >> 1 : r.y0) - 1;
          const int cnt = 1 + 2 * (c ? ctbc : r.ctb);
          uint16_t* dst = c == 0 ? m.top_y : m.top_c[c - 1];
#pragma unroll 1
          for (int i = lane; i < cnt; i += 32) {
            const int gx = gx0 + i;
            int v = 0;
            if (gx >= 0 && gx < cw) v = b8 ? (int)__ldcg(static_cast<const uint8_t*>(pic->rec[c]) + (size_t)gy * st + gx)
                                           : (int)__ldcg(static_cast<const uint16_t*>(pic->rec[c]) + (size_t)gy * st + gx);
            dst[i] = (uint16_t)v;
          }
        }
        __syncwarp();
      }
      const SliceInfo sl = r.slices[ci.slice_idx];
      for (unsigned ti = 0; ti < ci.tu_count; ti++) {
        const TuCmd cmd = next_cmd;
        if (ti + 1 < ci.tu_count) next_cmd = ld_tu<LIVE>(&tus[ci.tu_start + ti + 1]);      // prefetch: hides one dependent HBM/L2 round trip per TU
        const int x4 = cmd.w0 & 0xfff, y4 = (cmd.w0 >> 12) & 0xfff, log2n = 2 + ((cmd.w0 >> 24) & 3);
        const int lmode = cmd.w1 & 63, cmode = (cmd.w1 >> 6) & 63, qpy = (int)((cmd.w1 >> 12) & 0xff) - 64;
        const int nl = cmd.w3 & 0x7ff, ncb = (cmd.w3 >> 11) & 0x3ff, ncr = (cmd.w3 >> 21) & 0x3ff;
        const CoefEntry* ce = coefs + cmd.w2;
        const int bx = (x4 << 2) - r.x0, by = (y4 << 2) - r.y0;
        process_tb<LIVE>(wm, mat, rs, 0, bx, by, log2n, lmode, ce, ((cmd.w0 >> 26) & 1) ? nl : 0, qpy + 6 * (r.bd - 8), (cmd.w0 >> 30) & 1, lane);
        if ((cmd.w0 >> 29) & 1) {
          int cbx, cby, clog;
          if (log2n > 2) { cbx = bx >> 1; cby = by >> 1; clog = log2n - 1; } else { cbx = (bx - 4) >> 1; cby = (by - 4) >> 1; clog = 2; }
          process_tb<LIVE>(wm, mat, rs, 1, cbx, cby, clog, cmode, ce + nl, ((cmd.w0 >> 27) & 1) ? ncb : 0, chroma_qp(qpy, sl.cb_qp_offset, r.bd), (cmd.w0 >> 31) & 1, lane);
          process_tb<LIVE>(wm, mat, rs, 2, cbx, cby, clog, cmode, ce + nl + ncb, ((cmd.w0 >> 28) & 1) ? ncr : 0, chroma_qp(qpy, sl.cr_qp_offset, r.bd), (cmd.w1 >> 20) & 1, lane);
        }
      }
      // write the finished CTB to HBM (coalesced rows), keep its last column as the next CTB's left halo
#pragma unroll 1
      for (int c = 0; c < nch; c++) {
        const int cw = c ? r.W >> 1 : r.W, chh = c ? r.H >> 1 : r.H, st = pic->rec_stride[c];
        const int gx0 = c ? r.x0 >> 1 : r.x0, gy0 = c ? r.y0 >> 1 : r.y0, sz = c ? ctbc : r.ctb, lg = c ? r.log2ctb - 1 : r.log2ctb;
        const uint16_t* tile = c == 0 ? m.tile_y : m.tile_c[c - 1];
        const int ts = c == 0 ? m.ts : m.tsc;
        const int w = min(sz, cw - gx0), h = min(sz, chh - gy0);
#pragma unroll 2
        for (int i = lane; i < sz * h; i += 32) {
          const int x = i & (sz - 1), y = i >> lg;
          if (x < w) {
            const uint16_t v = tile[y * ts + x];
            if (b8) static_cast<uint8_t*>(pic->rec[c])[(size_t)(gy0 + y) * st + gx0 + x] = (uint8_t)v;
            else static_cast<uint16_t*>(pic->rec[c])[(size_t)(gy0 + y) * st + gx0 + x] = v;
          }
        }
        uint16_t* left = c == 0 ? m.left_y : m.left_c[c - 1];
#pragma unroll 1
        for (int y = lane; y < sz; y += 32) left[y] = tile[y * ts + sz - 1];
      }
      __syncwarp();                                         // all lanes' stores precede lane 0's release store (cumulativity)
      if (lane == 0) st_release(&prog[r.ry], (unsigned)(r.rx + 1));
    }
  }
}

int launch_recon(const DeviceBatch& b, cudaStream_t s) {
  if (b.nrows <= 0) return B200_OK;
  const size_t smem = MAT_BYTES + warp_mem_bytes(b.max_log2_ctb) * WARPS;
  const bool live = b.entropy_progress != nullptr;
  auto kern = live ? hevc_recon_kernel<true> : hevc_recon_kernel<false>;
  B200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(MAT_BYTES + warp_mem_bytes(6) * WARPS)));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int occ = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WARPS * 32, smem);
  if (occ < 1) occ = 1;
  if (b.blocks_per_sm > 0 && b.blocks_per_sm < occ) occ = b.blocks_per_sm;
  const int want = (b.nrows + WARPS - 1) / WARPS;
  const int grid = want < sms * occ ? want : sms * occ;
  kern<<<grid, WARPS * 32, smem, s>>>(b);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(B200_E_CUDA, "recon launch: %s", cudaGetErrorString(e));
  return B200_OK;
}

}  // namespace b200
