// b200_color.cu -- K6: fused colour post-stage for sm_100a.
//
// One pass over HBM does what the reference does in 2-4 passes with calloc'ed intermediates
// (libheif/color-conversion/colorconversion.cc:450-487 runs each op into a fresh image):
//   geometric transform (rotate_ccw / mirror / crop, libheif/image/pixelimage.cc:1175-1546) by addressing,
//   nearest-neighbour chroma upsampling (cx = x >> shiftH, cy = y >> shiftV, yuv2rgb.cc:226-229),
//   YCbCr -> RGB in the reference's integer arithmetic (yuv2rgb.cc:386-424) or float arithmetic
//   (yuv2rgb.cc:270-282, :696-709), optional ">> (bpp-8)" (hdr_sdr.cc:147-200), and the interleave /
//   endianness step (rgb2rgb.cc:71-150, yuv2rgb.cc:715-729).
//
// Layout: every CTA owns a 64x64 output tile.  The corresponding source window (also 64x64 because all
// transforms are axis permutations) is staged in shared memory with 16-byte loads, then each thread turns a
// 16-pixel strip into RGB and writes it with 128-bit stores (48 B of RGB24 = 3 x uint4).
// HBM-bound byte work: algorithmic traffic 4.5 B/px (8-bit -> RGB24), 9 B/px (16-bit -> RRGGBB).
//
// Bit-exactness: float expressions use __fmul_rn/__fadd_rn explicitly (no FMA contraction; the x86-64
// reference build has no FMA), evaluation order is the reference's, rounding is (int32)(fx + 0.5f).
#include "b200_internal.h"

namespace b200 {

constexpr int TILE = 64;
constexpr int SROWS = TILE + 2;     // an odd source origin needs one extra chroma row/column
constexpr int SPAD = 4;             // padding elements per shared row

struct K6Args {
  const void *y, *cb, *cr, *a;
  long long ys, cs, as;             // strides in bytes
  void* out[3];
  long long os;                     // output stride in bytes
  int src_w, src_h;
  int out_w, out_h;
  int m[6];
  int sh, sv;                       // chroma subsampling shifts; -1 = monochrome
  int bpp;
  int full_range;
  int int_mode;                     // 1: Op_YCbCr420_to_RGB24/32 integer arithmetic
  float cf[4];                      // r_cr, g_cb, g_cr, b_cb
  int ci[4];                        // lround(256*cf)
  int out_fmt;                      // b200_chroma value
  int sdr_shift;                    // Op_to_sdr_planes applied to the RGB result
  int pre_shift;                    // Op_to_sdr_planes applied to the YCbCr planes first (then integer op)
  int alpha_fill;                   // alpha value when the target wants alpha and the input has none
  int out_bytes;                    // bytes per output sample (1 or 2)
  int special;                      // matrix_coefficients branches of the generic op (yuv2rgb.cc:222-262): 1 = 0 (GBR), 2 = 8 (YCgCo), 3 = 16 (YCgCo-Re)
};

__device__ __forceinline__ int clip_f(float fx, int maxv) {      // common_utils.h:108-114 clip_f_u16
  int x = __float2int_rz(__fadd_rn(fx, 0.5f));
  return x < 0 ? 0 : (x > maxv ? maxv : x);
}
__device__ __forceinline__ int clip_u8(int x) { return x < 0 ? 0 : (x > 255 ? 255 : x); }

template <typename T>
__device__ __forceinline__ void convert_px(const K6Args& p, int Y, int Cb, int Cr, int& r, int& g, int& b) {
  if (p.sh < 0) { r = g = b = Y; return; }
  if (p.int_mode) {                                               // yuv2rgb.cc:401-417
    int cb = Cb - 128, cr = Cr - 128;
    r = clip_u8(Y + ((p.ci[0] * cr + 128) >> 8));
    g = clip_u8(Y + ((p.ci[1] * cb + p.ci[2] * cr + 128) >> 8));
    b = clip_u8(Y + ((p.ci[3] * cb + 128) >> 8));
    return;
  }
  const int half = 1 << (p.bpp - 1), maxv = (1 << p.bpp) - 1;     // yuv2rgb.cc:270-282
  if (p.special) {                                                // yuv2rgb.cc:222-262
    if (p.special == 1) {                                         // GBR: copy, or range-expand
      if (p.full_range) { r = Cr; g = Y; b = Cb; }
      else {
        const float lro = (float)(16 << (p.bpp - 8));
        r = clip_f(__fmul_rn(__fsub_rn((float)Cr, lro), 1.1429f), maxv);
        g = clip_f(__fmul_rn(__fsub_rn((float)Y, lro), 1.1689f), maxv);
        b = clip_f(__fmul_rn(__fsub_rn((float)Cb, lro), 1.1429f), maxv);
      }
    } else if (p.special == 2) {                                  // YCgCo; clip_int_u8 also for > 8 bit (reference quirk, :240-242)
      const int cb = Cb - half, cr = Cr - half;
      r = clip_u8(Y - cb + cr); g = clip_u8(Y + cb); b = clip_u8(Y - cb - cr);
    } else {                                                      // YCgCo-Re: int16 arithmetic, x4
      const short yy = (short)Y, cb = (short)((short)Cb - (short)half), cr = (short)((short)Cr - (short)half);
      const short t = (short)(yy - (cb >> 1)), gg = (short)(t + cb), bb = (short)(t - (cr >> 1)), rr = (short)(bb + cr);
      const int rv = rr * 4, gv = gg * 4, bv = bb * 4;
      r = rv < 0 ? 0 : (rv > maxv ? maxv : rv); g = gv < 0 ? 0 : (gv > maxv ? maxv : gv); b = bv < 0 ? 0 : (bv > maxv ? maxv : bv);
    }
    if (p.sdr_shift) { r >>= p.sdr_shift; g >>= p.sdr_shift; b >>= p.sdr_shift; }
    return;
  }
  float yv = (float)Y, cb = (float)(Cb - half), cr = (float)(Cr - half);
  if (!p.full_range) {
    yv = __fmul_rn(__fsub_rn(yv, (float)(16 << (p.bpp - 8))), 1.1689f);
    cb = __fmul_rn(cb, 1.1429f);
    cr = __fmul_rn(cr, 1.1429f);
  }
  r = clip_f(__fadd_rn(yv, __fmul_rn(p.cf[0], cr)), maxv);
  g = clip_f(__fadd_rn(__fadd_rn(yv, __fmul_rn(p.cf[1], cb)), __fmul_rn(p.cf[2], cr)), maxv);
  b = clip_f(__fadd_rn(yv, __fmul_rn(p.cf[3], cb)), maxv);
  if (p.sdr_shift) { r >>= p.sdr_shift; g >>= p.sdr_shift; b >>= p.sdr_shift; }
}

// stage a w x h window (origin x0,y0, clipped to pw x ph) of a plane into shared memory
template <typename T>
__device__ __forceinline__ void stage_plane(T (*dst)[TILE + SPAD], const void* base, long long stride, int x0, int y0,
                                            int w, int h, int pw, int ph) {
  const int tid = threadIdx.x;
  constexpr int VEC = 16 / sizeof(T);
  const bool vec_ok = (x0 % VEC == 0) && (w % VEC == 0) && (x0 + w <= pw) && (stride % 16 == 0) &&
                      ((reinterpret_cast<uintptr_t>(base) & 15) == 0);
  if (vec_ok) {
    const int per_row = w / VEC;
    for (int i = tid; i < per_row * h; i += blockDim.x) {
      int r = i / per_row, c = i - r * per_row;
      int sy = y0 + r;
      if (sy >= ph) continue;
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(static_cast<const char*>(base) + (long long)sy * stride) + (x0 / VEC + c));
      T tmp[VEC];
      *reinterpret_cast<uint4*>(tmp) = v;
#pragma unroll
      for (int k = 0; k < VEC; k++) dst[r][c * VEC + k] = tmp[k];
    }
  } else {
    for (int i = tid; i < w * h; i += blockDim.x) {
      int r = i / w, c = i - r * w;
      int sy = y0 + r, sx = x0 + c;
      if (sy < ph && sx < pw) dst[r][c] = reinterpret_cast<const T*>(static_cast<const char*>(base) + (long long)sy * stride)[sx];
    }
  }
}

template <typename T, bool HAS_ALPHA>
__global__ void __launch_bounds__(256) k6_color_kernel(const K6Args p) {
  __shared__ __align__(16) T sY[TILE][TILE + SPAD];
  __shared__ __align__(16) T sCb[SROWS][TILE + SPAD];
  __shared__ __align__(16) T sCr[SROWS][TILE + SPAD];
  __shared__ __align__(16) T sA[HAS_ALPHA ? TILE : 1][TILE + SPAD];

  const int ox = blockIdx.x * TILE, oy = blockIdx.y * TILE;
  const int tw = min(TILE, p.out_w - ox), th = min(TILE, p.out_h - oy);
  // source window = image of the output tile's corners
  const int ax = p.m[0] * ox + p.m[1] * oy + p.m[2], ay = p.m[3] * ox + p.m[4] * oy + p.m[5];
  const int bx = p.m[0] * (ox + tw - 1) + p.m[1] * (oy + th - 1) + p.m[2];
  const int by = p.m[3] * (ox + tw - 1) + p.m[4] * (oy + th - 1) + p.m[5];
  const int s0x = min(ax, bx), s0y = min(ay, by);
  const int sw = abs(ax - bx) + 1, sh_ = abs(ay - by) + 1;

  // round the window out to 16-sample columns where the picture allows so the vector path is taken
  int lx0 = s0x & ~15, lw = ((s0x + sw + 15) & ~15) - lx0;
  if (lw > TILE || lx0 + lw > p.src_w) { lx0 = s0x; lw = sw; }
  stage_plane<T>(sY, p.y, p.ys, lx0, s0y, lw, sh_, p.src_w, p.src_h);
  if (HAS_ALPHA) stage_plane<T>(sA, p.a, p.as, lx0, s0y, lw, sh_, p.src_w, p.src_h);
  int c0x = 0, c0y = 0;
  if (p.sh >= 0) {
    const int cw_pl = (p.src_w + (1 << p.sh) - 1) >> p.sh, ch_pl = (p.src_h + (1 << p.sv) - 1) >> p.sv;
    c0x = lx0 >> p.sh; c0y = s0y >> p.sv;
    int cw = ((lx0 + lw - 1) >> p.sh) - c0x + 1, chh = ((s0y + sh_ - 1) >> p.sv) - c0y + 1;
    int cl0 = c0x & ~15, clw = ((c0x + cw + 15) & ~15) - cl0;
    if (clw > TILE || cl0 + clw > cw_pl) { cl0 = c0x; clw = cw; }
    stage_plane<T>(sCb, p.cb, p.cs, cl0, c0y, clw, chh, cw_pl, ch_pl);
    stage_plane<T>(sCr, p.cr, p.cs, cl0, c0y, clw, chh, cw_pl, ch_pl);
    c0x = cl0;
  }
  __syncthreads();

  const int row = threadIdx.x >> 2, strip = threadIdx.x & 3;
  const int y = oy + row, x_begin = ox + strip * 16;
  if (row >= th || x_begin >= p.out_w) return;
  const int npx = min(16, p.out_w - x_begin);

  int R[16], G[16], B[16], A[16];
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const int x = x_begin + (i < npx ? i : 0);
    const int sx = p.m[0] * x + p.m[1] * y + p.m[2], sy = p.m[3] * x + p.m[4] * y + p.m[5];
    const int Y = sY[sy - s0y][sx - lx0] >> p.pre_shift;
    int Cb = 0, Cr = 0;
    if (p.sh >= 0) { Cb = sCb[(sy >> p.sv) - c0y][(sx >> p.sh) - c0x] >> p.pre_shift; Cr = sCr[(sy >> p.sv) - c0y][(sx >> p.sh) - c0x] >> p.pre_shift; }
    convert_px<T>(p, Y, Cb, Cr, R[i], G[i], B[i]);
    if (HAS_ALPHA) A[i] = sA[sy - s0y][sx - lx0] >> (p.sdr_shift + p.pre_shift); else A[i] = p.alpha_fill;
  }

  char* orow = static_cast<char*>(p.out[0]) + (long long)y * p.os;
  const int fmt = p.out_fmt;
  if (fmt == B200_CHROMA_444) {                      // planar RGB (yuv2rgb.cc Op_YCbCr_to_RGB output)
    for (int c = 0; c < 3; c++) {
      const int* v = c == 0 ? R : (c == 1 ? G : B);
      char* prow = static_cast<char*>(p.out[c]) + (long long)y * p.os;
      if (p.out_bytes == 1) for (int i = 0; i < npx; i++) reinterpret_cast<uint8_t*>(prow)[x_begin + i] = (uint8_t)v[i];
      else for (int i = 0; i < npx; i++) reinterpret_cast<uint16_t*>(prow)[x_begin + i] = (uint16_t)v[i];
    }
    return;
  }
  // interleaved formats: build the 16-pixel strip in registers, then 128-bit stores
  const int nch = (fmt == B200_CHROMA_INTERLEAVED_RGB || fmt == B200_CHROMA_INTERLEAVED_RRGGBB_BE || fmt == B200_CHROMA_INTERLEAVED_RRGGBB_LE) ? 3 : 4;
  const int bps = (fmt == B200_CHROMA_INTERLEAVED_RGB || fmt == B200_CHROMA_INTERLEAVED_RGBA) ? 1 : 2;
  const int le = (fmt == B200_CHROMA_INTERLEAVED_RRGGBB_LE || fmt == B200_CHROMA_INTERLEAVED_RRGGBBAA_LE);
  __align__(16) uint8_t buf[16 * 8];
  if (bps == 1) {
#pragma unroll
    for (int i = 0; i < 16; i++) {
      if (nch == 3) { buf[3 * i] = (uint8_t)R[i]; buf[3 * i + 1] = (uint8_t)G[i]; buf[3 * i + 2] = (uint8_t)B[i]; }
      else { buf[4 * i] = (uint8_t)R[i]; buf[4 * i + 1] = (uint8_t)G[i]; buf[4 * i + 2] = (uint8_t)B[i]; buf[4 * i + 3] = (uint8_t)A[i]; }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const int v[4] = {R[i], G[i], B[i], A[i]};
#pragma unroll
      for (int c = 0; c < 4; c++) {
        if (c < nch) {
          buf[(nch * i + c) * 2 + (le ? 1 : 0)] = (uint8_t)(v[c] >> 8);     // yuv2rgb.cc:715-729
          buf[(nch * i + c) * 2 + (le ? 0 : 1)] = (uint8_t)(v[c] & 0xff);
        }
      }
    }
  }
  const int bpp_out = nch * bps;
  char* dst = orow + (long long)x_begin * bpp_out;
  const int nbytes = npx * bpp_out;
  if (npx == 16 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    const uint4* src = reinterpret_cast<const uint4*>(buf);
    const int nvec = bpp_out;                      // 16 px * bpp_out bytes / 16
    for (int k = 0; k < nvec; k++) __stcs(reinterpret_cast<uint4*>(dst) + k, src[k]);
  } else {
    for (int k = 0; k < nbytes; k++) dst[k] = buf[k];
  }
}


// ---------------------------------------------------------------------------------------------- bilinear chroma
// Op_YCbCr420_bilinear_to_YCbCr444<T> (chroma_sampling.cc:623-700): one thread per full-resolution chroma sample.
// Interior: weights 9/3/3/1 (+8)/16; borders: 3/1 (+2)/4 with the reference's source indexing (cx/2, cy/2); corners copied.
template <typename T>
__global__ void __launch_bounds__(256) bilinear_420_to_444_kernel(const T* __restrict__ in_cb, const T* __restrict__ in_cr, long long cs,
                                                                  T* __restrict__ out_cb, T* __restrict__ out_cr, long long os, int w, int h) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w || y >= h) return;
  const bool weven = (w & 1) == 0, heven = (h & 1) == 0;
  for (int c = 0; c < 2; c++) {
    const T* in = c ? in_cr : in_cb;
    T* out = c ? out_cr : out_cb;
    int v;
    const bool top = y == 0, left = x == 0, right = weven && x == w - 1, bottom = heven && y == h - 1;
    if ((top || bottom) && (left || right)) v = in[(long long)(top ? 0 : h / 2 - 1) * cs + (left ? 0 : w / 2 - 1)];
    else if (top || bottom) {
      const int cx = (x - 1) >> 1, row = top ? 0 : h / 2 - 1;
      const int a = in[(long long)row * cs + cx / 2], b = in[(long long)row * cs + cx / 2 + 1];
      v = (x & 1) ? (3 * a + b + 2) / 4 : (a + 3 * b + 2) / 4;
    } else if (left || right) {
      const int cy = (y - 1) >> 1, col = left ? 0 : w / 2 - 1;
      const int a = in[(long long)(cy / 2) * cs + col], b = in[(long long)(cy / 2 + 1) * cs + col];
      v = (y & 1) ? (3 * a + b + 2) / 4 : (a + 3 * b + 2) / 4;
    } else {
      const int cx = (x - 1) >> 1, cy = (y - 1) >> 1;
      const int c00 = in[(long long)cy * cs + cx], c01 = in[(long long)cy * cs + cx + 1], c10 = in[(long long)(cy + 1) * cs + cx], c11 = in[(long long)(cy + 1) * cs + cx + 1];
      const int wx0 = (x & 1) ? 3 : 1, wx1 = 4 - wx0, wy0 = (y & 1) ? 3 : 1, wy1 = 4 - wy0;
      v = (c00 * wx0 * wy0 + c01 * wx1 * wy0 + c10 * wx0 * wy1 + c11 * wx1 * wy1 + 8) / 16;
    }
    out[(long long)y * os + x] = (T)v;
  }
}

// ---------------------------------------------------------------------------------------------- fast path
// Identity geometry, 4:2:0, no alpha, 16-byte aligned rows: no shared-memory staging.  One thread = 16 x 2 output
// pixels: 2 x 16 luma samples and 8 + 8 chroma samples come in with 128-bit / 64-bit loads, the chroma terms are
// computed once per 2x2 block (identical values to the per-pixel evaluation of the reference: same float products,
// same evaluation order), and each row leaves as NCH * BPS 128-bit streaming stores.
template <typename T, int INT_MODE, int NCH, int BPS, int LE>
__global__ void __launch_bounds__(256) k6_direct_kernel(const K6Args p) {
  const int x16 = blockIdx.x * blockDim.x + threadIdx.x;
  const int y2 = blockIdx.y * blockDim.y + threadIdx.y;
  if (x16 * 16 >= p.out_w || y2 * 2 >= p.out_h) return;
  const int x = x16 * 16, y = y2 * 2;
  T Y[2][16], Cb[8], Cr[8];
  if (sizeof(T) == 1) {
    *reinterpret_cast<uint4*>(Y[0]) = __ldcs(reinterpret_cast<const uint4*>(static_cast<const char*>(p.y) + (long long)y * p.ys + x));
    *reinterpret_cast<uint4*>(Y[1]) = __ldcs(reinterpret_cast<const uint4*>(static_cast<const char*>(p.y) + (long long)(y + 1) * p.ys + x));
    *reinterpret_cast<uint2*>(Cb) = __ldcs(reinterpret_cast<const uint2*>(static_cast<const char*>(p.cb) + (long long)y2 * p.cs + x / 2));
    *reinterpret_cast<uint2*>(Cr) = __ldcs(reinterpret_cast<const uint2*>(static_cast<const char*>(p.cr) + (long long)y2 * p.cs + x / 2));
  } else {
#pragma unroll
    for (int r = 0; r < 2; r++) {
      const uint4* src = reinterpret_cast<const uint4*>(static_cast<const char*>(p.y) + (long long)(y + r) * p.ys + x * 2);
      reinterpret_cast<uint4*>(Y[r])[0] = __ldcs(src); reinterpret_cast<uint4*>(Y[r])[1] = __ldcs(src + 1);
    }
    *reinterpret_cast<uint4*>(Cb) = __ldcs(reinterpret_cast<const uint4*>(static_cast<const char*>(p.cb) + (long long)y2 * p.cs + x));
    *reinterpret_cast<uint4*>(Cr) = __ldcs(reinterpret_cast<const uint4*>(static_cast<const char*>(p.cr) + (long long)y2 * p.cs + x));
  }
  const int bpp = p.bpp, half = 1 << (bpp - 1), maxv = (1 << bpp) - 1, pre = p.pre_shift, post = p.sdr_shift;
  const float lro = (float)(16 << (bpp - 8));
  const bool full = p.full_range != 0;
  // per chroma sample terms
  int ri[8], gi[8], bi[8]; float rf[8], g1f[8], g2f[8], bf[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int cbv = (int)Cb[i] >> pre, crv = (int)Cr[i] >> pre;
    if (INT_MODE) {
      const int cb = cbv - 128, cr = crv - 128;
      ri[i] = (p.ci[0] * cr + 128) >> 8; gi[i] = (p.ci[1] * cb + p.ci[2] * cr + 128) >> 8; bi[i] = (p.ci[3] * cb + 128) >> 8;
    } else {
      float cb = (float)(cbv - half), cr = (float)(crv - half);
      if (!full) { cb = __fmul_rn(cb, 1.1429f); cr = __fmul_rn(cr, 1.1429f); }
      rf[i] = __fmul_rn(p.cf[0], cr); g1f[i] = __fmul_rn(p.cf[1], cb); g2f[i] = __fmul_rn(p.cf[2], cr); bf[i] = __fmul_rn(p.cf[3], cb);
    }
  }
#pragma unroll
  for (int r = 0; r < 2; r++) {
    __align__(16) uint8_t buf[16 * NCH * BPS];
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const int yv0 = (int)Y[r][i] >> pre;
      int R, G, B;
      if (INT_MODE) { R = clip_u8(yv0 + ri[i >> 1]); G = clip_u8(yv0 + gi[i >> 1]); B = clip_u8(yv0 + bi[i >> 1]); }
      else {
        float yv = (float)yv0;
        if (!full) yv = __fmul_rn(__fsub_rn(yv, lro), 1.1689f);
        R = clip_f(__fadd_rn(yv, rf[i >> 1]), maxv) >> post;
        G = clip_f(__fadd_rn(__fadd_rn(yv, g1f[i >> 1]), g2f[i >> 1]), maxv) >> post;
        B = clip_f(__fadd_rn(yv, bf[i >> 1]), maxv) >> post;
      }
      if (BPS == 1) {
        buf[NCH * i] = (uint8_t)R; buf[NCH * i + 1] = (uint8_t)G; buf[NCH * i + 2] = (uint8_t)B;
        if (NCH == 4) buf[NCH * i + 3] = (uint8_t)p.alpha_fill;
      } else {
        const int v[4] = {R, G, B, p.alpha_fill};
#pragma unroll
        for (int c = 0; c < NCH; c++) { buf[(NCH * i + c) * 2 + (LE ? 1 : 0)] = (uint8_t)(v[c] >> 8); buf[(NCH * i + c) * 2 + (LE ? 0 : 1)] = (uint8_t)(v[c] & 0xff); }
      }
    }
    uint4* dst = reinterpret_cast<uint4*>(static_cast<char*>(p.out[0]) + (long long)(y + r) * p.os + (long long)x * NCH * BPS);
#pragma unroll
    for (int k = 0; k < NCH * BPS; k++) __stcs(dst + k, reinterpret_cast<const uint4*>(buf)[k]);
  }
}

template <typename T, int INT_MODE>
static bool launch_direct(const K6Args& a, cudaStream_t s) {
  dim3 blk(32, 8), grid((a.out_w / 16 + 31) / 32, (a.out_h / 2 + 7) / 8);
  switch (a.out_fmt) {
    case B200_CHROMA_INTERLEAVED_RGB: k6_direct_kernel<T, INT_MODE, 3, 1, 0><<<grid, blk, 0, s>>>(a); return true;
    case B200_CHROMA_INTERLEAVED_RGBA: k6_direct_kernel<T, INT_MODE, 4, 1, 0><<<grid, blk, 0, s>>>(a); return true;
    case B200_CHROMA_INTERLEAVED_RRGGBB_LE: if (INT_MODE) return false; k6_direct_kernel<T, 0, 3, 2, 1><<<grid, blk, 0, s>>>(a); return true;
    case B200_CHROMA_INTERLEAVED_RRGGBB_BE: if (INT_MODE) return false; k6_direct_kernel<T, 0, 3, 2, 0><<<grid, blk, 0, s>>>(a); return true;
    case B200_CHROMA_INTERLEAVED_RRGGBBAA_LE: if (INT_MODE) return false; k6_direct_kernel<T, 0, 4, 2, 1><<<grid, blk, 0, s>>>(a); return true;
    case B200_CHROMA_INTERLEAVED_RRGGBBAA_BE: if (INT_MODE) return false; k6_direct_kernel<T, 0, 4, 2, 0><<<grid, blk, 0, s>>>(a); return true;
    default: return false;
  }
}

// ---------------------------------------------------------------------------------------------- host
// nclx.cc:84-173, evaluated in float exactly as the reference does
static void primaries_of(int idx, float p[8], bool& defined) {
  defined = true;   // order: gx, gy, bx, by, rx, ry, wx, wy (nclx.cc:31-42 constructor argument order)
  switch (idx) {
    case 1: { const float v[8] = {0.300f, 0.600f, 0.150f, 0.060f, 0.640f, 0.330f, 0.3127f, 0.3290f}; memcpy(p, v, sizeof v); break; }
    case 4: { const float v[8] = {0.21f, 0.71f, 0.14f, 0.08f, 0.67f, 0.33f, 0.310f, 0.316f}; memcpy(p, v, sizeof v); break; }
    case 5: { const float v[8] = {0.29f, 0.60f, 0.15f, 0.06f, 0.64f, 0.33f, 0.3127f, 0.3290f}; memcpy(p, v, sizeof v); break; }
    case 6: case 7: { const float v[8] = {0.310f, 0.595f, 0.155f, 0.070f, 0.630f, 0.340f, 0.3127f, 0.3290f}; memcpy(p, v, sizeof v); break; }
    case 8: { const float v[8] = {0.243f, 0.692f, 0.145f, 0.049f, 0.681f, 0.319f, 0.310f, 0.316f}; memcpy(p, v, sizeof v); break; }
    case 9: { const float v[8] = {0.170f, 0.797f, 0.131f, 0.046f, 0.708f, 0.292f, 0.3127f, 0.3290f}; memcpy(p, v, sizeof v); break; }
    case 10: { const float v[8] = {0.0f, 1.0f, 0.0f, 0.0f, 1.0f, 0.0f, 0.333333f, 0.33333f}; memcpy(p, v, sizeof v); break; }
    case 11: { const float v[8] = {0.265f, 0.690f, 0.150f, 0.060f, 0.680f, 0.320f, 0.314f, 0.351f}; memcpy(p, v, sizeof v); break; }
    case 12: { const float v[8] = {0.265f, 0.690f, 0.150f, 0.060f, 0.680f, 0.320f, 0.3127f, 0.3290f}; memcpy(p, v, sizeof v); break; }
    case 22: { const float v[8] = {0.295f, 0.605f, 0.155f, 0.077f, 0.630f, 0.340f, 0.3127f, 0.3290f}; memcpy(p, v, sizeof v); break; }
    default: defined = false; for (int i = 0; i < 8; i++) p[i] = 0.0f;
  }
}

static void kr_kb(int matrix, int primaries, float* pKr, float* pKb);
void ycbcr_to_rgb_coefficients(int matrix, int primaries, float out[4]) {
  float kr, kb; kr_kb(matrix, primaries, &kr, &kb);
  if (kb != 0 || kr != 0) {
    out[0] = 2 * (-kr + 1);
    out[1] = 2 * kb * (-kb + 1) / (kb + kr - 1);
    out[2] = 2 * kr * (-kr + 1) / (kb + kr - 1);
    out[3] = 2 * (-kb + 1);
  } else { out[0] = 1.402f; out[1] = -0.344136f; out[2] = -0.714136f; out[3] = 1.772f; }
}
// RGB -> YCbCr coefficients of the same matrix (nclx.cc:177-200)
static void rgb_to_ycbcr_coefficients(int matrix, int primaries, float c[3][3]) {
  float Kr, Kb; kr_kb(matrix, primaries, &Kr, &Kb);
  if (Kb != 0 || Kr != 0) {
    c[0][0] = Kr; c[0][1] = 1 - Kr - Kb; c[0][2] = Kb;
    c[1][0] = -Kr / (1 - Kb) / 2; c[1][1] = -(1 - Kr - Kb) / (1 - Kb) / 2; c[1][2] = 0.5f;
    c[2][0] = 0.5f; c[2][1] = -(1 - Kr - Kb) / (1 - Kr) / 2; c[2][2] = -Kb / (1 - Kr) / 2;
  } else {
    c[0][0] = 0.299f; c[0][1] = 0.587f; c[0][2] = 0.114f; c[1][0] = -0.168735f; c[1][1] = -0.331264f; c[1][2] = 0.5f;
    c[2][0] = 0.5f; c[2][1] = -0.418688f; c[2][2] = -0.081312f;
  }
}
static void kr_kb(int matrix, int primaries, float* pKr, float* pKb) {
  volatile float Kr = 0.0f, Kb = 0.0f;              // volatile: keep every intermediate rounded to float
  if (matrix == 12 || matrix == 13) {
    float p[8]; bool def;
    primaries_of(primaries, p, def);
    const float gx = p[0], gy = p[1], bx = p[2], by = p[3], rx = p[4], ry = p[5], wx = p[6], wy = p[7];
    float zr = 1 - (rx + ry), zg = 1 - (gx + gy), zb = 1 - (bx + by), zw = 1 - (wx + wy);
    float denom = wy * (rx * (gy * zb - by * zg) + gx * (by * zr - ry * zb) + bx * (ry * zg - gy * zr));
    if (denom != 0.0f) {
      Kr = (ry * (wx * (gy * zb - by * zg) + wy * (bx * zg - gx * zb) + zw * (gx * by - bx * gy))) / denom;
      Kb = (by * (wx * (ry * zg - gy * zr) + wy * (gx * zr - rx * zg) + zw * (rx * gy - gx * ry))) / denom;
    }
  } else {
    switch (matrix) {
      case 1: Kr = 0.2126f; Kb = 0.0722f; break;
      case 4: Kr = 0.30f; Kb = 0.11f; break;
      case 5: case 6: Kr = 0.299f; Kb = 0.114f; break;
      case 7: Kr = 0.212f; Kb = 0.087f; break;
      case 9: case 10: Kr = 0.2627f; Kb = 0.0593f; break;
      default: break;
    }
  }
  *pKr = Kr; *pKb = Kb;
}

// Plane-wise geometry (rotate / mirror / crop on a 4:2:0 picture as the reference's ComponentStorage code does it) for
// the transforms that precede the 4:4:4 conversion point: luma-resolution planes follow the affine map, chroma planes
// follow it at half resolution (those transforms keep the 2x2 chroma grid aligned, otherwise the conversion point would
// have been earlier).
template <typename T>
__global__ void plane_geometry_kernel(const T* __restrict__ in, long long in_stride, T* __restrict__ out, long long out_stride, int out_w, int out_h,
                                      int m0, int m1, int m2, int m3, int m4, int m5, int shx, int shy) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
  if (u >= out_w || v >= out_h) return;
  const int sx = (m0 * (u << shx) + m1 * (v << shy) + m2) >> shx, sy = (m3 * (u << shx) + m4 * (v << shy) + m5) >> shy;
  out[(long long)v * out_stride + u] = in[(long long)sy * in_stride + sx];
}

// Op_YCbCr422_bilinear_to_YCbCr444<T> (chroma_sampling.cc:784-905): chroma (w + 1) / 2 x h -> w x h, both planes
template <typename T>
__global__ void bilinear_422_to_444_kernel(const T* __restrict__ cb, const T* __restrict__ cr, long long in_stride, T* __restrict__ ocb, T* __restrict__ ocr,
                                           long long out_stride, int w, int h) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w || y >= h) return;
  const T* rb = cb + (long long)y * in_stride; const T* rr = cr + (long long)y * in_stride;
  unsigned vb, vr;
  if (x == 0) { vb = rb[0]; vr = rr[0]; }
  else if ((w & 1) == 0 && x == w - 1) { vb = rb[w / 2 - 1]; vr = rr[w / 2 - 1]; }
  else if (x & 1) { const int c = x >> 1; vb = ((unsigned)rb[c] * 3 + rb[c + 1] + 2) / 4; vr = ((unsigned)rr[c] * 3 + rr[c + 1] + 2) / 4; }
  else { const int c = (x - 1) >> 1; vb = ((unsigned)rb[c] + (unsigned)rb[c + 1] * 3 + 2) / 4; vr = ((unsigned)rr[c] + (unsigned)rr[c + 1] * 3 + 2) / 4; }
  ocb[(long long)y * out_stride + x] = (T)vb; ocr[(long long)y * out_stride + x] = (T)vr;
}

// The 4:4:4 conversion point of a LIMITED-range picture: the reference's target profile there is full range, and the path
// its planner finds is Op_YCbCr_to_RGB<T> (yuv2rgb.cc:263-279) followed by Op_RGB_to_YCbCr<T> (rgb2yuv.cc:226-300) with
// the same matrix.  One pass over the three 4:4:4 planes, float arithmetic in the reference's order.
struct RangeArgs { float cf[4]; float c[3][3]; int bpp; };
template <typename T>
__global__ void range_limited_to_full_444_kernel(const T* __restrict__ y, long long ys, const T* __restrict__ cb, const T* __restrict__ cr, long long cs,
                                                 T* __restrict__ oy, long long oys, T* __restrict__ ocb, T* __restrict__ ocr, long long ocs, int w, int h, RangeArgs a) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, yy = blockIdx.y;
  if (x >= w || yy >= h) return;
  const int half = 1 << (a.bpp - 1), maxv = (1 << a.bpp) - 1; const float lro = (float)(16 << (a.bpp - 8));
  float yv = (float)y[(long long)yy * ys + x], cbv = (float)((int)cb[(long long)yy * cs + x] - half), crv = (float)((int)cr[(long long)yy * cs + x] - half);
  yv = __fmul_rn(__fsub_rn(yv, lro), 1.1689f); cbv = __fmul_rn(cbv, 1.1429f); crv = __fmul_rn(crv, 1.1429f);
  const float r = (float)clip_f(__fadd_rn(yv, __fmul_rn(a.cf[0], crv)), maxv);
  const float g = (float)clip_f(__fadd_rn(__fadd_rn(yv, __fmul_rn(a.cf[1], cbv)), __fmul_rn(a.cf[2], crv)), maxv);
  const float b = (float)clip_f(__fadd_rn(yv, __fmul_rn(a.cf[3], cbv)), maxv);
  auto dot = [&](int k) { return __fadd_rn(__fadd_rn(__fmul_rn(r, a.c[k][0]), __fmul_rn(g, a.c[k][1])), __fmul_rn(b, a.c[k][2])); };
  oy[(long long)yy * oys + x] = (T)clip_f(dot(0), maxv);
  ocb[(long long)yy * ocs + x] = (T)clip_f(__fadd_rn(dot(1), (float)half), maxv);
  ocr[(long long)yy * ocs + x] = (T)clip_f(__fadd_rn(dot(2), (float)half), maxv);
}

int launch_color(const b200_planes* in, const b200_geometry* g, const b200_color_options* opt, void* out, void* out_g,
                 void* out_b, size_t out_stride, cudaStream_t stream, int* pipeline);

// 4:2:0 / 4:2:2 picture -> (plane-wise geometry `pre`) -> Op_YCbCr420_bilinear_to_YCbCr444 / Op_YCbCr422_bilinear_to_YCbCr444
// -> (limited range: range conversion through RGB) -> the rest of the chain and the colour conversion from 4:4:4.
// Serves the reference's 4:4:4 conversion point and bilinear upsampling after geometry.
static int convert_via_444(const b200_planes* in, const b200_geometry* g, const b200_color_options* opt, void* out, void* out_g, void* out_b,
                           size_t out_stride, cudaStream_t stream, int* pipeline, bool range_convert) {
  const int bps = in->bit_depth > 8 ? 2 : 1;
  const bool c422 = in->chroma == B200_CHROMA_422;
  const int shy = c422 ? 0 : 1;
  const int pw = g->pre_w, ph = g->pre_h, pcw = (pw + 1) / 2, pch = c422 ? ph : (ph + 1) / 2;
  const bool pre_identity = g->pre[0] == 1 && g->pre[1] == 0 && g->pre[2] == 0 && g->pre[3] == 0 && g->pre[4] == 1 && g->pre[5] == 0 && pw == in->width && ph == in->height;
  const size_t pitch = (((size_t)pw * bps) + 255) & ~(size_t)255, cpitch = (((size_t)pcw * bps) + 255) & ~(size_t)255;
  const size_t n_full = pitch * ph, n_c = cpitch * pch;
  char* tmp = nullptr;
  B200_CUDA_CHECK(cudaMallocAsync(&tmp, 4 * n_full + 2 * n_c, stream));    // Y', A', Cb444, Cr444, Cb', Cr'
  char *ty = tmp, *ta = tmp + n_full, *u_cb = tmp + 2 * n_full, *u_cr = tmp + 3 * n_full, *tcb = tmp + 4 * n_full, *tcr = tmp + 4 * n_full + n_c;
  b200_planes p = *in;
  if (!pre_identity) {
    const int* q = g->pre;
    auto run = [&](const void* src, size_t sstride, void* dst, size_t dstride, int w, int h, int sx, int sy) {
      dim3 grid((w + 255) / 256, h);
      if (bps == 1) plane_geometry_kernel<uint8_t><<<grid, 256, 0, stream>>>((const uint8_t*)src, (long long)sstride, (uint8_t*)dst, (long long)dstride, w, h, q[0], q[1], q[2], q[3], q[4], q[5], sx, sy);
      else plane_geometry_kernel<uint16_t><<<grid, 256, 0, stream>>>((const uint16_t*)src, (long long)sstride / 2, (uint16_t*)dst, (long long)dstride / 2, w, h, q[0], q[1], q[2], q[3], q[4], q[5], sx, sy);
    };
    run(in->y, in->y_stride, ty, pitch, pw, ph, 0, 0);
    run(in->cb, in->c_stride, tcb, cpitch, pcw, pch, 1, shy);
    run(in->cr, in->c_stride, tcr, cpitch, pcw, pch, 1, shy);
    if (in->alpha) run(in->alpha, in->alpha_stride, ta, pitch, pw, ph, 0, 0);
    p.y = ty; p.y_stride = pitch; p.cb = tcb; p.cr = tcr; p.c_stride = cpitch;
    if (in->alpha) { p.alpha = ta; p.alpha_stride = pitch; }
    p.width = pw; p.height = ph;
  }
  dim3 grid((pw + 255) / 256, ph);
  if (c422) {
    if (bps == 1) bilinear_422_to_444_kernel<uint8_t><<<grid, 256, 0, stream>>>((const uint8_t*)p.cb, (const uint8_t*)p.cr, (long long)p.c_stride, (uint8_t*)u_cb, (uint8_t*)u_cr, (long long)pitch, pw, ph);
    else bilinear_422_to_444_kernel<uint16_t><<<grid, 256, 0, stream>>>((const uint16_t*)p.cb, (const uint16_t*)p.cr, (long long)p.c_stride / 2, (uint16_t*)u_cb, (uint16_t*)u_cr, (long long)pitch / 2, pw, ph);
  } else {
    if (bps == 1) bilinear_420_to_444_kernel<uint8_t><<<grid, 256, 0, stream>>>((const uint8_t*)p.cb, (const uint8_t*)p.cr, (long long)p.c_stride, (uint8_t*)u_cb, (uint8_t*)u_cr, (long long)pitch, pw, ph);
    else bilinear_420_to_444_kernel<uint16_t><<<grid, 256, 0, stream>>>((const uint16_t*)p.cb, (const uint16_t*)p.cr, (long long)p.c_stride / 2, (uint16_t*)u_cb, (uint16_t*)u_cr, (long long)pitch / 2, pw, ph);
  }
  p.cb = u_cb; p.cr = u_cr; p.c_stride = pitch; p.chroma = B200_CHROMA_444;
  if (range_convert) {
    // limited -> full range through RGB; Y goes to the scratch plane (the caller's luma plane is never written)
    RangeArgs ra; ra.bpp = in->bit_depth;
    ycbcr_to_rgb_coefficients(in->matrix_coefficients, in->colour_primaries, ra.cf);
    rgb_to_ycbcr_coefficients(in->matrix_coefficients, in->colour_primaries, ra.c);
    if (bps == 1) range_limited_to_full_444_kernel<uint8_t><<<grid, 256, 0, stream>>>((const uint8_t*)p.y, (long long)p.y_stride, (const uint8_t*)u_cb, (const uint8_t*)u_cr, (long long)pitch, (uint8_t*)ty, (long long)pitch, (uint8_t*)u_cb, (uint8_t*)u_cr, (long long)pitch, pw, ph, ra);
    else range_limited_to_full_444_kernel<uint16_t><<<grid, 256, 0, stream>>>((const uint16_t*)p.y, (long long)p.y_stride / 2, (const uint16_t*)u_cb, (const uint16_t*)u_cr, (long long)pitch / 2, (uint16_t*)ty, (long long)pitch / 2, (uint16_t*)u_cb, (uint16_t*)u_cr, (long long)pitch / 2, pw, ph, ra);
    p.y = ty; p.y_stride = pitch; p.width = pw; p.height = ph; p.full_range = 1;
  }
  b200_geometry rest = *g; rest.detour = 0; rest.chroma = B200_CHROMA_444;
  b200_color_options o2 = *opt; o2.chroma_upsampling = 0;
  int rc = launch_color(&p, &rest, &o2, out, out_g, out_b, out_stride, stream, pipeline);
  if (pipeline) *pipeline |= B200_PIPE_BILINEAR;
  cudaFreeAsync(tmp, stream);
  return rc;
}

// Mirror of the reference planner's choice for the supported states (colorconversion.cc:279-435; the
// measured pipelines are tabulated in SURVEY.md Appendix A and re-checked by tests/test_color_parity.py).
int launch_color(const b200_planes* in, const b200_geometry* g, const b200_color_options* opt, void* out, void* out_g,
                 void* out_b, size_t out_stride, cudaStream_t stream, int* pipeline) {
  if (!in || !g || !opt || !out) return set_error(B200_E_INVALID, "null argument");
  if (in->bit_depth < 8 || in->bit_depth > 16) return set_error(B200_E_UNSUPPORTED, "bit depth %d", in->bit_depth);
  const int fmt = opt->out_chroma;
  const bool interleaved8 = fmt == B200_CHROMA_INTERLEAVED_RGB || fmt == B200_CHROMA_INTERLEAVED_RGBA;
  const bool interleaved16 = fmt >= B200_CHROMA_INTERLEAVED_RRGGBB_BE && fmt <= B200_CHROMA_INTERLEAVED_RRGGBBAA_LE;
  if (!interleaved8 && !interleaved16 && fmt != B200_CHROMA_444) return set_error(B200_E_UNSUPPORTED, "output chroma %d", fmt);
  const int mc = in->matrix_coefficients;
  if (in->chroma != B200_CHROMA_MONO && (mc == 11 || mc == 14))
    return set_error(B200_E_UNSUPPORTED, "matrix_coefficients %d: the reference has no YCbCr->RGB operation for it either (yuv2rgb.cc:107-112)", mc);
  if (interleaved16 && in->bit_depth == 8) return set_error(B200_E_UNSUPPORTED, "8-bit input to RRGGBB output");
  const bool subsampled = in->chroma == B200_CHROMA_420 || in->chroma == B200_CHROMA_422;
  if (g->detour && !subsampled) {
    // the chain was composed with the rules of a subsampled format but the picture is not subsampled: one affine map
    b200_geometry t = *g; t.detour = 0;
    t.m[0] = g->pre[0] * g->m[0] + g->pre[1] * g->m[3]; t.m[1] = g->pre[0] * g->m[1] + g->pre[1] * g->m[4]; t.m[2] = g->pre[0] * g->m[2] + g->pre[1] * g->m[5] + g->pre[2];
    t.m[3] = g->pre[3] * g->m[0] + g->pre[4] * g->m[3]; t.m[4] = g->pre[3] * g->m[1] + g->pre[4] * g->m[4]; t.m[5] = g->pre[3] * g->m[2] + g->pre[4] * g->m[5] + g->pre[5];
    return launch_color(in, &t, opt, out, out_g, out_b, out_stride, stream, pipeline);
  }
  const bool geom_identity = g->m[0] == 1 && g->m[1] == 0 && g->m[2] == 0 && g->m[3] == 0 && g->m[4] == 1 && g->m[5] == 0 && g->out_w == in->width && g->out_h == in->height && !g->detour;
  if (subsampled && !geom_identity && g->chroma != in->chroma)
    return set_error(B200_E_INVALID, "geometry was composed for chroma format %d, the picture has %d (b200_geometry_init)", g->chroma, in->chroma);
  if (g->detour) {
    // 4:4:4 conversion point of the reference (see b200_geometry)
    // (a limited-range picture is also range-converted there: the conversion's target profile is full range)
    if (!in->full_range && (mc == 0 || mc == 8 || mc == 16))
      return set_error(B200_E_UNSUPPORTED, "limited-range picture with matrix_coefficients %d at the 4:4:4 conversion point", mc);
    return convert_via_444(in, g, opt, out, out_g, out_b, out_stride, stream, pipeline, !in->full_range);
  }
  if (opt->chroma_upsampling == 1 && in->chroma == B200_CHROMA_420) {
    // heif_color_conversion_options.only_use_preferred_chroma_algorithm with bilinear upsampling: the reference runs
    // Op_YCbCr420_bilinear_to_YCbCr444 first and converts from 4:4:4 with the generic float op.
    const bool identity = g->m[0] == 1 && g->m[1] == 0 && g->m[2] == 0 && g->m[3] == 0 && g->m[4] == 1 && g->m[5] == 0 && g->out_w == in->width && g->out_h == in->height;
    if (!identity) {
      // geometry happens on the planes BEFORE the colour conversion in the reference, so the bilinear op sees the
      // transformed 4:2:0 picture: all of the chain is `pre`, nothing is left after the upsampling
      b200_geometry t = *g; t.detour = 1;
      for (int i = 0; i < 6; i++) t.pre[i] = g->m[i];
      t.pre_w = g->out_w; t.pre_h = g->out_h;
      t.m[0] = 1; t.m[1] = 0; t.m[2] = 0; t.m[3] = 0; t.m[4] = 1; t.m[5] = 0;
      return convert_via_444(in, &t, opt, out, out_g, out_b, out_stride, stream, pipeline, false);
    }
    const int bps = in->bit_depth > 8 ? 2 : 1;
    const size_t pitch = (((size_t)in->width * bps) + 255) & ~(size_t)255;
    char* tmp = nullptr;
    B200_CUDA_CHECK(cudaMallocAsync(&tmp, 2 * pitch * in->height, stream));
    dim3 grid((in->width + 255) / 256, in->height);
    if (bps == 1) bilinear_420_to_444_kernel<uint8_t><<<grid, 256, 0, stream>>>((const uint8_t*)in->cb, (const uint8_t*)in->cr, (long long)in->c_stride, (uint8_t*)tmp, (uint8_t*)(tmp + pitch * in->height), (long long)pitch, in->width, in->height);
    else bilinear_420_to_444_kernel<uint16_t><<<grid, 256, 0, stream>>>((const uint16_t*)in->cb, (const uint16_t*)in->cr, (long long)in->c_stride / 2, (uint16_t*)tmp, (uint16_t*)(tmp + pitch * in->height), (long long)pitch / 2, in->width, in->height);
    b200_planes up = *in; up.cb = tmp; up.cr = tmp + pitch * in->height; up.c_stride = pitch; up.chroma = B200_CHROMA_444;
    b200_color_options o2 = *opt; o2.chroma_upsampling = 0;
    int rc = launch_color(&up, g, &o2, out, out_g, out_b, out_stride, stream, pipeline);
    if (pipeline) *pipeline |= B200_PIPE_BILINEAR;
    cudaFreeAsync(tmp, stream);
    return rc;
  }
  K6Args a{};
  a.y = in->y; a.cb = in->cb; a.cr = in->cr; a.a = in->alpha;
  a.ys = (long long)in->y_stride; a.cs = (long long)in->c_stride; a.as = (long long)in->alpha_stride;
  a.out[0] = out; a.out[1] = out_g; a.out[2] = out_b; a.os = (long long)out_stride;
  a.src_w = in->width; a.src_h = in->height; a.out_w = g->out_w; a.out_h = g->out_h;
  for (int i = 0; i < 6; i++) a.m[i] = g->m[i];
  switch (in->chroma) {
    case B200_CHROMA_MONO: a.sh = a.sv = -1; break;
    case B200_CHROMA_420: a.sh = 1; a.sv = 1; break;
    case B200_CHROMA_422: a.sh = 1; a.sv = 0; break;
    case B200_CHROMA_444: a.sh = 0; a.sv = 0; break;
    default: return set_error(B200_E_INVALID, "input chroma %d", in->chroma);
  }
  a.bpp = in->bit_depth; a.full_range = in->full_range ? 1 : 0;
  ycbcr_to_rgb_coefficients(mc, in->colour_primaries, a.cf);
  for (int i = 0; i < 4; i++) a.ci[i] = (int)lroundf(256 * a.cf[i]);      // yuv2rgb.cc:377-380
  a.out_fmt = fmt;
  const bool want_alpha = fmt == B200_CHROMA_INTERLEAVED_RGBA || fmt == B200_CHROMA_INTERLEAVED_RRGGBBAA_BE || fmt == B200_CHROMA_INTERLEAVED_RRGGBBAA_LE;
  const bool has_alpha = in->alpha != nullptr && want_alpha;
  int pipe = 0;
  // integer fast path: 4:2:0, 8 bit, full range, interleaved 8-bit target (yuv2rgb.cc:300-340, :440-478)
  // matrix_coefficients 0 / 8: the dedicated 4:2:0 ops are not selected (the generic op with its special branch runs);
  // 16: they are, and convert with the default coefficients (they have no YCgCo-Re branch) -- planner behaviour pinned in
  // tests/test_color_oracle.py: test_special_matrices_match_reference
  const bool dedicated_ok = !(mc == 0 || mc == 8);
  a.int_mode = (in->chroma == B200_CHROMA_420 && in->bit_depth == 8 && in->full_range && interleaved8 && dedicated_ok) ? 1 : 0;
  pipe |= a.int_mode ? B200_PIPE_INT420 : B200_PIPE_FLOAT;
  a.sdr_shift = 0; a.pre_shift = 0;
  a.out_bytes = in->bit_depth > 8 ? 2 : 1;
  if (in->bit_depth > 8 && interleaved8) {
    a.out_bytes = 1; pipe |= B200_PIPE_SDR_SHIFT;
    if (in->chroma == B200_CHROMA_420 && in->full_range && dedicated_ok) {
      // the reference planner shifts the YCbCr planes to 8 bit first and then takes the integer op
      a.pre_shift = in->bit_depth - 8; a.int_mode = 1; a.bpp = 8; pipe = B200_PIPE_SDR_SHIFT | B200_PIPE_INT420;
    } else a.sdr_shift = in->bit_depth - 8;
  }
  const bool rrggbb_direct = in->chroma == B200_CHROMA_420 && in->bit_depth > 8 && interleaved16 && dedicated_ok;     // Op_YCbCr420_to_RRGGBBaa
  a.special = 0;
  if (in->chroma != B200_CHROMA_MONO && (mc == 0 || mc == 8 || mc == 16) && !a.int_mode && !rrggbb_direct) a.special = mc == 0 ? 1 : (mc == 8 ? 2 : 3);
  a.alpha_fill = a.out_bytes == 1 ? 0xFF : (1 << in->bit_depth) - 1;
  if (fmt == B200_CHROMA_444 && (!out_g || !out_b)) return set_error(B200_E_INVALID, "planar output needs three planes");
  if (pipeline) *pipeline = pipe;
  if (g->out_w <= 0 || g->out_h <= 0) return B200_OK;
  // fast path: identity geometry, 4:2:0, no alpha, everything 16-byte aligned (the grid / single-image decode case)
  const bool identity = g->m[0] == 1 && g->m[1] == 0 && g->m[2] == 0 && g->m[3] == 0 && g->m[4] == 1 && g->m[5] == 0 && g->out_w == in->width && g->out_h == in->height;
  const uintptr_t al = (uintptr_t)in->y | (uintptr_t)in->cb | (uintptr_t)in->cr | (uintptr_t)out | (uintptr_t)in->y_stride | (uintptr_t)in->c_stride | (uintptr_t)out_stride;
  if (identity && in->chroma == B200_CHROMA_420 && !has_alpha && !a.special && (al & 15) == 0 && in->width % 16 == 0 && in->height % 2 == 0 && fmt != B200_CHROMA_444 && !(a.sdr_shift && a.out_bytes == 2)) {
    bool done;
    if (in->bit_depth == 8) done = a.int_mode ? launch_direct<uint8_t, 1>(a, stream) : launch_direct<uint8_t, 0>(a, stream);
    else done = a.int_mode ? launch_direct<uint16_t, 1>(a, stream) : launch_direct<uint16_t, 0>(a, stream);
    if (done) {
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) return set_error(B200_E_CUDA, "k6 direct launch: %s", cudaGetErrorString(e));
      return B200_OK;
    }
  }
  dim3 grid((g->out_w + TILE - 1) / TILE, (g->out_h + TILE - 1) / TILE);
  if (in->bit_depth == 8) {
    if (has_alpha) k6_color_kernel<uint8_t, true><<<grid, 256, 0, stream>>>(a);
    else k6_color_kernel<uint8_t, false><<<grid, 256, 0, stream>>>(a);
  } else {
    if (has_alpha) k6_color_kernel<uint16_t, true><<<grid, 256, 0, stream>>>(a);
    else k6_color_kernel<uint16_t, false><<<grid, 256, 0, stream>>>(a);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(B200_E_CUDA, "k6 launch: %s", cudaGetErrorString(e));
  return B200_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Encoder-side colour op: interleaved RGB / RGBA 8 bit -> planar YCbCr 4:2:0 / 4:2:2 / 4:4:4 8 bit
// (Op_RGB24_32_to_YCbCr, rgb2yuv.cc:575-808: float arithmetic in the reference's order, chroma of 4:2:0 from the integer
// mean of the 2x2 RGB quad -- 2x1 / 1x2 / 1x1 on an odd right column / bottom row --, chroma of 4:2:2 from the left pixel).
// One thread converts a 4 x 2 pixel block: 12 / 16 bytes per row in (vector loads when the rows are aligned), 4 luma bytes
// per row and 2..4 chroma bytes per plane out.  HBM bound: 3 or 4 bytes read, 1.5 .. 4 bytes written per pixel.
// ---------------------------------------------------------------------------------------------------------------
struct RgbToYccArgs {
  const uint8_t* in; long long in_stride;
  uint8_t *y, *cb, *cr, *a; long long ys, cs, as;
  int w, h, chroma, full, vec_in, vec_out;
  float c[3][3];
};
__device__ __forceinline__ int clip_round(float v, int maxv) {
  const int x = __float2int_rz(__fadd_rn(v, 0.5f));
  return x < 0 ? 0 : (x > maxv ? maxv : x);
}
__device__ __forceinline__ float dot3(int r, int g, int b, const float* c) {
  return __fadd_rn(__fadd_rn(__fmul_rn((float)r, c[0]), __fmul_rn((float)g, c[1])), __fmul_rn((float)b, c[2]));
}
__device__ __forceinline__ void put_chroma(const RgbToYccArgs& p, int r, int g, int b, uint8_t* ocb, uint8_t* ocr) {
  const float cb = dot3(r, g, b, p.c[1]), cr = dot3(r, g, b, p.c[2]);
  if (p.full) { *ocb = (uint8_t)clip_round(__fadd_rn(cb, 128.0f), 255); *ocr = (uint8_t)clip_round(__fadd_rn(cr, 128.0f), 255); }
  else {
    *ocb = (uint8_t)clip_round(__fadd_rn(__fmul_rn(cb, 0.875f), 128.0f), 255);
    *ocr = (uint8_t)clip_round(__fadd_rn(__fmul_rn(cr, 0.875f), 128.0f), 255);
  }
}
__device__ __forceinline__ void store_bytes(uint8_t* dst, const uint8_t* v, int n, bool vec) {
  if (vec && n == 4) *(uint32_t*)dst = (uint32_t)v[0] | ((uint32_t)v[1] << 8) | ((uint32_t)v[2] << 16) | ((uint32_t)v[3] << 24);
  else if (vec && n == 2) *(uint16_t*)dst = (uint16_t)((unsigned)v[0] | ((unsigned)v[1] << 8));
  else for (int i = 0; i < n; i++) dst[i] = v[i];
}
template <int BPP>
__global__ void __launch_bounds__(256) rgb_to_ycbcr_kernel(const RgbToYccArgs p) {
  const int x0 = (blockIdx.x * 64 + threadIdx.x) * 4, y0 = (blockIdx.y * 4 + threadIdx.y) * 2;
  if (x0 >= p.w || y0 >= p.h) return;
  const int nx = min(4, p.w - x0), ny = min(2, p.h - y0);
  uint8_t px[2][4][4];
#pragma unroll
  for (int j = 0; j < 2; j++) {
    if (j >= ny) break;
    const uint8_t* row = p.in + (long long)(y0 + j) * p.in_stride + (long long)x0 * BPP;
    if (p.vec_in && nx == 4) {
      if (BPP == 4) {
        const uint4 v = __ldg((const uint4*)row);
        const uint32_t q[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; i++) { px[j][i][0] = q[i] & 255; px[j][i][1] = (q[i] >> 8) & 255; px[j][i][2] = (q[i] >> 16) & 255; px[j][i][3] = q[i] >> 24; }
      } else {
        const uint32_t a = __ldg((const uint32_t*)row), b = __ldg((const uint32_t*)row + 1), c = __ldg((const uint32_t*)row + 2);
        px[j][0][0] = a & 255; px[j][0][1] = (a >> 8) & 255; px[j][0][2] = (a >> 16) & 255;
        px[j][1][0] = a >> 24; px[j][1][1] = b & 255; px[j][1][2] = (b >> 8) & 255;
        px[j][2][0] = (b >> 16) & 255; px[j][2][1] = b >> 24; px[j][2][2] = c & 255;
        px[j][3][0] = (c >> 8) & 255; px[j][3][1] = (c >> 16) & 255; px[j][3][2] = c >> 24;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; i++) {
        if (i >= nx) break;
#pragma unroll
        for (int k = 0; k < BPP; k++) px[j][i][k] = __ldg(row + i * BPP + k);
      }
    }
  }
  const bool vo = p.vec_out != 0;
  // luma (rgb2yuv.cc:640-665) and alpha (:667-691: copied, or 0xff when the source has none)
#pragma unroll
  for (int j = 0; j < 2; j++) {
    if (j >= ny) break;
    uint8_t yv[4], av[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (i >= nx) break;
      const float f = dot3(px[j][i][0], px[j][i][1], px[j][i][2], p.c[0]);
      yv[i] = p.full ? (uint8_t)clip_round(f, 255) : (uint8_t)(clip_round(__fmul_rn(f, 0.85547f), 219) + 16);
      av[i] = BPP == 4 ? px[j][i][3] : 0xff;
    }
    store_bytes(p.y + (long long)(y0 + j) * p.ys + x0, yv, nx, vo);
    if (p.a) store_bytes(p.a + (long long)(y0 + j) * p.as + x0, av, nx, vo);
  }
  if (p.chroma == B200_CHROMA_444) {
#pragma unroll
    for (int j = 0; j < 2; j++) {
      if (j >= ny) break;
      uint8_t vb[4], vr[4];
#pragma unroll
      for (int i = 0; i < 4; i++) { if (i >= nx) break; put_chroma(p, px[j][i][0], px[j][i][1], px[j][i][2], &vb[i], &vr[i]); }
      store_bytes(p.cb + (long long)(y0 + j) * p.cs + x0, vb, nx, vo);
      store_bytes(p.cr + (long long)(y0 + j) * p.cs + x0, vr, nx, vo);
    }
  } else if (p.chroma == B200_CHROMA_422) {
    const int nc = (nx + 1) >> 1;
#pragma unroll
    for (int j = 0; j < 2; j++) {
      if (j >= ny) break;
      uint8_t vb[2], vr[2];
#pragma unroll
      for (int i = 0; i < 2; i++) { if (i >= nc) break; put_chroma(p, px[j][2 * i][0], px[j][2 * i][1], px[j][2 * i][2], &vb[i], &vr[i]); }
      store_bytes(p.cb + (long long)(y0 + j) * p.cs + (x0 >> 1), vb, nc, vo);
      store_bytes(p.cr + (long long)(y0 + j) * p.cs + (x0 >> 1), vr, nc, vo);
    }
  } else {
    const int nc = (nx + 1) >> 1;
    uint8_t vb[2], vr[2];
#pragma unroll
    for (int i = 0; i < 2; i++) {
      if (i >= nc) break;
      const int qx = (2 * i + 1 < nx) ? 2 : 1, n = qx * ny;
      int s[3];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        int t = px[0][2 * i][k];
        if (qx == 2) t += px[0][2 * i + 1][k];
        if (ny == 2) { t += px[1][2 * i][k]; if (qx == 2) t += px[1][2 * i + 1][k]; }
        s[k] = t / n;
      }
      put_chroma(p, s[0], s[1], s[2], &vb[i], &vr[i]);
    }
    store_bytes(p.cb + (long long)(y0 >> 1) * p.cs + (x0 >> 1), vb, nc, vo);
    store_bytes(p.cr + (long long)(y0 >> 1) * p.cs + (x0 >> 1), vr, nc, vo);
  }
}

int launch_rgb_to_ycbcr(const void* rgb, size_t rgb_stride, int has_alpha, const b200_planes* out, cudaStream_t stream) {
  if (!rgb || !out || !out->y) return set_error(B200_E_INVALID, "null argument");
  if (out->bit_depth != 8) return set_error(B200_E_UNSUPPORTED, "RGB -> YCbCr: %d-bit target (the 8-bit interleaved op only)", out->bit_depth);
  if (out->chroma != B200_CHROMA_420 && out->chroma != B200_CHROMA_422 && out->chroma != B200_CHROMA_444)
    return set_error(B200_E_UNSUPPORTED, "RGB -> YCbCr: target chroma %d", out->chroma);
  if (!out->cb || !out->cr) return set_error(B200_E_INVALID, "RGB -> YCbCr: chroma planes missing");
  int mc = out->matrix_coefficients, cp = out->colour_primaries;
  if (mc == 0 || mc == 8 || mc == 11 || mc == 14)
    return set_error(B200_E_UNSUPPORTED, "matrix_coefficients %d: not converted by this operation in the reference either (rgb2yuv.cc:536-539)", mc);
  if (mc == 2) mc = 6;              // unspecified -> sRGB defaults, as convert_colorspace() does with the target profile
  if (cp == 2) cp = 1;              // (nclx.cc:360-373 through colorconversion.cc:513-515)
  if (out->width <= 0 || out->height <= 0) return B200_OK;
  const int bpp = has_alpha ? 4 : 3;
  if (rgb_stride < (size_t)out->width * bpp) return set_error(B200_E_INVALID, "RGB stride %zu < row of %d pixels", rgb_stride, out->width);
  RgbToYccArgs a;
  a.in = (const uint8_t*)rgb; a.in_stride = (long long)rgb_stride;
  a.y = (uint8_t*)out->y; a.cb = (uint8_t*)out->cb; a.cr = (uint8_t*)out->cr; a.a = (uint8_t*)out->alpha;
  a.ys = (long long)out->y_stride; a.cs = (long long)out->c_stride; a.as = (long long)out->alpha_stride;
  a.w = out->width; a.h = out->height; a.chroma = out->chroma; a.full = out->full_range ? 1 : 0;
  rgb_to_ycbcr_coefficients(mc, cp, a.c);
  const uintptr_t ai = (uintptr_t)rgb | (uintptr_t)rgb_stride;
  a.vec_in = (ai & (has_alpha ? 15 : 3)) == 0;
  uintptr_t ao = (uintptr_t)out->y | (uintptr_t)out->cb | (uintptr_t)out->cr | (uintptr_t)out->y_stride | (uintptr_t)out->c_stride;
  if (out->alpha) ao |= (uintptr_t)out->alpha | (uintptr_t)out->alpha_stride;
  a.vec_out = (ao & 3) == 0;
  const dim3 block(64, 4), grid((a.w + 255) / 256, (a.h + 7) / 8);
  if (has_alpha) rgb_to_ycbcr_kernel<4><<<grid, block, 0, stream>>>(a);
  else rgb_to_ycbcr_kernel<3><<<grid, block, 0, stream>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(B200_E_CUDA, "RGB -> YCbCr launch: %s", cudaGetErrorString(e));
  return B200_OK;
}


}  // namespace b200
