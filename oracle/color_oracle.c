/*
 * oracle/color_oracle.c -- TEST INFRASTRUCTURE ONLY. Never linked, imported or executed by the product.
 *
 * Plain C restatement of the reference's post-stage, written in the reference's own ORDER of operations
 * (geometry on the separate planes first, then colour conversion op by op), so that it is an independent
 * formulation of what the fused CUDA kernel computes by addressing:
 *   HeifPixelImage::rotate_ccw / mirror_inplace / crop      libheif/image/pixelimage.cc:1175-1546
 *   Op_YCbCr_to_RGB<T> (float arithmetic)                   libheif/color-conversion/yuv2rgb.cc:92-292
 *   Op_YCbCr420_to_RGB24 / RGB32 (integer arithmetic)       yuv2rgb.cc:345-426, :481-562
 *   Op_YCbCr420_to_RRGGBBaa                                 yuv2rgb.cc:622-734
 *   Op_RGB_to_RGB24_32, Op_to_sdr_planes                    rgb2rgb.cc:71-150, hdr_sdr.cc:147-200
 *   Op_YCbCr420_bilinear_to_YCbCr444<T>                     chroma_sampling.cc:501-724 (incl. its border indexing)
 *   get_YCbCr_to_RGB_coefficients / get_Kr_Kb               nclx.cc:84-173
 *   HeifPixelImage::overlay, scale_nearest_neighbor         pixelimage.cc:1637-1780, :1783-1972
 * PINNED against the unmodified reference (oracle/_ref/libheif_ref.so through ref_postprocess() in
 * oracle/ref_plugin.cc) and the reference's golden table tests/conversion.cc:697-724 by tests/test_color_oracle.py.
 *
 * Build note: compiled with -O2 for baseline x86-64 (no FMA), like the reference; float expressions are
 * written in the reference's evaluation order.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  uint16_t* p[4];     /* Y, Cb, Cr, A (samples widened to uint16) */
  int w, h, cw, ch;   /* luma and chroma plane sizes */
  int chroma;         /* 0 mono, 1 420, 2 422, 3 444 */
  int has_alpha;
} co_image;

static void co_free(co_image* im) { for (int i = 0; i < 4; i++) { free(im->p[i]); im->p[i] = NULL; } }

static uint16_t* plane_rot(const uint16_t* in, int w, int h, int deg) {      /* pixelimage.cc:1303-1332 */
  int ow = (deg == 180) ? w : h, oh = (deg == 180) ? h : w;
  uint16_t* out = (uint16_t*)malloc((size_t)ow * oh * 2 + 2);
  if (deg == 270) { for (int x = 0; x < h; x++) for (int y = 0; y < w; y++) out[y * ow + x] = in[(h - 1 - x) * w + y]; }
  else if (deg == 180) { for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) out[y * ow + x] = in[(h - 1 - y) * w + (w - 1 - x)]; }
  else { for (int x = 0; x < h; x++) for (int y = 0; y < w; y++) out[y * ow + x] = in[x * w + (w - 1 - y)]; }
  return out;
}

static void plane_mirror(uint16_t* d, int w, int h, int direction) {           /* pixelimage.cc:1336-1355 */
  if (direction == 1) { for (int y = 0; y < h; y++) for (int x = 0; x < w / 2; x++) { uint16_t t = d[y * w + x]; d[y * w + x] = d[y * w + w - 1 - x]; d[y * w + w - 1 - x] = t; } }
  else { for (int y = 0; y < h / 2; y++) for (int x = 0; x < w; x++) { uint16_t t = d[y * w + x]; d[y * w + x] = d[(h - 1 - y) * w + x]; d[(h - 1 - y) * w + x] = t; } }
}

void co_bilinear_420_to_444(const uint16_t* in, int w, int h, uint16_t* out);

/* Op_YCbCr422_bilinear_to_YCbCr444<T> (chroma_sampling.cc:784-905) for one chroma plane: in (w+1)/2 x h, out w x h. */
void co_bilinear_422_to_444(const uint16_t* in, int w, int h, uint16_t* out) {
  const int cw = (w + 1) / 2;
  for (int y = 0; y < h; y++) out[(size_t)y * w] = in[(size_t)y * cw];                                   /* left border */
  if (w % 2 == 0) for (int y = 0; y < h; y++) out[(size_t)y * w + w - 1] = in[(size_t)y * cw + w / 2 - 1];  /* right border */
  for (int y = 0; y < h; y++) for (int x = 1; x < w - 1; x += 2) {
    const int cx = x / 2;
    const unsigned c00 = in[(size_t)y * cw + cx], c01 = in[(size_t)y * cw + cx + 1];
    out[(size_t)y * w + x] = (uint16_t)((c00 * 3 + c01 + 2) / 4);
    out[(size_t)y * w + x + 1] = (uint16_t)((c00 + c01 * 3 + 2) / 4);
  }
}

/* rotate_ccw / mirror_inplace / crop on subsampled images the plane-wise code cannot handle first convert to 4:4:4
 * (pixelimage.cc:1187-1215, 1370-1396, 1458-1481): convert_colorspace(YCbCr, 4:4:4, nclx_profile() /+ undefined, but
 * full_range_flag = true +/, bpp, default options) = Op_YCbCr420_bilinear_to_YCbCr444 for a full-range 4:2:0 image.  A
 * limited-range image is additionally range-converted through RGB (the default target profile is full range); a 4:2:2
 * image takes Op_YCbCr422_bilinear_to_YCbCr444. */
static int co_detour_needed(const co_image* im, const int* o) {
  const int ow = im->w & 1, oh = im->h & 1;
  if (im->chroma == 2) {
    if (o[0] == 1) return (o[1] == 90 || o[1] == 270) || (o[1] == 180 && oh);
    if (o[0] == 2) return o[1] == 1 && ow;
    if (o[0] == 3) return o[1] & 1;
  } else if (im->chroma == 1) {
    if (o[0] == 1) return (o[1] == 90 && ow) || (o[1] == 180 && (ow || oh)) || (o[1] == 270 && oh);
    if (o[0] == 2) return ow || oh;
    if (o[0] == 3) return (o[1] & 1) || (o[3] & 1);
  }
  return 0;
}

static int clip_f_u16(float fx, int maxi);
static void co_kr_kb(int matrix, int primaries, float* pKr, float* pKb);
void co_coefficients(int matrix, int primaries, float out[4]);
static int co_geometry(co_image* im, const int* ops, int nops, int* full_range, int bpp, int cp, int mc) {
  for (int i = 0; i < nops; i++) {
    const int* o = ops + 5 * i;
    int np = im->chroma ? 3 : 1;
    if (co_detour_needed(im, o)) {
      for (int c = 1; c <= 2; c++) {
        uint16_t* up = (uint16_t*)malloc((size_t)im->w * im->h * 2 + 2);
        if (im->chroma == 1) co_bilinear_420_to_444(im->p[c], im->w, im->h, up); else co_bilinear_422_to_444(im->p[c], im->w, im->h, up);
        free(im->p[c]); im->p[c] = up;
      }
      im->chroma = 3; im->cw = im->w; im->ch = im->h;
      if (!*full_range) {
        /* The target profile of the detour is nclx_profile() with full_range_flag = true, so a limited-range picture is
           also range-converted; the path the planner finds goes through RGB after the upsampling: Op_YCbCr_to_RGB<T>
           (generic float op on the 4:4:4 picture, yuv2rgb.cc:263-279) then Op_RGB_to_YCbCr<T> to full-range 4:4:4 with
           the same matrix (rgb2yuv.cc:226-300, coefficients nclx.cc:177-200). */
        if (mc == 0 || mc == 8 || mc == 11 || mc == 14 || mc == 16) return -1;
        const int half = 1 << (bpp - 1), maxv = (1 << bpp) - 1; const float lro = (float)(16 << (bpp - 8));
        float cf[4]; co_coefficients(mc, cp, cf);
        float Kr, Kb; co_kr_kb(mc, cp, &Kr, &Kb);
        float c[3][3];
        if (Kb != 0 || Kr != 0) {
          c[0][0] = Kr; c[0][1] = 1 - Kr - Kb; c[0][2] = Kb;
          c[1][0] = -Kr / (1 - Kb) / 2; c[1][1] = -(1 - Kr - Kb) / (1 - Kb) / 2; c[1][2] = 0.5f;
          c[2][0] = 0.5f; c[2][1] = -(1 - Kr - Kb) / (1 - Kr) / 2; c[2][2] = -Kb / (1 - Kr) / 2;
        } else {
          c[0][0] = 0.299f; c[0][1] = 0.587f; c[0][2] = 0.114f; c[1][0] = -0.168735f; c[1][1] = -0.331264f; c[1][2] = 0.5f;
          c[2][0] = 0.5f; c[2][1] = -0.418688f; c[2][2] = -0.081312f;
        }
        for (size_t i = 0; i < (size_t)im->w * im->h; i++) {
          float yv = (float)im->p[0][i], cbv = (float)(im->p[1][i] - half), crv = (float)(im->p[2][i] - half);
          yv = (yv - lro) * 1.1689f; cbv = cbv * 1.1429f; crv = crv * 1.1429f;
          const float r = (float)clip_f_u16(yv + cf[0] * crv, maxv), g = (float)clip_f_u16(yv + cf[1] * cbv + cf[2] * crv, maxv), b = (float)clip_f_u16(yv + cf[3] * cbv, maxv);
          im->p[0][i] = (uint16_t)clip_f_u16(r * c[0][0] + g * c[0][1] + b * c[0][2], maxv);
          im->p[1][i] = (uint16_t)clip_f_u16((r * c[1][0] + g * c[1][1] + b * c[1][2]) + half, maxv);
          im->p[2][i] = (uint16_t)clip_f_u16((r * c[2][0] + g * c[2][1] + b * c[2][2]) + half, maxv);
        }
        *full_range = 1;
      }

    }
    if (o[0] == 1 && o[1] != 0) {
      for (int c = 0; c < 4; c++) {
        if (!im->p[c]) continue;
        int pw = (c == 1 || c == 2) ? im->cw : im->w, ph = (c == 1 || c == 2) ? im->ch : im->h;
        uint16_t* n = plane_rot(im->p[c], pw, ph, o[1]);
        free(im->p[c]); im->p[c] = n;
      }
      if (o[1] != 180) { int t = im->w; im->w = im->h; im->h = t; t = im->cw; im->cw = im->ch; im->ch = t; }
      (void)np;
    } else if (o[0] == 2) {
      for (int c = 0; c < 4; c++) if (im->p[c]) plane_mirror(im->p[c], (c == 1 || c == 2) ? im->cw : im->w, (c == 1 || c == 2) ? im->ch : im->h, o[1]);
    } else if (o[0] == 3) {                                                    /* crop, pixelimage.cc:1433-1546 (even origin only) */
      int l = o[1], r = o[2], t = o[3], b = o[4];
      int sh = (im->chroma == 1 || im->chroma == 2) ? 1 : 0, sv = im->chroma == 1 ? 1 : 0;
      int nw = r - l + 1, nh = b - t + 1, ncw = im->chroma ? (nw + sh) >> sh : 0, nch = im->chroma ? (nh + sv) >> sv : 0;
      for (int c = 0; c < 4; c++) {
        if (!im->p[c]) continue;
        int isc = (c == 1 || c == 2);
        int pw = isc ? im->cw : im->w, ow = isc ? ncw : nw, oh = isc ? nch : nh, x0 = isc ? l >> sh : l, y0 = isc ? t >> sv : t;
        uint16_t* n = (uint16_t*)malloc((size_t)ow * oh * 2 + 2);
        for (int y = 0; y < oh; y++) memcpy(n + (size_t)y * ow, im->p[c] + (size_t)(y + y0) * pw + x0, (size_t)ow * 2);
        free(im->p[c]); im->p[c] = n;
      }
      im->w = nw; im->h = nh; im->cw = ncw; im->ch = nch;
    }
  }
  return 0;
}

/* nclx.cc:84-140 (get_Kr_Kb) */
static void co_kr_kb(int matrix, int primaries, float* pKr, float* pKb) {
  float Kr = 0.0f, Kb = 0.0f;
  if (matrix == 12 || matrix == 13) {
    float gx, gy, bx, by, rx, ry, wx, wy; int ok = 1;
    switch (primaries) {
      case 1: gx = 0.300f; gy = 0.600f; bx = 0.150f; by = 0.060f; rx = 0.640f; ry = 0.330f; wx = 0.3127f; wy = 0.3290f; break;
      case 4: gx = 0.21f; gy = 0.71f; bx = 0.14f; by = 0.08f; rx = 0.67f; ry = 0.33f; wx = 0.310f; wy = 0.316f; break;
      case 5: gx = 0.29f; gy = 0.60f; bx = 0.15f; by = 0.06f; rx = 0.64f; ry = 0.33f; wx = 0.3127f; wy = 0.3290f; break;
      case 6: case 7: gx = 0.310f; gy = 0.595f; bx = 0.155f; by = 0.070f; rx = 0.630f; ry = 0.340f; wx = 0.3127f; wy = 0.3290f; break;
      case 8: gx = 0.243f; gy = 0.692f; bx = 0.145f; by = 0.049f; rx = 0.681f; ry = 0.319f; wx = 0.310f; wy = 0.316f; break;
      case 9: gx = 0.170f; gy = 0.797f; bx = 0.131f; by = 0.046f; rx = 0.708f; ry = 0.292f; wx = 0.3127f; wy = 0.3290f; break;
      case 10: gx = 0.0f; gy = 1.0f; bx = 0.0f; by = 0.0f; rx = 1.0f; ry = 0.0f; wx = 0.333333f; wy = 0.33333f; break;
      case 11: gx = 0.265f; gy = 0.690f; bx = 0.150f; by = 0.060f; rx = 0.680f; ry = 0.320f; wx = 0.314f; wy = 0.351f; break;
      case 12: gx = 0.265f; gy = 0.690f; bx = 0.150f; by = 0.060f; rx = 0.680f; ry = 0.320f; wx = 0.3127f; wy = 0.3290f; break;
      case 22: gx = 0.295f; gy = 0.605f; bx = 0.155f; by = 0.077f; rx = 0.630f; ry = 0.340f; wx = 0.3127f; wy = 0.3290f; break;
      default: gx = gy = bx = by = rx = ry = wx = wy = 0.0f; ok = 0;
    }
    (void)ok;
    float zr = 1 - (rx + ry), zg = 1 - (gx + gy), zb = 1 - (bx + by), zw = 1 - (wx + wy);
    float denom = wy * (rx * (gy * zb - by * zg) + gx * (by * zr - ry * zb) + bx * (ry * zg - gy * zr));
    if (denom != 0.0f) {
      Kr = (ry * (wx * (gy * zb - by * zg) + wy * (bx * zg - gx * zb) + zw * (gx * by - bx * gy))) / denom;
      Kb = (by * (wx * (ry * zg - gy * zr) + wy * (gx * zr - rx * zg) + zw * (rx * gy - gx * ry))) / denom;
    }
  } else switch (matrix) {
    case 1: Kr = 0.2126f; Kb = 0.0722f; break;
    case 4: Kr = 0.30f; Kb = 0.11f; break;
    case 5: case 6: Kr = 0.299f; Kb = 0.114f; break;
    case 7: Kr = 0.212f; Kb = 0.087f; break;
    case 9: case 10: Kr = 0.2627f; Kb = 0.0593f; break;
    default: break;
  }
  *pKr = Kr; *pKb = Kb;
}

/* nclx.cc:143-173 */
void co_coefficients(int matrix, int primaries, float out[4]) {
  float Kr, Kb; co_kr_kb(matrix, primaries, &Kr, &Kb);
  if (Kb != 0 || Kr != 0) {
    out[0] = 2 * (-Kr + 1); out[1] = 2 * Kb * (-Kb + 1) / (Kb + Kr - 1);
    out[2] = 2 * Kr * (-Kr + 1) / (Kb + Kr - 1); out[3] = 2 * (-Kb + 1);
  } else { out[0] = 1.402f; out[1] = -0.344136f; out[2] = -0.714136f; out[3] = 1.772f; }
}

static int clip_f_u16(float fx, int maxi) { int x = (int)(fx + 0.5f); return x < 0 ? 0 : (x > maxi ? maxi : x); }   /* common_utils.h:108-114 */
static int clip_int_u8(int x) { return x < 0 ? 0 : (x > 255 ? 255 : x); }

/* chroma_sampling.cc:623-700 restated loop for loop, including the border loops that index the source with cx/2, cy/2
   (golden table tests/conversion.cc:697-724).  in: (w+1)/2 x (h+1)/2, out: w x h. */
void co_bilinear_420_to_444(const uint16_t* in, int w, int h, uint16_t* out) {
  const int cs = (w + 1) / 2, os = w;
  out[0] = in[0];
  for (int cx = 0; cx < (w - 1) / 2; cx++) {
    out[2 * cx + 1] = (uint16_t)((3 * in[cx / 2] + 1 * in[cx / 2 + 1] + 2) / 4);
    out[2 * cx + 2] = (uint16_t)((1 * in[cx / 2] + 3 * in[cx / 2 + 1] + 2) / 4);
  }
  if (w % 2 == 0) out[w - 1] = in[w / 2 - 1];
  for (int cy = 0; cy < (h - 1) / 2; cy++) {
    out[(2 * cy + 1) * os] = (uint16_t)((3 * in[cy / 2 * cs] + 1 * in[(cy / 2 + 1) * cs] + 2) / 4);
    out[(2 * cy + 2) * os] = (uint16_t)((1 * in[cy / 2 * cs] + 3 * in[(cy / 2 + 1) * cs] + 2) / 4);
  }
  if (h % 2 == 0) out[(h - 1) * os] = in[(h / 2 - 1) * cs];
  if (w % 2 == 0) for (int cy = 0; cy < (h - 1) / 2; cy++) {
    out[(2 * cy + 1) * os + w - 1] = (uint16_t)((3 * in[cy / 2 * cs + w / 2 - 1] + 1 * in[(cy / 2 + 1) * cs + w / 2 - 1] + 2) / 4);
    out[(2 * cy + 2) * os + w - 1] = (uint16_t)((1 * in[cy / 2 * cs + w / 2 - 1] + 3 * in[(cy / 2 + 1) * cs + w / 2 - 1] + 2) / 4);
  }
  if (h % 2 == 0) for (int cx = 0; cx < (w - 1) / 2; cx++) {
    out[(h - 1) * os + 2 * cx + 1] = (uint16_t)((3 * in[(h / 2 - 1) * cs + cx / 2] + 1 * in[(h / 2 - 1) * cs + cx / 2 + 1] + 2) / 4);
    out[(h - 1) * os + 2 * cx + 2] = (uint16_t)((1 * in[(h / 2 - 1) * cs + cx / 2] + 3 * in[(h / 2 - 1) * cs + cx / 2 + 1] + 2) / 4);
  }
  if (w % 2 == 0 && h % 2 == 0) out[(h - 1) * os + w - 1] = in[(h / 2 - 1) * cs + w / 2 - 1];
  for (int y = 1; y < h - 1; y += 2) for (int x = 1; x < w - 1; x += 2) {
    int cx = x / 2, cy = y / 2;
    int c00 = in[cy * cs + cx], c01 = in[cy * cs + cx + 1], c10 = in[(cy + 1) * cs + cx], c11 = in[(cy + 1) * cs + cx + 1];
    out[y * os + x] = (uint16_t)((c00 * 9 + c01 * 3 + c10 * 3 + c11 + 8) / 16);
    out[y * os + x + 1] = (uint16_t)((c00 * 3 + c01 * 9 + c10 + c11 * 3 + 8) / 16);
    out[(y + 1) * os + x] = (uint16_t)((c00 * 3 + c01 + c10 * 9 + c11 * 3 + 8) / 16);
    out[(y + 1) * os + x + 1] = (uint16_t)((c00 + c01 * 3 + c10 * 3 + c11 * 9 + 8) / 16);
  }
}

/*
 * The whole post-stage on tightly packed uint16 planes.
 *   ops: nops x 5 ints {kind (1 rotate_ccw, 2 mirror, 3 crop), a, b, c, d}
 *   out_chroma: 10 RGB, 11 RGBA, 12/13 RRGGBB(AA)_BE, 14/15 RRGGBB(AA)_LE, 3 planar RGB (uint8 or uint16 by depth)
 * Returns bytes written to out (rows tightly packed; planar: R,G,B planes one after the other) or <0.
 */
long co_postprocess2(const uint16_t* y, const uint16_t* cb, const uint16_t* cr, const uint16_t* alpha, int w, int h, int chroma,
                     int bpp, int cp, int mc, int full_range, const int* ops, int nops, int out_chroma, int bilinear,
                     uint8_t* out, int* out_w, int* out_h) {
  co_image im; memset(&im, 0, sizeof im);
  int sh = (chroma == 1 || chroma == 2) ? 1 : 0, sv = chroma == 1 ? 1 : 0;
  im.w = w; im.h = h; im.chroma = chroma; im.cw = chroma ? (w + sh) >> sh : 0; im.ch = chroma ? (h + sv) >> sv : 0;
  const uint16_t* src[4] = {y, cb, cr, alpha};
  for (int c = 0; c < 4; c++) {
    if (!src[c]) continue;
    size_t n = (size_t)((c == 1 || c == 2) ? im.cw * im.ch : w * h);
    im.p[c] = (uint16_t*)malloc(n * 2 + 2); memcpy(im.p[c], src[c], n * 2);
  }
  if (co_geometry(&im, ops, nops, &full_range, bpp, cp, mc)) { co_free(&im); return -2; }
  w = im.w; h = im.h;
  if (im.chroma != chroma) { chroma = im.chroma; sh = (chroma == 1 || chroma == 2) ? 1 : 0; sv = chroma == 1 ? 1 : 0; }   /* 4:4:4 detour taken */
  if (bilinear && im.chroma == 1) {          /* only_use_preferred_chroma_algorithm: Op_YCbCr420_bilinear_to_YCbCr444 first, then the generic float op */
    for (int c = 1; c <= 2; c++) { uint16_t* up = (uint16_t*)malloc((size_t)w * h * 2 + 2); co_bilinear_420_to_444(im.p[c], w, h, up); free(im.p[c]); im.p[c] = up; }
    im.chroma = 3; im.cw = w; im.ch = h; chroma = 3; sh = 0; sv = 0;
  }
  *out_w = w; *out_h = h;
  float cf[4]; co_coefficients(mc, cp, cf);
  const int interleaved8 = out_chroma == 10 || out_chroma == 11;
  /* >8-bit full-range 4:2:0 to an 8-bit interleaved target: the reference planner runs Op_to_sdr_planes on the
     YCbCr planes FIRST and then the integer op (observed with ref_postprocess); every other >8-bit case
     converts in float at full depth and shifts afterwards. */
  /* matrix_coefficients 0 (GBR), 8 (YCgCo), 16 (YCgCo-Re): special branches of the GENERIC op only (yuv2rgb.cc:225-262).
     The dedicated 4:2:0 ops have no such branch: 0 and 8 are excluded from them (so the generic op always runs), 16 is
     not -- there they convert with the default coefficients as if the matrix were unspecified (observed with ref_postprocess). */
  const int special_matrix = (mc == 0 || mc == 8 || mc == 16) && chroma != 0;
  const int dedicated_ok = !(mc == 0 || mc == 8);                 /* Op_YCbCr420_to_RGB24/32, Op_YCbCr420_to_RRGGBBaa */
  const int pre_shift = (chroma == 1 && bpp > 8 && full_range && interleaved8 && dedicated_ok) ? bpp - 8 : 0;
  if (pre_shift) {
    for (int c = 0; c < 4; c++) if (im.p[c]) {
      size_t n = (size_t)((c == 1 || c == 2) ? im.cw * im.ch : im.w * im.h);
      for (size_t i = 0; i < n; i++) im.p[c][i] >>= pre_shift;
    }
    bpp = 8;
  }
  const int int_mode = chroma == 1 && bpp == 8 && full_range && interleaved8 && dedicated_ok;     /* yuv2rgb.cc:300-340 */
  const int rrggbb_direct = chroma == 1 && bpp > 8 && (out_chroma >= 12 && out_chroma <= 15) && dedicated_ok;   /* Op_YCbCr420_to_RRGGBBaa */
  const int special = special_matrix && !int_mode && !rrggbb_direct;
  const int ci[4] = {(int)lroundf(256 * cf[0]), (int)lroundf(256 * cf[1]), (int)lroundf(256 * cf[2]), (int)lroundf(256 * cf[3])};
  const int half = 1 << (bpp - 1), maxv = (1 << bpp) - 1;
  const float lro = (float)(16 << (bpp - 8));
  const int sdr_shift = (bpp > 8 && interleaved8) ? bpp - 8 : 0;
  const int want_alpha = out_chroma == 11 || out_chroma == 13 || out_chroma == 15;
  const int nch = want_alpha ? 4 : 3;
  const int le = out_chroma == 14 || out_chroma == 15;
  size_t pos = 0;
  const int out16 = (out_chroma >= 12 && out_chroma <= 15) || (out_chroma == 3 && bpp > 8);
  for (int yy = 0; yy < h; yy++) for (int xx = 0; xx < w; xx++) {
    int Y = im.p[0][yy * w + xx], r, g, b;
    if (!chroma) r = g = b = Y;
    else {
      int Cb = im.p[1][(yy >> sv) * im.cw + (xx >> sh)], Cr = im.p[2][(yy >> sv) * im.cw + (xx >> sh)];
      if (special) {
        if (mc == 0) {
          if (full_range) { r = Cr; g = Y; b = Cb; }
          else { r = clip_f_u16(((float)Cr - lro) * 1.1429f, maxv); g = clip_f_u16(((float)Y - lro) * 1.1689f, maxv); b = clip_f_u16(((float)Cb - lro) * 1.1429f, maxv); }
        } else if (mc == 8) {                     /* clip_int_u8 also for > 8 bit: reference quirk (yuv2rgb.cc:240-242) */
          int cbv = Cb - half, crv = Cr - half;
          r = clip_int_u8(Y - cbv + crv); g = clip_int_u8(Y + cbv); b = clip_int_u8(Y - cbv - crv);
        } else {                                  /* 16: YCgCo-Re, int16 arithmetic, x4 */
          int16_t yy = (int16_t)Y, cbv = (int16_t)((int16_t)Cb - (int16_t)half), crv = (int16_t)((int16_t)Cr - (int16_t)half);
          int16_t t = (int16_t)(yy - (cbv >> 1)), gg = (int16_t)(t + cbv), bb = (int16_t)(t - (crv >> 1)), rr = (int16_t)(bb + crv);
          int rv = rr * 4, gv = gg * 4, bv = bb * 4;
          r = rv < 0 ? 0 : (rv > maxv ? maxv : rv); g = gv < 0 ? 0 : (gv > maxv ? maxv : gv); b = bv < 0 ? 0 : (bv > maxv ? maxv : bv);
        }
        r >>= sdr_shift; g >>= sdr_shift; b >>= sdr_shift;
      } else if (int_mode) {
        int cbv = Cb - 128, crv = Cr - 128;
        r = clip_int_u8(Y + ((ci[0] * crv + 128) >> 8));
        g = clip_int_u8(Y + ((ci[1] * cbv + ci[2] * crv + 128) >> 8));
        b = clip_int_u8(Y + ((ci[3] * cbv + 128) >> 8));
      } else {
        float yv = (float)Y, cbv = (float)(Cb - half), crv = (float)(Cr - half);
        if (!full_range) { yv = (yv - lro) * 1.1689f; cbv = cbv * 1.1429f; crv = crv * 1.1429f; }
        r = clip_f_u16(yv + cf[0] * crv, maxv);
        g = clip_f_u16(yv + cf[1] * cbv + cf[2] * crv, maxv);
        b = clip_f_u16(yv + cf[3] * cbv, maxv);
        r >>= sdr_shift; g >>= sdr_shift; b >>= sdr_shift;
      }
    }
    int a = im.p[3] ? im.p[3][yy * w + xx] >> sdr_shift : (out16 ? maxv : 255);
    if (out_chroma == 3) {
      size_t pl = (size_t)w * h;
      if (out16) { ((uint16_t*)out)[yy * w + xx] = (uint16_t)r; ((uint16_t*)out)[pl + yy * w + xx] = (uint16_t)g; ((uint16_t*)out)[2 * pl + yy * w + xx] = (uint16_t)b; }
      else { out[yy * w + xx] = (uint8_t)r; out[pl + yy * w + xx] = (uint8_t)g; out[2 * pl + yy * w + xx] = (uint8_t)b; }
    } else if (!out16) {
      out[pos++] = (uint8_t)r; out[pos++] = (uint8_t)g; out[pos++] = (uint8_t)b; if (want_alpha) out[pos++] = (uint8_t)a;
    } else {
      int v[4] = {r, g, b, a};
      for (int c = 0; c < nch; c++) { out[pos + (le ? 1 : 0)] = (uint8_t)(v[c] >> 8); out[pos + (le ? 0 : 1)] = (uint8_t)(v[c] & 0xff); pos += 2; }
    }
  }
  co_free(&im);
  if (out_chroma == 3) return (long)((size_t)w * h * 3 * (out16 ? 2 : 1));
  return (long)pos;
}

long co_postprocess(const uint16_t* y, const uint16_t* cb, const uint16_t* cr, const uint16_t* alpha, int w, int h, int chroma,
                    int bpp, int cp, int mc, int full_range, const int* ops, int nops, int out_chroma,
                    uint8_t* out, int* out_w, int* out_h) {
  return co_postprocess2(y, cb, cr, alpha, w, h, chroma, bpp, cp, mc, full_range, ops, nops, out_chroma, 0, out, out_w, out_h);
}

/* ------------------------------------------------------------------------------------------------------------------
 * a11: overlay compositing, restated from HeifPixelImage::fill_RGB_16bit (libheif/image/pixelimage.cc:1549-1621) and
 * HeifPixelImage::overlay (pixelimage.cc:1637-1780), including the loop bounds of the original: after clipping against
 * the left / top canvas border the copy loops run `for (y = in_y0; y < in_h - in_y0; ...)` and, on the alpha path,
 * `for (x = in_x0; x < in_w - in_x0; ...)` with x added to BOTH the source offset in_x0 and the destination offset.
 * Planes are packed 8-bit, w bytes per row.
 * ------------------------------------------------------------------------------------------------------------------ */
void co_fill_rgb16(uint8_t* canvas /* 3 planes */, int cw, int ch, const uint16_t bkg[4]) {
  for (int c = 0; c < 3; c++) memset(canvas + (size_t)c * cw * ch, (uint8_t)(bkg[c] >> 8), (size_t)cw * ch);
}

static uint32_t co_negate(int32_t x) { return x == INT32_MIN ? (uint32_t)INT32_MAX + 1u : (uint32_t)(-x); }

void co_overlay(uint8_t* canvas /* R,G,B */, int cw, int ch, const uint8_t* ov /* R,G,B[,A] */, int ow_, int oh_, int has_alpha, int32_t dx, int32_t dy) {
  const uint8_t* alpha_p = has_alpha ? ov + (size_t)3 * ow_ * oh_ : NULL;
  for (int c = 0; c < 3; c++) {                      /* the canvas of decode_overlay_image has no alpha plane (overlay.cc:312-322) */
    const uint8_t* in_p = ov + (size_t)c * ow_ * oh_;
    uint8_t* out_p = canvas + (size_t)c * cw * ch;
    uint32_t in_w = (uint32_t)ow_, in_h = (uint32_t)oh_;
    const uint32_t out_w = (uint32_t)cw, out_h = (uint32_t)ch;
    const size_t in_stride = (size_t)ow_, out_stride = (size_t)cw, alpha_stride = (size_t)ow_;
    if (dx > 0 && (uint32_t)dx >= out_w) return;
    else if (dx < 0 && in_w <= co_negate(dx)) return;
    if (dy > 0 && (uint32_t)dy >= out_h) return;
    else if (dy < 0 && in_h <= co_negate(dy)) return;
    uint32_t in_x0, in_y0, out_x0, out_y0;
    if (dx + (int64_t)in_w > out_w) in_w = (uint32_t)((int64_t)out_w - dx);
    if (dy + (int64_t)in_h > out_h) in_h = (uint32_t)((int64_t)out_h - dy);
    if (dx < 0) { in_x0 = co_negate(dx); out_x0 = 0; in_w = in_w - in_x0; } else { in_x0 = 0; out_x0 = (uint32_t)dx; }
    if (dy < 0) { in_y0 = co_negate(dy); out_y0 = 0; in_h = in_h - in_y0; } else { in_y0 = 0; out_y0 = (uint32_t)dy; }
    for (uint32_t y = in_y0; y < in_h; y++) {
      if (!has_alpha) memcpy(out_p + out_x0 + (out_y0 + y - in_y0) * out_stride, in_p + in_x0 + y * in_stride, in_w);
      else for (uint32_t x = in_x0; x < in_w; x++) {
        uint8_t* outptr = &out_p[out_x0 + (out_y0 + y - in_y0) * out_stride + x];
        const uint8_t in_val = in_p[in_x0 + y * in_stride + x];
        const uint8_t alpha_val = alpha_p[in_x0 + y * alpha_stride + x];
        *outptr = (uint8_t)((in_val * alpha_val + *outptr * (255 - alpha_val)) / 255);
      }
    }
  }
}

/* a12: HeifPixelImage::scale_nearest_neighbor (pixelimage.cc:1783-1972) for ONE plane: the source index is derived from
 * the IMAGE sizes (m_width / width), also for subsampled chroma planes; `comps` interleaved samples of `bps` bytes. */
void co_scale_nearest_plane(const uint8_t* in, size_t in_stride, uint8_t* out, size_t out_stride, uint32_t out_w, uint32_t out_h,
                            uint32_t img_w_in, uint32_t img_h_in, uint32_t img_w_out, uint32_t img_h_out, int comps, int bps) {
  for (uint32_t y = 0; y < out_h; y++) {
    const uint32_t iy = (uint32_t)((uint64_t)y * img_h_in / img_h_out);
    for (uint32_t x = 0; x < out_w; x++) {
      const uint32_t ix = (uint32_t)((uint64_t)x * img_w_in / img_w_out);
      memcpy(out + y * out_stride + (size_t)x * comps * bps, in + iy * in_stride + (size_t)ix * comps * bps, (size_t)comps * bps);
    }
  }
}


/* ---- encoder side of the colour stage: Op_RGB24_32_to_YCbCr (color-conversion/rgb2yuv.cc:575-808), the op
   heif_context_encode_image() runs on interleaved 8-bit RGB / RGBA input before the encoder plugin sees it.  Restated loop for
   loop: float products in the reference's order, clip_f_u8 / clip_f_u16 rounding (common_utils.h:108-122), 4:2:0 averaging with
   its odd-width / odd-height border loops, left-aligned 4:2:2.  out_chroma: 1 = 4:2:0, 2 = 4:2:2, 3 = 4:4:4. */
static int co_clip_f_u8(float fx) { int x = (int)(fx + 0.5f); return x < 0 ? 0 : (x > 255 ? 255 : x); }
static void co_set_chroma(uint8_t* ocb, uint8_t* ocr, int r, int g, int b, float c[3][3], int full) {   /* rgb2yuv.cc:556-572 */
  float cb = r * c[1][0] + g * c[1][1] + b * c[1][2];
  float cr = r * c[2][0] + g * c[2][1] + b * c[2][2];
  if (full) { *ocb = (uint8_t)co_clip_f_u8(cb + 128); *ocr = (uint8_t)co_clip_f_u8(cr + 128); }
  else { *ocb = (uint8_t)co_clip_f_u8(cb * 0.875f + 128.0f); *ocr = (uint8_t)co_clip_f_u8(cr * 0.875f + 128.0f); }
}
int co_rgb_to_ycbcr(const uint8_t* in, size_t in_stride, int w, int h, int has_alpha, int out_chroma, int mc, int cp, int full,
                    uint8_t* oy, uint8_t* ocb, uint8_t* ocr, uint8_t* oa) {
  const int bpp = has_alpha ? 4 : 3, sh = (out_chroma == 1 || out_chroma == 2) ? 2 : 1, sv = out_chroma == 1 ? 2 : 1;
  const int cw = (w + sh - 1) / sh;
  if (mc == 2) mc = 6;                                       /* convert_colorspace() works on a copy of the target profile with the */
  if (cp == 2) cp = 1;                                       /* unspecified values replaced (nclx.cc:360-373, colorconversion.cc:513-515) */
  float Kr, Kb; co_kr_kb(mc, cp, &Kr, &Kb);
  float c[3][3];
  if (Kb != 0 || Kr != 0) {                                  /* nclx.cc:176-200 */
    c[0][0] = Kr; c[0][1] = 1 - Kr - Kb; c[0][2] = Kb;
    c[1][0] = -Kr / (1 - Kb) / 2; c[1][1] = -(1 - Kr - Kb) / (1 - Kb) / 2; c[1][2] = 0.5f;
    c[2][0] = 0.5f; c[2][1] = -(1 - Kr - Kb) / (1 - Kr) / 2; c[2][2] = -Kb / (1 - Kr) / 2;
  } else {
    c[0][0] = 0.299f; c[0][1] = 0.587f; c[0][2] = 0.114f; c[1][0] = -0.168735f; c[1][1] = -0.331264f; c[1][2] = 0.5f;
    c[2][0] = 0.5f; c[2][1] = -0.418688f; c[2][2] = -0.081312f;
  }
  for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
    const uint8_t* p = in + (size_t)y * in_stride + (size_t)x * bpp;
    float yv = p[0] * c[0][0] + p[1] * c[0][1] + p[2] * c[0][2];
    oy[(size_t)y * w + x] = full ? (uint8_t)co_clip_f_u8(yv) : (uint8_t)(clip_f_u16(yv * 0.85547f, 219) + 16);
    if (oa) oa[(size_t)y * w + x] = has_alpha ? p[3] : 0xff;
  }
  if (sh == 1 && sv == 1) {
    for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) { const uint8_t* p = in + (size_t)y * in_stride + (size_t)x * bpp; co_set_chroma(ocb + (size_t)y * cw + x, ocr + (size_t)y * cw + x, p[0], p[1], p[2], c, full); }
  } else if (sh == 2 && sv == 2) {
    for (int y = 0; y < (h & ~1); y += 2) for (int x = 0; x < (w & ~1); x += 2) {
      const uint8_t* p = in + (size_t)y * in_stride + (size_t)x * bpp;
      int r = (uint8_t)((p[0] + p[bpp + 0] + p[in_stride + 0] + p[bpp + in_stride + 0]) / 4);
      int g = (uint8_t)((p[1] + p[bpp + 1] + p[in_stride + 1] + p[bpp + in_stride + 1]) / 4);
      int b = (uint8_t)((p[2] + p[bpp + 2] + p[in_stride + 2] + p[bpp + in_stride + 2]) / 4);
      co_set_chroma(ocb + (size_t)(y / 2) * cw + x / 2, ocr + (size_t)(y / 2) * cw + x / 2, r, g, b, c, full);
    }
    if (w & 1) {                                              /* right column */
      const int x = w - 1;
      for (int y = 0; y < h; y += 2) {
        const uint8_t* p = in + (size_t)y * in_stride + (size_t)x * bpp;
        int r, g, b;
        if (y + 1 < h) { r = (uint8_t)((p[0] + p[in_stride + 0]) / 2); g = (uint8_t)((p[1] + p[in_stride + 1]) / 2); b = (uint8_t)((p[2] + p[in_stride + 2]) / 2); }
        else { r = p[0]; g = p[1]; b = p[2]; }
        co_set_chroma(ocb + (size_t)(y / 2) * cw + x / 2, ocr + (size_t)(y / 2) * cw + x / 2, r, g, b, c, full);
      }
    }
    if (h & 1) {                                              /* bottom row */
      const int y = h - 1;
      for (int x = 0; x < w; x += 2) {
        const uint8_t* p = in + (size_t)y * in_stride + (size_t)x * bpp;
        int r, g, b;
        if (x + 1 < w) { r = (uint8_t)((p[0] + p[bpp + 0]) / 2); g = (uint8_t)((p[1] + p[bpp + 1]) / 2); b = (uint8_t)((p[2] + p[bpp + 2]) / 2); }
        else { r = p[0]; g = p[1]; b = p[2]; }
        co_set_chroma(ocb + (size_t)(y / 2) * cw + x / 2, ocr + (size_t)(y / 2) * cw + x / 2, r, g, b, c, full);
      }
    }
  } else {                                                    /* 4:2:2, left-aligned chroma (rgb2yuv.cc:760-787) */
    for (int y = 0; y < h; y++) for (int x = 0; x < w; x += 2) {
      const uint8_t* p = in + (size_t)y * in_stride + (size_t)x * bpp;
      co_set_chroma(ocb + (size_t)y * cw + x / 2, ocr + (size_t)y * cw + x / 2, p[0], p[1], p[2], c, full);
    }
  }
  return 0;
}
