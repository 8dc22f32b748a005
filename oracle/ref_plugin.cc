/*
 * oracle/ref_plugin.cc -- TEST INFRASTRUCTURE ONLY. Never linked, imported or executed by the product.
 *
 * CPU reference decoder plugin for the reference libheif built by oracle/Makefile
 * (oracle/_ref/libheif_ref.so): the libde265 role (libheif/plugins/decoder_libde265.cc) filled by
 * FFmpeg (oracle/ffhevc.c) -- or, selectable, by the C restatement (oracle/hevc_oracle.c) -- so that
 * heif_decode_image() of the UNMODIFIED reference runs end to end on this machine.  Member order of the
 * plugin table follows decoder_libde265.cc:497-517; plane hand-over follows :97-171; nclx from VUI :426-448.
 *
 * Env: B200_ORACLE_BACKEND=ffmpeg|restatement (default ffmpeg), B200_ORACLE_DUMP_DIR=<dir> dumps every
 * pushed access unit (used once to harvest the HEVC streams of the reference's fixtures).
 */
#include <libheif/heif.h>
#include <libheif/heif_plugin.h>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

extern "C" {
typedef struct { int width, height, cw, ch, bit_depth, chroma, full_range_name; uint16_t* plane[3]; } ffhevc_picture;
int ffhevc_init(const char* dir);
int ffhevc_decode(const uint8_t* data, size_t size, int threads, ffhevc_picture* out);
void ffhevc_free_picture(ffhevc_picture* p);
typedef struct { int width, height, cw, ch, bit_depth, chroma_format, vui_colour_present, colour_primaries,
  transfer_characteristics, matrix_coeffs, full_range, video_signal_present; uint16_t* plane[3]; } hevc_oracle_picture;
int hevc_oracle_decode(const uint8_t* data, size_t size, int stage, hevc_oracle_picture* out);
void hevc_oracle_free_picture(hevc_oracle_picture* p);
int hevc_oracle_parse_vui(const uint8_t* data, size_t size, int out[4]);
}

namespace {
struct Dec { std::vector<uint8_t> data; int strict = 0; int threads = 1; };
std::atomic<int> g_dump_counter{0};
const char kOk[] = "Success";

heif_error ok() { return heif_error{heif_error_Ok, heif_suberror_Unspecified, kOk}; }
heif_error fail(const char* m) { return heif_error{heif_error_Decoder_plugin_error, heif_suberror_Unspecified, m}; }

const char* name() { return "b200 test oracle (FFmpeg libavcodec / C restatement)"; }
void init() {} void deinit() {}
int supports(heif_compression_format f) { return f == heif_compression_HEVC ? 500 : 0; }
int supports2(const heif_decoder_plugin_compressed_format_description* d) { return supports(d->format); }
heif_error new_dec2(void** out, const heif_decoder_plugin_options* o) {
  Dec* d = new Dec; d->strict = o->strict_decoding; d->threads = o->num_threads > 0 ? o->num_threads : 1; *out = d; return ok();
}
heif_error new_dec(void** out) { heif_decoder_plugin_options o{}; o.format = heif_compression_HEVC; return new_dec2(out, &o); }
void free_dec(void* p) { delete (Dec*)p; }
heif_error push2(void* p, const void* data, size_t n, uintptr_t) { Dec* d = (Dec*)p; d->data.insert(d->data.end(), (const uint8_t*)data, (const uint8_t*)data + n); return ok(); }
heif_error push(void* p, const void* data, size_t n) { return push2(p, data, n, 0); }
heif_error flush(void*) { return ok(); }
void set_strict(void* p, int f) { ((Dec*)p)->strict = f; }

heif_error decode2(void* p, heif_image** out, uintptr_t* ud, const heif_security_limits* limits) {
  Dec* d = (Dec*)p;
  *out = nullptr;
  if (ud) *ud = 0;
  if (d->data.empty()) return ok();
  if (const char* dir = getenv("B200_ORACLE_DUMP_DIR")) {
    char path[4096]; snprintf(path, sizeof path, "%s/au_%04d.bin", dir, g_dump_counter++);
    if (FILE* f = fopen(path, "wb")) { fwrite(d->data.data(), 1, d->data.size(), f); fclose(f); }
  }
  const char* be = getenv("B200_ORACLE_BACKEND");
  int w, h, cw, ch, bd, chroma; uint16_t* planes[3];
  ffhevc_picture fp{}; hevc_oracle_picture op{};
  bool restate = be && strcmp(be, "restatement") == 0;
  if (restate) {
    if (hevc_oracle_decode(d->data.data(), d->data.size(), 0, &op) != 0) { d->data.clear(); return fail("hevc_oracle_decode failed"); }
    w = op.width; h = op.height; cw = op.cw; ch = op.ch; bd = op.bit_depth; chroma = op.chroma_format;
    for (int i = 0; i < 3; i++) planes[i] = op.plane[i];
  } else {
    if (ffhevc_decode(d->data.data(), d->data.size(), d->threads, &fp) != 0) { d->data.clear(); return fail("ffhevc_decode failed"); }
    w = fp.width; h = fp.height; cw = fp.cw; ch = fp.ch; bd = fp.bit_depth; chroma = fp.chroma;
    for (int i = 0; i < 3; i++) planes[i] = fp.plane[i];
  }
  int vui[4]; hevc_oracle_parse_vui(d->data.data(), d->data.size(), vui);
  d->data.clear();
  heif_image* img = nullptr;
  heif_error err = heif_image_create(w, h, chroma == 0 ? heif_colorspace_monochrome : heif_colorspace_YCbCr, (heif_chroma)chroma, &img);
  if (err.code) return err;
  const heif_channel chans[3] = {heif_channel_Y, heif_channel_Cb, heif_channel_Cr};
  for (int c = 0; c < (chroma ? 3 : 1) && !err.code; c++) {
    int pw = c ? cw : w, ph = c ? ch : h;
    err = heif_image_add_plane_safe(img, chans[c], pw, ph, bd, limits);
    if (err.code) break;
    size_t stride; uint8_t* dst = heif_image_get_plane2(img, chans[c], &stride);
    for (int y = 0; y < ph; y++) {
      const uint16_t* src = planes[c] + (size_t)y * pw;
      if (bd == 8) for (int x = 0; x < pw; x++) dst[y * stride + x] = (uint8_t)src[x];
      else memcpy(dst + y * stride, src, (size_t)pw * 2);
    }
  }
  if (restate) hevc_oracle_free_picture(&op); else ffhevc_free_picture(&fp);
  if (err.code) { heif_image_release(img); return err; }
  heif_color_profile_nclx* nclx = heif_nclx_color_profile_alloc();
  heif_nclx_color_profile_set_color_primaries(nclx, (uint16_t)vui[0]);
  heif_nclx_color_profile_set_transfer_characteristics(nclx, (uint16_t)vui[1]);
  heif_nclx_color_profile_set_matrix_coefficients(nclx, (uint16_t)vui[2]);
  nclx->full_range_flag = (uint8_t)vui[3];
  heif_image_set_nclx_color_profile(img, nclx);
  heif_nclx_color_profile_free(nclx);
  *out = img;
  return ok();
}
heif_error decode_next(void* p, heif_image** out, const heif_security_limits* l) { return decode2(p, out, nullptr, l); }
heif_error decode_img(void* p, heif_image** out) { return decode2(p, out, nullptr, nullptr); }

const heif_decoder_plugin g_plugin = {5, name, init, deinit, supports, new_dec, free_dec, push, decode_img, set_strict,
                                      "b200-oracle", decode_next, LIBHEIF_MAKE_VERSION(1, 21, 0), supports2, new_dec2, push2, flush, decode2};
}  // namespace

extern "C" int b200_oracle_register(const char* avcodec_dir) {
  if (avcodec_dir && ffhevc_init(avcodec_dir) != 0) return -1;
  heif_error e = heif_register_decoder_plugin(&g_plugin);
  return e.code;
}

// ------------------------------------------------------------------------------------------------
// Direct access to the UNMODIFIED reference post-stage (internal C++ API of libheif_ref.so) so that the
// C restatement (color_oracle.c) and the CUDA path can be pinned on arbitrary synthetic planes:
// HeifPixelImage::rotate_ccw / mirror_inplace / crop (libheif/image/pixelimage.cc:1175-1546) followed by
// convert_colorspace (libheif/color-conversion/colorconversion.cc:490-623) with the options
// heif_decode_image uses (api/libheif/heif_decoding.cc:78-81; target nclx = sRGB defaults, context.cc:1545-1558).
#include "image/pixelimage.h"
#include "color-conversion/colorconversion.h"
#include "security_limits.h"

extern "C" int ref_postprocess(const void* y, const void* cb, const void* cr, const void* alpha, int w, int h, int chroma, int bpp,
                               int has_nclx, int cp, int tc, int mc, int full_range, const int* ops, int nops,
                               int out_colorspace, int out_chroma, int only_preferred, int upsampling, int hdr_to_8bit,
                               uint8_t* out, size_t out_capacity, int* out_w, int* out_h, int* out_rowbytes, int* out_planes) {
  const heif_security_limits* limits = heif_get_global_security_limits();
  auto img = std::make_shared<HeifPixelImage>();
  img->create(w, h, chroma == 0 ? heif_colorspace_monochrome : heif_colorspace_YCbCr, (heif_chroma)chroma);
  const int sh = (chroma == 1 || chroma == 2) ? 1 : 0, sv = chroma == 1 ? 1 : 0;
  const int cw = (w + sh) >> sh, ch = (h + sv) >> sv, bps = bpp > 8 ? 2 : 1;
  struct { heif_channel c; const void* p; int w, h; } pl[4] = {{heif_channel_Y, y, w, h}, {heif_channel_Cb, cb, cw, ch}, {heif_channel_Cr, cr, cw, ch}, {heif_channel_Alpha, alpha, w, h}};
  for (auto& q : pl) {
    if (!q.p) continue;
    if (img->add_channel(q.c, q.w, q.h, bpp, limits)) return -1;
    size_t stride; uint8_t* dst = img->get_channel_memory(q.c, &stride);
    for (int r = 0; r < q.h; r++) memcpy(dst + r * stride, (const uint8_t*)q.p + (size_t)r * q.w * bps, (size_t)q.w * bps);
  }
  if (has_nclx) {
    nclx_profile p;
    p.set_colour_primaries((uint16_t)cp); p.set_transfer_characteristics((uint16_t)tc);
    p.set_matrix_coefficients((uint16_t)mc); p.set_full_range_flag(full_range != 0);
    img->set_color_profile_nclx(p);
  }
  for (int i = 0; i < nops; i++) {
    const int* o = ops + 5 * i;
    Result<std::shared_ptr<HeifPixelImage>> r = Error::Ok;
    if (o[0] == 1) r = img->rotate_ccw(o[1], limits);
    else if (o[0] == 2) r = img->mirror_inplace((heif_transform_mirror_direction)o[1], limits);
    else if (o[0] == 3) r = img->crop(o[1], o[2], o[3], o[4], limits);
    else return -2;
    if (!r) return -3;
    img = *r;
  }
  heif_color_conversion_options copt{};
  copt.version = 1;
  copt.preferred_chroma_downsampling_algorithm = heif_chroma_downsampling_average;
  copt.preferred_chroma_upsampling_algorithm = (heif_chroma_upsampling_algorithm)upsampling;
  copt.only_use_preferred_chroma_algorithm = (uint8_t)only_preferred;
  nclx_profile target; target.set_sRGB_defaults();
  auto res = convert_colorspace(img, (heif_colorspace)out_colorspace, (heif_chroma)out_chroma, target, hdr_to_8bit ? 8 : 0, copt, nullptr, limits);
  if (!res) return -4;
  auto o = *res;
  *out_w = o->get_width(); *out_h = o->get_height();
  size_t pos = 0;
  std::vector<heif_channel> chans;
  if (o->has_channel(heif_channel_interleaved)) chans = {heif_channel_interleaved};
  else { chans = {heif_channel_R, heif_channel_G, heif_channel_B}; if (o->has_channel(heif_channel_Alpha)) chans.push_back(heif_channel_Alpha); }
  *out_planes = (int)chans.size();
  for (heif_channel c : chans) {
    size_t stride; const uint8_t* src = o->get_channel_memory(c, &stride);
    int obpp = o->get_bits_per_pixel(c);
    size_t rowb;
    if (c == heif_channel_interleaved) {
      heif_chroma oc = o->get_chroma_format();
      int bytes = oc == heif_chroma_interleaved_RGB ? 3 : oc == heif_chroma_interleaved_RGBA ? 4 :
                  (oc == heif_chroma_interleaved_RRGGBB_BE || oc == heif_chroma_interleaved_RRGGBB_LE) ? 6 : 8;
      rowb = (size_t)*out_w * bytes;
    } else rowb = (size_t)*out_w * (obpp > 8 ? 2 : 1);
    *out_rowbytes = (int)rowb;
    if (pos + rowb * *out_h > out_capacity) return -5;
    for (int r = 0; r < *out_h; r++) { memcpy(out + pos, src + r * stride, rowb); pos += rowb; }
  }
  return 0;
}

// ---- encoder side: the reference's own conversion of interleaved 8-bit RGB / RGBA to YCbCr, as
// heif_context_encode_image() performs it before an encoder plugin sees the picture (convert_colorspace picks Op_RGB24_32_to_YCbCr).
// Packed planes out: Y (w x h), Cb, Cr (cw x ch), optional alpha.
extern "C" int ref_rgb_to_ycbcr(const uint8_t* rgb, int w, int h, int has_alpha, int out_chroma, int cp, int tc, int mc, int full_range,
                                uint8_t* oy, uint8_t* ocb, uint8_t* ocr, uint8_t* oa) {
  const heif_security_limits* limits = heif_get_global_security_limits();
  auto img = std::make_shared<HeifPixelImage>();
  img->create(w, h, heif_colorspace_RGB, has_alpha ? heif_chroma_interleaved_RGBA : heif_chroma_interleaved_RGB);
  if (img->add_channel(heif_channel_interleaved, w, h, 8, limits)) return -1;
  { size_t stride; uint8_t* dst = img->get_channel_memory(heif_channel_interleaved, &stride); const int bpp = has_alpha ? 4 : 3;
    for (int r = 0; r < h; r++) memcpy(dst + r * stride, rgb + (size_t)r * w * bpp, (size_t)w * bpp); }
  nclx_profile target;
  target.set_colour_primaries((uint16_t)cp); target.set_transfer_characteristics((uint16_t)tc);
  target.set_matrix_coefficients((uint16_t)mc); target.set_full_range_flag(full_range != 0);
  heif_color_conversion_options copt{};
  copt.version = 1;
  copt.preferred_chroma_downsampling_algorithm = heif_chroma_downsampling_average;
  copt.preferred_chroma_upsampling_algorithm = heif_chroma_upsampling_bilinear;
  copt.only_use_preferred_chroma_algorithm = 0;
  auto res = convert_colorspace(img, heif_colorspace_YCbCr, (heif_chroma)out_chroma, target, 8, copt, nullptr, limits);
  if (!res) return -4;
  auto o = *res;
  struct { heif_channel c; uint8_t* p; } pl[4] = {{heif_channel_Y, oy}, {heif_channel_Cb, ocb}, {heif_channel_Cr, ocr}, {heif_channel_Alpha, oa}};
  for (auto& q : pl) {
    if (!q.p) continue;
    if (!o->has_channel(q.c)) { if (q.c == heif_channel_Alpha) continue; return -5; }
    size_t stride; const uint8_t* src = o->get_channel_memory(q.c, &stride);
    const int pw = o->get_width(q.c), ph = o->get_height(q.c);
    for (int r = 0; r < ph; r++) memcpy(q.p + (size_t)r * pw, src + r * stride, (size_t)pw);
  }
  return 0;
}

// ---- a11 / a12: the reference's own overlay compositing and nearest-neighbour scaler, driven on caller-provided planes.
// Packed 8-bit planes R,G,B[,A] (row-major, w bytes per row) in, canvas planes R,G,B out.
extern "C" int ref_overlay(int cw, int ch, const uint16_t bkg[4], int noverlays, const uint8_t* const* ov_data, const int* ov_w, const int* ov_h,
                           const int* ov_alpha, const int* dx, const int* dy, uint8_t* out /* 3 * cw * ch */) {
  const heif_security_limits* limits = heif_get_global_security_limits();
  auto img = std::make_shared<HeifPixelImage>();
  img->create(cw, ch, heif_colorspace_RGB, heif_chroma_444);
  for (heif_channel c : {heif_channel_R, heif_channel_G, heif_channel_B}) if (img->add_channel(c, cw, ch, 8, limits)) return -1;
  if (img->fill_RGB_16bit(bkg[0], bkg[1], bkg[2], bkg[3])) return -2;
  for (int i = 0; i < noverlays; i++) {
    auto ov = std::make_shared<HeifPixelImage>();
    ov->create(ov_w[i], ov_h[i], heif_colorspace_RGB, heif_chroma_444);
    const heif_channel chans[4] = {heif_channel_R, heif_channel_G, heif_channel_B, heif_channel_Alpha};
    const int n = ov_alpha[i] ? 4 : 3;
    for (int k = 0; k < n; k++) {
      if (ov->add_channel(chans[k], ov_w[i], ov_h[i], 8, limits)) return -3;
      size_t stride; uint8_t* dst = ov->get_channel_memory(chans[k], &stride);
      const uint8_t* src = ov_data[i] + (size_t)k * ov_w[i] * ov_h[i];
      for (int r = 0; r < ov_h[i]; r++) memcpy(dst + r * stride, src + (size_t)r * ov_w[i], (size_t)ov_w[i]);
    }
    Error e = img->overlay(ov, dx[i], dy[i]);
    if (e && !(e.error_code == heif_error_Invalid_input && e.sub_error_code == heif_suberror_Overlay_image_outside_of_canvas)) return -4;
  }
  size_t pos = 0;
  for (heif_channel c : {heif_channel_R, heif_channel_G, heif_channel_B}) {
    size_t stride; const uint8_t* src = img->get_channel_memory(c, &stride);
    for (int r = 0; r < ch; r++) { memcpy(out + pos, src + r * stride, (size_t)cw); pos += (size_t)cw; }
  }
  return 0;
}

// planes: colorspace 0 = YCbCr (chroma 1/2/3), 1 = RGB planar 4:4:4, 2 = monochrome, 3 = interleaved (chroma = heif_chroma_interleaved_*);
// samples uint8 (bpp <= 8) or uint16.  `in` / `out` hold the planes packed one after the other (Y,Cb,Cr | R,G,B | Y | interleaved, then alpha).
extern "C" int ref_scale_nn(int colorspace, int chroma, int bpp, int has_alpha, int w, int h, int ow, int oh, const uint8_t* in, uint8_t* out, size_t out_capacity) {
  const heif_security_limits* limits = heif_get_global_security_limits();
  auto img = std::make_shared<HeifPixelImage>();
  const heif_colorspace cs = colorspace == 0 ? heif_colorspace_YCbCr : colorspace == 2 ? heif_colorspace_monochrome : heif_colorspace_RGB;
  img->create(w, h, cs, (heif_chroma)chroma);
  std::vector<heif_channel> chans;
  if (colorspace == 0) chans = {heif_channel_Y, heif_channel_Cb, heif_channel_Cr};
  else if (colorspace == 1) chans = {heif_channel_R, heif_channel_G, heif_channel_B};
  else if (colorspace == 2) chans = {heif_channel_Y};
  else chans = {heif_channel_interleaved};
  if (has_alpha && colorspace != 3) chans.push_back(heif_channel_Alpha);
  const int bps = bpp > 8 ? 2 : 1;
  size_t pos = 0;
  for (heif_channel c : chans) {
    uint32_t pw = w, ph = h;
    if (c == heif_channel_Cb || c == heif_channel_Cr) get_subsampled_size(w, h, c, (heif_chroma)chroma, &pw, &ph);
    if (img->add_channel(c, pw, ph, bpp, limits)) return -1;
    size_t stride; uint8_t* dst = img->get_channel_memory(c, &stride);
    const int comps = c == heif_channel_interleaved ? num_interleaved_components_per_plane((heif_chroma)chroma) : 1;
    const size_t rowb = (size_t)pw * comps * bps;
    for (uint32_t r = 0; r < ph; r++) { memcpy(dst + r * stride, in + pos, rowb); pos += rowb; }
  }
  std::shared_ptr<HeifPixelImage> o;
  if (img->scale_nearest_neighbor(o, ow, oh, limits)) return -2;
  pos = 0;
  for (heif_channel c : chans) {
    const uint32_t pw = o->get_width(c), ph = o->get_height(c);
    size_t stride; const uint8_t* src = o->get_channel_memory(c, &stride);
    const int comps = c == heif_channel_interleaved ? num_interleaved_components_per_plane((heif_chroma)chroma) : 1;
    const size_t rowb = (size_t)pw * comps * bps;
    if (pos + rowb * ph > out_capacity) return -3;
    for (uint32_t r = 0; r < ph; r++) { memcpy(out + pos, src + r * stride, rowb); pos += rowb; }
  }
  return (int)(pos > 0x7fffffff ? 0x7fffffff : pos);
}
