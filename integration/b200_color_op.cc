// integration/b200_color_op.cc -- SURVEY.md 8(f) N2: the fused GPU colour stage inside heif_decode_image().
//
// libheif keeps its colour-conversion operation pool private (libheif/color-conversion/colorconversion.cc:225-270; there is
// no registration API), so this file is compiled INTO a second build of the reference core
// (oracle/_ref/libheif_ref_b200.so, recipe in oracle/Makefile; the parity oracle libheif_ref.so stays unmodified) together
// with the three-line patch integration/colorconversion_b200.patch that calls b200_register_color_ops() from
// ColorConversionPipeline::init_ops().  It adds ONE ColorConversionOperation with SpeedCosts_Hardware
// (colorconversion.h:57-64): YCbCr 4:2:0 / 4:2:2 / 4:4:4, 8..16 bit, optional alpha plane -> interleaved RGB / RGBA /
// RRGGBB[AA]_{LE,BE} in one step, executed by libb200heif.so (b200_color_convert_host: H2D, the fused K6 kernel, D2H).
// The planner then prefers it over its own one- or two-operation chains; the bytes are the same (tests/test_n2_patch.py).
#include "color-conversion/colorconversion.h"
#include "image/pixelimage.h"
#include "../include/b200_heif.h"

namespace {

class Op_B200_YCbCr_to_interleaved : public ColorConversionOperation {
 public:
  std::vector<ColorStateWithCost> state_after_conversion(const ColorState& in, const ColorState& target, const heif_color_conversion_options& options,
                                                         const heif_color_conversion_options_ext&) const override {
    if (in.colorspace != heif_colorspace_YCbCr) return {};
    if (in.chroma != heif_chroma_420 && in.chroma != heif_chroma_422 && in.chroma != heif_chroma_444) return {};
    if (in.bits_per_pixel < 8 || in.bits_per_pixel > 16) return {};
    if (in.has_alpha && in.get_alpha_bits_per_pixel() != in.bits_per_pixel) return {};
    const int matrix = in.nclx.get_matrix_coefficients();
    if (matrix == 11 || matrix == 14) return {};
    if (in.chroma != heif_chroma_444 && options.only_use_preferred_chroma_algorithm) {
      // nearest neighbour is what the kernel does by default; bilinear is mirrored for 4:2:0 only
      const bool nn = options.preferred_chroma_upsampling_algorithm == heif_chroma_upsampling_nearest_neighbor;
      const bool bil420 = options.preferred_chroma_upsampling_algorithm == heif_chroma_upsampling_bilinear && in.chroma == heif_chroma_420;
      if (!nn && !bil420) return {};
    }
    if (target.colorspace != heif_colorspace_RGB) return {};
    std::vector<ColorStateWithCost> states;
    ColorState out;
    out.colorspace = heif_colorspace_RGB;
    switch (target.chroma) {
      case heif_chroma_interleaved_RGB: case heif_chroma_interleaved_RGBA:
        out.chroma = target.chroma; out.bits_per_pixel = 8; out.has_alpha = target.chroma == heif_chroma_interleaved_RGBA;
        break;
      case heif_chroma_interleaved_RRGGBB_LE: case heif_chroma_interleaved_RRGGBB_BE:
      case heif_chroma_interleaved_RRGGBBAA_LE: case heif_chroma_interleaved_RRGGBBAA_BE:
        if (in.bits_per_pixel <= 8) return {};
        out.chroma = target.chroma; out.bits_per_pixel = in.bits_per_pixel;
        out.has_alpha = target.chroma == heif_chroma_interleaved_RRGGBBAA_LE || target.chroma == heif_chroma_interleaved_RRGGBBAA_BE;
        break;
      default: return {};
    }
    states.emplace_back(out, SpeedCosts_Hardware);
    return states;
  }

  Result<std::shared_ptr<HeifPixelImage>> convert_colorspace(const std::shared_ptr<const HeifPixelImage>& input, const ColorState& in_state, const ColorState& target,
                                                             const heif_color_conversion_options& options, const heif_color_conversion_options_ext&,
                                                             const heif_security_limits* limits) const override {
    const uint32_t width = input->get_width(), height = input->get_height();
    const int bpp = input->get_bits_per_pixel(heif_channel_Y);
    b200_planes pl{};
    size_t ys = 0, cbs = 0, crs = 0, as = 0;
    pl.y = input->get_channel_memory(heif_channel_Y, &ys);
    pl.cb = input->get_channel_memory(heif_channel_Cb, &cbs);
    pl.cr = input->get_channel_memory(heif_channel_Cr, &crs);
    if (!pl.y || !pl.cb || !pl.cr || cbs != crs) return Error::InternalError;
    pl.y_stride = ys; pl.c_stride = cbs;
    const bool want_alpha = target.chroma == heif_chroma_interleaved_RGBA || target.chroma == heif_chroma_interleaved_RRGGBBAA_LE || target.chroma == heif_chroma_interleaved_RRGGBBAA_BE;
    if (want_alpha && input->has_channel(heif_channel_Alpha)) { pl.alpha = input->get_channel_memory(heif_channel_Alpha, &as); pl.alpha_stride = as; }
    pl.width = (int)width; pl.height = (int)height; pl.bit_depth = bpp;
    pl.chroma = in_state.chroma == heif_chroma_420 ? B200_CHROMA_420 : (in_state.chroma == heif_chroma_422 ? B200_CHROMA_422 : B200_CHROMA_444);
    // the reference's ops read the IMAGE's own nclx (yuv2rgb.cc:208-215), not the planner's defaulted copy
    pl.colour_primaries = 2; pl.transfer_characteristics = 2; pl.matrix_coefficients = 2; pl.full_range = 1;
    if (input->has_nclx_color_profile()) {
      auto p = input->get_color_profile_nclx();
      pl.colour_primaries = p.get_colour_primaries(); pl.transfer_characteristics = p.get_transfer_characteristics();
      pl.matrix_coefficients = p.get_matrix_coefficients(); pl.full_range = p.get_full_range_flag() ? 1 : 0;
    }
    b200_color_options opt{};
    opt.out_chroma = (int)target.chroma;          // heif_chroma values == B200_CHROMA_* values
    opt.chroma_upsampling = (in_state.chroma == heif_chroma_420 && options.only_use_preferred_chroma_algorithm &&
                             options.preferred_chroma_upsampling_algorithm == heif_chroma_upsampling_bilinear) ? 1 : 0;
    auto outimg = std::make_shared<HeifPixelImage>();
    outimg->create(width, height, heif_colorspace_RGB, target.chroma);
    const int out_bits = (target.chroma == heif_chroma_interleaved_RGB || target.chroma == heif_chroma_interleaved_RGBA) ? 8 : bpp;
    if (auto err = outimg->add_channel(heif_channel_interleaved, width, height, out_bits, limits)) return err;
    size_t os = 0;
    uint8_t* out = outimg->get_channel_memory(heif_channel_interleaved, &os);
    b200_geometry g; b200_geometry_init((int)width, (int)height, pl.chroma, &g);
    const int rc = b200_color_convert_host(&pl, &g, &opt, out, nullptr, nullptr, os, nullptr);
    if (rc) return Error(heif_error_Unsupported_feature, heif_suberror_Unsupported_color_conversion, b200_last_error());
    return outimg;
  }
};

}  // namespace

void b200_register_color_ops(std::vector<std::shared_ptr<ColorConversionOperation>>& ops) {
  ops.emplace_back(std::make_shared<Op_B200_YCbCr_to_interleaved>());
}
