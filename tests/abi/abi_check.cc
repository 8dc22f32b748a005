// Compile-time cross-check of include/b200_heif_plugin_abi.h against the reference's own headers
// (libheif/api/libheif/heif_plugin.h, heif_error.h, heif_library.h).  Built and run by tests/test_plugin_abi.py
// only where /root/reference exists.
#include <libheif/heif.h>
#include <libheif/heif_plugin.h>
#include <cstddef>
#include "../../include/b200_heif_plugin_abi.h"

#define SAME_OFF(A, B, m) static_assert(offsetof(A, m) == offsetof(B, m), "offset of " #m)
static_assert(sizeof(b200h_error) == sizeof(heif_error), "heif_error");
static_assert(offsetof(b200h_error, message) == offsetof(heif_error, message), "heif_error.message");
static_assert(sizeof(b200h_decoder_plugin) == sizeof(heif_decoder_plugin), "decoder plugin size");
SAME_OFF(b200h_decoder_plugin, heif_decoder_plugin, plugin_api_version); SAME_OFF(b200h_decoder_plugin, heif_decoder_plugin, get_plugin_name);
SAME_OFF(b200h_decoder_plugin, heif_decoder_plugin, does_support_format); SAME_OFF(b200h_decoder_plugin, heif_decoder_plugin, new_decoder);
SAME_OFF(b200h_decoder_plugin, heif_decoder_plugin, push_data); SAME_OFF(b200h_decoder_plugin, heif_decoder_plugin, decode_image);
SAME_OFF(b200h_decoder_plugin, heif_decoder_plugin, set_strict_decoding); SAME_OFF(b200h_decoder_plugin, heif_decoder_plugin, id_name);
SAME_OFF(b200h_decoder_plugin, heif_decoder_plugin, decode_next_image); SAME_OFF(b200h_decoder_plugin, heif_decoder_plugin, minimum_required_libheif_version);
SAME_OFF(b200h_decoder_plugin, heif_decoder_plugin, does_support_format2); SAME_OFF(b200h_decoder_plugin, heif_decoder_plugin, new_decoder2);
SAME_OFF(b200h_decoder_plugin, heif_decoder_plugin, push_data2); SAME_OFF(b200h_decoder_plugin, heif_decoder_plugin, flush_data);
SAME_OFF(b200h_decoder_plugin, heif_decoder_plugin, decode_next_image2);
static_assert(sizeof(b200h_encoder_plugin) == sizeof(heif_encoder_plugin), "encoder plugin size");
SAME_OFF(b200h_encoder_plugin, heif_encoder_plugin, compression_format); SAME_OFF(b200h_encoder_plugin, heif_encoder_plugin, id_name);
SAME_OFF(b200h_encoder_plugin, heif_encoder_plugin, priority); SAME_OFF(b200h_encoder_plugin, heif_encoder_plugin, supports_lossless_compression);
SAME_OFF(b200h_encoder_plugin, heif_encoder_plugin, new_encoder); SAME_OFF(b200h_encoder_plugin, heif_encoder_plugin, list_parameters);
SAME_OFF(b200h_encoder_plugin, heif_encoder_plugin, get_parameter_string); SAME_OFF(b200h_encoder_plugin, heif_encoder_plugin, query_input_colorspace);
SAME_OFF(b200h_encoder_plugin, heif_encoder_plugin, encode_image); SAME_OFF(b200h_encoder_plugin, heif_encoder_plugin, get_compressed_data);
SAME_OFF(b200h_encoder_plugin, heif_encoder_plugin, query_input_colorspace2); SAME_OFF(b200h_encoder_plugin, heif_encoder_plugin, query_encoded_size);
SAME_OFF(b200h_encoder_plugin, heif_encoder_plugin, minimum_required_libheif_version); SAME_OFF(b200h_encoder_plugin, heif_encoder_plugin, start_sequence_encoding);
SAME_OFF(b200h_encoder_plugin, heif_encoder_plugin, get_compressed_data2); SAME_OFF(b200h_encoder_plugin, heif_encoder_plugin, does_indicate_keyframes);
static_assert(sizeof(b200h_encoder_parameter) == sizeof(heif_encoder_parameter), "encoder parameter size");
static_assert(offsetof(b200h_encoder_parameter, has_default) == offsetof(heif_encoder_parameter, has_default), "has_default");
static_assert(sizeof(b200h_decoder_options) == sizeof(heif_decoder_plugin_options), "decoder options");
static_assert(offsetof(b200h_decoder_options, limits) == offsetof(heif_decoder_plugin_options, limits), "options.limits");
static_assert(offsetof(b200h_security_limits, max_image_size_pixels) == offsetof(heif_security_limits, max_image_size_pixels), "limits");
static_assert(sizeof(b200h_plugin_info) == sizeof(heif_plugin_info), "plugin_info");
static_assert(B200H_ERR_DECODER_PLUGIN == heif_error_Decoder_plugin_error && B200H_ERR_ENCODER_PLUGIN == heif_error_Encoder_plugin_error &&
              B200H_ERR_UNSUPPORTED_FEATURE == heif_error_Unsupported_feature && B200H_ERR_MEMORY == heif_error_Memory_allocation_error &&
              B200H_ERR_USAGE == heif_error_Usage_error, "error codes");
static_assert(B200H_SUBERR_SECURITY_LIMIT == heif_suberror_Security_limit_exceeded && B200H_SUBERR_UNSUPPORTED_CODEC == heif_suberror_Unsupported_codec &&
              B200H_SUBERR_END_OF_DATA == heif_suberror_End_of_data && B200H_SUBERR_UNSUPPORTED_BIT_DEPTH == heif_suberror_Unsupported_bit_depth, "suberrors");
static_assert(B200H_COMPRESSION_HEVC == heif_compression_HEVC && B200H_COLORSPACE_YCBCR == heif_colorspace_YCbCr && B200H_COLORSPACE_MONOCHROME == heif_colorspace_monochrome &&
              B200H_CHANNEL_Y == heif_channel_Y && B200H_CHANNEL_CB == heif_channel_Cb && B200H_CHANNEL_CR == heif_channel_Cr && heif_chroma_420 == 1 &&
              heif_plugin_type_decoder == 1 && heif_plugin_type_encoder == 0, "enums");
static_assert(LIBHEIF_MAKE_VERSION(1, 21, 0) == ((1u << 24) | (21u << 16)), "version macro");
struct NclxPublic { uint8_t version; int color_primaries; int transfer_characteristics; int matrix_coefficients; uint8_t full_range_flag; };
static_assert(offsetof(NclxPublic, full_range_flag) == offsetof(heif_color_profile_nclx, full_range_flag), "nclx.full_range_flag");
static_assert(offsetof(NclxPublic, matrix_coefficients) == offsetof(heif_color_profile_nclx, matrix_coefficients), "nclx.matrix");
int main() { return 0; }
