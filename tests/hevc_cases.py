"""Shared list of HEVC streams for the parser / decoder parity tests: the reference's own fixtures (harvested
through oracle/ref_plugin.cc's dump hook from examples/example.heic, tests/data/rainbow-451x461.heic and
fuzzing/data/corpus/*.heic) plus synthetic streams from the product's encoder covering every supported tool."""
import glob
import os

from libheif_b200 import hevc_enc

HERE = os.path.dirname(os.path.abspath(__file__))


def fixture_streams():
    out = []
    for f in sorted(glob.glob(os.path.join(HERE, "golden", "streams", "*.au"))):
        out.append((os.path.basename(f), open(f, "rb").read()))
    return out


# (name, width, height, bit_depth, chroma, encoder options)
SYNTH = [
    ("ctb16_basic", 64, 64, 8, True, dict(log2_ctb_size=4)),
    ("ctb16_nofilters", 64, 64, 8, True, dict(log2_ctb_size=4, sao=0, cu_qp_delta=0, sign_data_hiding=0, deblocking_disabled=1)),
    ("ctb32", 128, 96, 8, True, dict(log2_ctb_size=5)),
    ("ctb64", 256, 256, 8, True, dict(log2_ctb_size=6)),
    ("ctb16_wpp_nosao", 200, 120, 8, True, dict(log2_ctb_size=4, wpp=1, sao=0)),
    ("ctb32_wpp_deep", 200, 120, 8, True, dict(log2_ctb_size=5, wpp=1, max_transform_hierarchy_depth_intra=3)),
    ("ctb64_wpp_random", 264, 200, 8, True, dict(log2_ctb_size=6, wpp=1, mode_decision=0)),
    ("slices", 256, 192, 8, True, dict(log2_ctb_size=5, slice_ctb_rows=2)),
    ("slices_nolf", 256, 192, 8, True, dict(log2_ctb_size=5, slice_ctb_rows=2, slice_loop_filter_across_slices=0)),
    ("dependent_slices", 256, 192, 8, True, dict(log2_ctb_size=5, slice_ctb_rows=3, dependent_slice_segments=1)),
    ("slices_wpp", 256, 192, 8, True, dict(log2_ctb_size=5, slice_ctb_rows=2, wpp=1)),
    ("main10", 160, 160, 10, True, dict(log2_ctb_size=5)),
    ("main12_wpp", 160, 160, 12, True, dict(log2_ctb_size=6, wpp=1)),
    ("mono8", 160, 96, 8, False, dict(log2_ctb_size=5)),
    ("mono10_ctb16_wpp", 160, 96, 10, False, dict(log2_ctb_size=4, wpp=1)),
    ("transform_skip", 128, 128, 8, True, dict(log2_ctb_size=5, transform_skip=1, mode_decision=0)),
    ("chroma_qp_offsets", 128, 128, 8, True, dict(log2_ctb_size=5, cb_qp_offset=3, cr_qp_offset=-4, slice_chroma_qp_offsets=1, slice_cb_qp_offset=-2, slice_cr_qp_offset=5)),
    ("deblock_offsets", 128, 128, 8, True, dict(log2_ctb_size=5, beta_offset_div2=2, tc_offset_div2=-3)),
    ("deblock_slice_override", 128, 128, 8, True, dict(log2_ctb_size=5, slice_deblocking_override=1, slice_beta_offset_div2=-4, slice_tc_offset_div2=5, slice_ctb_rows=2)),
    ("deblock_slice_disabled", 128, 128, 8, True, dict(log2_ctb_size=5, slice_deblocking_override=1, slice_deblocking_disabled=1)),
    ("odd_size_random", 130, 70, 8, True, dict(log2_ctb_size=5, mode_decision=0, max_transform_hierarchy_depth_intra=2)),
    ("big_qp22", 1000, 600, 8, True, dict(log2_ctb_size=6, wpp=1, qp=22, dqp_range=6, diff_cu_qp_delta_depth=2)),
    ("big_qp37_vui", 1000, 600, 8, True, dict(log2_ctb_size=5, qp=37, strong_intra_smoothing=0, vui_present=1, colour_description_present=1,
                                              colour_primaries=1, transfer_characteristics=13, matrix_coefficients=6, full_range=1)),
    ("lowqp_deep", 72, 40, 8, True, dict(log2_ctb_size=6, qp=10, dqp_range=10, mode_decision=0, max_transform_hierarchy_depth_intra=4)),
    ("highqp", 72, 40, 8, True, dict(log2_ctb_size=6, qp=48, dqp_range=3, mode_decision=0)),
    ("tile_1024_like", 512, 512, 8, True, dict(log2_ctb_size=5, qp=27, wpp=1)),
    # scaling lists (7.3.4 / 7.4.5): default lists, lists coded in the SPS, lists coded in the PPS (explicit, copied, defaulted matrices)
    ("scaling_default", 192, 128, 8, True, dict(log2_ctb_size=5, scaling_lists=1, qp=24)),
    ("scaling_sps_ctb64", 256, 192, 8, True, dict(log2_ctb_size=6, scaling_lists=2, qp=26, wpp=1, max_transform_hierarchy_depth_intra=2)),
    ("scaling_pps_main10_tskip", 160, 160, 10, True, dict(log2_ctb_size=5, scaling_lists=3, qp=22, transform_skip=1, mode_decision=0)),
    # PCM coding units (7.3.8.7; pcm=2: pcm_loop_filter_disabled_flag) and cu_transquant_bypass_flag (transquant_bypass=2: every unit, i.e. lossless)
    ("pcm", 128, 96, 8, True, dict(log2_ctb_size=5, pcm=1)),
    ("pcm_nolf_wpp_ctb64", 192, 128, 8, True, dict(log2_ctb_size=6, pcm=2, wpp=1, mode_decision=0)),
    ("pcm_nolf_main10", 160, 96, 10, True, dict(log2_ctb_size=5, pcm=2)),
    ("pcm_mono", 136, 72, 8, False, dict(log2_ctb_size=4, pcm=1)),
    ("bypass_mixed", 128, 96, 8, True, dict(log2_ctb_size=5, transquant_bypass=1)),
    ("bypass_lossless", 96, 64, 8, True, dict(log2_ctb_size=4, transquant_bypass=2)),
    ("bypass_tskip_main12", 160, 96, 12, True, dict(log2_ctb_size=5, transquant_bypass=1, transform_skip=1, mode_decision=0)),
    ("pcm_bypass_scaling_slices_wpp", 256, 192, 8, True, dict(log2_ctb_size=5, pcm=1, transquant_bypass=1, scaling_lists=2, slice_ctb_rows=2, wpp=1)),
    # HEVC tiles (6.5.1): uniform and explicit spacing, one slice over all tiles / one slice per tile, loop filters across tile boundaries on and off
    ("tiles_2x2", 256, 192, 8, True, dict(log2_ctb_size=5, tile_cols=2, tile_rows=2)),
    ("tiles_3x2_explicit_nolf", 320, 256, 8, True, dict(log2_ctb_size=5, tile_cols=3, tile_rows=2, tiles_uniform=0, loop_filter_across_tiles=0)),
    ("tiles_2x3_slice_per_tile_main10", 264, 200, 10, True, dict(log2_ctb_size=5, tile_cols=2, tile_rows=3, slice_per_tile=1, loop_filter_across_tiles=0, slice_loop_filter_across_slices=0)),
    ("tiles_4x1_ctb64", 512, 128, 8, True, dict(log2_ctb_size=6, tile_cols=4, tile_rows=1)),
    ("tiles_5x4_ctb16_mono_nolf", 208, 176, 8, False, dict(log2_ctb_size=4, tile_cols=5, tile_rows=4, loop_filter_across_tiles=0)),
    ("tiles_3x3_pcm_bypass_scaling", 200, 168, 8, True, dict(log2_ctb_size=4, tile_cols=3, tile_rows=3, tiles_uniform=0, mode_decision=0, pcm=1, transquant_bypass=1, scaling_lists=2, sao=0)),
]

# More feature combinations, used by the CPU tests only (FFmpeg pin of the restatement, host front-end vs restatement):
# they widen the pin of the oracle and of the syntax decoder the GPU kernel shares with the host front-end.
SYNTH_CPU_EXTRA = [
    ("x_main10_slices_wpp_tskip", 264, 136, 10, True, dict(log2_ctb_size=5, slice_ctb_rows=2, wpp=1, transform_skip=1, mode_decision=0)),
    ("x_ctb64_dependent_wpp", 320, 256, 8, True, dict(log2_ctb_size=6, slice_ctb_rows=2, dependent_slice_segments=1, wpp=1)),
    ("x_mono12_ctb64", 200, 136, 12, False, dict(log2_ctb_size=6, qp=20, dqp_range=8)),
    ("x_ctb16_deep_qp14", 136, 72, 8, True, dict(log2_ctb_size=4, qp=14, max_transform_hierarchy_depth_intra=2, dqp_range=5, diff_cu_qp_delta_depth=1, sao=0)),
    ("x_main12_highqp_nolf", 160, 96, 12, True, dict(log2_ctb_size=5, qp=45, deblocking_disabled=1, sign_data_hiding=0)),
    ("x_odd_8bit_wpp_random", 74, 58, 8, True, dict(log2_ctb_size=5, wpp=1, mode_decision=0, max_transform_hierarchy_depth_intra=3, cb_qp_offset=-5, cr_qp_offset=7)),
    ("x_main10_ctb64_qg8", 256, 128, 10, True, dict(log2_ctb_size=6, diff_cu_qp_delta_depth=3, dqp_range=12, qp=30)),
    ("x_scaling_sps_mono_qp12", 136, 72, 8, False, dict(log2_ctb_size=4, scaling_lists=2, qp=12, seed=77)),
    ("x_scaling_pps_slices_wpp", 256, 192, 8, True, dict(log2_ctb_size=5, scaling_lists=3, slice_ctb_rows=2, wpp=1, qp=30, seed=991)),
    ("x_bypass_mixed_nosao", 128, 96, 8, True, dict(log2_ctb_size=5, transquant_bypass=1, sao=0, seed=5)),
    ("x_pcm_nolf_nosao_ctb16", 136, 72, 8, True, dict(log2_ctb_size=4, pcm=2, sao=0, seed=6)),
    ("x_lossless_main10_ctb64_wpp", 200, 136, 10, True, dict(log2_ctb_size=6, transquant_bypass=2, wpp=1)),
    ("x_lossless_mono_nosao", 72, 40, 8, False, dict(log2_ctb_size=5, transquant_bypass=2, sao=0)),
    ("x_tiles_2x2_main12_dqp_slice_per_tile", 256, 192, 12, True, dict(log2_ctb_size=6, tile_cols=2, tile_rows=2, dqp_range=8, diff_cu_qp_delta_depth=2, slice_per_tile=1)),
    ("x_tiles_1x3_slices_nolf_across_slices", 136, 200, 8, True, dict(log2_ctb_size=5, tile_cols=1, tile_rows=3, slice_per_tile=1, slice_loop_filter_across_slices=0)),
    ("x_tiles_7x1_random_deep", 232, 72, 8, True, dict(log2_ctb_size=5, tile_cols=7, tile_rows=1, tiles_uniform=0, mode_decision=0, max_transform_hierarchy_depth_intra=3, transform_skip=1)),
    # 4:2:2 and 4:4:4 coded pictures (chroma column: chroma_format_idc)
    ("x_422_basic", 128, 96, 8, 2, dict(log2_ctb_size=5)),
    ("x_444_basic", 128, 96, 8, 3, dict(log2_ctb_size=5)),
    ("x_422_main10_random_tskip", 136, 72, 10, 2, dict(log2_ctb_size=5, mode_decision=0, max_transform_hierarchy_depth_intra=2, transform_skip=1)),
    ("x_444_main10_ctb16_random_tskip", 136, 72, 10, 3, dict(log2_ctb_size=4, mode_decision=0, max_transform_hierarchy_depth_intra=2, transform_skip=1)),
    ("x_422_main12_ctb64_deep_dqp", 200, 136, 12, 2, dict(log2_ctb_size=6, mode_decision=0, max_transform_hierarchy_depth_intra=4, qp=18, dqp_range=8)),
    ("x_444_ctb64_deep_dqp", 200, 136, 8, 3, dict(log2_ctb_size=6, mode_decision=0, max_transform_hierarchy_depth_intra=4, qp=18, dqp_range=8)),
    ("x_422_wpp_slices", 128, 128, 8, 2, dict(log2_ctb_size=5, wpp=1, slice_ctb_rows=2, qp=35)),
    ("x_444_wpp_slices", 128, 128, 8, 3, dict(log2_ctb_size=5, wpp=1, slice_ctb_rows=2, qp=35)),
    ("x_422_pcm_bypass_nosao", 96, 64, 8, 2, dict(log2_ctb_size=5, transquant_bypass=1, pcm=1, sao=0)),
    ("x_444_pcm_bypass_nosao", 96, 64, 8, 3, dict(log2_ctb_size=4, transquant_bypass=1, pcm=1, sao=0)),
    ("x_422_scaling_sps", 160, 96, 8, 2, dict(log2_ctb_size=5, scaling_lists=2, qp=20)),
    ("x_444_scaling_pps_ctb64", 192, 128, 8, 3, dict(log2_ctb_size=6, scaling_lists=3, mode_decision=0, max_transform_hierarchy_depth_intra=1, qp=14)),
    ("x_422_tiles_nolf", 128, 96, 8, 2, dict(log2_ctb_size=5, tile_cols=2, tile_rows=2, loop_filter_across_tiles=0)),
    ("x_444_lossless", 72, 40, 8, 3, dict(log2_ctb_size=5, transquant_bypass=2, sao=0)),
    ("x_422_lossless_main10", 72, 40, 10, 2, dict(log2_ctb_size=5, transquant_bypass=2)),
    ("x_444_qp_offsets_pcm_nolf", 130, 70, 10, 3, dict(log2_ctb_size=5, mode_decision=0, cb_qp_offset=5, cr_qp_offset=-7, pcm=2)),
    ("x_slices_every_row_nolf_across", 192, 160, 8, True, dict(log2_ctb_size=5, slice_ctb_rows=1, loop_filter_across_slices=0, slice_loop_filter_across_slices=0)),
]

_cache = {}


def synth_params(name):
    for c in SYNTH + SYNTH_CPU_EXTRA:
        if c[0] == name:
            return c
    return None


def synth_source(name):
    """the picture the synthetic stream `name` was encoded from (y, cb, cr)"""
    (_, w, h, bd, chroma, _) = synth_params(name)
    return hevc_enc.synthetic_image(0xB200 + w + h, w, h, bd, chroma)


def synth_stream(name):
    if name not in _cache:
        for (nm, w, h, bd, chroma, opts) in SYNTH + SYNTH_CPU_EXTRA:
            if nm == name:
                y, cb, cr = hevc_enc.synthetic_image(0xB200 + w + h, w, h, bd, chroma)
                _cache[name] = hevc_enc.encode_intra(y, cb, cr, bit_depth=bd, **opts)
    return _cache[name]


def all_streams():
    return fixture_streams() + [(s[0], synth_stream(s[0])) for s in SYNTH]


def cpu_extra_streams():
    return [(s[0], synth_stream(s[0])) for s in SYNTH_CPU_EXTRA]
