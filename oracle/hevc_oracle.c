/*
 * oracle/hevc_oracle.c -- TEST INFRASTRUCTURE ONLY. Never linked, imported or executed by the product.
 *
 * Plain scalar C restatement of the HEVC (ITU-T H.265) intra-picture decoding process, i.e. of the
 * arithmetic the reference obtains from libde265 behind
 *   libheif/plugins/decoder_libde265.cc:322-368 (push NALs), :386-457 (de265_decode / get_next_picture),
 *   :97-171 (plane hand-over, conformance-cropped), :426-448 (VUI -> nclx).
 * libde265 is a third-party dependency that is absent from /root/reference (not vendored, no submodule,
 * version unpinned: cmake/modules/FindLIBDE265.cmake:1-43), so this file restates the PUBLISHED
 * algorithm (H.265 clause numbers are cited at each function) and is PINNED against an independent
 * conforming decoder (FFmpeg libavcodec 62, oracle/ffhevc.c) on the reference's own HEVC fixtures
 * (examples/example.heic, tests/data/rainbow-451x461.heic, fuzzing/data/corpus/*.heic) and on every
 * synthetic stream used by the tests -- see tests/test_oracle_hevc.py.
 *
 * Scope: I slices only (IDR/CRA/BLA still pictures), 4:2:0 and 4:0:0, 8..12 bit, CTB 16/32/64,
 * TB 4..32, SAO, deblocking, sign-data hiding, cu_qp_delta, transform-skip 4x4, WPP, multiple slices and
 * dependent slice segments.  Not handled (returns HO_UNSUPPORTED): P/B slices, 4:2:2/4:4:4, tiles,
 * scaling lists, PCM, cu_transquant_bypass, range-extension coding tools.
 *
 * Deliberately simple: bit-serial CABAC exactly as written in clause 9.3.4.3, whole-picture passes for
 * deblocking and SAO, uint16 planes for every bit depth.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define HO_OK 0
#define HO_ERROR (-1)
#define HO_UNSUPPORTED (-2)

typedef struct {
  int width, height;   /* conformance-cropped luma size */
  int cw, ch;          /* chroma plane size, 0 for 4:0:0 */
  int bit_depth;       /* luma bit depth (== chroma bit depth enforced) */
  int chroma_format;   /* 0 or 1 */
  int vui_colour_present, colour_primaries, transfer_characteristics, matrix_coeffs, full_range;
  int video_signal_present;
  uint16_t* plane[3];  /* tightly packed rows */
} hevc_oracle_picture;

/* ------------------------------------------------------------------------------------------ bits */
typedef struct { const uint8_t* d; size_t n; size_t pos; /* bit position */ } bitrd;

static unsigned rd_bit(bitrd* b) {
  if ((b->pos >> 3) >= b->n) { b->pos++; return 0; }
  unsigned v = (b->d[b->pos >> 3] >> (7 - (b->pos & 7))) & 1;
  b->pos++;
  return v;
}
static unsigned rd_bits(bitrd* b, int n) { unsigned v = 0; while (n-- > 0) v = (v << 1) | rd_bit(b); return v; }
static unsigned rd_ue(bitrd* b) {
  int z = 0;
  while (rd_bit(b) == 0 && z < 32) z++;
  return z ? ((1u << z) - 1 + rd_bits(b, z)) : 0;
}
static int rd_se(bitrd* b) { unsigned k = rd_ue(b); return (k & 1) ? (int)((k + 1) >> 1) : -(int)(k >> 1); }

/* 7.4.2 / 7.3.1.1: strip emulation_prevention_three_byte */
static size_t nal_to_rbsp(const uint8_t* in, size_t n, uint8_t* out) {
  size_t o = 0; int zeros = 0;
  for (size_t i = 0; i < n; i++) {
    if (zeros >= 2 && in[i] == 3) { zeros = 0; continue; }
    out[o++] = in[i];
    zeros = in[i] == 0 ? zeros + 1 : 0;
  }
  return o;
}

/* ------------------------------------------------------------------------------- parameter sets */
typedef struct {
  int valid, chroma_format_idc, width, height;
  int conf_l, conf_r, conf_t, conf_b;
  int bit_depth, bit_depth_c;
  int log2_max_poc_lsb;
  int log2_min_cb, log2_ctb, log2_min_tb, log2_max_tb;
  int max_th_depth_intra;
  int scaling_list_enabled, amp, sao, pcm, strong_intra_smoothing;
  int num_st_rps, long_term_present, temporal_mvp;
  int st_rps_num_delta[65];
  int num_lt_sps;
  int vui_signal, vui_full_range, vui_colour, vui_cp, vui_tc, vui_mc;
  int range_ext_unsupported;
  int sl_data_present; uint8_t sl[4][6][64], sl_dc[4][6];   /* ScalingList[sizeId][matrixId][i], scaling_list_dc_coef (7.3.4) */
  int pcm_bd_y, pcm_bd_c, log2_min_pcm, log2_max_pcm, pcm_loop_filter_disabled;
} sps_t;

typedef struct {
  int valid, sps_id;
  int dependent_slices, output_flag_present, num_extra_bits, sign_hiding, cabac_init_present;
  int init_qp, constrained_intra, transform_skip, cu_qp_delta, diff_cu_qp_delta_depth;
  int cb_qp_offset, cr_qp_offset, slice_chroma_qp_offsets_present;
  int transquant_bypass, tiles, wpp;
  int num_tile_cols, num_tile_rows, uniform_spacing, lf_across_tiles;   /* 7.3.2.3: tiles */
  int col_width[20], row_height[22];
  int lf_across_slices, deblock_control, deblock_override_enabled, deblock_disabled, beta_offset, tc_offset;
  int scaling_list_present, lists_modification, slice_ext_present;
  int log2_sao_scale_luma, log2_sao_scale_chroma;
  int range_ext_unsupported;
  uint8_t sl[4][6][64], sl_dc[4][6];
} pps_t;

/* ---- scaling lists: 7.3.4 scaling_list_data, 7.4.5 (defaults: Table 7-5 = 16 everywhere, Table 7-6 below in diagonal scan order) */
static const uint8_t default_sl8[2][64] = {
  {16,16,16,16,16,16,16,16,16,16,17,16,17,16,17,18,17,18,18,17,18,21,19,20,21,20,19,21,24,22,22,24,24,22,22,24,25,25,27,30,27,25,25,29,31,35,35,31,29,36,41,44,41,36,47,54,54,47,65,70,65,88,88,115},
  {16,16,16,16,16,16,16,16,16,16,17,17,17,17,17,18,18,18,18,18,18,20,20,20,20,20,20,20,24,24,24,24,24,24,24,24,25,25,25,25,25,25,25,28,28,28,28,28,28,33,33,33,33,33,41,41,41,41,54,54,54,71,71,91}};
static void sl_default(uint8_t sl[4][6][64], uint8_t dc[4][6], int size_id, int matrix_id) {
  for (int i = 0; i < 64; i++) sl[size_id][matrix_id][i] = size_id == 0 ? 16 : default_sl8[matrix_id < 3 ? 0 : 1][i];
  dc[size_id][matrix_id] = 16;
}
static int parse_scaling_list_data(bitrd* b, uint8_t sl[4][6][64], uint8_t dc[4][6]) {
  for (int sizeId = 0; sizeId < 4; sizeId++) for (int matrixId = 0; matrixId < 6; matrixId++) sl_default(sl, dc, sizeId, matrixId);
  for (int sizeId = 0; sizeId < 4; sizeId++)
    for (int matrixId = 0; matrixId < 6; matrixId += (sizeId == 3) ? 3 : 1) {
      int pred_mode_flag = rd_bit(b);
      if (!pred_mode_flag) {
        unsigned delta = rd_ue(b);                                  /* scaling_list_pred_matrix_id_delta */
        int step = sizeId == 3 ? 3 : 1;
        if (delta * step > (unsigned)matrixId) return HO_ERROR;
        if (delta == 0) sl_default(sl, dc, sizeId, matrixId);
        else { int ref = matrixId - (int)delta * step; memcpy(sl[sizeId][matrixId], sl[sizeId][ref], 64); dc[sizeId][matrixId] = dc[sizeId][ref]; }
      } else {
        int nextCoef = 8, coefNum = sizeId == 0 ? 16 : 64;
        if (sizeId > 1) { int v = rd_se(b); if (v < -7 || v > 247) return HO_ERROR; nextCoef = v + 8; dc[sizeId][matrixId] = (uint8_t)nextCoef; }
        for (int i = 0; i < coefNum; i++) {
          int d = rd_se(b); if (d < -128 || d > 127) return HO_ERROR;
          nextCoef = (nextCoef + d + 256) % 256;
          sl[sizeId][matrixId][i] = (uint8_t)nextCoef;
        }
      }
    }
  return HO_OK;
}

static void skip_profile_tier_level(bitrd* b, int max_sub_layers_minus1) {
  rd_bits(b, 8); rd_bits(b, 32); rd_bits(b, 4); rd_bits(b, 32); rd_bits(b, 11); rd_bit(b); /* 88 bits */
  rd_bits(b, 8);                                                                               /* level */
  int pp[8], lp[8];
  for (int i = 0; i < max_sub_layers_minus1; i++) { pp[i] = rd_bit(b); lp[i] = rd_bit(b); }
  if (max_sub_layers_minus1 > 0) for (int i = max_sub_layers_minus1; i < 8; i++) rd_bits(b, 2);
  for (int i = 0; i < max_sub_layers_minus1; i++) {
    if (pp[i]) { rd_bits(b, 32); rd_bits(b, 32); rd_bits(b, 24); }
    if (lp[i]) rd_bits(b, 8);
  }
}

/* E.2.2 hrd_parameters: skipped field by field */
static void skip_hrd(bitrd* b, int common, int max_sub) {
  int nal = 0, vcl = 0, subpic = 0;
  if (common) {
    nal = rd_bit(b); vcl = rd_bit(b);
    if (nal || vcl) {
      subpic = rd_bit(b);
      if (subpic) { rd_bits(b, 8); rd_bits(b, 5); rd_bit(b); rd_bits(b, 5); }
      rd_bits(b, 4); rd_bits(b, 4);
      if (subpic) rd_bits(b, 4);
      rd_bits(b, 5); rd_bits(b, 5); rd_bits(b, 5);
    }
  }
  for (int i = 0; i <= max_sub; i++) {
    int general = rd_bit(b), within = 1, low_delay = 0, cpb_cnt = 0;
    if (!general) within = rd_bit(b);
    if (within) rd_ue(b); else low_delay = rd_bit(b);
    if (!low_delay) cpb_cnt = rd_ue(b);
    for (int k = 0; k < nal + vcl; k++)
      for (int c = 0; c <= cpb_cnt; c++) { rd_ue(b); rd_ue(b); if (subpic) { rd_ue(b); rd_ue(b); } rd_bit(b); }
  }
}

/* 7.3.7 st_ref_pic_set: only parsed to reach the syntax that follows it */
static void parse_st_rps(bitrd* b, sps_t* s, int idx, int num) {
  int inter = idx ? rd_bit(b) : 0;
  if (inter) {
    int delta_idx = 1;
    if (idx == num) delta_idx = rd_ue(b) + 1;
    rd_bit(b); rd_ue(b);
    int ref = idx - delta_idx; if (ref < 0) ref = 0;
    int cnt = 0;
    for (int j = 0; j <= s->st_rps_num_delta[ref]; j++) {
      int used = rd_bit(b), use_delta = 1;
      if (!used) use_delta = rd_bit(b);
      if (used || use_delta) cnt++;
    }
    s->st_rps_num_delta[idx] = cnt;
  } else {
    int nn = rd_ue(b), np = rd_ue(b);
    for (int i = 0; i < nn + np; i++) { rd_ue(b); rd_bit(b); }
    s->st_rps_num_delta[idx] = nn + np;
  }
}

/* 7.3.2.2 */
static int parse_sps(const uint8_t* rbsp, size_t n, sps_t* s) {
  bitrd b = {rbsp, n, 16};
  memset(s, 0, sizeof *s);
  rd_bits(&b, 4);
  int msl = rd_bits(&b, 3);
  rd_bit(&b);
  skip_profile_tier_level(&b, msl);
  rd_ue(&b);
  s->chroma_format_idc = rd_ue(&b);
  if (s->chroma_format_idc == 3) rd_bit(&b);
  s->width = rd_ue(&b); s->height = rd_ue(&b);
  if (rd_bit(&b)) { s->conf_l = rd_ue(&b); s->conf_r = rd_ue(&b); s->conf_t = rd_ue(&b); s->conf_b = rd_ue(&b); }
  s->bit_depth = 8 + rd_ue(&b); s->bit_depth_c = 8 + rd_ue(&b);
  s->log2_max_poc_lsb = 4 + rd_ue(&b);
  int sub_info = rd_bit(&b);
  for (int i = sub_info ? 0 : msl; i <= msl; i++) { rd_ue(&b); rd_ue(&b); rd_ue(&b); }
  s->log2_min_cb = 3 + rd_ue(&b);
  s->log2_ctb = s->log2_min_cb + rd_ue(&b);
  s->log2_min_tb = 2 + rd_ue(&b);
  s->log2_max_tb = s->log2_min_tb + rd_ue(&b);
  rd_ue(&b);
  s->max_th_depth_intra = rd_ue(&b);
  s->scaling_list_enabled = rd_bit(&b);
  s->sl_data_present = 0;
  if (s->scaling_list_enabled) {
    s->sl_data_present = rd_bit(&b);
    if (s->sl_data_present && parse_scaling_list_data(&b, s->sl, s->sl_dc) != HO_OK) return HO_ERROR;
  }
  s->amp = rd_bit(&b); s->sao = rd_bit(&b); s->pcm = rd_bit(&b);
  if (s->pcm) {                                                  /* 7.3.2.2: pcm_enabled_flag */
    s->pcm_bd_y = 1 + rd_bits(&b, 4); s->pcm_bd_c = 1 + rd_bits(&b, 4);
    s->log2_min_pcm = 3 + rd_ue(&b); s->log2_max_pcm = s->log2_min_pcm + rd_ue(&b);
    s->pcm_loop_filter_disabled = rd_bit(&b);
    if (s->pcm_bd_y > s->bit_depth || s->pcm_bd_c > s->bit_depth_c || s->log2_max_pcm > 5 || s->log2_max_pcm > s->log2_ctb || s->log2_min_pcm < s->log2_min_cb) return HO_ERROR;
  }
  s->num_st_rps = rd_ue(&b);
  if (s->num_st_rps > 64) return HO_ERROR;
  for (int i = 0; i < s->num_st_rps; i++) parse_st_rps(&b, s, i, s->num_st_rps);
  s->long_term_present = rd_bit(&b);
  if (s->long_term_present) {
    s->num_lt_sps = rd_ue(&b);
    for (int i = 0; i < s->num_lt_sps; i++) { rd_bits(&b, s->log2_max_poc_lsb); rd_bit(&b); }
  }
  s->temporal_mvp = rd_bit(&b);
  s->strong_intra_smoothing = rd_bit(&b);
  if (rd_bit(&b)) { /* E.2.1 vui_parameters, up to the colour description */
    if (rd_bit(&b)) { if (rd_bits(&b, 8) == 255) { rd_bits(&b, 16); rd_bits(&b, 16); } }
    if (rd_bit(&b)) rd_bit(&b);
    s->vui_signal = rd_bit(&b);
    if (s->vui_signal) {
      rd_bits(&b, 3);
      s->vui_full_range = rd_bit(&b);
      s->vui_colour = rd_bit(&b);
      if (s->vui_colour) { s->vui_cp = rd_bits(&b, 8); s->vui_tc = rd_bits(&b, 8); s->vui_mc = rd_bits(&b, 8); }
    }
    if (rd_bit(&b)) { rd_ue(&b); rd_ue(&b); }              /* chroma_loc_info */
    rd_bit(&b); rd_bit(&b); rd_bit(&b);                     /* neutral_chroma, field_seq, frame_field_info */
    if (rd_bit(&b)) { rd_ue(&b); rd_ue(&b); rd_ue(&b); rd_ue(&b); }   /* default display window */
    if (rd_bit(&b)) {                                        /* vui_timing_info */
      rd_bits(&b, 32); rd_bits(&b, 32);
      if (rd_bit(&b)) rd_ue(&b);
      if (rd_bit(&b)) skip_hrd(&b, 1, msl);
    }
    if (rd_bit(&b)) { rd_bits(&b, 3); rd_ue(&b); rd_ue(&b); rd_ue(&b); rd_ue(&b); rd_ue(&b); }   /* bitstream_restriction */
  }
  if (rd_bit(&b)) {                                          /* sps_extension_present_flag, 7.3.2.2.2 */
    int range = rd_bit(&b); rd_bits(&b, 7);
    if (range) {
      /* transform_skip_rotation, transform_skip_context, implicit_rdpcm, explicit_rdpcm, extended_precision,
         intra_smoothing_disabled, high_precision_offsets, persistent_rice_adaptation, cabac_bypass_alignment */
      int f[9]; for (int i = 0; i < 9; i++) f[i] = rd_bit(&b);
      if (f[0] || f[1] || f[2] || f[4] || f[5] || f[7] || f[8]) return HO_UNSUPPORTED;
    }
  }
  if (s->chroma_format_idc > 3) return HO_ERROR;
  if (s->bit_depth != s->bit_depth_c || s->bit_depth > 12) return HO_UNSUPPORTED;
  if (s->log2_ctb > 6 || s->log2_ctb < 4 || s->log2_max_tb > 5) return HO_ERROR;
  s->valid = 1;
  return HO_OK;
}

/* 7.3.2.3 */
static int parse_pps(const uint8_t* rbsp, size_t n, pps_t* p) {
  bitrd b = {rbsp, n, 16};
  memset(p, 0, sizeof *p);
  rd_ue(&b);
  p->sps_id = rd_ue(&b);
  p->dependent_slices = rd_bit(&b);
  p->output_flag_present = rd_bit(&b);
  p->num_extra_bits = rd_bits(&b, 3);
  p->sign_hiding = rd_bit(&b);
  p->cabac_init_present = rd_bit(&b);
  rd_ue(&b); rd_ue(&b);
  p->init_qp = 26 + rd_se(&b);
  p->constrained_intra = rd_bit(&b);
  p->transform_skip = rd_bit(&b);
  p->cu_qp_delta = rd_bit(&b);
  if (p->cu_qp_delta) p->diff_cu_qp_delta_depth = rd_ue(&b);
  p->cb_qp_offset = rd_se(&b); p->cr_qp_offset = rd_se(&b);
  p->slice_chroma_qp_offsets_present = rd_bit(&b);
  rd_bit(&b); rd_bit(&b);
  p->transquant_bypass = rd_bit(&b);
  p->tiles = rd_bit(&b);
  p->wpp = rd_bit(&b);
  p->num_tile_cols = p->num_tile_rows = 1; p->lf_across_tiles = 1;
  if (p->tiles) {
    if (p->wpp) return HO_UNSUPPORTED;                           /* tiles together with wavefronts: not restated */
    p->num_tile_cols = 1 + rd_ue(&b); p->num_tile_rows = 1 + rd_ue(&b);
    if (p->num_tile_cols > 20 || p->num_tile_rows > 22) return HO_ERROR;
    p->uniform_spacing = rd_bit(&b);
    if (!p->uniform_spacing) {
      for (int i = 0; i + 1 < p->num_tile_cols; i++) p->col_width[i] = 1 + rd_ue(&b);
      for (int i = 0; i + 1 < p->num_tile_rows; i++) p->row_height[i] = 1 + rd_ue(&b);
    }
    p->lf_across_tiles = rd_bit(&b);
  }
  p->lf_across_slices = rd_bit(&b);
  p->deblock_control = rd_bit(&b);
  if (p->deblock_control) {
    p->deblock_override_enabled = rd_bit(&b);
    p->deblock_disabled = rd_bit(&b);
    if (!p->deblock_disabled) { p->beta_offset = 2 * rd_se(&b); p->tc_offset = 2 * rd_se(&b); }
  }
  p->scaling_list_present = rd_bit(&b);
  if (p->scaling_list_present && parse_scaling_list_data(&b, p->sl, p->sl_dc) != HO_OK) return HO_ERROR;
  p->lists_modification = rd_bit(&b);
  rd_ue(&b);
  p->slice_ext_present = rd_bit(&b);
  if (rd_bit(&b)) { /* pps_extension_present_flag */
    int range = rd_bit(&b); rd_bits(&b, 7);
    if (range) {
      if (p->transform_skip) { if (rd_ue(&b) != 0) return HO_UNSUPPORTED; } /* log2_max_transform_skip_block_size_minus2 */
      if (rd_bit(&b)) return HO_UNSUPPORTED;  /* cross_component_prediction */
      if (rd_bit(&b)) return HO_UNSUPPORTED;  /* chroma_qp_offset_list */
      p->log2_sao_scale_luma = rd_ue(&b);
      p->log2_sao_scale_chroma = rd_ue(&b);
    }
  }
  if (getenv("HO_DEBUG")) fprintf(stderr, "pps: cu_qp_delta=%d depth=%d sign_hiding=%d tskip=%d wpp=%d cb=%d cr=%d init_qp=%d\n", p->cu_qp_delta, p->diff_cu_qp_delta_depth, p->sign_hiding, p->transform_skip, p->wpp, p->cb_qp_offset, p->cr_qp_offset, p->init_qp);
  p->valid = 1;
  return HO_OK;
}

/* ----------------------------------------------------------------------------------- decoder state */
enum { CTX_SAO_MERGE = 0, CTX_SAO_TYPE = 1, CTX_SPLIT_CU = 2, CTX_PART_MODE = 5, CTX_PREV_INTRA = 6,
       CTX_CHROMA_PRED = 7, CTX_SPLIT_TR = 8, CTX_CBF_LUMA = 11, CTX_CBF_CHROMA = 13, CTX_QP_DELTA = 18,
       CTX_TSKIP = 20, CTX_LAST_X = 22, CTX_LAST_Y = 40, CTX_CSBF = 58, CTX_SIG = 62, CTX_GT1 = 104,
       CTX_GT2 = 128, CTX_TQ_BYPASS = 134, CTX_COUNT = 135 };

/* Tables 9-5..9-37, initType 0 (I slices) */
static const uint8_t ctx_init_I[CTX_COUNT] = {
  153,                                  /* sao_merge */
  200,                                  /* sao_type_idx */
  139, 141, 157,                        /* split_cu_flag */
  184,                                  /* part_mode */
  184,                                  /* prev_intra_luma_pred_flag */
  63,                                   /* intra_chroma_pred_mode */
  153, 138, 138,                        /* split_transform_flag */
  111, 141,                             /* cbf_luma */
  94, 138, 182, 154, 154,               /* cbf_cb / cbf_cr */
  154, 154,                             /* cu_qp_delta_abs */
  139, 139,                             /* transform_skip_flag luma, chroma */
  110, 110, 124, 125, 140, 153, 125, 127, 140, 109, 111, 143, 127, 111, 79, 108, 123, 63, /* last x */
  110, 110, 124, 125, 140, 153, 125, 127, 140, 109, 111, 143, 127, 111, 79, 108, 123, 63, /* last y */
  91, 171, 134, 141,                    /* coded_sub_block_flag */
  111, 111, 125, 110, 110, 94, 124, 108, 124, 107, 125, 141, 179, 153, 125, 107, 125, 141, 179, 153, 125,
  107, 125, 141, 179, 153, 125, 140, 139, 182, 182, 152, 136, 152, 136, 153, 136, 139, 111, 136, 139, 111, /* sig */
  140, 92, 137, 138, 140, 152, 138, 139, 153, 74, 149, 92, 139, 107, 122, 152, 140, 179, 166, 182, 140, 227, 122, 197, /* gt1 */
  138, 153, 136, 167, 152, 152,         /* gt2 */
  154                                   /* cu_transquant_bypass_flag */
};

/* Table 9-46 */
static const uint8_t range_tab_lps[64][4] = {
  {128,176,208,240},{128,167,197,227},{128,158,187,216},{123,150,178,205},{116,142,169,195},{111,135,160,185},
  {105,128,152,175},{100,122,144,166},{95,116,137,158},{90,110,130,150},{85,104,123,142},{81,99,117,135},
  {77,94,111,128},{73,89,105,122},{69,85,100,116},{66,80,95,110},{62,76,90,104},{59,72,86,99},{56,69,81,94},
  {53,65,77,89},{51,62,73,85},{48,59,69,80},{46,56,66,76},{43,53,63,72},{41,50,59,69},{39,48,56,65},
  {37,45,54,62},{35,43,51,59},{33,41,48,56},{32,39,46,53},{30,37,43,50},{29,35,41,48},{27,33,39,45},
  {26,31,37,43},{24,30,35,41},{23,28,33,39},{22,27,32,37},{21,26,30,35},{20,24,29,33},{19,23,27,31},
  {18,22,26,30},{17,21,25,28},{16,20,23,27},{15,19,22,25},{14,18,21,24},{14,17,20,23},{13,16,19,22},
  {12,15,18,21},{12,14,17,20},{11,14,16,19},{11,13,15,18},{10,12,15,17},{10,12,14,16},{9,11,13,15},
  {9,11,12,14},{8,10,12,14},{8,9,11,13},{7,9,11,12},{7,9,10,12},{7,8,10,11},{6,8,9,11},{6,7,9,10},
  {6,7,8,9},{2,2,2,2}};
/* Table 9-47 */
static const uint8_t trans_lps[64] = {0,0,1,2,2,4,4,5,6,7,8,9,9,11,11,12,13,13,15,15,16,16,18,18,19,19,21,21,22,22,23,24,
  24,25,26,26,27,27,28,29,29,30,30,30,31,32,32,33,33,33,34,34,35,35,35,36,36,36,37,37,37,38,38,63};

typedef struct { uint8_t state, mps; } cabac_ctx;

typedef struct {
  int type[3];          /* SaoTypeIdx */
  int band_pos[3];
  int eo_class[3];
  int offset[3][5];     /* SaoOffsetVal[0..4] */
} sao_params;

typedef struct {
  sps_t sps[16]; pps_t pps[64];
  const sps_t* s; const pps_t* p;
  int W, H, Wc, Hc;             /* coded plane sizes */
  int cfmt, sx, sy;             /* ChromaArrayType; log2 of SubWidthC / SubHeightC (Table 6-1) */
  int ctb, wctb, hctb;
  int min_cb_w, min_cb_h;       /* in min-CB units */
  int w4, h4;                   /* in 4x4 units */
  uint16_t* pl[3];              /* reconstruction */
  int stride[3];
  /* per 4x4 luma block */
  uint16_t* slice_of4;          /* 0 = not yet decoded, else 1 + slice index (slice, not segment) */
  uint8_t* ipm4;                /* intra luma mode */
  int8_t* qp4;                  /* QpY */
  uint8_t* tu_edge4;            /* bit0: left edge is a transform edge, bit1: top edge */
  uint8_t* cd4;                 /* coding quadtree depth */
  uint8_t* nofilt4;             /* 1: samples the in-loop filters must leave unchanged (cu_transquant_bypass; pcm with pcm_loop_filter_disabled) */
  int cu_bypass;                /* cu_transquant_bypass_flag of the current coding unit */
  /* per slice tables */
  int nslices;
  struct { int addr_rs, lf_across, deblock_disabled, beta_offset, tc_offset, cb_off, cr_off; } sl[1024];
  sao_params* sao;              /* per CTB */
  uint8_t* ctb_slice_sao;       /* bit0 luma on, bit1 chroma on (from slice header) */
  /* slice segment state */
  int slice_qp, sao_luma, sao_chroma, slice_idx, slice_addr_rs, cur_cb_off, cur_cr_off;
  /* CABAC */
  bitrd br; unsigned range, offset;
  cabac_ctx ctx[CTX_COUNT], ctx_wpp[CTX_COUNT];
  /* QP state */
  int is_cu_qp_delta_coded, cu_qp_delta_val, qpy_prev_qg, last_cu_qpy, first_qg_in_row;
  int col_bd[21], row_bd[23], ntc, ntr;   /* 6.5.1: tile column / row boundaries in CTBs */
  int* ts2rs; int* rs2ts; int* tile_of_ctb;   /* CtbAddrTsToRs, CtbAddrRsToTs, TileId (indexed by raster address) */
  int cur_tile;
  int cur_qpy;
  int picture_started;
  int err;
  unsigned long long coef_hash, coef_count, tu_count;   /* debug: order-independent digest of the parsed levels */
  uint8_t* cmode4;                                      /* debug: chroma mode per 4x4 */
} dec_t;

/* ------------------------------------------------------------------------------------------ CABAC */
static void cabac_init_engine(dec_t* d) { d->range = 510; d->offset = rd_bits(&d->br, 9); }    /* 9.3.2.5 */

static void cabac_init_contexts(dec_t* d) {                                                        /* 9.3.2.2 */
  int qp = d->slice_qp < 0 ? 0 : (d->slice_qp > 51 ? 51 : d->slice_qp);
  for (int i = 0; i < CTX_COUNT; i++) {
    int iv = ctx_init_I[i];
    int m = (iv >> 4) * 5 - 45, n = ((iv & 15) << 3) - 16;
    int pre = ((m * qp) >> 4) + n;
    pre = pre < 1 ? 1 : (pre > 126 ? 126 : pre);
    d->ctx[i].mps = pre > 63;
    d->ctx[i].state = d->ctx[i].mps ? pre - 64 : 63 - pre;
  }
}

static int dec_bin(dec_t* d, int ci) {                                                             /* 9.3.4.3.2 */
  cabac_ctx* c = &d->ctx[ci];
  unsigned lps = range_tab_lps[c->state][(d->range >> 6) & 3];
  int bin;
  d->range -= lps;
  if (d->offset >= d->range) {
    bin = !c->mps;
    d->offset -= d->range; d->range = lps;
    if (c->state == 0) c->mps = 1 - c->mps;
    c->state = trans_lps[c->state];
  } else {
    bin = c->mps;
    if (c->state < 62) c->state++;
  }
  while (d->range < 256) { d->range <<= 1; d->offset = (d->offset << 1) | rd_bit(&d->br); }     /* 9.3.4.3.3 */
  return bin;
}
static int dec_bypass(dec_t* d) {                                                                  /* 9.3.4.3.4 */
  d->offset = (d->offset << 1) | rd_bit(&d->br);
  if (d->offset >= d->range) { d->offset -= d->range; return 1; }
  return 0;
}
static int dec_terminate(dec_t* d) {                                                               /* 9.3.4.3.5 */
  d->range -= 2;
  if (d->offset >= d->range) return 1;
  while (d->range < 256) { d->range <<= 1; d->offset = (d->offset << 1) | rd_bit(&d->br); }
  return 0;
}
static unsigned dec_bypass_bits(dec_t* d, int n) { unsigned v = 0; while (n-- > 0) v = (v << 1) | dec_bypass(d); return v; }

/* After a terminating bin equal to 1 the 9-bit window of 9.3.4.3 has consumed every bit the encoder's
   flush wrote (9.3.4.5), the last of which is the '1' that plays the role of alignment_bit_equal_to_one:
   the next sub-stream starts at the next byte boundary (checked against entry_point_offset on
   examples/example.heic, which has a row ending exactly on a byte boundary). */
static void cabac_byte_align_and_restart(dec_t* d) {
  d->br.pos = (d->br.pos + 7) & ~(size_t)7;
  cabac_init_engine(d);
}

static inline int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int iabs(int v) { return v < 0 ? -v : v; }

static int tile_of_xy(const dec_t* d, int x, int y) { return d->tile_of_ctb[(y >> d->s->log2_ctb) * d->wctb + (x >> d->s->log2_ctb)]; }

/* 6.4.1 z-scan availability, restated with the "already decoded in the same slice" map */
static int avail4(const dec_t* d, int x, int y) {
  if (x < 0 || y < 0 || x >= d->W || y >= d->H) return 0;
  unsigned s = d->slice_of4[(y >> 2) * d->w4 + (x >> 2)];
  if (d->p->tiles && tile_of_xy(d, x, y) != d->cur_tile) return 0;            /* a different tile is never available */
  return s != 0 && s == (unsigned)(d->slice_idx + 1);
}

/* scan orders 6.5.3 (diagonal), 6.5.4 (horizontal), 6.5.5 (vertical); index [log2 size 1..3][scanIdx][pos] */
static uint8_t scan_x[4][3][64], scan_y[4][3][64];
static int scan_ready;
static void init_scans(void) {
  if (scan_ready) return;
  for (int l = 0; l <= 3; l++) {
    int n = 1 << l, i = 0, x = 0, y = 0, stop = 0;
    while (!stop) {                       /* up-right diagonal */
      while (y >= 0) { if (x < n && y < n) { scan_x[l][0][i] = x; scan_y[l][0][i] = y; i++; } y--; x++; }
      y = x; x = 0;
      if (i >= n * n) stop = 1;
    }
    i = 0;
    for (y = 0; y < n; y++) for (x = 0; x < n; x++) { scan_x[l][1][i] = x; scan_y[l][1][i] = y; i++; }
    i = 0;
    for (x = 0; x < n; x++) for (y = 0; y < n; y++) { scan_x[l][2][i] = x; scan_y[l][2][i] = y; i++; }
  }
  scan_ready = 1;
}

/* ------------------------------------------------------------------------------ inverse transforms */
static const int8_t dct_t[32] = {64, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67,
                                 64, 61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9, 4};
/* transMatrix (8.6.4.2, equation 8-?) entry for an nTbS-point transform: row k, column n */
static int dct_coef(int log2n, int k, int n) {
  if (k == 0) return 64;
  int j = ((k << (5 - log2n)) * (2 * n + 1)) & 127, sgn = 1;
  if (j > 64) j = 128 - j;
  if (j > 32) { j = 64 - j; sgn = -1; }
  return sgn * dct_t[j];
}
static const int8_t dst4[4][4] = {{29, 55, 74, 84}, {74, 74, 0, -74}, {84, -29, -74, 55}, {55, -84, 74, -29}};

/* 8.6.2 + 8.6.4: residual from scaled coefficients; coef is row-major [y][x] */
static void inverse_transform(const int16_t* coef, int* res, int log2n, int bit_depth, int use_dst, int tskip) {
  int n = 1 << log2n;
  int bd_shift = 20 - bit_depth;
  if (tskip) {                                            /* 8.6.4.2 with transform_skip_flag: r = d << 7 */
    for (int i = 0; i < n * n; i++) res[i] = (((int)coef[i] << 7) + (1 << (bd_shift - 1))) >> bd_shift;
    return;
  }
  int tmp[32 * 32];
  for (int x = 0; x < n; x++)                              /* first stage: columns */
    for (int y = 0; y < n; y++) {
      int e = 0;
      for (int k = 0; k < n; k++) {
        int c = coef[k * n + x];
        if (c) e += c * (use_dst ? dst4[k][y] : dct_coef(log2n, k, y));
      }
      tmp[y * n + x] = clip3(-32768, 32767, (e + 64) >> 7);
    }
  for (int y = 0; y < n; y++)                              /* second stage: rows */
    for (int x = 0; x < n; x++) {
      int e = 0;
      for (int k = 0; k < n; k++) {
        int c = tmp[y * n + k];
        if (c) e += c * (use_dst ? dst4[k][x] : dct_coef(log2n, k, x));
      }
      res[y * n + x] = (e + (1 << (bd_shift - 1))) >> bd_shift;
    }
}

/* --------------------------------------------------------------------------------- intra prediction */
static const int8_t intra_angle[35] = {0, 0, 32, 26, 21, 17, 13, 9, 5, 2, 0, -2, -5, -9, -13, -17, -21, -26, -32,
                                       -26, -21, -17, -13, -9, -5, -2, 0, 2, 5, 9, 13, 17, 21, 26, 32};
static const int16_t inv_angle[35] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, -4096, -1638, -910, -630, -482, -390, -315, -256,
                                      -315, -390, -482, -630, -910, -1638, -4096, 0, 0, 0, 0, 0, 0, 0, 0, 0};

/* 8.4.4.2: predict one nTbS x nTbS block of component c at plane position (x0,y0) */
static void intra_predict(dec_t* d, int c, int x0, int y0, int log2n, int mode) {
  const int n = 1 << log2n, bd = d->s->bit_depth;
  uint16_t* pl = d->pl[c];
  const int st = d->stride[c];
  const int shx = c ? d->sx : 0, shy = c ? d->sy : 0;     /* chroma subsampling shifts */
  int refbuf[4 * 32 + 1], fbuf[4 * 32 + 1];
  uint8_t av[4 * 32 + 1];
  /* linear neighbour array: index 0 = p[-1][2n-1] (bottom of left column) ... 2n = p[-1][-1] ... 4n = p[2n-1][-1] */
  int any = 0;
  for (int i = 0; i <= 4 * n; i++) {
    int px, py;
    if (i < 2 * n) { px = x0 - 1; py = y0 + 2 * n - 1 - i; }
    else if (i == 2 * n) { px = x0 - 1; py = y0 - 1; }
    else { px = x0 + (i - 2 * n - 1); py = y0 - 1; }
    av[i] = avail4(d, px << shx, py << shy);        /* 8.4.4.2.2 via 6.4.1 on the luma location */
    if (av[i]) { refbuf[i] = pl[py * st + px]; any = 1; }
  }
  if (!any) { for (int i = 0; i <= 4 * n; i++) refbuf[i] = 1 << (bd - 1); }   /* 8.4.4.2.2 substitution */
  else {
    int first = 0;
    while (!av[first]) first++;
    for (int i = 0; i < first; i++) refbuf[i] = refbuf[first];
    for (int i = first + 1; i <= 4 * n; i++) if (!av[i]) refbuf[i] = refbuf[i - 1];
  }
  int* ref = refbuf;
  /* 8.4.4.2.3 filtering of neighbouring samples (luma only for 4:2:0) */
  if ((c == 0 || d->cfmt == 3) && mode != 1 && n != 4) {     /* luma, and chroma when ChromaArrayType == 3 */
    int dist = iabs(mode - 26) < iabs(mode - 10) ? iabs(mode - 26) : iabs(mode - 10);
    int thr = n == 8 ? 7 : (n == 16 ? 1 : 0);
    if (dist > thr) {
      int corner = ref[2 * n], bl = ref[0], tr = ref[4 * n];
      if (d->s->strong_intra_smoothing && c == 0 && n == 32 &&
          iabs(corner + tr - 2 * ref[2 * n + n]) < (1 << (bd - 5)) &&
          iabs(corner + bl - 2 * ref[2 * n - n]) < (1 << (bd - 5))) {
        fbuf[2 * n] = corner; fbuf[0] = bl; fbuf[4 * n] = tr;
        for (int y = 0; y < 63; y++) fbuf[2 * n - 1 - y] = ((63 - y) * corner + (y + 1) * bl + 32) >> 6;
        for (int x = 0; x < 63; x++) fbuf[2 * n + 1 + x] = ((63 - x) * corner + (x + 1) * tr + 32) >> 6;
      } else {
        fbuf[0] = ref[0]; fbuf[4 * n] = ref[4 * n];
        for (int i = 1; i < 4 * n; i++) fbuf[i] = (ref[i - 1] + 2 * ref[i] + ref[i + 1] + 2) >> 2;
      }
      ref = fbuf;
    }
  }
#define LEFT(y) ref[2 * n - 1 - (y)]   /* p[-1][y], y = -1 .. 2n-1 */
#define TOP(x) ref[2 * n + 1 + (x)]    /* p[x][-1], x = -1 .. 2n-1 */
  const int maxv = (1 << bd) - 1;
  if (mode == 0) {                                        /* 8.4.4.2.4 planar */
    for (int y = 0; y < n; y++) for (int x = 0; x < n; x++)
      pl[(y0 + y) * st + x0 + x] = ((n - 1 - x) * LEFT(y) + (x + 1) * TOP(n) + (n - 1 - y) * TOP(x) + (y + 1) * LEFT(n) + n) >> (log2n + 1);
  } else if (mode == 1) {                                 /* 8.4.4.2.5 DC */
    int sum = n;
    for (int i = 0; i < n; i++) sum += LEFT(i) + TOP(i);
    int dc = sum >> (log2n + 1);
    for (int y = 0; y < n; y++) for (int x = 0; x < n; x++) pl[(y0 + y) * st + x0 + x] = dc;
    if (c == 0 && n < 32) {
      pl[y0 * st + x0] = (LEFT(0) + 2 * dc + TOP(0) + 2) >> 2;
      for (int x = 1; x < n; x++) pl[y0 * st + x0 + x] = (TOP(x) + 3 * dc + 2) >> 2;
      for (int y = 1; y < n; y++) pl[(y0 + y) * st + x0] = (LEFT(y) + 3 * dc + 2) >> 2;
    }
  } else {                                                /* 8.4.4.2.6 angular */
    int ang = intra_angle[mode], ia = inv_angle[mode];
    int rbuf[3 * 32 + 2]; int* r = rbuf + 32;             /* r[-n .. 2n] */
    if (mode >= 18) {
      for (int x = 0; x <= n; x++) r[x] = TOP(x - 1);
      if (ang < 0) { int last = (n * ang) >> 5; if (last < -1) for (int x = last; x <= -1; x++) r[x] = LEFT(-1 + ((x * ia + 128) >> 8)); }
      else for (int x = n + 1; x <= 2 * n; x++) r[x] = TOP(x - 1);
      for (int y = 0; y < n; y++) {
        int idx = ((y + 1) * ang) >> 5, f = ((y + 1) * ang) & 31;
        for (int x = 0; x < n; x++)
          pl[(y0 + y) * st + x0 + x] = f ? ((32 - f) * r[x + idx + 1] + f * r[x + idx + 2] + 16) >> 5 : r[x + idx + 1];
      }
      if (mode == 26 && c == 0 && n < 32)
        for (int y = 0; y < n; y++) pl[(y0 + y) * st + x0] = clip3(0, maxv, TOP(0) + ((LEFT(y) - LEFT(-1)) >> 1));
    } else {
      for (int x = 0; x <= n; x++) r[x] = LEFT(x - 1);
      if (ang < 0) { int last = (n * ang) >> 5; if (last < -1) for (int x = last; x <= -1; x++) r[x] = TOP(-1 + ((x * ia + 128) >> 8)); }
      else for (int x = n + 1; x <= 2 * n; x++) r[x] = LEFT(x - 1);
      for (int x = 0; x < n; x++) {
        int idx = ((x + 1) * ang) >> 5, f = ((x + 1) * ang) & 31;
        for (int y = 0; y < n; y++)
          pl[(y0 + y) * st + x0 + x] = f ? ((32 - f) * r[y + idx + 1] + f * r[y + idx + 2] + 16) >> 5 : r[y + idx + 1];
      }
      if (mode == 10 && c == 0 && n < 32)
        for (int x = 0; x < n; x++) pl[y0 * st + x0 + x] = clip3(0, maxv, LEFT(0) + ((TOP(x) - TOP(-1)) >> 1));
    }
  }
#undef LEFT
#undef TOP
}

/* ------------------------------------------------------------------------------- residual_coding */
static const uint8_t sig_ctx_map4[16] = {0, 1, 4, 5, 2, 3, 4, 5, 6, 6, 8, 8, 7, 7, 8, 8};
static const uint8_t qpc_tab[14] = {29, 30, 31, 32, 33, 33, 34, 34, 35, 35, 36, 36, 37, 37};
static const uint8_t level_scale[6] = {40, 45, 51, 57, 64, 72};

static int chroma_qp(const dec_t* d, int qpy, int off) {                     /* 8.6.1, ChromaArrayType 1 */
  int qbd = 6 * (d->s->bit_depth - 8);
  int qpi = clip3(-qbd, 57, qpy + off);
  int qpc = d->cfmt != 1 ? (qpi < 51 ? qpi : 51) : (qpi < 30 ? qpi : (qpi >= 43 ? qpi - 6 : qpc_tab[qpi - 30]));   /* Table 8-10 only when ChromaArrayType == 1 */
  return qpc + qbd;
}

/* 7.3.8.11 residual_coding + 8.6.3 scaling + 8.6.4 transform + 8.6.6 reconstruction */
static void residual_coding(dec_t* d, int x0, int y0, int log2n, int c, int pred_mode) {
  const int n = 1 << log2n;
  int16_t coef[32 * 32];
  memset(coef, 0, sizeof(int16_t) * n * n);
  int tskip = 0;
  if (d->p->transform_skip && log2n == 2 && !d->cu_bypass) tskip = dec_bin(d, CTX_TSKIP + (c ? 1 : 0));   /* 7.3.8.11 */
  /* last significant coefficient position, 9.3.4.2.3 */
  int cmax = (log2n << 1) - 1, ctx_off, ctx_shift;
  if (c == 0) { ctx_off = 3 * (log2n - 2) + ((log2n - 1) >> 2); ctx_shift = (log2n + 1) >> 2; }
  else { ctx_off = 15; ctx_shift = log2n - 2; }
  int lx = 0, ly = 0;
  while (lx < cmax && dec_bin(d, CTX_LAST_X + ctx_off + (lx >> ctx_shift))) lx++;
  while (ly < cmax && dec_bin(d, CTX_LAST_Y + ctx_off + (ly >> ctx_shift))) ly++;
  if (lx > 3) { int nb = (lx >> 1) - 1; lx = (1 << nb) * (2 + (lx & 1)) + dec_bypass_bits(d, nb); }
  if (ly > 3) { int nb = (ly >> 1) - 1; ly = (1 << nb) * (2 + (ly & 1)) + dec_bypass_bits(d, nb); }
  /* scanIdx, 7.4.9.11 */
  int scan = 0;
  if (log2n == 2 || (log2n == 3 && (c == 0 || d->cfmt == 3))) {
    if (pred_mode >= 6 && pred_mode <= 14) scan = 2;
    else if (pred_mode >= 22 && pred_mode <= 30) scan = 1;
  }
  if (scan == 2) { int t = lx; lx = ly; ly = t; }
  const int l2sb = log2n - 2;
  const uint8_t *sbx = scan_x[l2sb][scan], *sby = scan_y[l2sb][scan], *px = scan_x[2][scan], *py = scan_y[2][scan];
  int last_sb = 0, last_pos = 0;
  { int nsb = 1 << (2 * l2sb), found = 0;
    for (int i = 0; i < nsb && !found; i++) for (int k = 0; k < 16; k++)
      if ((sbx[i] << 2) + px[k] == lx && (sby[i] << 2) + py[k] == ly) { last_sb = i; last_pos = k; found = 1; break; }
  }
  uint8_t csbf[8][8];
  memset(csbf, 0, sizeof csbf);
  int greater1_ctx_carry = 1;        /* "lastGreater1Ctx" carried between sub-blocks, 9.3.4.2.6 */
  int first_sb_processed = 1;
  for (int i = last_sb; i >= 0; i--) {
    int xs = sbx[i], ys = sby[i];
    int infer_dc = 0, coded;
    if (i < last_sb && i > 0) {
      int cs = 0;
      if (xs + 1 < (1 << l2sb)) cs |= csbf[ys][xs + 1];
      if (ys + 1 < (1 << l2sb)) cs |= csbf[ys + 1][xs];
      coded = dec_bin(d, CTX_CSBF + (cs ? 1 : 0) + (c ? 2 : 0));
      infer_dc = 1;
    } else coded = 1;
    csbf[ys][xs] = coded;
    if (!coded) continue;
    int sig[16]; memset(sig, 0, sizeof sig);
    int prev_csbf = 0;
    if (xs + 1 < (1 << l2sb)) prev_csbf |= csbf[ys][xs + 1];
    if (ys + 1 < (1 << l2sb)) prev_csbf |= csbf[ys + 1][xs] << 1;
    int start = (i == last_sb) ? last_pos - 1 : 15;
    if (i == last_sb) sig[last_pos] = 1;
    for (int k = start; k >= 0; k--) {
      int xc = (xs << 2) + px[k], yc = (ys << 2) + py[k];
      if (k > 0 || !infer_dc) {
        int sc;                                                /* 9.3.4.2.5 */
        if (log2n == 2) sc = sig_ctx_map4[(yc << 2) + xc];
        else if (xc + yc == 0) sc = 0;
        else {
          int xp = xc & 3, yp = yc & 3;
          if (prev_csbf == 0) sc = (xp + yp == 0) ? 2 : (xp + yp < 3) ? 1 : 0;
          else if (prev_csbf == 1) sc = yp == 0 ? 2 : (yp == 1 ? 1 : 0);
          else if (prev_csbf == 2) sc = xp == 0 ? 2 : (xp == 1 ? 1 : 0);
          else sc = 2;
          if (c == 0) { if (xs || ys) sc += 3; sc += (log2n == 3) ? (scan == 0 ? 9 : 15) : 21; }
          else sc += (log2n == 3) ? 9 : 12;
        }
        sig[k] = dec_bin(d, CTX_SIG + (c == 0 ? sc : 27 + sc));
        if (sig[k]) infer_dc = 0;
      } else sig[k] = 1;  /* k == 0 && infer_dc: inferred significant */
    }
    /* greater1 / greater2, 9.3.4.2.6-7 */
    int g1[16] = {0}, g2[16] = {0};
    int first_sig = 16, last_sig = -1, ng1 = 0, last_g1_pos = -1;
    int ctx_set = (i == 0 || c > 0) ? 0 : 2;
    if (!first_sb_processed && greater1_ctx_carry == 0) ctx_set++;
    first_sb_processed = 0;
    int g1ctx = 1, any_sig = 0;
    for (int k = 15; k >= 0; k--) if (sig[k]) {
      any_sig = 1;
      if (ng1 < 8) {
        g1[k] = dec_bin(d, CTX_GT1 + ctx_set * 4 + (g1ctx > 3 ? 3 : g1ctx) + (c ? 16 : 0));
        ng1++;
        if (g1[k]) { g1ctx = 0; if (last_g1_pos < 0) last_g1_pos = k; }
        else if (g1ctx > 0) g1ctx++;
      }
      if (last_sig < 0) last_sig = k;
      first_sig = k;
    }
    if (any_sig) greater1_ctx_carry = g1ctx;
    int sign_hidden = d->p->sign_hiding && !d->cu_bypass && (last_sig - first_sig > 3);
    if (last_g1_pos >= 0) g2[last_g1_pos] = dec_bin(d, CTX_GT2 + ctx_set + (c ? 4 : 0));
    int sign[16] = {0};
    for (int k = 15; k >= 0; k--) if (sig[k] && (!sign_hidden || k != first_sig)) sign[k] = dec_bypass(d);
    int nsig = 0, sum_abs = 0, rice = 0;
    for (int k = 15; k >= 0; k--) if (sig[k]) {
      int base = 1 + g1[k] + g2[k];
      int absl = base;
      if (base == ((nsig < 8) ? ((k == last_g1_pos) ? 3 : 2) : 1)) {
        int pre = 0;                                         /* 9.3.3.11 coeff_abs_level_remaining */
        while (pre < 32 && dec_bypass(d)) pre++;
        if (pre > 20) { d->err = HO_ERROR; return; }          /* a value far outside the 16-bit range of TransCoeffLevel (7.4.9.11): corrupt data */
        int rem;
        if (pre <= 3) rem = (pre << rice) + dec_bypass_bits(d, rice);
        else rem = (((1 << (pre - 3)) + 3 - 1) << rice) + dec_bypass_bits(d, pre - 3 + rice);
        absl = base + rem;
        if (absl > 3 * (1 << rice)) rice = rice < 4 ? rice + 1 : 4;
      }
      int v = sign[k] ? -absl : absl;
      if (sign_hidden) { sum_abs += absl; if (k == first_sig && (sum_abs & 1)) v = -v; }
      int xc = (xs << 2) + px[k], yc = (ys << 2) + py[k];
      coef[yc * n + xc] = (int16_t)clip3(-32768, 32767, v);
      { unsigned long long hh = ((unsigned long long)(x0 << (c ? d->sx : 0)) * 1000003ULL + (unsigned long long)(y0 << (c ? d->sy : 0))) * 1000003ULL + (unsigned long long)c; hh = hh * 1000003ULL + (unsigned long long)(yc * n + xc); hh = hh * 1000003ULL + (unsigned long long)(unsigned short)coef[yc * n + xc]; hh ^= hh >> 29; hh *= 0x9E3779B97F4A7C15ULL; d->coef_hash += hh; d->coef_count++; }
      nsig++;
    }
  }
  int16_t raw[32 * 32]; memcpy(raw, coef, sizeof(int16_t) * n * n);
  /* 8.6.3 / 8.6.4.2 scaling: m = 16 without scaling lists, else ScalingFactor[sizeId][matrixId][x][y] (7.4.5): the list in
     effect (PPS, else SPS, else default) mapped through the up-right diagonal scan, 8x8 lists replicated for 16x16 / 32x32
     with scaling_list_dc_coef at (0, 0); matrixId = cIdx for intra blocks (32x32: matrixId 0) */
  const int bd = d->s->bit_depth;
  int qp = c == 0 ? d->cur_qpy + 6 * (bd - 8)
                  : chroma_qp(d, d->cur_qpy, (c == 1 ? d->p->cb_qp_offset + d->cur_cb_off : d->p->cr_qp_offset + d->cur_cr_off));
  int bd_shift = bd + log2n - 5;
  int scale = level_scale[qp % 6] << (qp / 6);
  const uint8_t (*sl)[6][64] = NULL; const uint8_t (*sldc)[6] = NULL;
  uint8_t dsl[4][6][64], ddc[4][6];
  if (d->s->scaling_list_enabled) {
    if (d->p->scaling_list_present) { sl = d->p->sl; sldc = d->p->sl_dc; }
    else if (d->s->sl_data_present) { sl = d->s->sl; sldc = d->s->sl_dc; }
    else { for (int a = 0; a < 4; a++) for (int bq = 0; bq < 6; bq++) sl_default(dsl, ddc, a, bq); sl = dsl; sldc = ddc; }
  }
  for (int k = 0; k < n * n; k++) if (coef[k]) {
    int m = 16;
    if (sl) {
      /* 32x32 chroma (ChromaArrayType == 3 only): the 16x16 list of its matrixId, replicated 4x4, with that list's DC (7.4.5) */
      const int sizeId = (log2n == 5 && c > 0) ? 2 : log2n - 2, matrixId = (log2n == 5 && c == 0) ? 0 : c, x = k & (n - 1), y = k >> log2n;
      const int l2 = sizeId == 0 ? 2 : 3, rep = log2n <= 3 ? 0 : log2n - 3;        /* replication shift of the 8x8 list */
      const int xs = x >> rep, ys = y >> rep;
      /* scan index i of (xs, ys) in the up-right diagonal scan of the (1 << l2) square (6.5.3) */
      int i = 0;
      { int xx = 0, yy = 0, found = 0, sz = 1 << l2;
        while (!found) { while (yy >= 0 && !found) { if (xx < sz && yy < sz) { if (xx == xs && yy == ys) found = 1; else i++; } if (!found) { yy--; xx++; } } if (!found) { yy = xx; xx = 0; } } }
      m = sl[sizeId][matrixId][i];
      if (log2n >= 4 && x == 0 && y == 0) m = sldc[sizeId][matrixId];
    }
    long long t = ((long long)coef[k] * m * scale + (1LL << (bd_shift - 1))) >> bd_shift;
    coef[k] = (int16_t)(t < -32768 ? -32768 : (t > 32767 ? 32767 : t));
  }
  int res[32 * 32];
  inverse_transform(coef, res, log2n, bd, c == 0 && log2n == 2, tskip);
  if (d->cu_bypass) for (int k = 0; k < n * n; k++) res[k] = raw[k];     /* 8.6.2: cu_transquant_bypass_flag -> r = TransCoeffLevel */
  uint16_t* pl = d->pl[c]; int st = d->stride[c], maxv = (1 << bd) - 1;
  for (int y = 0; y < n; y++) for (int x = 0; x < n; x++) {
    uint16_t* q = &pl[(y0 + y) * st + x0 + x];
    *q = (uint16_t)clip3(0, maxv, *q + res[y * n + x]);
  }
}

/* ------------------------------------------------------------------------------------ QP, 8.6.1 */
static void derive_qpy(dec_t* d, int xcb, int ycb) {
  int qgmask = (1 << (d->s->log2_ctb - d->p->diff_cu_qp_delta_depth)) - 1;
  int xqg = xcb & ~qgmask, yqg = ycb & ~qgmask;
  int prev = d->qpy_prev_qg;
  int ctbmask = ~((1 << d->s->log2_ctb) - 1);
  int qa = prev, qb = prev;
  if (avail4(d, xqg - 1, yqg) && ((xqg - 1) & ctbmask) == (xqg & ctbmask)) qa = d->qp4[(yqg >> 2) * d->w4 + ((xqg - 1) >> 2)];
  if (avail4(d, xqg, yqg - 1) && ((yqg - 1) & ctbmask) == (yqg & ctbmask)) qb = d->qp4[((yqg - 1) >> 2) * d->w4 + (xqg >> 2)];
  int pred = (qa + qb + 1) >> 1;
  int qbd = 6 * (d->s->bit_depth - 8);
  d->cur_qpy = ((pred + d->cu_qp_delta_val + 52 + 2 * qbd) % (52 + qbd)) - qbd;
  if (getenv("HO_DEBUG2") && ycb >= 0 && ycb < 128 && xcb < 128) fprintf(stderr, "qp cu(%d,%d) qg(%d,%d) prev=%d qa=%d qb=%d delta=%d -> %d slice_qp=%d\n", xcb, ycb, xqg, yqg, prev, qa, qb, d->cu_qp_delta_val, d->cur_qpy, d->slice_qp);
}

/* ---------------------------------------------------------------------------- transform tree etc. */
typedef struct { int x0, y0, log2cb, part_nxn, luma_mode[4], chroma_mode[4]; } cu_t;   /* chroma_mode: per prediction unit (one per PU only when ChromaArrayType == 3) */

static void mark_tu(dec_t* d, int x0, int y0, int log2n) {
  int n4 = 1 << (log2n - 2), bx = x0 >> 2, by = y0 >> 2;
  for (int y = 0; y < n4 && by + y < d->h4; y++) for (int x = 0; x < n4 && bx + x < d->w4; x++) {
    int i = (by + y) * d->w4 + bx + x;
    d->slice_of4[i] = (uint16_t)(d->slice_idx + 1);
    d->qp4[i] = (int8_t)d->cur_qpy;
    if (x == 0) d->tu_edge4[i] |= 1;
    if (y == 0) d->tu_edge4[i] |= 2;
  }
}

/* cbf_cb / cbf_cr: the chroma flags that apply to this unit ([1]: lower block of a 4:2:2 unit) -- its own, or for a 4x4 luma block in
   4:2:0 / 4:2:2 those of the parent 8x8 node (7.3.8.10) */
static void transform_unit(dec_t* d, cu_t* cu, int x0, int y0, int xb, int yb, int log2n, int depth, int blk,
                           int cbf_luma, const int cbf_cb[2], const int cbf_cr[2]) {
  const int cfmt = d->cfmt;
  int cbf_chroma = cfmt ? (cbf_cb[0] | cbf_cb[1] | cbf_cr[0] | cbf_cr[1]) : 0;
  (void)depth;
  if ((cbf_luma || cbf_chroma) && d->p->cu_qp_delta && !d->is_cu_qp_delta_coded) {   /* 7.3.8.10 */
    int v = 0;
    while (v < 5 && dec_bin(d, CTX_QP_DELTA + (v ? 1 : 0))) v++;
    if (v == 5) { int k = 0; while (k < 16 && dec_bypass(d)) { v += 1 << k; k++; } v += dec_bypass_bits(d, k); } /* EG0 */
    if (v && dec_bypass(d)) v = -v;
    d->is_cu_qp_delta_coded = 1;
    d->cu_qp_delta_val = v;
    derive_qpy(d, cu->x0, cu->y0);
  }
  int pu = cu->part_nxn ? ((y0 >= cu->y0 + (1 << (cu->log2cb - 1))) ? 2 : 0) + ((x0 >= cu->x0 + (1 << (cu->log2cb - 1))) ? 1 : 0) : 0;
  int lmode = cu->luma_mode[pu];
  int cmode = cu->chroma_mode[cfmt == 3 ? pu : 0];
  d->tu_count++;
  { int n4 = 1 << (log2n - 2); for (int yy = 0; yy < n4; yy++) for (int xx = 0; xx < n4; xx++) d->cmode4[((y0 >> 2) + yy) * d->w4 + (x0 >> 2) + xx] = (uint8_t)cmode; }
  /* luma: 8.4.4.1 predict, then residual */
  intra_predict(d, 0, x0, y0, log2n, lmode);
  if (cbf_luma) residual_coding(d, x0, y0, log2n, 0, lmode);
  mark_tu(d, x0, y0, log2n);
  if (cfmt) {
    const int nb = cfmt == 2 ? 2 : 1;               /* 4:2:2: two square blocks, the lower one predicted from the reconstructed upper one */
    if (log2n > 2 || cfmt == 3) {
      const int lc = cfmt == 3 ? log2n : log2n - 1, cx = x0 >> d->sx, cy = y0 >> d->sy;
      for (int c = 1; c <= 2; c++) for (int t = 0; t < nb; t++) {
        intra_predict(d, c, cx, cy + (t << lc), lc, cmode);
        if ((c == 1 ? cbf_cb : cbf_cr)[t]) residual_coding(d, cx, cy + (t << lc), lc, c, cmode);
      }
    } else if (blk == 3) {
      const int cx = xb >> d->sx, cy = yb >> d->sy;
      for (int c = 1; c <= 2; c++) for (int t = 0; t < nb; t++) {
        intra_predict(d, c, cx, cy + (t << 2), 2, cmode);
        if ((c == 1 ? cbf_cb : cbf_cr)[t]) residual_coding(d, cx, cy + (t << 2), 2, c, cmode);
      }
    }
  }
}

static void transform_tree(dec_t* d, cu_t* cu, int x0, int y0, int xb, int yb, int log2n, int depth, int blk,
                           const int parent_cbf_cb[2], const int parent_cbf_cr[2], int max_depth) {
  const sps_t* s = d->s;
  int split;
  int intra_split = cu->part_nxn;
  if (log2n <= s->log2_max_tb && log2n > s->log2_min_tb && depth < max_depth && !(intra_split && depth == 0))
    split = dec_bin(d, CTX_SPLIT_TR + 5 - log2n);
  else
    split = (log2n > s->log2_max_tb || (intra_split && depth == 0)) ? 1 : 0;
  int cbf_cb[2] = {0, 0}, cbf_cr[2] = {0, 0};
  if (d->cfmt) {
    if (log2n > 2 || d->cfmt == 3) {                 /* 7.3.8.8 */
      const int two = d->cfmt == 2 && (!split || log2n == 3);
      if (depth == 0 || parent_cbf_cb[0]) { cbf_cb[0] = dec_bin(d, CTX_CBF_CHROMA + depth); if (two) cbf_cb[1] = dec_bin(d, CTX_CBF_CHROMA + depth); }
      if (depth == 0 || parent_cbf_cr[0]) { cbf_cr[0] = dec_bin(d, CTX_CBF_CHROMA + depth); if (two) cbf_cr[1] = dec_bin(d, CTX_CBF_CHROMA + depth); }
    } else { cbf_cb[0] = parent_cbf_cb[0]; cbf_cb[1] = parent_cbf_cb[1]; cbf_cr[0] = parent_cbf_cr[0]; cbf_cr[1] = parent_cbf_cr[1]; }   /* 4x4 luma: the parent's */
  }
  if (split) {
    int h = 1 << (log2n - 1);
    for (int k = 0; k < 4; k++)
      transform_tree(d, cu, x0 + (k & 1) * h, y0 + (k >> 1) * h, x0, y0, log2n - 1, depth + 1, k, cbf_cb, cbf_cr, max_depth);
  } else {
    int cbf_luma = dec_bin(d, CTX_CBF_LUMA + (depth == 0 ? 1 : 0));
    transform_unit(d, cu, x0, y0, xb, yb, log2n, depth, blk, cbf_luma, cbf_cb, cbf_cr);
  }
}

/* 8.4.2 luma intra prediction mode */
static int derive_luma_mode(dec_t* d, int x, int y, int prev_flag, int mpm_idx, int rem) {
  int ca = 1, cb = 1;
  if (avail4(d, x - 1, y)) ca = d->ipm4[(y >> 2) * d->w4 + ((x - 1) >> 2)];
  if (avail4(d, x, y - 1) && (y - 1) >= ((y >> d->s->log2_ctb) << d->s->log2_ctb)) cb = d->ipm4[((y - 1) >> 2) * d->w4 + (x >> 2)];
  int cand[3];
  if (ca == cb) {
    if (ca < 2) { cand[0] = 0; cand[1] = 1; cand[2] = 26; }
    else { cand[0] = ca; cand[1] = 2 + ((ca + 29) % 32); cand[2] = 2 + ((ca - 2 + 1) % 32); }
  } else {
    cand[0] = ca; cand[1] = cb;
    if (ca != 0 && cb != 0) cand[2] = 0; else if (ca != 1 && cb != 1) cand[2] = 1; else cand[2] = 26;
  }
  if (prev_flag) return cand[mpm_idx];
  if (cand[0] > cand[1]) { int t = cand[0]; cand[0] = cand[1]; cand[1] = t; }
  if (cand[0] > cand[2]) { int t = cand[0]; cand[0] = cand[2]; cand[2] = t; }
  if (cand[1] > cand[2]) { int t = cand[1]; cand[1] = cand[2]; cand[2] = t; }
  int m = rem;
  for (int i = 0; i < 3; i++) if (m >= cand[i]) m++;
  return m;
}

/* 7.3.8.5 coding_unit */
static void coding_unit(dec_t* d, int x0, int y0, int log2cb, int cq_depth) {
  const sps_t* s = d->s;
  cu_t cu; memset(&cu, 0, sizeof cu);
  cu.x0 = x0; cu.y0 = y0; cu.log2cb = log2cb;
  int n = 1 << log2cb;
  d->cu_bypass = d->p->transquant_bypass ? dec_bin(d, CTX_TQ_BYPASS) : 0;                  /* cu_transquant_bypass_flag */
  if (log2cb == s->log2_min_cb) cu.part_nxn = !dec_bin(d, CTX_PART_MODE);
  if (cu.part_nxn && log2cb == 3 && s->log2_min_tb > 2) { d->err = HO_ERROR; return; }
  for (int yy = 0; yy < n; yy += 4) for (int xx = 0; xx < n; xx += 4)
    if (x0 + xx < d->W && y0 + yy < d->H) d->nofilt4[((y0 + yy) >> 2) * d->w4 + ((x0 + xx) >> 2)] = (uint8_t)d->cu_bypass;
  if (s->pcm && !cu.part_nxn && log2cb >= s->log2_min_pcm && log2cb <= s->log2_max_pcm && dec_terminate(d)) {
    /* pcm_flag = 1 (7.3.8.5): pcm_alignment_zero_bits, pcm_sample(), then the arithmetic decoder starts over (9.3.2.5) */
    d->br.pos = (d->br.pos + 7) & ~(size_t)7;
    for (int c = 0; c < (s->chroma_format_idc ? 3 : 1); c++) {
      int shx = c ? d->sx : 0, shy = c ? d->sy : 0, pbd = c ? s->pcm_bd_c : s->pcm_bd_y, bdc = c ? s->bit_depth_c : s->bit_depth, mw = n >> shx, mh = n >> shy;
      for (int y = 0; y < mh; y++) for (int x = 0; x < mw; x++) {
        unsigned v = rd_bits(&d->br, pbd);
        d->pl[c][((y0 >> shy) + y) * d->stride[c] + (x0 >> shx) + x] = (uint16_t)(v << (bdc - pbd));       /* 8.4.4.1: recSamples = pcm_sample << (BitDepth - PcmBitDepth) */
        { unsigned long long hy = (unsigned long long)(y0 + ((d->cfmt == 2 && c && y >= mw) ? mw : 0));   /* 4:2:2: the lower chroma block's own position */
          unsigned long long hh = ((unsigned long long)x0 * 1000003ULL + hy) * 1000003ULL + (unsigned long long)c; hh = hh * 1000003ULL + (unsigned long long)(y * mw + x); hh = hh * 1000003ULL + (unsigned long long)(unsigned short)(v << (bdc - pbd)); hh ^= hh >> 29; hh *= 0x9E3779B97F4A7C15ULL; d->coef_hash += hh; d->coef_count++; }
      }
    }
    cabac_init_engine(d);
    if (!d->p->cu_qp_delta) d->cur_qpy = d->slice_qp; else derive_qpy(d, x0, y0);
    d->tu_count++;
    for (int yy = 0; yy < n; yy += 4) for (int xx = 0; xx < n; xx += 4)
      if (x0 + xx < d->W && y0 + yy < d->H) {
        int idx = ((y0 + yy) >> 2) * d->w4 + ((x0 + xx) >> 2);
        d->ipm4[idx] = 1; d->cd4[idx] = (uint8_t)cq_depth; d->cmode4[idx] = 1;     /* a PCM unit is INTRA_DC for the mode derivation of its neighbours (8.4.2) */
        if (s->pcm_loop_filter_disabled) d->nofilt4[idx] = 1;
      }
    mark_tu(d, x0, y0, log2cb);
    d->last_cu_qpy = d->cur_qpy;
    return;
  }
  int np = cu.part_nxn ? 4 : 1, pb = cu.part_nxn ? n / 2 : n;
  int prev[4], mpm[4] = {0}, rem[4] = {0};
  for (int i = 0; i < np; i++) prev[i] = dec_bin(d, CTX_PREV_INTRA);
  for (int i = 0; i < np; i++) {
    if (prev[i]) { mpm[i] = dec_bypass(d); if (mpm[i]) mpm[i] += dec_bypass(d); }
    else rem[i] = dec_bypass_bits(d, 5);
  }
  /* cqt depth + "decoded" marks are needed by later CUs; modes are derived PU by PU in order */
  for (int i = 0; i < np; i++) {
    int px = x0 + (i & 1) * pb, py = y0 + (i >> 1) * pb;
    int m = derive_luma_mode(d, px, py, prev[i], mpm[i], rem[i]);
    cu.luma_mode[i] = m;
    for (int yy = 0; yy < pb; yy += 4) for (int xx = 0; xx < pb; xx += 4)
      if (px + xx < d->W && py + yy < d->H) d->ipm4[((py + yy) >> 2) * d->w4 + ((px + xx) >> 2)] = (uint8_t)m;
    /* earlier PUs of this CU count as available for the mode derivation of the following ones (6.4.2):
       mark them now; the marks are cleared below and set again TU by TU once pixels are reconstructed */
    for (int yy = 0; yy < pb; yy += 4) for (int xx = 0; xx < pb; xx += 4)
      if (px + xx < d->W && py + yy < d->H) d->slice_of4[((py + yy) >> 2) * d->w4 + ((px + xx) >> 2)] = (uint16_t)(d->slice_idx + 1);
  }
  if (d->cfmt) {
    /* intra_chroma_pred_mode: one per prediction unit when ChromaArrayType == 3, else one (7.3.8.5); 8.4.3, and Table 8-3
       (the 4:2:2 mapping of the mode to the half-width sample grid) */
    static const uint8_t tab[4] = {0, 26, 10, 1};
    static const uint8_t mode422[35] = {0, 1, 2, 2, 2, 2, 3, 5, 7, 8, 10, 12, 13, 15, 17, 18, 19, 20, 21, 22, 23, 23, 24, 24, 25, 25, 26, 27, 27, 28, 28, 29, 29, 30, 31};   /* Table 8-3 as corrected in the later editions (modeIdc 11 -> 12, 14 -> 17); every entry checked against FFmpeg */
    for (int i = 0; i < (d->cfmt == 3 ? np : 1); i++) {
      int v = 4, m;
      if (dec_bin(d, CTX_CHROMA_PRED)) v = dec_bypass_bits(d, 2);
      if (v == 4) m = cu.luma_mode[i];
      else { m = tab[v]; if (m == cu.luma_mode[i]) m = 34; }
      if (d->cfmt == 2) m = mode422[m];
      cu.chroma_mode[i] = m;
    }
  }
  /* undo temporary marks */
  for (int yy = 0; yy < n; yy += 4) for (int xx = 0; xx < n; xx += 4)
    if (x0 + xx < d->W && y0 + yy < d->H) {
      int idx = ((y0 + yy) >> 2) * d->w4 + ((x0 + xx) >> 2);
      d->slice_of4[idx] = 0;
      d->cd4[idx] = (uint8_t)cq_depth;
    }
  /* QpY for a CU without coded delta (so far): predicted value */
  if (!d->p->cu_qp_delta) d->cur_qpy = d->slice_qp;
  else derive_qpy(d, x0, y0);
  int max_depth = s->max_th_depth_intra + cu.part_nxn;
  { const int none[2] = {0, 0}; transform_tree(d, &cu, x0, y0, x0, y0, log2cb, 0, 0, none, none, max_depth); }
  /* the whole CU carries its final QpY (8.6.1; used by deblocking and by QP prediction) */
  for (int yy = 0; yy < n; yy += 4) for (int xx = 0; xx < n; xx += 4)
    if (x0 + xx < d->W && y0 + yy < d->H) d->qp4[((y0 + yy) >> 2) * d->w4 + ((x0 + xx) >> 2)] = (int8_t)d->cur_qpy;
  d->last_cu_qpy = d->cur_qpy;
}

/* 7.3.8.4 coding_quadtree */
static void coding_quadtree(dec_t* d, int x0, int y0, int log2cb, int depth) {
  const sps_t* s = d->s;
  if (d->err) return;
  int n = 1 << log2cb, split;
  if (x0 + n <= d->W && y0 + n <= d->H && log2cb > s->log2_min_cb) {
    int inc = 0;
    if (avail4(d, x0 - 1, y0) && d->cd4[(y0 >> 2) * d->w4 + ((x0 - 1) >> 2)] > depth) inc++;
    if (avail4(d, x0, y0 - 1) && d->cd4[((y0 - 1) >> 2) * d->w4 + (x0 >> 2)] > depth) inc++;
    split = dec_bin(d, CTX_SPLIT_CU + inc);
  } else split = log2cb > s->log2_min_cb;
  if (d->p->cu_qp_delta && log2cb >= s->log2_ctb - d->p->diff_cu_qp_delta_depth) {
    d->is_cu_qp_delta_coded = 0; d->cu_qp_delta_val = 0;
    /* a quantization group starts at the node whose size equals Log2MinCuQpDeltaSize, or at a larger
       unsplit coding block: fix qPY_PREV there (8.6.1) */
    if (!split || log2cb == s->log2_ctb - d->p->diff_cu_qp_delta_depth) {
      if (d->first_qg_in_row) { d->qpy_prev_qg = d->slice_qp; d->first_qg_in_row = 0; }
      else d->qpy_prev_qg = d->last_cu_qpy;
    }
  }
  if (split) {
    int h = n >> 1;
    for (int k = 0; k < 4; k++) {
      int x1 = x0 + (k & 1) * h, y1 = y0 + (k >> 1) * h;
      if (x1 < d->W && y1 < d->H) coding_quadtree(d, x1, y1, log2cb - 1, depth + 1);
    }
  } else coding_unit(d, x0, y0, log2cb, depth);
}

/* 7.3.8.3 sao */
static void parse_sao(dec_t* d, int rx, int ry) {
  const sps_t* s = d->s;
  int addr = ry * d->wctb + rx;
  sao_params* sp = &d->sao[addr];
  memset(sp, 0, sizeof *sp);
  d->ctb_slice_sao[addr] = (uint8_t)((d->sao_luma ? 1 : 0) | (d->sao_chroma ? 2 : 0));
  if (!d->sao_luma && !d->sao_chroma) return;
  int merge_left = 0, merge_up = 0;
  if (rx > 0 && addr - 1 >= d->slice_addr_rs && d->tile_of_ctb[addr - 1] == d->tile_of_ctb[addr]) merge_left = dec_bin(d, CTX_SAO_MERGE);     /* leftCtbInSliceSeg && leftCtbInTile */
  if (ry > 0 && !merge_left && addr - d->wctb >= d->slice_addr_rs && d->tile_of_ctb[addr - d->wctb] == d->tile_of_ctb[addr]) merge_up = dec_bin(d, CTX_SAO_MERGE);
  if (merge_left) { *sp = d->sao[addr - 1]; return; }
  if (merge_up) { *sp = d->sao[addr - d->wctb]; return; }
  int ncomp = s->chroma_format_idc ? 3 : 1;
  for (int c = 0; c < ncomp; c++) {
    if ((c == 0 && !d->sao_luma) || (c > 0 && !d->sao_chroma)) continue;
    if (c < 2) {
      int t = 0;
      if (dec_bin(d, CTX_SAO_TYPE)) t = dec_bypass(d) ? 2 : 1;
      sp->type[c] = t;
    } else { sp->type[2] = sp->type[1]; }
    if (!sp->type[c]) continue;
    int bd = s->bit_depth;
    int cmax = (1 << ((bd < 10 ? bd : 10) - 5)) - 1;
    int absv[4];
    for (int i = 0; i < 4; i++) { int v = 0; while (v < cmax && dec_bypass(d)) v++; absv[i] = v; }
    int scale = c == 0 ? d->p->log2_sao_scale_luma : d->p->log2_sao_scale_chroma;
    if (sp->type[c] == 1) {
      for (int i = 0; i < 4; i++) if (absv[i] && dec_bypass(d)) absv[i] = -absv[i];
      sp->band_pos[c] = dec_bypass_bits(d, 5);
      for (int i = 0; i < 4; i++) sp->offset[c][i + 1] = absv[i] * (1 << scale);
    } else {
      if (c == 0) sp->eo_class[0] = dec_bypass_bits(d, 2);
      else if (c == 1) sp->eo_class[1] = dec_bypass_bits(d, 2);
      else sp->eo_class[2] = sp->eo_class[1];
      sp->offset[c][1] = absv[0] << scale; sp->offset[c][2] = absv[1] << scale;
      sp->offset[c][3] = -(absv[2] << scale); sp->offset[c][4] = -(absv[3] << scale);
    }
  }
}

/* --------------------------------------------------------------------------------- slice decoding */
static int ceil_log2(unsigned v) { int r = 0; while ((1u << r) < v) r++; return r; }

static int alloc_picture(dec_t* d) {
  const sps_t* s = d->s;
  d->W = s->width; d->H = s->height;
  d->ctb = 1 << s->log2_ctb;
  d->wctb = (d->W + d->ctb - 1) >> s->log2_ctb; d->hctb = (d->H + d->ctb - 1) >> s->log2_ctb;
  d->w4 = (d->W + 3) >> 2; d->h4 = (d->H + 3) >> 2;
  d->cfmt = s->chroma_format_idc; d->sx = (d->cfmt == 1 || d->cfmt == 2) ? 1 : 0; d->sy = d->cfmt == 1 ? 1 : 0;
  d->Wc = d->cfmt ? d->W >> d->sx : 0; d->Hc = d->cfmt ? d->H >> d->sy : 0;
  d->stride[0] = d->W; d->stride[1] = d->stride[2] = d->Wc;
  d->pl[0] = (uint16_t*)calloc((size_t)d->W * d->H, 2);
  for (int c = 1; c < 3; c++) d->pl[c] = d->Wc ? (uint16_t*)calloc((size_t)d->Wc * d->Hc, 2) : NULL;
  size_t n4 = (size_t)d->w4 * d->h4;
  d->slice_of4 = (uint16_t*)calloc(n4, 2);
  d->ipm4 = (uint8_t*)calloc(n4, 1); d->qp4 = (int8_t*)calloc(n4, 1);
  d->tu_edge4 = (uint8_t*)calloc(n4, 1); d->cd4 = (uint8_t*)calloc(n4, 1); d->nofilt4 = (uint8_t*)calloc(n4, 1);
  d->cmode4 = (uint8_t*)calloc(n4, 1);
  d->sao = (sao_params*)calloc((size_t)d->wctb * d->hctb, sizeof(sao_params));
  d->ctb_slice_sao = (uint8_t*)calloc((size_t)d->wctb * d->hctb, 1);
  d->nslices = 0;
  d->picture_started = 1;
  /* 6.5.1: tile boundaries, CtbAddrRsToTs / TsToRs, TileId */
  { const pps_t* p = d->p; int nctb = d->wctb * d->hctb;
    d->ntc = p->num_tile_cols; d->ntr = p->num_tile_rows;
    if (d->ntc > d->wctb || d->ntr > d->hctb) return HO_ERROR;
    d->col_bd[0] = d->row_bd[0] = 0;
    for (int i = 0; i < d->ntc; i++) {
      int w = (!p->tiles || p->uniform_spacing) ? ((i + 1) * d->wctb) / d->ntc - (i * d->wctb) / d->ntc : (i + 1 < d->ntc ? p->col_width[i] : d->wctb - d->col_bd[i]);
      if (w <= 0) return HO_ERROR;
      d->col_bd[i + 1] = d->col_bd[i] + w; }
    for (int i = 0; i < d->ntr; i++) {
      int h = (!p->tiles || p->uniform_spacing) ? ((i + 1) * d->hctb) / d->ntr - (i * d->hctb) / d->ntr : (i + 1 < d->ntr ? p->row_height[i] : d->hctb - d->row_bd[i]);
      if (h <= 0) return HO_ERROR;
      d->row_bd[i + 1] = d->row_bd[i] + h; }
    if (d->col_bd[d->ntc] != d->wctb || d->row_bd[d->ntr] != d->hctb) return HO_ERROR;
    d->ts2rs = (int*)calloc((size_t)nctb, sizeof(int)); d->rs2ts = (int*)calloc((size_t)nctb, sizeof(int)); d->tile_of_ctb = (int*)calloc((size_t)nctb, sizeof(int));
    int ts = 0;
    for (int tr = 0; tr < d->ntr; tr++) for (int tc = 0; tc < d->ntc; tc++)
      for (int y = d->row_bd[tr]; y < d->row_bd[tr + 1]; y++) for (int x = d->col_bd[tc]; x < d->col_bd[tc + 1]; x++) {
        int rs = y * d->wctb + x; d->ts2rs[ts] = rs; d->rs2ts[rs] = ts; d->tile_of_ctb[rs] = tr * d->ntc + tc; ts++; }
  }
  return HO_OK;
}

/* 7.3.6.1 slice_segment_header + 7.3.8.1 slice_segment_data */
static int decode_slice(dec_t* d, const uint8_t* rbsp, size_t n, int nal_type) {
  bitrd b = {rbsp, n, 16};
  int first = rd_bit(&b);
  if (nal_type >= 16 && nal_type <= 23) rd_bit(&b);
  int pps_id = rd_ue(&b);
  if (pps_id > 63 || !d->pps[pps_id].valid) return HO_ERROR;
  const pps_t* p = &d->pps[pps_id];
  if (p->sps_id > 15 || !d->sps[p->sps_id].valid) return HO_ERROR;
  const sps_t* s = &d->sps[p->sps_id];
  if (first) { if (d->picture_started) return HO_ERROR; d->s = s; d->p = p; int rc = alloc_picture(d); if (rc) return rc; }
  else if (!d->picture_started) return HO_ERROR;
  d->p = p;
  int dependent = 0, seg_addr = 0;
  if (!first) {
    if (p->dependent_slices) dependent = rd_bit(&b);
    seg_addr = rd_bits(&b, ceil_log2((unsigned)(d->wctb * d->hctb)));
  }
  if (!dependent) {
    rd_bits(&b, p->num_extra_bits);
    int slice_type = rd_ue(&b);
    if (slice_type != 2) return HO_UNSUPPORTED;
    if (p->output_flag_present) rd_bit(&b);
    if (nal_type != 19 && nal_type != 20) {
      rd_bits(&b, s->log2_max_poc_lsb);
      int sps_rps = rd_bit(&b);
      if (!sps_rps) { sps_t tmp = *s; parse_st_rps(&b, &tmp, s->num_st_rps, s->num_st_rps); }
      else if (s->num_st_rps > 1) rd_bits(&b, ceil_log2((unsigned)s->num_st_rps));
      if (s->long_term_present) {
        int nsps = 0, npics;
        if (s->num_lt_sps > 0) nsps = rd_ue(&b);
        npics = rd_ue(&b);
        for (int i = 0; i < nsps + npics; i++) {
          if (i < nsps) { if (s->num_lt_sps > 1) rd_bits(&b, ceil_log2((unsigned)s->num_lt_sps)); }
          else { rd_bits(&b, s->log2_max_poc_lsb); rd_bit(&b); }
          if (rd_bit(&b)) rd_ue(&b);
        }
      }
      if (s->temporal_mvp) rd_bit(&b);
    }
    d->sao_luma = d->sao_chroma = 0;
    if (s->sao) { d->sao_luma = rd_bit(&b); if (s->chroma_format_idc) d->sao_chroma = rd_bit(&b); }
    d->slice_qp = p->init_qp + rd_se(&b);
    d->cur_cb_off = d->cur_cr_off = 0;
    if (p->slice_chroma_qp_offsets_present) { d->cur_cb_off = rd_se(&b); d->cur_cr_off = rd_se(&b); }
    int dis = p->deblock_disabled, beta = p->beta_offset, tc = p->tc_offset, override = 0;
    if (p->deblock_override_enabled) override = rd_bit(&b);
    if (override) { dis = rd_bit(&b); if (!dis) { beta = 2 * rd_se(&b); tc = 2 * rd_se(&b); } }
    int across = p->lf_across_slices;
    if (p->lf_across_slices && (d->sao_luma || d->sao_chroma || !dis)) across = rd_bit(&b);
    if (d->nslices >= 1023) return HO_ERROR;
    d->slice_idx = d->nslices++;
    d->slice_addr_rs = seg_addr;
    d->sl[d->slice_idx].addr_rs = seg_addr; d->sl[d->slice_idx].lf_across = across;
    d->sl[d->slice_idx].deblock_disabled = dis; d->sl[d->slice_idx].beta_offset = beta; d->sl[d->slice_idx].tc_offset = tc;
    d->sl[d->slice_idx].cb_off = d->cur_cb_off; d->sl[d->slice_idx].cr_off = d->cur_cr_off;
  } else if (d->nslices == 0) return HO_ERROR;
  if (p->wpp || p->tiles) {
    int ne = rd_ue(&b);
    if (ne > 0) { int len = rd_ue(&b) + 1; for (int i = 0; i < ne; i++) { unsigned v = rd_bits(&b, len); if (getenv("HO_DEBUG") && i < 4) fprintf(stderr, "entry[%d]=%u\n", i, v + 1); } }
  }
  if (p->slice_ext_present) { int len = rd_ue(&b); for (int i = 0; i < len; i++) rd_bits(&b, 8); }
  rd_bit(&b);                                  /* byte_alignment(): alignment_bit_equal_to_one */
  b.pos = (b.pos + 7) & ~(size_t)7;
  /* slice_segment_data */
  d->br = b;
  if (!dependent) cabac_init_contexts(d);
  /* (dependent segments continue with the context state left by the previous segment, 9.3.1) */
  cabac_init_engine(d);
  int total = d->wctb * d->hctb;
  if (seg_addr >= total) return HO_ERROR;
  int ctb_ts = d->rs2ts[seg_addr], ctb_addr = seg_addr;                 /* the CTBs of a segment are consecutive in TILE scan (6.5.1) */
  d->first_qg_in_row = 1;
  if (dependent) d->first_qg_in_row = 0; /* qPY_PREV continues across a dependent slice segment */
  if (!dependent) d->last_cu_qpy = d->slice_qp;
  for (;;) {
    ctb_addr = d->ts2rs[ctb_ts];
    int rx = ctb_addr % d->wctb, ry = ctb_addr / d->wctb;
    d->cur_tile = d->tile_of_ctb[ctb_addr];
    if (p->wpp && rx == 0 && ctb_addr != seg_addr) {
      /* 9.3.1: synchronise with the state stored after the 2nd CTB of the row above */
      if (avail4(d, (rx + 1) << s->log2_ctb, (ry - 1) << s->log2_ctb)) memcpy(d->ctx, d->ctx_wpp, sizeof d->ctx);
      else cabac_init_contexts(d);
      d->first_qg_in_row = 1;
    }
    if (p->wpp && rx == 0 && ctb_addr == seg_addr && dependent && ry > 0) {
      if (avail4(d, 1 << s->log2_ctb, (ry - 1) << s->log2_ctb)) memcpy(d->ctx, d->ctx_wpp, sizeof d->ctx);
      d->first_qg_in_row = 1;
    }
    if (s->sao) parse_sao(d, rx, ry);
    coding_quadtree(d, rx << s->log2_ctb, ry << s->log2_ctb, s->log2_ctb, 0);
    if (d->err) return d->err;
    if (p->wpp && rx == 1) memcpy(d->ctx_wpp, d->ctx, sizeof d->ctx);   /* 9.3.2.2 storage after the 2nd CTB of a row */
    int end = dec_terminate(d);              /* end_of_slice_segment_flag */
    ctb_ts++; ctb_addr++;
    if (end) break;
    if (ctb_ts >= total) return HO_ERROR;
    if (p->tiles && d->tile_of_ctb[d->ts2rs[ctb_ts]] != d->tile_of_ctb[d->ts2rs[ctb_ts - 1]]) {
      /* the next CTB starts a tile: end_of_subset_one_bit, byte alignment, and 9.3.1 initialisation of the context variables */
      if (!dec_terminate(d)) return HO_ERROR;
      cabac_byte_align_and_restart(d);
      cabac_init_contexts(d);
      d->first_qg_in_row = 1;              /* first quantization group in a tile: qPY_PREV = SliceQpY (8.6.1) */
    }
    if (p->wpp && ctb_addr % d->wctb == 0) {
      if (!dec_terminate(d)) return HO_ERROR; /* end_of_subset_one_bit */
      if (getenv("HO_DEBUG")) fprintf(stderr, "row end at ctb %d: bitpos=%zu (byte %zu rem %zu), data start byte %zu, last bytes %02x %02x %02x\n", ctb_addr, d->br.pos, d->br.pos >> 3, d->br.pos & 7, b.pos >> 3, d->br.d[(d->br.pos >> 3) - 1], d->br.d[d->br.pos >> 3], d->br.d[(d->br.pos >> 3) + 1]);
      cabac_byte_align_and_restart(d);
    }
    if (d->br.pos > (n + 8) * 8) return HO_ERROR;
  }
  return HO_OK;
}

/* ---------------------------------------------------------------------------------- deblocking 8.7.2 */
static const uint8_t tc_tab[54] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,1,1,1,1,1,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,5,5,6,6,7,8,9,10,11,13,14,16,18,20,22,24};
static const uint8_t beta_tab[52] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,6,7,8,9,10,11,12,13,14,15,16,17,18,20,22,24,26,28,30,32,34,36,38,40,42,44,46,48,50,52,54,56,58,60,62,64};

/* is the edge at luma (x,y) (left edge if vert, else top edge of the 4x4 block) to be filtered?  8.7.2.3 */
static int edge_filtered(const dec_t* d, int x, int y, int vert) {
  int i = (y >> 2) * d->w4 + (x >> 2);
  if (!(d->tu_edge4[i] & (vert ? 1 : 2))) return 0;
  if (vert ? x == 0 : y == 0) return 0;
  int j = vert ? i - 1 : i - d->w4;
  int sq = d->slice_of4[i] - 1, sp = d->slice_of4[j] - 1;
  if (d->sl[sq].deblock_disabled) return 0;
  if (sq != sp && !d->sl[sq].lf_across) return 0;
  if (d->p->tiles && !d->p->lf_across_tiles && tile_of_xy(d, x, y) != tile_of_xy(d, vert ? x - 1 : x, vert ? y : y - 1)) return 0;   /* tile boundary */
  return 1;
}

static void deblock_luma_edge(dec_t* d, int x, int y, int vert) {   /* one 4-sample segment, 8.7.2.5.3/6/7 */
  uint16_t* pl = d->pl[0]; int st = d->stride[0], bd = d->s->bit_depth;
  int i = (y >> 2) * d->w4 + (x >> 2), j = vert ? i - 1 : i - d->w4;
  int sq = d->slice_of4[i] - 1;
  int qpl = (d->qp4[i] + d->qp4[j] + 1) >> 1;
  int beta = beta_tab[clip3(0, 51, qpl + d->sl[sq].beta_offset)] * (1 << (bd - 8));
  int tc = tc_tab[clip3(0, 53, qpl + 2 + d->sl[sq].tc_offset)] * (1 << (bd - 8));
  int xs = vert ? 1 : st, ls = vert ? st : 1;   /* step across the edge / along the edge */
  uint16_t* q = pl + y * st + x;
#define P(k, l) ((int)q[-(k + 1) * xs + (l) * ls])
#define Q(k, l) ((int)q[(k) * xs + (l) * ls])
  int dp0 = iabs(P(2, 0) - 2 * P(1, 0) + P(0, 0)), dp3 = iabs(P(2, 3) - 2 * P(1, 3) + P(0, 3));
  int dq0 = iabs(Q(2, 0) - 2 * Q(1, 0) + Q(0, 0)), dq3 = iabs(Q(2, 3) - 2 * Q(1, 3) + Q(0, 3));
  int dpq0 = dp0 + dq0, dpq3 = dp3 + dq3, dp = dp0 + dp3, dq = dq0 + dq3, dd = dpq0 + dpq3;
  if (dd >= beta) return;
  int s0 = 2 * dpq0 < (beta >> 2) && iabs(P(3, 0) - P(0, 0)) + iabs(Q(0, 0) - Q(3, 0)) < (beta >> 3) && iabs(P(0, 0) - Q(0, 0)) < ((5 * tc + 1) >> 1);
  int s3 = 2 * dpq3 < (beta >> 2) && iabs(P(3, 3) - P(0, 3)) + iabs(Q(0, 3) - Q(3, 3)) < (beta >> 3) && iabs(P(0, 3) - Q(0, 3)) < ((5 * tc + 1) >> 1);
  int strong = s0 && s3;
  int dep = dp < ((beta + (beta >> 1)) >> 3), deq = dq < ((beta + (beta >> 1)) >> 3);
  int maxv = (1 << bd) - 1;
  const int keep_p = d->nofilt4[j], keep_q = d->nofilt4[i];      /* nDp / nDq = 0 (8.7.2.5.7): pcm + pcm_loop_filter_disabled, cu_transquant_bypass */
  uint16_t save[4][8];
  for (int l = 0; l < 4; l++) for (int k = 0; k < 8; k++) save[l][k] = q[(k - 4) * xs + l * ls];
  for (int l = 0; l < 4; l++) {
    int p0 = P(0, l), p1 = P(1, l), p2 = P(2, l), p3 = P(3, l), q0 = Q(0, l), q1 = Q(1, l), q2 = Q(2, l), q3 = Q(3, l);
    if (strong) {
      q[-1 * xs + l * ls] = (uint16_t)clip3(p0 - 2 * tc, p0 + 2 * tc, (p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
      q[-2 * xs + l * ls] = (uint16_t)clip3(p1 - 2 * tc, p1 + 2 * tc, (p2 + p1 + p0 + q0 + 2) >> 2);
      q[-3 * xs + l * ls] = (uint16_t)clip3(p2 - 2 * tc, p2 + 2 * tc, (2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
      q[0 * xs + l * ls] = (uint16_t)clip3(q0 - 2 * tc, q0 + 2 * tc, (p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
      q[1 * xs + l * ls] = (uint16_t)clip3(q1 - 2 * tc, q1 + 2 * tc, (p0 + q0 + q1 + q2 + 2) >> 2);
      q[2 * xs + l * ls] = (uint16_t)clip3(q2 - 2 * tc, q2 + 2 * tc, (p0 + q0 + q1 + 3 * q2 + 2 * q3 + 4) >> 3);
    } else {
      int delta = (9 * (q0 - p0) - 3 * (q1 - p1) + 8) >> 4;
      if (iabs(delta) < tc * 10) {
        delta = clip3(-tc, tc, delta);
        q[-1 * xs + l * ls] = (uint16_t)clip3(0, maxv, p0 + delta);
        q[0 * xs + l * ls] = (uint16_t)clip3(0, maxv, q0 - delta);
        if (dep) { int dl = clip3(-(tc >> 1), tc >> 1, (((p2 + p0 + 1) >> 1) - p1 + delta) >> 1); q[-2 * xs + l * ls] = (uint16_t)clip3(0, maxv, p1 + dl); }
        if (deq) { int dl = clip3(-(tc >> 1), tc >> 1, (((q2 + q0 + 1) >> 1) - q1 - delta) >> 1); q[1 * xs + l * ls] = (uint16_t)clip3(0, maxv, q1 + dl); }
      }
    }
  }
  for (int l = 0; l < 4; l++) for (int k = 0; k < 8; k++) if (k < 4 ? keep_p : keep_q) q[(k - 4) * xs + l * ls] = save[l][k];
#undef P
#undef Q
}

/* one 4-luma-sample segment of a chroma edge; (x, y) in luma samples.  Along the edge it covers 4 >> (sub-sampling along the edge) samples */
static void deblock_chroma_edge(dec_t* d, int c, int x, int y, int vert) {
  uint16_t* pl = d->pl[c]; int st = d->stride[c], bd = d->s->bit_depth;
  int i = (y >> 2) * d->w4 + (x >> 2), j = vert ? i - 1 : i - d->w4;
  int sq = d->slice_of4[i] - 1;
  int off = c == 1 ? d->p->cb_qp_offset : d->p->cr_qp_offset;
  int qpi = ((d->qp4[i] + d->qp4[j] + 1) >> 1) + off;
  int qpc = d->cfmt != 1 ? (qpi < 51 ? qpi : 51) : (qpi < 30 ? qpi : (qpi >= 43 ? qpi - 6 : qpc_tab[qpi - 30]));     /* 8.7.2.5.5 */
  int tc = tc_tab[clip3(0, 53, qpc + 2 + d->sl[sq].tc_offset)] * (1 << (bd - 8));
  int xs = vert ? 1 : st, ls = vert ? st : 1, maxv = (1 << bd) - 1;
  uint16_t* q = pl + (y >> d->sy) * st + (x >> d->sx);
  int len = 4 >> (vert ? d->sy : d->sx);
  for (int l = 0; l < len; l++) {
    int p0 = q[-xs + l * ls], p1 = q[-2 * xs + l * ls], q0 = q[l * ls], q1 = q[xs + l * ls];
    int delta = clip3(-tc, tc, ((((q0 - p0) << 2) + p1 - q1 + 4) >> 3));
    if (!d->nofilt4[j]) q[-xs + l * ls] = (uint16_t)clip3(0, maxv, p0 + delta);
    if (!d->nofilt4[i]) q[l * ls] = (uint16_t)clip3(0, maxv, q0 - delta);
  }
}

static void deblock_picture(dec_t* d) {
  for (int vert = 1; vert >= 0; vert--) {          /* all vertical edges first, then horizontal */
    for (int y = 0; y < d->H; y += 4) for (int x = 0; x < d->W; x += 4) {
      if ((vert ? x : y) & 7) continue;             /* 8x8 luma grid */
      if (!edge_filtered(d, x, y, vert)) continue;
      deblock_luma_edge(d, x, y, vert);
      /* chroma edges lie on the 8x8 CHROMA sample grid (8.7.2.5): every 8 << sub-sampling luma samples across the edge */
      if (d->cfmt && ((vert ? x : y) & ((8 << (vert ? d->sx : d->sy)) - 1)) == 0) {
        deblock_chroma_edge(d, 1, x, y, vert);
        deblock_chroma_edge(d, 2, x, y, vert);
      }
    }
  }
}

/* ----------------------------------------------------------------------------------------- SAO 8.7.3 */
static void sao_picture(dec_t* d) {
  const sps_t* s = d->s;
  int ncomp = s->chroma_format_idc ? 3 : 1, bd = s->bit_depth, maxv = (1 << bd) - 1;
  for (int c = 0; c < ncomp; c++) {
    int w = c ? d->Wc : d->W, h = c ? d->Hc : d->H, st = d->stride[c], shx = c ? d->sx : 0, shy = c ? d->sy : 0;
    uint16_t* src = (uint16_t*)malloc((size_t)w * h * 2);
    memcpy(src, d->pl[c], (size_t)w * h * 2);      /* deblocked picture, read-only */
    int csx = d->ctb >> shx, csy = d->ctb >> shy;
    for (int ry = 0; ry < d->hctb; ry++) for (int rx = 0; rx < d->wctb; rx++) {
      const sao_params* sp = &d->sao[ry * d->wctb + rx];
      int on = d->ctb_slice_sao[ry * d->wctb + rx] & (c ? 2 : 1);
      if (!on || sp->type[c] == 0) continue;
      for (int y = ry * csy; y < (ry + 1) * csy && y < h; y++) for (int x = rx * csx; x < (rx + 1) * csx && x < w; x++) {
        int v = src[y * st + x], idx;
        if (d->nofilt4[((y << shy) >> 2) * d->w4 + ((x << shx) >> 2)]) continue;     /* 8.7.3: SaoTypeIdx treated as 0 for these samples */
        if (sp->type[c] == 1) {
          int k = ((v >> (bd - 5)) - sp->band_pos[c]) & 31;
          idx = k < 4 ? k + 1 : 0;
        } else {
          static const int8_t hp[4][2] = {{-1, 1}, {0, 0}, {-1, 1}, {1, -1}}, vp[4][2] = {{0, 0}, {-1, 1}, {-1, 1}, {-1, 1}};
          int e = sp->eo_class[c];
          int xa = x + hp[e][0], ya = y + vp[e][0], xb = x + hp[e][1], yb = y + vp[e][1];
          if (xa < 0 || xb < 0 || ya < 0 || yb < 0 || xa >= w || xb >= w || ya >= h || yb >= h) continue;
          /* slice boundary rule of 8.7.3.2 */
          int cur = d->slice_of4[((y << shy) >> 2) * d->w4 + ((x << shx) >> 2)] - 1;
          int sa = d->slice_of4[((ya << shy) >> 2) * d->w4 + ((xa << shx) >> 2)] - 1;
          int sb = d->slice_of4[((yb << shy) >> 2) * d->w4 + ((xb << shx) >> 2)] - 1;
          int skip = 0;
          if (sa != cur) { if (sa < cur ? !d->sl[cur].lf_across : !d->sl[sa].lf_across) skip = 1; }
          if (sb != cur) { if (sb < cur ? !d->sl[cur].lf_across : !d->sl[sb].lf_across) skip = 1; }
          if (d->p->tiles && !d->p->lf_across_tiles) {           /* 8.7.3.2: a neighbouring sample in a different tile */
            int tcur = tile_of_xy(d, x << shx, y << shy);
            if (tile_of_xy(d, xa << shx, ya << shy) != tcur || tile_of_xy(d, xb << shx, yb << shy) != tcur) skip = 1;
          }
          if (skip) continue;
          int a = src[ya * st + xa], b2 = src[yb * st + xb];
          int ei = 2 + (v > a) - (v < a) + (v > b2) - (v < b2);
          idx = ei == 2 ? 0 : (ei < 2 ? ei + 1 : ei);
        }
        if (getenv("HO_SAO_DBG")) { int dc, dx, dy; sscanf(getenv("HO_SAO_DBG"), "%d,%d,%d", &dc, &dx, &dy); if (dc == c && dx == x && dy == y) fprintf(stderr, "sao c=%d (%d,%d) type=%d class=%d band=%d idx=%d off=[%d %d %d %d] v=%d\n", c, x, y, sp->type[c], sp->eo_class[c], sp->band_pos[c], idx, sp->offset[c][1], sp->offset[c][2], sp->offset[c][3], sp->offset[c][4], v); }
        d->pl[c][y * st + x] = (uint16_t)clip3(0, maxv, v + sp->offset[c][idx]);
      }
    }
    free(src);
  }
}

/* --------------------------------------------------------------------------------------- top level */
static void free_dec(dec_t* d) {
  for (int c = 0; c < 3; c++) free(d->pl[c]);
  free(d->cmode4); free(d->slice_of4); free(d->ipm4); free(d->qp4); free(d->tu_edge4); free(d->cd4); free(d->nofilt4); free(d->sao); free(d->ctb_slice_sao); free(d->ts2rs); free(d->rs2ts); free(d->tile_of_ctb);
}

void hevc_oracle_free_picture(hevc_oracle_picture* p) { for (int c = 0; c < 3; c++) { free(p->plane[c]); p->plane[c] = NULL; } }

/*
 * data: [uint32 BE length][NAL]... exactly what libheif pushes into a decoder plugin
 * (libheif/codecs/decoder.cc:275-308; decoder_libde265.cc:322-368 splits on the 4-byte length).
 * stage: 0 = final picture (after SAO, conformance-cropped) ; 1 = reconstruction before deblocking ;
 *        2 = after deblocking, before SAO (stages 1/2 are cropped too).
 */
int hevc_oracle_decode(const uint8_t* data, size_t size, int stage, hevc_oracle_picture* out) {
  init_scans();
  memset(out, 0, sizeof *out);
  dec_t* d = (dec_t*)calloc(1, sizeof(dec_t));
  uint8_t* rbsp = (uint8_t*)malloc(size + 16);
  int rc = HO_OK;
  size_t p = 0;
  while (p + 4 <= size && rc == HO_OK) {
    uint32_t n = ((uint32_t)data[p] << 24) | (data[p + 1] << 16) | (data[p + 2] << 8) | data[p + 3];
    p += 4;
    if (n > size - p) { rc = HO_ERROR; break; }
    if (n >= 2) {
      int type = (data[p] >> 1) & 0x3f;
      size_t rn = nal_to_rbsp(data + p, n, rbsp);
      memset(rbsp + rn, 0, 8);
      if (type == 33) { sps_t s; rc = parse_sps(rbsp, rn, &s); if (rc == HO_OK) { bitrd b = {rbsp, rn, 16}; rd_bits(&b, 4); int msl = rd_bits(&b, 3); rd_bit(&b); skip_profile_tier_level(&b, msl); unsigned id = rd_ue(&b); if (id < 16) d->sps[id] = s; } }
      else if (type == 34) { pps_t pp; rc = parse_pps(rbsp, rn, &pp); if (rc == HO_OK) { bitrd b = {rbsp, rn, 16}; unsigned id = rd_ue(&b); if (id < 64) d->pps[id] = pp; } }
      else if (type <= 9) rc = HO_UNSUPPORTED;                       /* non-IRAP pictures */
      else if (type >= 16 && type <= 21) rc = decode_slice(d, rbsp, rn, type);
    }
    p += n;
  }
  if (rc == HO_OK && !d->picture_started) rc = HO_ERROR;
  if (rc == HO_OK) {
    for (size_t i = 0; i < (size_t)d->w4 * d->h4; i++) if (d->slice_of4[i] == 0) { rc = HO_ERROR; break; } /* incomplete picture */
  }
  if (rc == HO_OK) {
    if (stage != 1) deblock_picture(d);
    if (stage == 0 && d->s->sao) sao_picture(d);
    const sps_t* s = d->s;
    int subx = s->chroma_format_idc ? 1 << d->sx : 1, suby = s->chroma_format_idc ? 1 << d->sy : 1;      /* conformance window units (7.4.3.2.1) */
    int x0 = s->conf_l * subx, y0 = s->conf_t * suby;
    int w = d->W - (s->conf_l + s->conf_r) * subx, h = d->H - (s->conf_t + s->conf_b) * suby;
    if (w <= 0 || h <= 0) rc = HO_ERROR;
    else {
      out->width = w; out->height = h; out->bit_depth = s->bit_depth; out->chroma_format = s->chroma_format_idc;
      out->cw = s->chroma_format_idc ? (w + subx - 1) / subx : 0; out->ch = s->chroma_format_idc ? (h + suby - 1) / suby : 0;
      out->video_signal_present = s->vui_signal; out->full_range = s->vui_full_range;
      out->vui_colour_present = s->vui_colour;
      out->colour_primaries = s->vui_colour ? s->vui_cp : 2;
      out->transfer_characteristics = s->vui_colour ? s->vui_tc : 2;
      out->matrix_coeffs = s->vui_colour ? s->vui_mc : 2;
      out->plane[0] = (uint16_t*)malloc((size_t)w * h * 2);
      for (int y = 0; y < h; y++) memcpy(out->plane[0] + (size_t)y * w, d->pl[0] + (size_t)(y + y0) * d->W + x0, (size_t)w * 2);
      if (s->chroma_format_idc) for (int c = 1; c < 3; c++) {
        out->plane[c] = (uint16_t*)malloc((size_t)out->cw * out->ch * 2);
        for (int y = 0; y < out->ch; y++)
          memcpy(out->plane[c] + (size_t)y * out->cw, d->pl[c] + (size_t)(y + y0 / suby) * d->Wc + x0 / subx, (size_t)out->cw * 2);
      }
    }
  }
  if (d->picture_started) free_dec(d);
  free(d); free(rbsp);
  return rc;
}

/* VUI colour description of the first SPS in a length-prefixed access unit; mirrors what the reference
   plugin forwards (decoder_libde265.cc:426-448: defaults 2/2/2, full_range 0 when absent). */
int hevc_oracle_parse_vui(const uint8_t* data, size_t size, int out[4]) {
  out[0] = out[1] = out[2] = 2; out[3] = 0;
  size_t p = 0;
  uint8_t* rbsp = (uint8_t*)malloc(size + 16);
  int rc = HO_ERROR;
  while (p + 4 <= size) {
    uint32_t n = ((uint32_t)data[p] << 24) | (data[p + 1] << 16) | (data[p + 2] << 8) | data[p + 3];
    p += 4;
    if (n > size - p) break;
    if (n >= 2 && ((data[p] >> 1) & 0x3f) == 33) {
      size_t rn = nal_to_rbsp(data + p, n, rbsp);
      memset(rbsp + rn, 0, 8);
      sps_t s;
      rc = parse_sps(rbsp, rn, &s);
      if (s.vui_colour) { out[0] = s.vui_cp; out[1] = s.vui_tc; out[2] = s.vui_mc; }
      if (s.vui_signal) out[3] = s.vui_full_range;
      break;
    }
    p += n;
  }
  free(rbsp);
  return rc;
}

/* Debug digest of the syntax-level decoding state, used by tests/test_parser.py to pin the product's host front-end
   (which emits a command stream instead of pixels) without a GPU: per 8x8 block QpY and filterEdgeFlags, per 4x4 block
   luma / chroma intra mode, and an order-independent hash over every parsed coefficient level. */
int hevc_oracle_debug_maps(const uint8_t* data, size_t size, int8_t* qp8, uint8_t* edge8, uint8_t* lmode4, uint8_t* cmode4,
                           unsigned long long* out3 /* hash, coef count, tu count */, int* dims /* W, H */) {
  init_scans();
  dec_t* d = (dec_t*)calloc(1, sizeof(dec_t));
  uint8_t* rbsp = (uint8_t*)malloc(size + 16);
  int rc = HO_OK; size_t p = 0;
  while (p + 4 <= size && rc == HO_OK) {
    uint32_t n = ((uint32_t)data[p] << 24) | (data[p + 1] << 16) | (data[p + 2] << 8) | data[p + 3];
    p += 4;
    if (n > size - p) { rc = HO_ERROR; break; }
    if (n >= 2) {
      int type = (data[p] >> 1) & 0x3f;
      size_t rn = nal_to_rbsp(data + p, n, rbsp);
      memset(rbsp + rn, 0, 8);
      if (type == 33) { sps_t s; rc = parse_sps(rbsp, rn, &s); if (rc == HO_OK) { bitrd b = {rbsp, rn, 16}; rd_bits(&b, 4); int msl = rd_bits(&b, 3); rd_bit(&b); skip_profile_tier_level(&b, msl); unsigned id = rd_ue(&b); if (id < 16) d->sps[id] = s; } }
      else if (type == 34) { pps_t pp; rc = parse_pps(rbsp, rn, &pp); if (rc == HO_OK) { bitrd b = {rbsp, rn, 16}; unsigned id = rd_ue(&b); if (id < 64) d->pps[id] = pp; } }
      else if (type >= 16 && type <= 21) rc = decode_slice(d, rbsp, rn, type);
    }
    p += n;
  }
  if (rc == HO_OK && d->picture_started) {
    dims[0] = d->W; dims[1] = d->H;
    int w8 = d->W >> 3, h8 = d->H >> 3;
    for (int by = 0; by < h8; by++) for (int bx = 0; bx < w8; bx++) {
      qp8[by * w8 + bx] = d->qp4[(by * 2) * d->w4 + bx * 2];
      edge8[by * w8 + bx] = (uint8_t)((edge_filtered(d, bx * 8, by * 8, 1) ? 1 : 0) | (edge_filtered(d, bx * 8, by * 8, 0) ? 2 : 0) |
                                      (d->nofilt4[(by * 2) * d->w4 + bx * 2] ? 4 : 0));      /* bit 2: in-loop filters leave the unit unchanged */
    }
    memcpy(lmode4, d->ipm4, (size_t)d->w4 * d->h4);
    memcpy(cmode4, d->cmode4, (size_t)d->w4 * d->h4);
    out3[0] = d->coef_hash; out3[1] = d->coef_count; out3[2] = d->tu_count;
  } else if (rc == HO_OK) rc = HO_ERROR;
  if (d->picture_started) free_dec(d);
  free(d); free(rbsp);
  return rc;
}
