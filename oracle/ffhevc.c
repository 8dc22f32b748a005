/*
 * oracle/ffhevc.c -- TEST INFRASTRUCTURE ONLY. Never linked, imported or executed by the product.
 *
 * Independent HEVC decoder used to PIN the HEVC restatement (oracle/hevc_oracle.c) and the CUDA path.
 *
 * Why FFmpeg: the reference delegates all HEVC arithmetic to libde265
 * (libheif/plugins/decoder_libde265.cc:30,181,360,402,410), a third-party library that is neither
 * vendored in /root/reference nor installed in this image, with no pinned version
 * (cmake/modules/FindLIBDE265.cmake:1-43).  H.265 decoding is normative (bit-exact for conforming
 * streams), so any conforming decoder yields libde265's planes.  The only HEVC decoder present in the
 * image is libavcodec 62.11.100 bundled inside the opencv-python-headless wheel; the reference itself
 * ships an FFmpeg decoder plugin (libheif/plugins/decoder_ffmpeg.cc:139-195 open, :266-330 NAL
 * re-framing, :574-707 decode) which this file restates with hand-declared prototypes (the wheel has
 * no libav headers).
 *
 * Struct offsets (libavcodec 62 / libavutil 60 ABI, x86-64), verified by decoding examples/example.heic:
 *   AVFrame:  data[i] @ 8*i, linesize[i] @ 64+4*i, width @ 104, height @ 108, format @ 116
 *   AVPacket: data @ 24, size @ 32
 *
 * Build: see oracle/Makefile (-> oracle/_ref/libffhevc.so).
 */
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <dirent.h>

typedef void* (*fn_find_decoder)(int);
typedef void* (*fn_alloc_ctx)(void*);
typedef int (*fn_open2)(void*, void*, void*);
typedef void* (*fn_packet_alloc)(void);
typedef void (*fn_packet_free)(void**);
typedef int (*fn_send_packet)(void*, void*);
typedef int (*fn_receive_frame)(void*, void*);
typedef void (*fn_free_ctx)(void**);
typedef void* (*fn_frame_alloc)(void);
typedef void (*fn_frame_free)(void**);
typedef const char* (*fn_pix_name)(int);
typedef int (*fn_opt_set_int)(void*, const char*, int64_t, int);
typedef void (*fn_log_level)(int);

static struct {
  int loaded;
  fn_find_decoder find_decoder; fn_alloc_ctx alloc_ctx; fn_open2 open2;
  fn_packet_alloc packet_alloc; fn_packet_free packet_free;
  fn_send_packet send_packet; fn_receive_frame receive_frame; fn_free_ctx free_ctx;
  fn_frame_alloc frame_alloc; fn_frame_free frame_free; fn_pix_name pix_name;
  fn_opt_set_int opt_set_int; fn_log_level log_level;
} F;

static void* open_prefixed(const char* dir, const char* prefix, int flags) {
  DIR* d = opendir(dir);
  if (!d) return NULL;
  struct dirent* e; void* h = NULL;
  while ((e = readdir(d))) {
    if (strncmp(e->d_name, prefix, strlen(prefix)) == 0) {
      char path[4096];
      snprintf(path, sizeof path, "%s/%s", dir, e->d_name);
      h = dlopen(path, flags);
      break;
    }
  }
  closedir(d);
  return h;
}

/* dir = .../site-packages/opencv_python_headless.libs ; returns 0 on success */
int ffhevc_init(const char* dir) {
  if (F.loaded) return 0;
  /* the wheel's libraries carry no RUNPATH: preload dependencies in order so NEEDED resolves by soname */
  open_prefixed(dir, "libdrm-", RTLD_NOW | RTLD_GLOBAL);
  open_prefixed(dir, "libcrypto-", RTLD_NOW | RTLD_GLOBAL);
  void* hu = open_prefixed(dir, "libavutil-", RTLD_NOW | RTLD_GLOBAL);
  open_prefixed(dir, "libswresample-", RTLD_NOW | RTLD_GLOBAL);
  open_prefixed(dir, "libvpx-", RTLD_NOW | RTLD_GLOBAL);
  void* hc = open_prefixed(dir, "libavcodec-", RTLD_NOW | RTLD_GLOBAL);
  if (!hu || !hc) { fprintf(stderr, "ffhevc_init: %s\n", dlerror()); return -1; }
#define SYM(h, field, name) do { *(void**)(&F.field) = dlsym(h, name); if (!F.field) return -2; } while (0)
  SYM(hc, find_decoder, "avcodec_find_decoder"); SYM(hc, alloc_ctx, "avcodec_alloc_context3");
  SYM(hc, open2, "avcodec_open2"); SYM(hc, packet_alloc, "av_packet_alloc");
  SYM(hc, packet_free, "av_packet_free"); SYM(hc, send_packet, "avcodec_send_packet");
  SYM(hc, receive_frame, "avcodec_receive_frame"); SYM(hc, free_ctx, "avcodec_free_context");
  SYM(hu, frame_alloc, "av_frame_alloc"); SYM(hu, frame_free, "av_frame_free");
  SYM(hu, pix_name, "av_get_pix_fmt_name"); SYM(hu, opt_set_int, "av_opt_set_int");
  SYM(hu, log_level, "av_log_set_level");
  F.log_level(16 /* AV_LOG_ERROR */);
  F.loaded = 1;
  return 0;
}

typedef struct {
  int width, height;      /* luma size after conformance cropping */
  int cw, ch;             /* chroma plane size (0 for 4:0:0) */
  int bit_depth;          /* 8, 10, 12 */
  int chroma;             /* 0 = 4:0:0, 1 = 4:2:0, 2 = 4:2:2, 3 = 4:4:4 */
  int full_range_name;    /* 1 if pix fmt is a yuvj* (full range) format */
  uint16_t* plane[3];     /* samples widened to uint16, tightly packed (malloc) */
} ffhevc_picture;

void ffhevc_free_picture(ffhevc_picture* p) {
  for (int i = 0; i < 3; i++) { free(p->plane[i]); p->plane[i] = NULL; }
}

/*
 * data: one access unit as libheif hands it to a decoder plugin -- a sequence of
 * [uint32 big-endian length][NAL] (libheif/codecs/decoder.cc:275-308, decoder_libde265.cc:322-368).
 * Re-framed to Annex-B start codes like decoder_ffmpeg.cc:266-330 does.
 */
static int g_skip_loop_filter;
/* debugging aid: 1 = FFmpeg skips deblocking and SAO (AVDISCARD_ALL), giving the pre-filter reconstruction */
void ffhevc_set_skip_loop_filter(int on) { g_skip_loop_filter = on; }

int ffhevc_decode(const uint8_t* data, size_t size, int threads, ffhevc_picture* out) {
  memset(out, 0, sizeof *out);
  if (!F.loaded) return -100;
  uint8_t* annexb = (uint8_t*)malloc(size + 64);
  size_t o = 0, p = 0;
  while (p + 4 <= size) {
    uint32_t n = ((uint32_t)data[p] << 24) | (data[p + 1] << 16) | (data[p + 2] << 8) | data[p + 3];
    p += 4;
    if (n > size - p) { free(annexb); return -1; }
    annexb[o++] = 0; annexb[o++] = 0; annexb[o++] = 0; annexb[o++] = 1;
    memcpy(annexb + o, data + p, n);
    o += n; p += n;
  }
  memset(annexb + o, 0, 64);
  void* codec = F.find_decoder(173 /* AV_CODEC_ID_HEVC */);
  if (!codec) { free(annexb); return -2; }
  void* ctx = F.alloc_ctx(codec);
  if (threads > 0) F.opt_set_int(ctx, "threads", threads, 0);
  if (g_skip_loop_filter) F.opt_set_int(ctx, "skip_loop_filter", 48 /* AVDISCARD_ALL */, 0);
  int rc = F.open2(ctx, codec, NULL);
  if (rc < 0) { F.free_ctx(&ctx); free(annexb); return -3; }
  void* pkt = F.packet_alloc();
  *(uint8_t**)((char*)pkt + 24) = annexb;
  *(int*)((char*)pkt + 32) = (int)o;
  void* frame = F.frame_alloc();
  rc = F.send_packet(ctx, pkt);
  int ret = -4;
  if (rc >= 0) {
    F.send_packet(ctx, NULL);
    rc = F.receive_frame(ctx, frame);
    if (rc >= 0) {
      uint8_t** fdata = (uint8_t**)frame;
      int* linesize = (int*)((char*)frame + 64);
      int w = *(int*)((char*)frame + 104), h = *(int*)((char*)frame + 108);
      int fmt = *(int*)((char*)frame + 116);
      const char* name = F.pix_name(fmt);
      int bd = 8, chroma = -1;
      if (!name) name = "";
      if (strstr(name, "gray")) chroma = 0;
      else if (strstr(name, "420")) chroma = 1;
      else if (strstr(name, "422")) chroma = 2;
      else if (strstr(name, "444")) chroma = 3;
      if (strstr(name, "10le")) bd = 10; else if (strstr(name, "12le")) bd = 12;
      else if (strstr(name, "9le")) bd = 9; else if (strstr(name, "14le")) bd = 14; else if (strstr(name, "16le")) bd = 16;
      out->full_range_name = strncmp(name, "yuvj", 4) == 0;
      if (chroma < 0) ret = -5;
      else {
        out->width = w; out->height = h; out->bit_depth = bd; out->chroma = chroma;
        out->cw = chroma == 0 ? 0 : (chroma == 3 ? w : (w + 1) / 2);
        out->ch = chroma == 0 ? 0 : (chroma == 1 ? (h + 1) / 2 : h);
        int np = chroma == 0 ? 1 : 3;
        for (int c = 0; c < np; c++) {
          int pw = c ? out->cw : w, ph = c ? out->ch : h;
          out->plane[c] = (uint16_t*)malloc((size_t)pw * ph * 2 + 2);
          for (int y = 0; y < ph; y++) {
            const uint8_t* row = fdata[c] + (size_t)y * linesize[c];
            uint16_t* dst = out->plane[c] + (size_t)y * pw;
            if (bd == 8) for (int x = 0; x < pw; x++) dst[x] = row[x];
            else memcpy(dst, row, (size_t)pw * 2);
          }
        }
        ret = 0;
      }
    }
  }
  F.frame_free(&frame);
  *(uint8_t**)((char*)pkt + 24) = NULL;
  *(int*)((char*)pkt + 32) = 0;
  F.packet_free(&pkt);
  F.free_ctx(&ctx);
  free(annexb);
  return ret;
}
