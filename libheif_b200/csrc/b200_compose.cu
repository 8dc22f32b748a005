// b200_compose.cu -- the two remaining post-stage operations of SURVEY 8(a): overlay compositing (a11) and
// nearest-neighbour plane scaling (a12).  Byte copies / one integer blend per sample: HBM-bound streaming kernels.
//
// Reference behaviour (paths relative to the libheif tree):
//   HeifPixelImage::fill_RGB_16bit           libheif/image/pixelimage.cc:1549-1621   canvas = background >> 8
//   HeifPixelImage::overlay                  libheif/image/pixelimage.cc:1637-1780   clipped copy or (in*a + out*(255-a)) / 255
//   ImageItem_Overlay::decode_overlay_image  libheif/image-items/overlay.cc:290-393  8-bit planar RGB canvas, children in ipma order
//   HeifPixelImage::scale_nearest_neighbor   libheif/image/pixelimage.cc:1783-1972   ix = x * W_in / W_out (64-bit), per plane
#include <cstdint>
#include <cuda_runtime.h>
#include "b200_internal.h"

namespace b200 {

struct OverlayArgs {
  uint8_t* out[3]; size_t out_stride[3];
  const uint8_t* in[3]; size_t in_stride[3];
  const uint8_t* alpha; size_t alpha_stride;
  uint32_t in_x0, in_y0, out_x0, out_y0;
  uint32_t x_begin, x_end;      // iteration range of the reference's inner loop (alpha path: [in_x0, in_w), copy path: [0, in_w))
  uint32_t y_begin, y_end;      // [in_y0, in_h)
};

// 16 bytes per thread on the copy path when everything is 16-byte aligned would be possible; overlays are small next to
// the decode, so one sample per thread keeps the reference's index arithmetic visible.
__global__ void overlay_kernel(const OverlayArgs a) {
  const uint32_t x = a.x_begin + blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t y = a.y_begin + blockIdx.y;
  if (x >= a.x_end || y >= a.y_end) return;
  const size_t orow = (size_t)(a.out_y0 + y - a.in_y0);
#pragma unroll
  for (int c = 0; c < 3; c++) {
    uint8_t* o = a.out[c] + a.out_x0 + orow * a.out_stride[c] + x;
    const uint8_t v = a.in[c][a.in_x0 + (size_t)y * a.in_stride[c] + x];
    if (!a.alpha) *o = v;
    else {
      const unsigned al = a.alpha[a.in_x0 + (size_t)y * a.alpha_stride + x];
      *o = (uint8_t)((v * al + *o * (255u - al)) / 255u);
    }
  }
}

__global__ void fill_kernel(uint8_t* p, size_t stride, int w, int h, uint8_t v) {
  const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 16, y = blockIdx.y;
  if (x >= w || y >= h) return;
  uint8_t* q = p + (size_t)y * stride + x;
  if (x + 16 <= w && ((reinterpret_cast<uintptr_t>(q) & 15) == 0)) { const unsigned u = v * 0x01010101u; *reinterpret_cast<uint4*>(q) = make_uint4(u, u, u, u); }
  else for (int i = 0; i < 16 && x + i < w; i++) q[i] = v;
}

template <int BPP>
__global__ void scale_nn_kernel(const uint8_t* __restrict__ in, size_t in_stride, uint8_t* __restrict__ out, size_t out_stride, uint32_t out_w, uint32_t out_h,
                                uint32_t wi, uint32_t hi, uint32_t wo, uint32_t ho) {
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= out_w || y >= out_h) return;
  const uint32_t iy = (uint32_t)((uint64_t)y * hi / ho), ix = (uint32_t)((uint64_t)x * wi / wo);
  const uint8_t* s = in + (size_t)iy * in_stride + (size_t)ix * BPP;
  uint8_t* d = out + (size_t)y * out_stride + (size_t)x * BPP;
#pragma unroll
  for (int i = 0; i < BPP; i++) d[i] = s[i];
}

}  // namespace b200

using namespace b200;

extern "C" int b200_overlay_fill_device(void* const planes[3], const size_t strides[3], int width, int height, const uint16_t background_rgba[4], void* stream) {
  if (!planes || !strides || !background_rgba || width <= 0 || height <= 0) return set_error(B200_E_INVALID, "overlay fill: bad argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  for (int c = 0; c < 3; c++) {
    if (!planes[c] || strides[c] < (size_t)width) return set_error(B200_E_INVALID, "overlay fill: bad plane %d", c);
    fill_kernel<<<dim3((unsigned)((width + 16 * 128 - 1) / (16 * 128)), (unsigned)height), 128, 0, s>>>(static_cast<uint8_t*>(planes[c]), strides[c], width, height,
                                                                                                     (uint8_t)(background_rgba[c] >> 8));
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(B200_E_CUDA, "overlay fill: %s", cudaGetErrorString(e));
  return B200_OK;
}

extern "C" int b200_overlay_device(void* const canvas[3], const size_t canvas_strides[3], int canvas_w, int canvas_h, const void* const overlay[4],
                                   const size_t overlay_strides[4], int overlay_w, int overlay_h, int32_t dx, int32_t dy, void* stream) {
  if (!canvas || !canvas_strides || !overlay || !overlay_strides || canvas_w <= 0 || canvas_h <= 0 || overlay_w <= 0 || overlay_h <= 0)
    return set_error(B200_E_INVALID, "overlay: bad argument");
  for (int c = 0; c < 3; c++) if (!canvas[c] || !overlay[c]) return set_error(B200_E_INVALID, "overlay: missing colour plane %d", c);
  // clipping exactly as pixelimage.cc:1687-1755 (all planes share the logical size, so one set of values serves R, G and B)
  auto negate = [](int32_t x) -> uint32_t { return x == INT32_MIN ? (uint32_t)INT32_MAX + 1u : (uint32_t)(-x); };
  uint32_t in_w = (uint32_t)overlay_w, in_h = (uint32_t)overlay_h;
  const uint32_t out_w = (uint32_t)canvas_w, out_h = (uint32_t)canvas_h;
  if (dx > 0 && (uint32_t)dx >= out_w) return B200_OK;               // completely outside: nothing drawn, not an error
  if (dx < 0 && in_w <= negate(dx)) return B200_OK;
  if (dy > 0 && (uint32_t)dy >= out_h) return B200_OK;
  if (dy < 0 && in_h <= negate(dy)) return B200_OK;
  if (dx + (int64_t)in_w > out_w) in_w = (uint32_t)((int64_t)out_w - dx);
  if (dy + (int64_t)in_h > out_h) in_h = (uint32_t)((int64_t)out_h - dy);
  OverlayArgs a{};
  if (dx < 0) { a.in_x0 = negate(dx); a.out_x0 = 0; in_w -= a.in_x0; } else { a.in_x0 = 0; a.out_x0 = (uint32_t)dx; }
  if (dy < 0) { a.in_y0 = negate(dy); a.out_y0 = 0; in_h -= a.in_y0; } else { a.in_y0 = 0; a.out_y0 = (uint32_t)dy; }
  for (int c = 0; c < 3; c++) {
    a.out[c] = static_cast<uint8_t*>(canvas[c]); a.out_stride[c] = canvas_strides[c];
    a.in[c] = static_cast<const uint8_t*>(overlay[c]); a.in_stride[c] = overlay_strides[c];
  }
  a.alpha = static_cast<const uint8_t*>(overlay[3]); a.alpha_stride = overlay[3] ? overlay_strides[3] : 0;
  // the reference's loops: rows y in [in_y0, in_h); columns x in [in_x0, in_w) on the alpha path, memcpy of in_w bytes otherwise
  a.y_begin = a.in_y0; a.y_end = in_h;
  a.x_begin = a.alpha ? a.in_x0 : 0; a.x_end = in_w;
  if (a.y_end <= a.y_begin || a.x_end <= a.x_begin) return B200_OK;
  const unsigned nx = a.x_end - a.x_begin, ny = a.y_end - a.y_begin;
  if (ny > 65535u * 1u && ny > 2147483647u) return set_error(B200_E_LIMIT, "overlay: too many rows");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  for (unsigned y0 = 0; y0 < ny; y0 += 65535u) {                     // gridDim.y limit
    OverlayArgs b = a; b.y_begin = a.y_begin + y0; b.y_end = (ny - y0 > 65535u) ? b.y_begin + 65535u : a.y_end;
    overlay_kernel<<<dim3((nx + 255) / 256, b.y_end - b.y_begin), 256, 0, s>>>(b);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(B200_E_CUDA, "overlay: %s", cudaGetErrorString(e));
  return B200_OK;
}

extern "C" int b200_scale_nearest_device(const void* in, size_t in_stride, void* out, size_t out_stride, uint32_t out_w, uint32_t out_h,
                                         uint32_t image_w_in, uint32_t image_h_in, uint32_t image_w_out, uint32_t image_h_out, int bytes_per_pixel, void* stream) {
  if (!in || !out || !out_w || !out_h || !image_w_in || !image_h_in || !image_w_out || !image_h_out) return set_error(B200_E_INVALID, "scale: bad argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const dim3 blk(64, 4), grid((out_w + 63) / 64, (out_h + 3) / 4);
  if (grid.y > 65535u) return set_error(B200_E_LIMIT, "scale: plane too tall");
  const uint8_t* i8 = static_cast<const uint8_t*>(in); uint8_t* o8 = static_cast<uint8_t*>(out);
#define B200_SCALE(N) case N: scale_nn_kernel<N><<<grid, blk, 0, s>>>(i8, in_stride, o8, out_stride, out_w, out_h, image_w_in, image_h_in, image_w_out, image_h_out); break;
  switch (bytes_per_pixel) { B200_SCALE(1) B200_SCALE(2) B200_SCALE(3) B200_SCALE(4) B200_SCALE(6) B200_SCALE(8)
    default: return set_error(B200_E_INVALID, "scale: bytes_per_pixel must be 1, 2, 3, 4, 6 or 8"); }
#undef B200_SCALE
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(B200_E_CUDA, "scale: %s", cudaGetErrorString(e));
  return B200_OK;
}
