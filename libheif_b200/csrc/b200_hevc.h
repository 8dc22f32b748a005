// b200_hevc.h -- internal interfaces of the HEVC intra decoder (host front-end + device back-end)
#pragma once
#include "b200_internal.h"
#include "b200_hevc_types.h"
#include <vector>

namespace b200 {

struct ParseLimits { uint64_t max_image_size_pixels = 0; };   // heif_security_limits.max_image_size_pixels (0 = unlimited)

struct ParsedPicture {
  PicDesc desc;
  std::vector<CtuInfo> ctus;
  std::vector<TuCmd> tus;
  std::vector<CoefEntry> coefs;
  std::vector<SliceInfo> slices;
  std::vector<int8_t> qp8;
  std::vector<uint8_t> edge8;
  // VUI colour description as the libde265 plugin reports it (decoder_libde265.cc:426-448)
  int colour_primaries = 2, transfer_characteristics = 2, matrix_coefficients = 2, full_range = 0;
};

// Host front-end: length-prefixed NAL units of one access unit -> command stream.  Thread-safe (no shared state).
int parse_access_unit(const uint8_t* data, size_t size, const ParseLimits& limits, ParsedPicture& out);

// Device back-end (b200_hevc_recon.cu / b200_hevc_filters.cu).  All arrays are batch-wide device buffers.
struct DeviceBatch {
  const PicDesc* pics; int npics;
  const CtuInfo* ctus; const TuCmd* tus; const CoefEntry* coefs; const SliceInfo* slices;
  const int8_t* qp8; const uint8_t* edge8;
  unsigned int* progress;        // one counter per CTB row of every picture, zeroed before launch
  unsigned int* ticket;          // work-distribution counter, zeroed before launch
  unsigned int* error_flag;      // set by a kernel that gave up waiting (zeroed before launch)
  const uint2* row_list;         // (picture, ctb row) in launch order
  int nrows;
  int max_log2_ctb;              // largest CTB size of the batch (sizes the per-warp shared memory)
};
int launch_recon(const DeviceBatch& b, cudaStream_t s);
int launch_deblock(const DeviceBatch& b, const PicDesc* host_pics, cudaStream_t s);
int launch_sao(const DeviceBatch& b, const PicDesc* host_pics, cudaStream_t s);

}  // namespace b200
