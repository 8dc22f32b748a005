// b200_hevc_scaling.h -- scaling lists (H.265 7.3.4 scaling_list_data, 7.4.5 semantics, Tables 7-5 / 7-6): the coded lists,
// their defaults and prediction, and the per-position scaling factors m[x][y] the dequantiser of the reconstruction kernel
// multiplies with (8.6.4.2).  Host code, shared by the header parser (b200_hevc_parse.cc) and the fixture encoder
// (b200_hevc_enc.cc).  Intra-only: the factors are kept for matrixId = cIdx (0..2).
#pragma once
#include <cstdint>
#include <cstring>

namespace b200 {
namespace sl {

// Table 7-6, in up-right diagonal scan order: matrixId 0..2 (intra) and 3..5 (inter)
static const uint8_t kDefault8[2][64] = {
    {16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 17, 16, 17, 16, 17, 18, 17, 18, 18, 17, 18, 21, 19, 20, 21, 20, 19, 21, 24, 22, 22, 24,
     24, 22, 22, 24, 25, 25, 27, 30, 27, 25, 25, 29, 31, 35, 35, 31, 29, 36, 41, 44, 41, 36, 47, 54, 54, 47, 65, 70, 65, 88, 88, 115},
    {16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 17, 17, 17, 17, 17, 18, 18, 18, 18, 18, 18, 20, 20, 20, 20, 20, 20, 20, 24, 24, 24, 24,
     24, 24, 24, 24, 25, 25, 25, 25, 25, 25, 25, 28, 28, 28, 28, 28, 28, 33, 33, 33, 33, 33, 41, 41, 41, 41, 54, 54, 54, 71, 71, 91}};

struct Lists {                      // ScalingList[sizeId][matrixId][i] in coded (diagonal scan) order + scaling_list_dc_coef
  uint8_t list[4][6][64];
  uint8_t dc[4][6];
};

inline void set_default(Lists& L, int size_id, int matrix_id) {
  if (size_id == 0) memset(L.list[0][matrix_id], 16, 16);                               // Table 7-5
  else memcpy(L.list[size_id][matrix_id], kDefault8[matrix_id < 3 ? 0 : 1], 64);
  L.dc[size_id][matrix_id] = 16;
}
inline void set_all_default(Lists& L) { for (int s = 0; s < 4; s++) for (int m = 0; m < 6; m++) set_default(L, s, m); }

// 6.5.3 up-right diagonal scan of a (1 << log2) square: position of scan index i
inline void diag_scan(int log2, uint8_t* xs, uint8_t* ys) {
  const int n = 1 << log2;
  int i = 0, x = 0, y = 0; bool stop = false;
  while (!stop) {
    while (y >= 0) { if (x < n && y < n) { xs[i] = (uint8_t)x; ys[i] = (uint8_t)y; i++; } y--; x++; }
    y = x; x = 0;
    if (i >= n * n) stop = true;
  }
}

// Scaling factors in raster order per colour component and block size: m[c][sizeId][y8 * 8 + x8] for 8x8 and larger (the
// 8x8 list is replicated 2x2 / 4x4 for 16x16 / 32x32, 7.4.5), m[c][0][y * 4 + x] for 4x4; dc[c][sizeId] replaces
// position (0, 0) of 16x16 and 32x32 blocks.
struct Factors { uint8_t m[3][4][64]; uint8_t dc[3][4]; };
static_assert(sizeof(Factors) == 780, "layout shared with the reconstruction kernel");

inline void derive(const Lists& L, Factors& F) {
  uint8_t x4[16], y4[16], x8[64], y8[64];
  diag_scan(2, x4, y4); diag_scan(3, x8, y8);
  memset(&F, 16, sizeof F);
  for (int c = 0; c < 3; c++) {
    for (int i = 0; i < 16; i++) F.m[c][0][y4[i] * 4 + x4[i]] = L.list[0][c][i];
    for (int s = 1; s < 4; s++) {
      // 32x32: only matrixId 0 (intra) / 3 (inter) are coded; a 32x32 CHROMA block exists only in 4:4:4 and takes the 16x16 list of its
      // matrixId, replicated 4x4, with that list's DC coefficient (7.4.5, ChromaArrayType == 3)
      const int ls = (s == 3 && c > 0) ? 2 : s, mid = (s == 3 && c == 0) ? 0 : c;
      for (int i = 0; i < 64; i++) F.m[c][s][y8[i] * 8 + x8[i]] = L.list[ls][mid][i];
      F.dc[c][s] = s >= 2 ? L.dc[ls][mid] : L.list[ls][mid][0];
    }
  }
}

}  // namespace sl
}  // namespace b200
