// ASan harness for the host front-end: parse_access_unit on mutated streams read from a directory of .au files
#include "b200_hevc.h"
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
namespace b200 { int set_error(int code, const char* fmt, ...) { (void)fmt; return code; } }
using namespace b200;
static uint32_t rng_state = 1;
static uint32_t rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }
int main(int argc, char** argv) {
  if (argc < 3) return 2;
  rng_state = (uint32_t)atoi(argv[1]);
  const int iters = atoi(argv[2]);
  std::vector<std::vector<uint8_t>> streams;
  for (int i = 3; i < argc; i++) { FILE* f = fopen(argv[i], "rb"); if (!f) continue; std::vector<uint8_t> b; uint8_t buf[65536]; size_t n; while ((n = fread(buf, 1, sizeof buf, f)) > 0) b.insert(b.end(), buf, buf + n); fclose(f); if (b.size() < 300000) streams.push_back(b); }
  ParseLimits lim{}; lim.max_image_size_pixels = 1u << 24;
  ParsedPicture pic;
  int ok = 0, bad = 0;
  if (iters == 0) {            // replay mode: parse every file as it is (crafted streams, e.g. the header-field injections of tests/test_parser.py)
    for (auto& b : streams) { const int rc = parse_access_unit(b.data(), b.size(), lim, pic); if (rc == 0) ok++; else bad++; }
    printf("%zu files: %d decoded, %d rejected\n", streams.size(), ok, bad);
    return 0;
  }
  for (int it = 0; it < iters; it++) {
    std::vector<uint8_t> b = streams[rnd() % streams.size()];
    const int k = 1 + rnd() % 6, mode = rnd() % 10;
    for (int j = 0; j < k && !b.empty(); j++) {
      const size_t p = rnd() % b.size();
      if (mode < 7) b[p] ^= (uint8_t)(1u << (rnd() % 8)); else if (mode < 9) b[p] = (uint8_t)rnd(); else b.erase(b.begin() + p, b.begin() + std::min(b.size(), p + 1 + rnd() % 7));
    }
    const int rc = parse_access_unit(b.data(), b.size(), lim, pic);
    if (rc == 0) ok++; else bad++;
  }
  printf("%d mutations: %d decoded, %d rejected\n", iters, ok, bad);
  return 0;
}
