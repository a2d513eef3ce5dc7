// tests/cpp/consumer.cpp -- a C++ user of the reference decode API compiled against include/brotlig_b200/BrotliG.h and
// linked with libbrotlig_b200.so, written the way sample/brotlig_cli.cpp calls the SDK (:424-425 sizes the output with
// DecompressedSize, :441-444 calls DecodeGPU in a -num-repeat loop and sums `time`, :461 calls BrotliG::DecodeCPU).
// usage: consumer <stream.brotlig> <expected-output-file> [abort_after_pages]
#include "brotlig_b200/BrotliG.h"

#include <cstdio>
#include <cstring>
#include <fstream>
#include <iterator>
#include <vector>

static int g_calls = 0, g_abort_after = -1;
static float g_last_progress = -1.f;
static bool Feedback(BROTLIG_MESSAGE_TYPE type, std::string message) {
  if (type != BROTLIG_PROGRESS) return false;
  g_last_progress = std::stof(message);
  ++g_calls;
  return g_abort_after >= 0 && g_calls > g_abort_after;
}

static std::vector<uint8_t> slurp(const char* path) {
  std::ifstream f(path, std::ios::binary);
  return std::vector<uint8_t>(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  std::vector<uint8_t> src = slurp(argv[1]), want = slurp(argv[2]);
  if (argc > 3) g_abort_after = atoi(argv[3]);
  src.resize(src.size() + 16);   // the reference asks for slack after the stream; harmless here
  const uint32_t src_size = (uint32_t)src.size() - 16;
  const uint32_t n = BrotliG::DecompressedSize(src.data());
  if (n != want.size()) { printf("FAIL DecompressedSize %u != %zu\n", n, want.size()); return 1; }
  std::vector<uint8_t> out(n, 0xA5);
  double time = 0;
  for (int rep = 0; rep < 2; ++rep) {                     // brotlig_cli.cpp:441-444: `time` accumulates over the repeats
    uint32_t out_size = n;
    const double before = time;
    const BROTLIG_ERROR rc = DecodeGPU(false, src_size, src.data(), &out_size, out.data(), time);
    if (rc != BROTLIG_OK || out_size != n) { printf("FAIL DecodeGPU rc %d size %u\n", (int)rc, out_size); return 1; }
    if (!(time > before)) { printf("FAIL time did not accumulate (%f -> %f)\n", before, time); return 1; }
    if (memcmp(out.data(), want.data(), n)) { printf("FAIL DecodeGPU output differs\n"); return 1; }
  }
  // BrotliG::DecodeCPU with the feedback callback: once per page, progress = 100 * page / pages; true stops the decode
  std::fill(out.begin(), out.end(), 0xA5);
  uint32_t out_size = n;
  const BROTLIG_ERROR rc = BrotliG::DecodeCPU(src_size, src.data(), &out_size, out.data(), Feedback);
  if (rc != BROTLIG_OK || out_size != n) { printf("FAIL DecodeCPU rc %d\n", (int)rc); return 1; }
  const uint32_t w0 = src[0] | (src[1] << 8) | (src[2] << 16) | ((uint32_t)src[3] << 24);
  const int pages = (int)(w0 >> 16);
  if (g_abort_after < 0) {
    if (g_calls != pages) { printf("FAIL %d feedback calls for %d pages\n", g_calls, pages); return 1; }
    if (memcmp(out.data(), want.data(), n)) { printf("FAIL DecodeCPU output differs\n"); return 1; }
  } else {
    if (g_calls != g_abort_after + 1) { printf("FAIL abort: %d calls\n", g_calls); return 1; }
    // the decoded prefix is right, and from some page boundary on everything is zero (never stale bytes)
    size_t good = 0;
    while (good < n && out[good] == want[good]) ++good;
    size_t z = good;
    while (z < n && out[z] == 0) ++z;
    if (z != n) { printf("FAIL abort: byte %zu is neither decoded nor zero\n", z); return 1; }
    if (good >= n) { printf("FAIL abort: the whole stream was decoded\n"); return 1; }
  }
  printf("CONSUMER OK pages %d feedback calls %d last progress %.3f kernel ms %.3f\n", pages, g_calls, g_last_progress, time);
  return 0;
}
