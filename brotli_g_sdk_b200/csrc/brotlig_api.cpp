// brotlig_api.cpp -- C++ shim with the reference's decode API names over the C ABI (see
// include/brotlig_b200/BrotliG.h). Decoder contexts are created lazily and pooled: concurrent callers each take
// their own context (the reference's DecodeCPU is re-entrant, BrotligDecoder.cpp:495-519), a finished call returns
// its context to the pool.
#include "../../include/brotlig_b200/BrotliG.h"

#include <mutex>
#include <stdexcept>
#include <vector>

#include "../../include/brotlig_b200.h"

namespace {
std::mutex g_mu;
std::vector<bgx_context*> g_free;

struct ContextLease {
  bgx_context* ctx = nullptr;
  ContextLease() {
    {
      std::lock_guard<std::mutex> lock(g_mu);
      if (!g_free.empty()) { ctx = g_free.back(); g_free.pop_back(); }
    }
    if (!ctx && bgx_create(&ctx, -1) != 0)
      throw std::runtime_error("brotlig_b200: no usable CUDA device / kernel image (built for sm_100a)");
  }
  ~ContextLease() {
    if (!ctx) return;
    std::lock_guard<std::mutex> lock(g_mu);
    g_free.push_back(ctx);
  }
};

int feedback_trampoline(void* user, uint32_t page, uint32_t num_pages) {
  BROTLIG_Feedback_Proc proc = reinterpret_cast<BROTLIG_Feedback_Proc>(user);
  const float progress = 100.f * ((float)page / (float)num_pages);          // BrotligDecoder.cpp:320
  return proc(BROTLIG_PROGRESS, std::to_string(progress)) ? 1 : 0;
}
}  // namespace

uint32_t BrotliG::DecompressedSize(uint8_t* src) { return bgx_decompressed_size(src); }

BROTLIG_ERROR DecodeGPU(bool, uint32_t input_size, const uint8_t* input, uint32_t* output_size, uint8_t* output,
                        double& time) {
  ContextLease lease;
  return static_cast<BROTLIG_ERROR>(bgx_decode_host(lease.ctx, input_size, input, output_size, output, &time));
}

BROTLIG_ERROR BrotliG::DecodeCPU(uint32_t input_size, const uint8_t* src, uint32_t* output_size, uint8_t* output,
                                 BROTLIG_Feedback_Proc feedbackProc) {
  try {
    ContextLease lease;
    double ms = 0;
    if (!feedbackProc) return static_cast<BROTLIG_ERROR>(bgx_decode_host(lease.ctx, input_size, src, output_size, output, &ms));
    // per-page feedback; a true return stops the decode between two page groups, the rest of the output stays zero
    // and the call returns BROTLIG_OK, as the reference does (BrotligDecoder.cpp:318-325,448,490)
    return static_cast<BROTLIG_ERROR>(bgx_decode_host_progress(lease.ctx, input_size, src, output_size, output, &ms, feedback_trampoline,
                                                               reinterpret_cast<void*>(feedbackProc), 0));
  } catch (const std::exception&) {
    return BROTLIG_ERROR_GENERIC;   // C linkage: never let an exception cross it
  }
}
