// brotlig_api.cpp -- C++ shim with the reference's decode API names over the C ABI (see
// include/brotlig_b200/BrotliG.h). One process-wide decoder context per process, created lazily.
#include "../../include/brotlig_b200/BrotliG.h"

#include <mutex>
#include <stdexcept>

#include "../../include/brotlig_b200.h"

namespace {
std::mutex g_mu;
bgx_context* g_ctx = nullptr;

bgx_context* context_or_throw() {
  std::lock_guard<std::mutex> lock(g_mu);
  if (!g_ctx && bgx_create(&g_ctx, -1) != 0)
    throw std::runtime_error("brotlig_b200: no usable CUDA device / kernel image (built for sm_100a)");
  return g_ctx;
}
}  // namespace

uint32_t BrotliG::DecompressedSize(uint8_t* src) { return bgx_decompressed_size(src); }

BROTLIG_ERROR DecodeGPU(bool, uint32_t input_size, const uint8_t* input, uint32_t* output_size, uint8_t* output,
                        double& time) {
  bgx_context* ctx = context_or_throw();
  std::lock_guard<std::mutex> lock(g_mu);
  return static_cast<BROTLIG_ERROR>(bgx_decode_host(ctx, input_size, input, output_size, output, &time));
}

BROTLIG_ERROR BrotliG::DecodeCPU(uint32_t input_size, const uint8_t* src, uint32_t* output_size, uint8_t* output,
                                 BROTLIG_Feedback_Proc feedbackProc) {
  double ms = 0;
  BROTLIG_ERROR rc;
  try {
    rc = DecodeGPU(false, input_size, src, output_size, output, ms);
  } catch (const std::exception&) {
    return BROTLIG_ERROR_GENERIC;   // C linkage: never let an exception cross it
  }
  if (rc == BROTLIG_OK && feedbackProc && feedbackProc(BROTLIG_PROGRESS, std::to_string(100.f))) return BROTLIG_ABORTED;
  return rc;
}
