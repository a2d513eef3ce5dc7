// include/brotlig_b200/BrotliG.h -- C++ host interface of brotli_g_sdk_b200 with the names, argument
// meaning and error behaviour of the reference SDK's decode API, so that code written against
//   /root/reference/inc/BrotliG.h, inc/BrotligDecoder.h:32-33   (DecompressedSize, DecodeCPU)
//   /root/reference/sample/BrotligGPUDecoder.h:24                (DecodeGPU)
//   /root/reference/inc/common/BrotligCommon.h:50-92             (BROTLIG_ERROR, BROTLIG_Feedback_Proc)
// can switch to this library by changing the include path and the library it links.
// Implemented in brotli_g_sdk_b200/csrc/brotlig_api.cpp as a thin shim over the C ABI
// (include/brotlig_b200.h); every decode runs on the GPU -- there is no CPU decode path in this library.
#pragma once
#include <cstdint>
#include <string>

typedef enum {
  BROTLIG_OK = 0,
  BROTLIG_ABORTED,
  BROTLIG_ERROR_MIN_PAGE_SIZE,
  BROTLIG_ERROR_MAX_PAGE_SIZE,
  BROTLIG_ERROR_MAX_NUM_PAGES,
  BROTLIG_ERROR_PRECON_MIN_TEX_WIDTH,
  BROTLIG_ERROR_PRECON_MAX_TEX_WIDTH,
  BROTLIG_ERROR_PRECON_MIN_TEX_HEIGHT,
  BROTLIG_ERROR_PRECON_MAX_TEX_HEIGHT,
  BROTLIG_ERROR_PRECON_MIN_TEX_PITCH,
  BROTLIG_ERROR_PRECON_MAX_TEX_PITCH,
  BROTLIG_ERROR_PRECON_MIN_TEX_MIPLEVELS,
  BROTLIG_ERROR_PRECON_MAX_TEX_MIPLEVELS,
  BROTLIG_ERROR_PRECON_INCORRECT_FORMAT,
  BROTLIG_ERROR_CORRUPT_STREAM,
  BROTLIG_ERROR_INCORRECT_STREAM_FORMAT,
  BROTLIG_ERROR_GENERIC
} BROTLIG_ERROR;

typedef enum { BROTLIG_PROGRESS, BROTLIG_WARNING } BROTLIG_MESSAGE_TYPE;

#define BROTLIG_API
// Return true from the callback to abort (reference: BrotligDecoder.cpp:318-325).
typedef bool(BROTLIG_API* BROTLIG_Feedback_Proc)(BROTLIG_MESSAGE_TYPE type, std::string message);

namespace BrotliG {
extern "C" {
// Size of the decompressed data, from the stream header only.
uint32_t BROTLIG_API DecompressedSize(uint8_t* src);
// Same contract as the reference's DecodeCPU: *output_size is the buffer size on entry and the
// decompressed size on return; feedbackProc (nullable) is called for every page with the progress in
// percent (the stream is then decoded in groups of pages and the pages of a finished group are reported
// in order); when it returns true no further group is decoded, the rest of the output stays zero and
// the call returns BROTLIG_OK -- the reference's behaviour (BrotligDecoder.cpp:318-325,448,490).
// Re-entrant: concurrent callers use separate decoder contexts.
BROTLIG_ERROR BROTLIG_API DecodeCPU(uint32_t input_size, const uint8_t* src, uint32_t* output_size, uint8_t* output,
                                    BROTLIG_Feedback_Proc feedbackProc);
}
}  // namespace BrotliG

// Same contract as the reference sample's DecodeGPU: `time` is INCREMENTED by the kernel-only device
// milliseconds; useWarpDevice is accepted and ignored (there is no software adapter on CUDA).
// Throws std::runtime_error if no usable CUDA device exists (the reference throws on D3D12 failures).
BROTLIG_ERROR DecodeGPU(bool useWarpDevice, uint32_t input_size, const uint8_t* input, uint32_t* output_size,
                        uint8_t* output, double& time);
