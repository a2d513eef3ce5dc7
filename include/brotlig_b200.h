/* brotlig_b200.h -- C ABI of the B200-native Brotli-G decompressor (libbrotlig_b200.so).
 *
 * This is the drop-in boundary of the project: plain C, plain pointers and sizes, no torch / CUDA types
 * in the signatures (a CUDA stream is passed as an opaque void*). Each entry point names the reference
 * interface it stands in for; INTEGRATION.md shows the bindings a maintainer of the reference would add.
 *
 * Error codes are the values of the reference's BROTLIG_ERROR enum
 * (/root/reference/inc/common/BrotligCommon.h:50-68): 0 = BROTLIG_OK, 14 = BROTLIG_ERROR_CORRUPT_STREAM,
 * 15 = BROTLIG_ERROR_INCORRECT_STREAM_FORMAT, 16 = BROTLIG_ERROR_GENERIC (also used for CUDA failures;
 * bgx_last_error() then describes the cause).
 *
 * There is NO CPU fallback anywhere behind this header: if no CUDA device is usable, bgx_create fails.
 */
#ifndef BROTLIG_B200_H
#define BROTLIG_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct bgx_context bgx_context;
typedef struct bgx_plan bgx_plan;

/* Creates a decoder bound to CUDA device `device` (-1 = current). Owns a CUDA stream and staging buffers. */
int bgx_create(bgx_context** ctx, int device);
void bgx_destroy(bgx_context* ctx);
const char* bgx_last_error(const bgx_context* ctx);

/* = BrotliG::DecompressedSize (/root/reference/inc/BrotligDecoder.h:32, src/BrotligDecoder.cpp:34-38). Host-only. */
uint32_t bgx_decompressed_size(const uint8_t* src);

/* = DecodeGPU(useWarpDevice, input_size, input, output_size, output, time)
 *   (/root/reference/sample/BrotligGPUDecoder.h:24, sample/BrotligGPUDecoder.cpp:260-748) and, with the
 *   same argument meaning, BrotliG::DecodeCPU (/root/reference/inc/BrotligDecoder.h:33).
 * Host pointers in and out. *output_size: in = capacity of `output`, out = uncompressed size.
 * *kernel_ms (nullable) is INCREMENTED by the device time of the decode kernels only, exactly like the
 * reference's `double& time` (BrotligGPUDecoder.cpp:729-746); uploads/downloads are excluded from it. */
int bgx_decode_host(bgx_context* ctx, uint32_t input_size, const uint8_t* input, uint32_t* output_size, uint8_t* output,
                    double* kernel_ms);

/* Batch form of the above: n independent streams decoded by ONE launch sequence (the reference's shader
 * consumes a list of streams per dispatch: src/decoder/BrotliGCompute.hlsl:1757-1775). Host pointers.
 * output_sizes[i]: in = capacity, out = uncompressed size. Uploads, kernels and downloads are pipelined over groups
 * of streams and, for streams that are large against the batch, over page ranges of a stream (pinned host
 * buffers make the copies asynchronous; pageable ones work, serialised by the driver). */
int bgx_decode_batch_host(bgx_context* ctx, uint32_t n, const uint8_t* const* inputs, const uint32_t* input_sizes,
                          uint8_t* const* outputs, uint32_t* output_sizes, double* kernel_ms);

/* bgx_decode_host with the per-page feedback of BrotliG::DecodeCPU (/root/reference/src/BrotligDecoder.cpp:318-325:
 * the callback runs after a page, a non-zero return stops the decode). The stream is decoded in groups of
 * `pages_per_group` pages (0 = a sixteenth of the stream, at least 16); after each group progress(user, page, num_pages)
 * is called for its pages in order. When it returns non-zero no further group is decoded, the bytes of the pages that
 * were not decoded are zero (the reference memsets the output first, :448) and the call returns 0 like the reference. */
typedef int (*bgx_progress_fn)(void* user, uint32_t page_index, uint32_t num_pages);
int bgx_decode_host_progress(bgx_context* ctx, uint32_t input_size, const uint8_t* input, uint32_t* output_size, uint8_t* output,
                             double* kernel_ms, bgx_progress_fn progress, void* user, uint32_t pages_per_group);

/* bgx_decode_batch_host over several devices of one box from ONE process: ctxs[] are contexts created on different
 * devices; whole streams are assigned to them (balanced by compressed size, no inter-GPU traffic: pages and streams are
 * independent, /root/reference/src/decoder/PageDecoder.cpp:126-153) and decoded concurrently, one host thread per device.
 * *kernel_ms is incremented by the slowest device's kernel time. (The one-process-per-GPU launcher with the NCCL
 * broadcast of a sharded stream is brotli_g_sdk_b200/multi_gpu.py.) */
int bgx_decode_batch_host_multi(bgx_context* const* ctxs, uint32_t n_ctx, uint32_t n, const uint8_t* const* inputs,
                                const uint32_t* input_sizes, uint8_t* const* outputs, uint32_t* output_sizes, double* kernel_ms);

/* ---- device-resident interface (no host<->device traffic in the decode path) ---- */
typedef struct bgx_stream {
  const uint8_t* d_src;      /* device pointer to the stream (16-byte aligned) */
  uint32_t src_size;         /* stream bytes */
  uint32_t src_capacity;     /* readable bytes at d_src, >= src_size rounded up to 16; the kernel never reads beyond it */
  uint8_t* d_dst;            /* device pointer: where page `page_begin` of the stream is written */
  uint32_t dst_capacity;     /* writable bytes at d_dst */
  uint32_t page_begin;       /* first page to decode */
  uint32_t page_count;       /* pages to decode; 0 = all from page_begin. (page ranges: multi-GPU sharding) */
  uint8_t header[16];        /* host copy of the first 16 bytes of the stream (StreamHeader + PreconditionHeader) */
} bgx_stream;

/* Parses the headers, builds the device work descriptors (one flat queue of pages over all streams). */
int bgx_plan_create(bgx_context* ctx, const bgx_stream* streams, uint32_t n, bgx_plan** plan);
/* Enqueues the decode on `cuda_stream` (a cudaStream_t; NULL = the context's stream). Does not synchronise. */
int bgx_plan_launch(bgx_context* ctx, bgx_plan* plan, void* cuda_stream);
/* Waits for the launch and returns 0 or BROTLIG_ERROR_CORRUPT_STREAM; *bad_pages (nullable) = pages with errors. */
int bgx_plan_finish(bgx_context* ctx, bgx_plan* plan, uint32_t* bad_pages);
void bgx_plan_destroy(bgx_plan* plan);

typedef struct bgx_plan_info {
  uint64_t pages, raw_pages_unknown;     /* raw_pages_unknown: always 0 (page sizes live in device memory) */
  uint64_t compressed_bytes;             /* sum of stream bytes covered by the plan (algorithmic read bytes) */
  uint64_t uncompressed_bytes;           /* sum of bytes produced (algorithmic write bytes) */
  uint32_t kernels_per_launch;           /* CUDA kernels one bgx_plan_launch enqueues */
  uint32_t grid_blocks, block_threads, smem_bytes_per_block, sm_count;
} bgx_plan_info;
void bgx_plan_get_info(const bgx_plan* plan, bgx_plan_info* info);

#ifdef __cplusplus
}
#endif
#endif
