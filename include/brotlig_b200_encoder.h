/* brotlig_b200_encoder.h -- C ABI of the CPU-side Brotli-G stream encoder shipped with brotli_g_sdk_b200.
 *
 * The reference encoder (BrotliG::Encode, /root/reference/inc/BrotligEncoder.h:34-37,
 * src/BrotligEncoder.cpp:663-692, src/encoder/PageEncoder.cpp:247-574) depends on google/brotli
 * v1.0.9 *internals* that are not vendored and cannot be built offline. This
 * library is a from-scratch encoder that emits the same wire format (stream header, precondition
 * header, page table, page header, 32 swizzled sub-streams, three prefix-code tables, rounds of
 * commands + literals) so that streams for tests and benchmarks can be produced anywhere. It is a
 * CPU library: it has no CUDA dependency and is NOT on the decode hot path.
 */
#ifndef BROTLIG_B200_ENCODER_H
#define BROTLIG_B200_ENCODER_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct bgxenc_options {
  uint32_t page_size;        /* 32768, 65536 or 131072 (0 => 65536). Replaces Encode()'s page_size argument. */
  int32_t  npostfix;         /* distance postfix bits 0..3, or -1 = choose per page (PageEncoder.cpp:324-377) */
  int32_t  ndirect_msb;      /* NDIRECT >> NPOSTFIX, 0..15, or -1 = choose per page */
  int32_t  max_chain;        /* LZ77 hash-chain probes per position (0 => 16) */
  int32_t  lazy;             /* 1 = one-step lazy matching */
  int32_t  use_ring_codes;   /* 1 = emit distance short codes 0..15 when they apply */
  int32_t  rle_mode;         /* code-length RLE: 0 = like the reference (BrotligUtils.cpp:118-228), 1 = none */
  int32_t  split_insert_over;/* >0: literal runs longer than this are split off as insert-only commands (test coverage) */
  int32_t  allow_raw;        /* 1 = store a page raw when that is not larger (reference behaviour) */
  int32_t  num_threads;      /* 0 = hardware concurrency */
  /* pre-conditioning (BrotligDataconditionParams, inc/common/BrotligDataConditioner.h:29-62) */
  int32_t  precondition;     /* 0/1 */
  int32_t  format;           /* 1..5 = BC1..BC5 */
  uint32_t width_blocks, height_blocks;   /* mip 0, in 4x4 blocks */
  uint32_t pitch_bytes;      /* mip 0 row pitch; 0 => tight (or 256-aligned when pitch_aligned) */
  uint32_t num_mips;         /* >= 1 */
  int32_t  swizzle;          /* 2x2 block-group swizzle */
  int32_t  pitch_aligned;    /* D3D12 256-byte pitch alignment of every mip */
  int32_t  delta_encode;     /* delta-code the colour end-point planes per page */
} bgxenc_options;

void     bgxenc_default_options(bgxenc_options* opt);
/* Upper bound of the stream size for input_size bytes (header + tables + raw pages). */
uint32_t bgxenc_max_compressed_size(uint32_t input_size, uint32_t page_size, int precondition);
/* Encodes src[0..size) into dst. *dst_size: in = capacity, out = stream bytes. Returns a BROTLIG_ERROR value (0 = OK). */
int      bgxenc_encode(const uint8_t* src, uint32_t size, uint8_t* dst, uint32_t* dst_size, const bgxenc_options* opt);
/* Forward BCn pre-conditioner alone (twin of BrotliG::Condition, src/common/BrotligDataConditioner.cpp:121-133). */
int      bgxenc_condition(const uint8_t* src, uint32_t size, uint8_t* dst, const bgxenc_options* opt);
/* Per-stream statistics of the last bgxenc_encode call on this thread (for tests/bench reporting). */
typedef struct bgxenc_stats {
  uint64_t pages, raw_pages, commands, literals, ring_code_hits[16], implicit_dist0, insert_only_cmds;
  uint64_t table_types[3][3];   /* [alphabet: icp,dist,lit][type: trivial,simple,complex] */
} bgxenc_stats;
void     bgxenc_last_stats(bgxenc_stats* out);

#ifdef __cplusplus
}
#endif
#endif
