// bgx_format.h -- Brotli-G wire-format constants and header helpers shared by the CUDA page
// decoder, the host launcher and the CPU-side encoder of brotli_g_sdk_b200.
//
// Everything here restates the *format* (not the code) defined by the reference SDK:
//   stream header            /root/reference/inc/DataStream.h:28-87
//   precondition header      /root/reference/inc/DataStream.h:89-108
//   page table semantics     /root/reference/src/BrotligDecoder.cpp:310-314, src/BrotligEncoder.cpp:579-605
//   page header / size table /root/reference/src/decoder/PageDecoder.cpp:79-121
//   alphabets and limits     /root/reference/inc/common/BrotligConstants.h:32-243
//   BCn block layouts        /root/reference/inc/common/BrotligDataConditioner.h:92-237
//   insert/copy length codes RFC 7932 section 5 (google/brotli v1.0.9 c/enc/command.h, un-vendored
//                            dependency of the reference: fetch_dependencies.py:79)
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <initializer_list>

#if defined(__CUDACC__)
#define BGX_HD __host__ __device__ __forceinline__
#else
#define BGX_HD inline
#endif

namespace bgx {

// ---- alphabets (BrotligConstants.h:34-45) ----
constexpr int kNumLitSymbols = 256;
constexpr int kNumCmdSymbolsBrotli = 704;       // RFC 7932 insert&copy alphabet
constexpr int kCmdSentinel = 704;               // "end of page"
constexpr int kNumCmdSymbols = 728;             // 704 + sentinel + 23 insert-only symbols (705..727)
constexpr int kNumDistSymbols = 544;
constexpr int kMaxCodeLen = 15;                 // BROTLIG_HUFFMAN_MAX_CODE_LENGTH
constexpr int kMaxCodeLenCodeLen = 9;           // BROTLIG_HUFFMAN_MAX_CODE_LENGTH_CODE_LENGTH
constexpr int kNumCodeLenCodes = 18;
constexpr int kRepeatPrev = 16;
constexpr int kRepeatZero = 17;
constexpr int kInitialRepeatLen = 8;

// ---- stream / page geometry ----
constexpr int kNumSubstreams = 32;              // BROLTIG_DEFAULT_NUM_BITSTREAMS
constexpr uint32_t kMinPageSize = 32u * 1024u;  // PageSize = kMinPageSize << PageSizeIdx
constexpr uint32_t kMaxPageSize = 128u * 1024u;
constexpr uint32_t kStreamId = 5;
constexpr uint32_t kStreamHeaderBytes = 8;
constexpr uint32_t kPreconHeaderBytes = 8;
constexpr uint32_t kMaxPagesPerStream = 65535;
// The reference decoders deliberately read a few bytes past the end of a sub-stream / page
// (BrotligDeswizzler.h:74-81); device buffers handed to the kernel must have this much slack.
constexpr uint32_t kInputSlackBytes = 16;

// ---- error codes (values of BROTLIG_ERROR, BrotligCommon.h:50-68) ----
enum : int {
  kOk = 0,
  kAborted = 1,
  kErrCorruptStream = 14,
  kErrIncorrectStreamFormat = 15,
  kErrGeneric = 16,
};

// ---- RFC 7932 section 5 length-code tables ----
BGX_HD uint32_t insert_base(uint32_t code) {
  const uint32_t t[24] = {0, 1, 2, 3, 4, 5, 6, 8, 10, 14, 18, 26, 34, 50, 66, 98, 130, 194, 322, 578, 1090, 2114, 6210, 22594};
  return t[code];
}
BGX_HD uint32_t insert_extra_bits(uint32_t code) {
  const uint8_t t[24] = {0, 0, 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 7, 8, 9, 10, 12, 14, 24};
  return t[code];
}
BGX_HD uint32_t copy_base(uint32_t code) {
  const uint32_t t[24] = {2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 14, 18, 22, 30, 38, 54, 70, 102, 134, 198, 326, 582, 1094, 2118};
  return t[code];
}
BGX_HD uint32_t copy_extra_bits(uint32_t code) {
  const uint8_t t[24] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 7, 8, 9, 10, 24};
  return t[code];
}
// insert&copy symbol (< 704) -> (insert code, copy code); cells of 64 symbols, RFC 7932 section 5 table.
BGX_HD uint32_t icp_insert_code(uint32_t sym) {
  const uint32_t cell = sym >> 6;                       // 0..10
  return (((0x298500u >> (2 * cell)) & 3u) << 3) | ((sym >> 3) & 7u);
}
BGX_HD uint32_t icp_copy_code(uint32_t sym) {
  const uint32_t cell = sym >> 6;
  return (((0x262444u >> (2 * cell)) & 3u) << 3) | (sym & 7u);
}

BGX_HD uint32_t floor_log2(uint32_t x) {  // x > 0
#if defined(__CUDA_ARCH__)
  return 31u - (uint32_t)__clz((int)x);
#else
  return 31u - (uint32_t)__builtin_clz(x);
#endif
}
// "Log2Floor" of the reference (BrotligUtils.cpp:49-56) is really the bit length.
BGX_HD uint32_t bit_length(uint32_t x) { return x ? floor_log2(x) + 1u : 0u; }

// ---- stream header (8 bytes, little-endian bit-fields; DataStream.h:28-36) ----
struct StreamInfo {
  uint32_t num_pages;
  uint32_t page_size;
  uint32_t last_page_size;     // 0 => last page is a full page
  uint32_t preconditioned;
  uint32_t uncompressed_size;
  uint32_t header_bytes;       // 8 or 16: offset of the page table from the stream start
};

BGX_HD uint32_t load_le32(const uint8_t* p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

// Returns kOk / kErrCorruptStream / kErrIncorrectStreamFormat exactly like
// DecodeCPUMultithreaded (BrotligDecoder.cpp:436-446).
BGX_HD int parse_stream_header(const uint8_t* s, StreamInfo* out) {
  const uint32_t w0 = load_le32(s), w1 = load_le32(s + 4);
  const uint32_t id = w0 & 0xffu, magic = (w0 >> 8) & 0xffu;
  if (id != (magic ^ 0xffu)) return kErrCorruptStream;
  if (id != kStreamId) return kErrIncorrectStreamFormat;
  out->num_pages = w0 >> 16;
  out->page_size = kMinPageSize << (w1 & 3u);
  out->last_page_size = (w1 >> 2) & 0x3ffffu;
  out->preconditioned = (w1 >> 20) & 1u;
  // The API carries sizes in 32 bits (BrotligDecoder.h:32-33), the header can describe up to 65535 x 128 KiB = 8 GiB:
  // the reference lets that product wrap (DataStream.h:60-64); a stream whose size does not fit is rejected here, and
  // so is a last page larger than a page.
  const uint64_t full = (uint64_t)out->num_pages * out->page_size;
  if (out->last_page_size > out->page_size) return kErrCorruptStream;
  const uint64_t size = full - (out->last_page_size && out->num_pages ? out->page_size - out->last_page_size : 0u);
  if (size > 0xffffffffull) return kErrCorruptStream;
  out->uncompressed_size = (uint32_t)size;
  out->header_bytes = kStreamHeaderBytes + (out->preconditioned ? kPreconHeaderBytes : 0u);
  return kOk;
}

// ---- pre-conditioning (BCn) layout, decode-side view ----
constexpr int kMaxSubBlocks = 6;
constexpr int kMaxMips = 32;

struct PreconLayout {
  uint32_t swizzle;
  uint32_t pitch_aligned;
  uint32_t format;                 // 1..5 = BC1..BC5
  uint32_t num_mips;
  uint32_t block_bytes;
  uint32_t num_sub;
  uint32_t sub_size[kMaxSubBlocks];
  uint32_t sub_off[kMaxSubBlocks];         // byte offset of the field inside a block
  uint32_t num_color_sub;
  uint32_t color_sub[4];
  uint32_t width_blocks[kMaxMips];
  uint32_t height_blocks[kMaxMips];
  uint32_t pitch_bytes[kMaxMips];
  uint32_t num_blocks[kMaxMips];
  uint32_t mip_off_bytes[kMaxMips + 1];
  uint32_t mip_off_blocks[kMaxMips + 1];
  uint32_t sub_stream_off[kMaxSubBlocks + 1];   // start of each field's plane in the conditioned buffer
  uint32_t total_blocks;
  uint32_t valid;                  // Initialize() succeeded (sizes consistent with the output size)
};

// Mirrors BrotligDataconditionParams::Initialize (BrotligDataConditioner.h:92-237) for the decode
// side, where width/height/pitch of mip 0 come from the PreconditionHeader (+1 applied by caller,
// BrotligDecoder.cpp:470-476). Returns false when the reference's Initialize does, and -- unlike the reference, which
// trusts these header fields and then scatters outside the texture -- when a pitch is smaller than a row of blocks,
// when the blocks do not fit the output, or when the sizes overflow 32 bits.
inline bool precon_layout_init(PreconLayout* L, uint32_t format, uint32_t width_blocks, uint32_t height_blocks,
                               uint32_t pitch_bytes, uint32_t num_mips, bool swizzle, bool pitch_aligned,
                               uint32_t out_size) {
  *L = PreconLayout{};
  L->swizzle = swizzle;
  L->pitch_aligned = pitch_aligned;
  L->format = format;
  auto set = [&](uint32_t bb, std::initializer_list<uint32_t> sizes, std::initializer_list<uint32_t> colors) {
    L->block_bytes = bb;
    L->num_sub = 0;
    for (uint32_t s : sizes) L->sub_size[L->num_sub++] = s;
    L->num_color_sub = 0;
    for (uint32_t c : colors) L->color_sub[L->num_color_sub++] = c;
  };
  switch (format) {
    case 1: set(8, {2, 2, 4}, {0, 1}); break;                  // BC1: c0, c1, indices
    case 2: set(16, {8, 2, 2, 4}, {1, 2}); break;              // BC2: alpha vector, c0, c1, indices
    case 3: set(16, {1, 1, 6, 2, 2, 4}, {3, 4}); break;        // BC3: a0, a1, alpha idx, c0, c1, idx
    case 4: set(8, {1, 1, 6}, {0, 1}); break;                  // BC4: r0, r1, idx
    case 5: set(16, {1, 1, 6, 1, 1, 6}, {0, 1, 3, 4}); break;  // BC5: r0, r1, idx, g0, g1, idx
    default: set(1, {1}, {}); break;
  }
  const uint32_t block_px = (format >= 1 && format <= 5) ? 4u : 1u;
  if (num_mips == 0) num_mips = 1;
  if (num_mips > (uint32_t)kMaxMips) return false;
  L->num_mips = num_mips;
  L->width_blocks[0] = width_blocks;
  L->height_blocks[0] = height_blocks;
  L->total_blocks = L->num_blocks[0] = width_blocks * height_blocks;   // (15 + 15 bits)
  auto round_up = [](uint32_t n, uint32_t m) { return ((n + m - 1) / m) * m; };
  L->pitch_bytes[0] = pitch_bytes ? pitch_bytes
                                  : (pitch_aligned ? round_up(width_blocks * L->block_bytes, 256u)
                                                   : width_blocks * L->block_bytes);
  uint32_t wpx = (width_blocks * block_px) / 2, hpx = (height_blocks * block_px) / 2;
  for (uint32_t mip = 1; mip <= num_mips; ++mip) {
    if (mip < num_mips) {
      L->width_blocks[mip] = (wpx + block_px - 1) / block_px;
      L->height_blocks[mip] = (hpx + block_px - 1) / block_px;
      L->num_blocks[mip] = L->width_blocks[mip] * L->height_blocks[mip];
      L->pitch_bytes[mip] = pitch_aligned ? round_up(L->width_blocks[mip] * L->block_bytes, 256u)
                                          : L->width_blocks[mip] * L->block_bytes;
      L->total_blocks += L->num_blocks[mip];
    }
    if (L->pitch_bytes[mip - 1] < L->width_blocks[mip - 1] * L->block_bytes) return false;   // rows would overlap / leave the texture
    const uint64_t end = (uint64_t)L->mip_off_bytes[mip - 1] + (uint64_t)L->pitch_bytes[mip - 1] * L->height_blocks[mip - 1];
    if (end > out_size) return false;
    L->mip_off_bytes[mip] = (uint32_t)end;
    L->mip_off_blocks[mip] = L->mip_off_blocks[mip - 1] + L->num_blocks[mip - 1];
    wpx /= 2;
    hpx /= 2;
  }
  // The reference bails out here with isInitialized=false but keeps decoding with whatever was
  // filled in so far (sub_off / sub_stream_off still zero). We report it; callers decide.
  if (L->mip_off_bytes[num_mips] != out_size) return false;
  if ((uint64_t)L->total_blocks * L->block_bytes > out_size) return false;   // the conditioned planes hold out_size bytes
  for (uint32_t sub = 1; sub <= L->num_sub; ++sub) {
    if (sub < L->num_sub) L->sub_off[sub] = L->sub_off[sub - 1] + L->sub_size[sub - 1];
    L->sub_stream_off[sub] = L->sub_stream_off[sub - 1];
    for (uint32_t mip = 0; mip < num_mips; ++mip) L->sub_stream_off[sub] += L->num_blocks[mip] * L->sub_size[sub - 1];
  }
  if (L->sub_stream_off[L->num_sub] != L->total_blocks * L->block_bytes) return false;
  L->valid = 1;
  return true;
}

struct PreconHeaderFields {
  uint32_t swizzled, pitch_aligned, width_blocks, height_blocks, format, num_mips, pitch_bytes;  // +1 applied
};
BGX_HD PreconHeaderFields parse_precon_header(const uint8_t* p) {
  const uint32_t w0 = load_le32(p), w1 = load_le32(p + 4);
  PreconHeaderFields f;
  f.swizzled = w0 & 1u;
  f.pitch_aligned = (w0 >> 1) & 1u;
  f.width_blocks = ((w0 >> 2) & 0x7fffu) + 1u;
  f.height_blocks = ((w0 >> 17) & 0x7fffu) + 1u;
  f.format = w1 & 0xffu;
  f.num_mips = ((w1 >> 8) & 0x1fu) + 1u;
  f.pitch_bytes = ((w1 >> 13) & 0x7ffffu) + 1u;
  return f;
}

}  // namespace bgx
