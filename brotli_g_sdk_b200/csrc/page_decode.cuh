// page_decode.cuh -- the Brotli-G page decoder for sm_100a: TWO WARPS DECODE ONE PAGE (a producer that owns
// the 32 bit readers, lane = sub-stream, and a consumer that assembles the output).
//
// This is a from-scratch CUDA design (not a translation of src/decoder/BrotliGCompute.hlsl). What it
// has to reproduce bit-for-bit is the behaviour of the reference CPU page decoder:
//   PageDecoder::Run            /root/reference/src/decoder/PageDecoder.cpp:65-268
//   LoadHuffmanTable            /root/reference/src/decoder/BrotligHuffmanTable.cpp:73-205
//   DecodeCommand / Translate.. /root/reference/src/decoder/PageDecoder.cpp:290-404
//   BrotligDeswizzler           /root/reference/inc/common/BrotligDeswizzler.h:43-206
//
// Design (see DESIGN.md for the full picture):
//   * per-page shared-memory arena (WarpSmem, ~13.4 KB, 16 pages resident per SM): three LSB-first primary
//     look-up tables (9/10/9 bits) + canonical "limit/base/sorted" arrays for the rare longer codes, a
//     512 B literal ring, a 2 KB output ring that write-combines the page before it goes to HBM as 16-byte
//     coalesced stores (near matches are served from the ring, far matches from L1/L2), a per-lane cp.async
//     staging ring for the compressed input and a ring of kQ round buffers between the two warps.
//   * the bit reader of a lane is a 64-bit window (w0,w1) + one prefetched word, refilled from the staging
//     ring with one predicated LDS; a peek is a single funnel shift; staging is topped up at a few explicit
//     places per round.
//   * table build is warp-parallel: ballot/scan over the run-length coded code lengths, match_any
//     ranking for the canonical order, cooperative LUT fill for short codes.
//   * PRODUCER, per round of <=32 commands: speculative command decode in every lane, parallel relaxation
//     of the distance ring, one 64-bit warp scan for output and literal positions, literal decode (two per
//     peek) into the literal ring; the round is published through an mbarrier.
//   * CONSUMER, per round: flattened literal inserts (lane t places literal t), then match copies: every
//     copy whose source is already final goes in one pass flattened over 4-byte pieces (two aligned words +
//     funnel shift per piece), the few that depend on copies of the same round follow in command order.
//   * rounds that produce more than kRoundMax bytes or need more literals than the literal ring
//     holds (long runs) are split by the producer into virtual rounds that fit.
//
// The file is also compiled by g++ against tests/emul/warp_emul.h (BGX_EMULATED) so that the very
// same code can be exercised on the CPU-only development box. That emulator is test infrastructure.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <type_traits>

#include "bgx_format.h"

#ifndef BGX_EMULATED
#include <cuda_runtime.h>
#define BGX_DEV __device__ __forceinline__
#define BGX_DEV_NOINLINE __device__ __forceinline__
#define BGX_COLD __device__ __noinline__     // rare paths: out of line, so that the round loops stay small
#else
#define BGX_DEV inline
#define BGX_DEV_NOINLINE inline
#define BGX_COLD inline
#endif

#ifndef BGX_WAIT_SLEEP_NS
#define BGX_WAIT_SLEEP_NS 0
#endif
#ifndef BGX_RAW_PATH
#define BGX_RAW_PATH 2   // 2: raw pages leave through TMA bulk stores (copy_page_cta_bulk); 1: register copies only
#endif

namespace bgxk {

#ifdef BGX_STATS   // emulator-only instrumentation (never defined in the CUDA build)
struct EmuStats { uint64_t rounds, slow_rounds, wavefronts, sum_max_cpy, copy_bytes, copies, ins_bytes, sum_max_ins, ring_iters,
                  lits, coop_copies, far_bytes, overlap_copies, dep_copies; };
inline EmuStats& emu_stats() { static EmuStats s; return s; }
#define BGX_STAT(expr) do { if (lane == 0) { expr; } } while (0)
#else
#define BGX_STAT(expr) do { } while (0)
#endif

constexpr unsigned kFull = 0xffffffffu;
constexpr int kCmdLutBits = 9;
#ifndef BGX_LITBITS
#define BGX_LITBITS 10
#endif
constexpr int kLitLutBits = BGX_LITBITS;
constexpr int kDistLutBits = 9;
#ifndef BGX_LITQ
#define BGX_LITQ 512
#endif
constexpr uint32_t kLitQ = BGX_LITQ;  // literal ring bytes (power of two): the literals of the rounds in flight
#ifndef BGX_RING
#define BGX_RING 2048
#endif
constexpr uint32_t kRing = BGX_RING;      // output ring bytes (power of two, >= kFlushChunk + kRoundMax)
constexpr uint32_t kRoundMax = 1024;  // largest round (bytes produced) the ring path accepts
#ifndef BGX_SPLIT_LITS
#define BGX_SPLIT_LITS 384
#endif
constexpr uint32_t kSplitLits = BGX_SPLIT_LITS;  // literals per virtual round when a long round is split (384 of the 512-byte literal ring:
                                                 // fewer, larger virtual rounds beat two rounds of 256 in flight -- textures +2.5 %)
static_assert((kLitQ & (kLitQ - 1)) == 0 && kSplitLits + 128 <= kLitQ, "a virtual round's literals (+ rows decoded ahead) must fit the literal ring");
constexpr uint32_t kFlushChunk = 512; // 32 lanes x 16 B
constexpr uint32_t kCoopLen = 32;     // inserts/copies at least this long are done by the whole warp
constexpr uint16_t kLongCode = 0xffffu;  // primary-LUT marker: code longer than the LUT index

// page status bits written to PageResult.status
enum : uint32_t {
  kPageOk = 0,
  kPageErrOverrun = 1,      // commands produce more bytes than the page holds
  kPageErrDistance = 2,     // distance 0 or reaching before the start of the page
  kPageErrLiterals = 4,     // a round needs more literals than it carries
  kPageErrTable = 8,        // malformed prefix-code description
  kPageErrHang = 16,        // the two warps of the page stopped handing rounds over (cannot happen on a valid stream)
  kPageErrLayout = 32,      // the shared-memory arena is not where this build expects it (never on a supported toolchain)
};

struct HuffAux {
  uint16_t limit[16];   // limit[L] = (first_code[L] + count[L]) << (15 - L): left-aligned exclusive upper bound
  uint16_t base[16];    // base[L]  = offset_in_sorted[L] - first_code[L]   (mod 2^16)
};

#ifndef BGX_Q
#define BGX_Q 2
#endif
constexpr uint32_t kQ = BGX_Q;   // rounds in flight between the producer and the consumer warp (power of two)
struct RoundBuf {            // one round of <= 32 commands, producer -> consumer (a round never produces more than
  uint2 cmd[32];             //   kRoundMax bytes: longer ones travel as several virtual rounds)
};                           //   .x resolved match distance
                             //   .y inclusive prefix sums over the commands: bytes produced (bits 0..11) | literals consumed
                             //      (bits 16..27), plus the round's flags (kPkLast / kPkAbort, the same in every lane)
enum : uint32_t { kPkLast = 1u << 12, kPkAbort = 1u << 13, kPkSums = 0x0fff0fffu };
struct PageCtl {             // hand-over state of the two warps of a page
  uint32_t err, is_delta;
  uint32_t phead[kQ];            // producer only: literal head at the start of the round in that slot
};

struct WarpSmem {
  // The first three members are placed so that the output ring sits on a 2048-byte and the literal ring on a 512-byte
  // boundary of the shared-memory WINDOW (static shared memory of a CTA starts 0x400 into its window on sm_90+), which
  // turns "base + (position & mask)" into one LOP3 ((position & mask) | base) everywhere. bgx_decode_pages_kernel checks
  // the two addresses once and refuses to decode (kPageErrLayout) if a toolchain ever lays the arena out differently.
  uint16_t lut_cmd[1 << kCmdLutBits];   // 1024 B
  uint8_t ring[kRing];              // output ring; while tables are read: the list of used symbols (u16 each) and,
                                    // unless that list may need the room, the 512 x u16 code-length-code LUT
  uint8_t litq[kLitQ];              // literal ring, indexed by page-global literal index
  uint16_t lut_lit[1 << kLitLutBits];
  uint16_t lut_dist[1 << kDistLutBits];
  uint16_t sorted_cmd[bgx::kNumCmdSymbols];
  uint16_t sorted_dist[bgx::kNumDistSymbols];
  uint8_t sorted_lit[bgx::kNumLitSymbols];
  HuffAux aux[3];
  uint32_t lenlut[48];              // [0..23] insert code, [24..47] copy code: base | extra_bits << 16
  alignas(16) uint32_t scratch[224]; // table build: cnt[16], next[16], 18 code-length-code lengths;
                                    // consumer: [0..31] insert table, [32..159] wave-1 copies (uint4), [160..223] the other copies (two words each)
  alignas(16) uint4 stage[32][4];   // compressed-input staging: per lane 4 slots x 16 B (cp.async ring)
  RoundBuf rb[kQ];
  alignas(8) uint64_t mbar[2 * kQ];   // full[kQ], empty[kQ]
  PageCtl ctl;
  uint32_t q_shared;                // the CTA's current slot of the page queue (bgx_decode_pages_kernel)
};
static_assert(offsetof(WarpSmem, ring) == 1024 && offsetof(WarpSmem, litq) == 3072, "ring / literal ring placement (see above)");
static_assert(kRing >= 2 * bgx::kNumCmdSymbols && kRing >= 1024 + 1024 && 2 * bgx::kNumLitSymbols <= 1024 && (1 << kLitLutBits) >= 512,
              "table phase: the list of used symbols fits the ring, the code-length-code LUT its upper half or the literal LUT");

// ---------------------------------------------------------------------------------------------
// Explicit shared-state-space accesses. A `saddr_t` is a 32-bit shared-window address on the device (so
// the hot loops issue plain `LDS/STS [R+imm]` instead of re-deriving the generic address of the arena for
// every access) and a host pointer in the emulator.
#ifdef BGX_EMULATED
typedef uintptr_t saddr_t;
BGX_DEV saddr_t saddr(const void* p) { return reinterpret_cast<uintptr_t>(p); }
BGX_DEV saddr_t saddr_pinned(const void* p) { return reinterpret_cast<uintptr_t>(p); }
BGX_DEV uint32_t lds_u8(saddr_t a) { return *reinterpret_cast<const uint8_t*>(a); }
BGX_DEV void sts_u8(saddr_t a, uint32_t v) { *reinterpret_cast<uint8_t*>(a) = (uint8_t)v; }
BGX_DEV uint32_t lds_u16(saddr_t a) { return *reinterpret_cast<const uint16_t*>(a); }
BGX_DEV uint32_t lds_u32(saddr_t a) { return *reinterpret_cast<const uint32_t*>(a); }
BGX_DEV uint2 lds_u32x2(saddr_t a) { return *reinterpret_cast<const uint2*>(a); }
BGX_DEV void sts_u32(saddr_t a, uint32_t v) { *reinterpret_cast<uint32_t*>(a) = v; }
BGX_DEV uint4 lds_u32x4(saddr_t a) { return *reinterpret_cast<const uint4*>(a); }
BGX_DEV void sts_u32x4(saddr_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint32_t* p = reinterpret_cast<uint32_t*>(a); p[0] = x; p[1] = y; p[2] = z; p[3] = w; }
BGX_DEV void sts_u32x2(saddr_t a, uint32_t x, uint32_t y) { uint32_t* p = reinterpret_cast<uint32_t*>(a); p[0] = x; p[1] = y; }
BGX_DEV uint32_t ldg_u8(const uint8_t* p) { return *p; }
BGX_DEV uint32_t ldg_u32(const uint32_t* p) { return *p; }
#else
typedef uint32_t saddr_t;
BGX_DEV saddr_t saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// the same, opaque to the optimiser: it keeps the address in a register instead of re-deriving it (S2R CgaCtaId +
// LEA) next to every access of a hot loop
BGX_DEV saddr_t saddr_pinned(const void* p) { uint32_t a = (uint32_t)__cvta_generic_to_shared(p); asm volatile("" : "+r"(a)); return a; }
BGX_DEV uint32_t lds_u8(saddr_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
BGX_DEV void sts_u8(saddr_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
BGX_DEV uint32_t lds_u16(saddr_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
BGX_DEV uint32_t lds_u32(saddr_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
BGX_DEV uint2 lds_u32x2(saddr_t a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
BGX_DEV void sts_u32(saddr_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
BGX_DEV uint4 lds_u32x4(saddr_t a) { uint4 v; asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; }
BGX_DEV void sts_u32x4(saddr_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) { asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory"); }
BGX_DEV void sts_u32x2(saddr_t a, uint32_t x, uint32_t y) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory"); }
BGX_DEV uint32_t ldg_u8(const uint8_t* p) { uint32_t v; asm volatile("ld.global.u8 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
BGX_DEV uint32_t ldg_u32(const uint32_t* p) { uint32_t v; asm volatile("ld.global.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
#endif

// keeps a loop-invariant value in its register (the optimiser otherwise re-derives cheap ones inside hot loops)
// base + 2 * index as ONE multiply-add (the compiler otherwise doubles, masks and adds)
BGX_DEV saddr_t saddr_scaled2(saddr_t base, uint32_t index) {
#ifdef BGX_EMULATED
  return base + 2u * index;
#else
  uint32_t a;
  asm("mad.lo.u32 %0, %1, 2, %2;" : "=r"(a) : "r"(index), "r"(base));
  return a;
#endif
}
BGX_DEV uint32_t pinned(uint32_t v) {
#ifndef BGX_EMULATED
  asm volatile("" : "+r"(v));
#endif
  return v;
}
// byte p of the output ring / literal index g of the literal ring
#ifdef BGX_EMULATED
BGX_DEV saddr_t ring_at(saddr_t ring_a, uint32_t p) { return ring_a + (p & (kRing - 1u)); }
BGX_DEV saddr_t litq_at(saddr_t litq_a, uint32_t g) { return litq_a + (g & (kLitQ - 1u)); }
#else   // the rings are aligned to their size in the shared window (WarpSmem): OR instead of ADD, one LOP3
BGX_DEV saddr_t ring_at(saddr_t ring_a, uint32_t p) { return ring_a | (p & (kRing - 1u)); }
BGX_DEV saddr_t litq_at(saddr_t litq_a, uint32_t g) { return litq_a | (g & (kLitQ - 1u)); }
#endif
BGX_DEV bool arena_layout_ok(const void* ring, const void* litq) {
#ifdef BGX_EMULATED
  (void)ring; (void)litq;
  return true;
#else
  return (((uint32_t)__cvta_generic_to_shared(ring)) & (kRing - 1u)) == 0u && (((uint32_t)__cvta_generic_to_shared(litq)) & (kLitQ - 1u)) == 0u;
#endif
}

// ---------------------------------------------------------------------------------------------
// Input staging + bit reader. Every lane reads its own sub-stream strictly sequentially, so the
// compressed bytes are streamed through a small per-lane ring in shared memory: 4 slots x 16 B,
// filled with cp.async (LDGSTS: global -> shared without registers) up to three slots AHEAD of
// the read position. A lane therefore never waits on an HBM/L2 round trip in the middle of a
// dependent decode chain; a window refill is one predicated LDS. Chunk indices are clamped to the
// stream buffer, so the deliberate over-read of the format (BrotligDeswizzler.h:74-81) never leaves it.
//
// The ring is topped up at a few explicit places (br_topup: once per round before the commands, once per
// literal pair, once per code-length symbol), not inside every refill. Invariant: after a top-up at word
// index k, chunks up to (k >> 2) + 3 are requested; a request waits for all EARLIER requests first. As
// long as a lane fetches at most 4 words between two top-ups (a command with every extra field is
// <= 102 bits, a literal pair <= 30), each chunk has landed one top-up before its first word is read,
// and a chunk is only overwritten after its last word went into the window.
BGX_DEV void cp_async16(saddr_t smem_dst, const void* gmem_src) {
#ifdef BGX_EMULATED
  memcpy(reinterpret_cast<void*>(smem_dst), gmem_src, 16);
#else
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_dst), "l"(gmem_src) : "memory");
#endif
}
BGX_DEV void cp_async_commit() {
#ifndef BGX_EMULATED
  asm volatile("cp.async.commit_group;\n" ::: "memory");
#endif
}
template <int N>
BGX_DEV void cp_async_wait() {
#ifndef BGX_EMULATED
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
#endif
}

struct BitRd {
  uint32_t w0, w1, nxt;   // 64-bit window {w1:w0} and the prefetched next word
  uint32_t bitpos;        // < 32 between operations
  uint32_t k4;            // 4 x index (from the lane's first 16-byte chunk) of the next word to fetch
};

struct PageIn {
  // header access (a few broadcast loads at the start of a page)
  const uint32_t* base;   // page start (4-byte aligned)
  uint32_t lim;           // largest word index that may be loaded through `base`
  // per-lane sub-stream staging
  const uint4* g16;       // 16-byte aligned address at or below the page start
  uint32_t lim16;         // largest chunk index that may be loaded through `g16`
  uint32_t c0;            // chunk index (relative to g16) of this lane's chunk 0
  saddr_t stage_a;        // this lane's 64-byte staging area (4 slots x 16 B)
  uint32_t swz;           // slot swizzle ((lane >> 1) & 3) << 4: spreads the lanes' slots over the banks
  uint32_t issued;        // chunks requested so far
#ifdef BGX_EMULATED
  uint32_t landed;        // emulator only: chunks known to have arrived (checks the top-up invariant)
#endif
};

BGX_DEV uint32_t ld_word(const PageIn& in, uint32_t idx) { return in.base[idx < in.lim ? idx : in.lim]; }

BGX_DEV void stage_issue(const PageIn& in, uint32_t chunk) {   // chunk: index from the lane's chunk 0
  const uint32_t g = in.c0 + chunk;
  cp_async16(in.stage_a + (((chunk << 4) ^ in.swz) & 0x30u), in.g16 + (g < in.lim16 ? g : in.lim16));
  cp_async_commit();
}
// Requests the chunks the reader may need before the next top-up (see the invariant above).
BGX_DEV void br_topup(const BitRd& r, PageIn& in) {
  while (in.issued < (r.k4 >> 4) + 4u) {
    cp_async_wait<0>();
#ifdef BGX_EMULATED
    in.landed = in.issued;
#endif
    stage_issue(in, in.issued);
    ++in.issued;
  }
}
// Hot-path variant: at most one new chunk can be due when no more than 4 words were fetched since the last top-up.
BGX_DEV void br_topup1(const BitRd& r, PageIn& in) {
#ifdef BGX_TOPUP_WHILE
  br_topup(r, in);
#else
  if (in.issued < (r.k4 >> 4) + 4u) {
    cp_async_wait<0>();
#ifdef BGX_EMULATED
    in.landed = in.issued;
#endif
    stage_issue(in, in.issued);
    ++in.issued;
  }
#ifdef BGX_EMULATED
  if (in.issued < (r.k4 >> 4) + 4u) { fprintf(stderr, "bit reader: more than one chunk due at a hot top-up\n"); abort(); }
#endif
#endif
}
BGX_DEV uint32_t stage_word(const PageIn& in, uint32_t k4) {
#ifdef BGX_EMULATED
  if ((k4 >> 4) >= in.landed) { fprintf(stderr, "bit reader: word %u read before its chunk landed (%u landed)\n", k4 >> 2, in.landed); abort(); }
#endif
  return lds_u32(in.stage_a + ((k4 ^ in.swz) & 0x3cu));
}

BGX_DEV void br_init(BitRd& r, PageIn& in, uint32_t byte_off) {
  // byte_off is relative to the page start; chunks are relative to g16
  const uint32_t rel = (uint32_t)(reinterpret_cast<uintptr_t>(in.base) - reinterpret_cast<uintptr_t>(in.g16)) + byte_off;
  in.c0 = rel >> 4;
  stage_issue(in, 0);
  stage_issue(in, 1);
  stage_issue(in, 2);
  stage_issue(in, 3);
  in.issued = 4;
  cp_async_wait<0>();
#ifdef BGX_EMULATED
  in.landed = 4;
#endif
  const uint32_t k0 = (rel & 15u) >> 2;
  r.w0 = stage_word(in, 4u * k0);
  r.w1 = stage_word(in, 4u * k0 + 4u);
  r.nxt = stage_word(in, 4u * k0 + 8u);
  r.k4 = 4u * k0 + 12u;
  r.bitpos = (rel & 3u) * 8u;
  br_topup(r, in);
}
BGX_DEV uint32_t br_peek(const BitRd& r) { return __funnelshift_r(r.w0, r.w1, r.bitpos); }  // 32 valid bits
BGX_DEV void br_skip(BitRd& r, const PageIn& in, uint32_t n) {   // n <= 32
  r.bitpos += n;
  if (r.bitpos >= 32u) {
    r.w0 = r.w1;
    r.w1 = r.nxt;
    r.nxt = stage_word(in, r.k4);
    r.k4 += 4u;
    r.bitpos -= 32u;
  }
}
BGX_DEV uint32_t umin32(uint32_t a, uint32_t b) { return a < b ? a : b; }
BGX_DEV uint32_t shr32(uint32_t v, uint32_t n) { return n >= 32u ? 0u : (v >> n); }
BGX_DEV uint32_t low_mask(uint32_t n) { return n >= 32u ? 0xffffffffu : ((1u << n) - 1u); }
BGX_DEV uint32_t br_read(BitRd& r, const PageIn& in, uint32_t n) {   // n <= 32
  const uint32_t v = br_peek(r) & low_mask(n);
  br_skip(r, in, n);
  return v;
}

// Rare wide fields (24-bit length extras, distance extras that do not fit the peek), out of line: skips `skip` bits,
// then reads n1 and n2 bits. The caller passes copies of its reader state and takes them back.
struct ColdBits { BitRd r; PageIn in; uint32_t second; };
BGX_COLD uint32_t cold_read_fields(ColdBits* c, uint32_t skip, uint32_t n1, uint32_t n2) {
  br_skip(c->r, c->in, skip);
  const uint32_t first = br_read(c->r, c->in, n1);
  c->second = br_read(c->r, c->in, n2);
  return first;
}
BGX_COLD uint32_t cold_udiv(uint32_t a, uint32_t b) { return a / b; }

BGX_DEV uint32_t warp_index() {
#ifdef BGX_EMULATED
  return (uint32_t)wemu::warp_id();
#else
  return threadIdx.x >> 5;
#endif
}

BGX_DEV uint32_t lane_id() {
#ifdef BGX_EMULATED
  return (uint32_t)wemu::lane();
#else
  return threadIdx.x & 31u;
#endif
}

BGX_DEV uint32_t warp_incl_scan(uint32_t v, uint32_t lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(kFull, v, d);
    if (lane >= (uint32_t)d) v += t;
  }
  return v;
}

BGX_DEV uint64_t warp_incl_scan64(uint64_t v, uint32_t lane) {   // two 32-bit scans in one shuffle chain
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint64_t t = __shfl_up_sync(kFull, v, d);
    if (lane >= (uint32_t)d) v += t;
  }
  return v;
}

// ---------------------------------------------------------------------------------------------
// prefix-code decode: primary LUT entry = symbol | length << 10; kLongCode => canonical search.
// The canonical search behind a kLongCode entry (codes longer than the LUT index).
template <int BITS, typename SortedT>
BGX_DEV uint32_t huff_decode_long(const HuffAux& aux, const SortedT* sorted, uint32_t nsym, uint32_t peek, uint32_t& len) {
  const uint32_t msb = __brev(peek) >> 17;   // first 15 stream bits as an MSB-first number
  uint32_t L = BITS + 1;
#pragma unroll 1
  while (L < 15u && msb >= aux.limit[L]) ++L;
  len = L;
  uint32_t idx = (uint16_t)(aux.base[L] + (msb >> (15u - L)));
  if (idx >= nsym) idx = nsym - 1;
  return sorted[idx];
}
template <int BITS, typename SortedT>
BGX_DEV uint32_t huff_decode(const uint16_t* lut, const HuffAux& aux, const SortedT* sorted, uint32_t nsym,
                             uint32_t peek, uint32_t& len) {
  const uint32_t e = lut[peek & ((1u << BITS) - 1u)];
  if (e != kLongCode) {
    len = e >> 10;
    return e & 0x3ffu;
  }
  return huff_decode_long<BITS, SortedT>(aux, sorted, nsym, peek, len);
}
// The same with the LUT given as a shared-window address held in a register (the literal loop: through the generic
// pointer the compiler re-derives the window base -- S2UR + ULEA -- in every trip).
template <int BITS, typename SortedT>
BGX_DEV uint32_t huff_decode_at(saddr_t lut_a, const HuffAux& aux, const SortedT* sorted, uint32_t nsym,
                                uint32_t peek, uint32_t& len) {
  const uint32_t e = lds_u16(saddr_scaled2(lut_a, peek & ((1u << BITS) - 1u)));
  if (e != kLongCode) {
    len = e >> 10;
    return e & 0x3ffu;
  }
  return huff_decode_long<BITS, SortedT>(aux, sorted, nsym, peek, len);
}

// ---------------------------------------------------------------------------------------------
// Table construction. It runs once per table and page, so it is written for SIZE: one out-of-line routine serves the
// three tables (LUT width, alphabet and the element size of `sorted` are run-time values), which keeps a page's table phase
// inside the instruction cache next to the round loops of the other pages on the SM.
// scratch[] words during the table phase
constexpr uint32_t kScrCntA = 0, kScrNext = 16, kScrClLen = 48, kScrDesc = 72;   // desc: 3 x 8 words
struct TableRef {
  uint16_t* lut;
  HuffAux* aux;
  void* sorted;            // uint16_t[alphabet], or uint8_t[alphabet] when sorted_u8
  uint32_t bits;           // LUT index width
  uint32_t alphabet;
  uint32_t sorted_u8;
};

// cooperative fill of the LUT entries owned by one code: all indices whose low `len` bits equal
// `rev` (the bit-reversed code).
BGX_DEV void lut_fill_coop(uint16_t* lut, uint32_t bits, uint32_t rev, uint32_t len, uint16_t entry, uint32_t lane) {
#pragma unroll 1
  for (uint32_t j = rev + (lane << len); j < (1u << bits); j += (32u << len)) lut[j] = entry;
}

// Builds LUT + canonical arrays of one prefix code from
//   list[0..used)  (symbol | length << 10) of every symbol with a non-zero length, in symbol order (shared memory)
//   cnt[1..15]      the number of symbols per length
// Canonical order = (length, symbol index), as GenerateHuffmanTable (BrotligHuffmanTable.cpp:44-71). Work is
// proportional to the symbols in use and to the LUT size, not to the alphabet (most of the 728 / 544 symbols of a
// page are unused):
//   1. one 64-bit warp scan over the lengths gives first codes, limits, offsets into `sorted` and the Kraft sum;
//   2. `sorted`: 32 list entries at a time, rank within equal lengths by match_any;
//   3. the LUT, filled by POSITION: in most-significant-bit-first order the codes of a canonical prefix code tile the
//      code space in `sorted` order, so the entries of length L are the contiguous positions [lo_L, hi_L) and entry m
//      of that range belongs to sorted[off_L + ((m - lo_L) >> (bits - L))]; lanes take consecutive positions (the
//      LUT index is the bit-reversed position, the stream being read least-significant bit first).
BGX_DEV uint32_t build_table(WarpSmem* sm, const uint16_t* list, const uint32_t* cnt, uint32_t used, const TableRef& t, uint32_t lane) {
  const uint32_t bits = t.bits;
  uint16_t* const lut = t.lut;
  HuffAux& aux = *t.aux;
  uint32_t* next = sm->scratch + kScrNext;  // [16] running position in `sorted` per length
  // ---- 1. lane L (1..15) owns length L. With the counts left-aligned to 15 bits, the first code of a length is the
  //      EXCLUSIVE prefix sum over the shorter lengths (code[L] = (code[L-1] + cnt[L-1]) << 1, left-aligned), its limit
  //      the inclusive one, and the total the Kraft sum.
  const uint32_t sh = 15u - (lane & 15u);
  const uint32_t c = (lane >= 1u && lane <= 15u) ? cnt[lane] : 0u;
  const uint64_t both = warp_incl_scan64(((uint64_t)(c << sh) << 32) | c, lane);
  const uint32_t lim = (uint32_t)(both >> 32), off = (uint32_t)both - c;
  const uint32_t first_la = lim - (c << sh);             // first code of length L, left-aligned to 15 bits
  __syncwarp();
  if (lane <= 15u) {
    aux.base[lane] = (uint16_t)(lane ? off - (first_la >> sh) : 0u);
    aux.limit[lane] = (uint16_t)(lane ? (lim > 0x8000u ? 0x8000u : lim) : 0u);
    next[lane] = off;
  }
  // A prefix code must be complete (Kraft sum exactly 1; RFC 7932 section 3.2), or consist of a single symbol. The
  // reference trusts the lengths (BrotligHuffmanTable.cpp:44-71,135-145): an over-subscribed set overwrites table
  // entries, an incomplete one leaves entries of the previous page in place. Both are rejected here.
  const uint32_t kraft = __shfl_sync(kFull, lim, 15);
  __syncwarp();
  if (kraft != 0x8000u && used != 1u) return kPageErrTable;
  // ---- 2. sorted
#pragma unroll 1
  for (uint32_t i0 = 0; i0 < used; i0 += 32) {
    const uint32_t i = i0 + lane;
    const uint32_t e = i < used ? list[i] : 0u;
    const uint32_t L = e >> 10;
    const uint32_t m = __match_any_sync(kFull, L);
    const uint32_t rank = __popc(m & ((1u << lane) - 1u));
    if (L) {
      const uint32_t idx = next[L] + rank;               // (< used <= alphabet: the counts are those of the list)
      if (t.sorted_u8) static_cast<uint8_t*>(t.sorted)[idx] = (uint8_t)e;
      else static_cast<uint16_t*>(t.sorted)[idx] = (uint16_t)(e & 0x3ffu);
    }
    __syncwarp();
    if (L && rank == 0) next[L] += __popc(m);
    __syncwarp();
  }
  // ---- 3. the LUT by position
  const uint32_t down = 15u - bits;
  const uint32_t lo_mine = first_la >> down, hi_mine = (lim > 0x8000u ? 0x8000u : lim) >> down;
#pragma unroll 1
  for (uint32_t L = 1; L <= bits; ++L) {
    const uint32_t lo = __shfl_sync(kFull, lo_mine, (int)L), hi = __shfl_sync(kFull, hi_mine, (int)L);
    const uint32_t off_l = __shfl_sync(kFull, off, (int)L);
    const uint32_t tag = L << 10, sh2 = bits - L;
#pragma unroll 1
    for (uint32_t m = lo + lane; m < hi; m += 32) {
      const uint32_t idx = off_l + ((m - lo) >> sh2);
      const uint32_t sym = t.sorted_u8 ? (uint32_t)static_cast<const uint8_t*>(t.sorted)[idx] : (uint32_t)static_cast<const uint16_t*>(t.sorted)[idx];
      lut[__brev(m) >> (32u - bits)] = (uint16_t)(sym | tag);
    }
  }
  // positions past the last code of <= bits bits start longer codes (a lone symbol: positions no valid stream reaches)
#pragma unroll 1
  for (uint32_t m = __shfl_sync(kFull, hi_mine, (int)bits) + lane; m < (1u << bits); m += 32) lut[__brev(m) >> (32u - bits)] = kLongCode;
  __syncwarp();
  return 0u;
}

// A prefix code travels from its description in the bit stream (read_table: everything that consumes bits) to LUT +
// canonical arrays (build_from_desc) as a TableDesc and, for a complex code, the list of its used symbols + the counts
// per length.
struct TableDesc {
  uint32_t type;           // 0 trivial, 1 simple, 2 complex
  uint32_t n;              // simple: symbols (2..4); complex: entries of the list
  uint32_t shape;          // simple: 0..3 = length shapes {1,1} {1,2,2} {2,2,2,2} {1,2,3,3}
  uint32_t sym[4];         // trivial / simple: the symbols in stored order
  uint32_t pad;
};
static_assert(sizeof(TableDesc) == 32 && kScrDesc + 24 <= 224, "table descriptors fit the scratch words");

// Reads one prefix-code description (trivial / simple / complex) into `d` (+ list / cnt).
// Returns 0 or kPageErrTable. Cursor conventions: every table starts at sub-stream 0 (lane 0).
BGX_DEV uint32_t read_table(WarpSmem* sm, BitRd& rd, PageIn& in, uint32_t alphabet, TableDesc* d, uint16_t* cl_lut, uint16_t* list,
                            uint32_t* cnt, uint32_t lane) {
  const uint32_t max_bits = bgx::bit_length(alphabet - 1);
  br_topup(rd, in);
  uint32_t hdr = 0;
  if (lane == 0) hdr = br_read(rd, in, 6);
  hdr = __shfl_sync(kFull, hdr, 0);
  const uint32_t type = hdr & 3u;
  if (type < 2u) {
    // trivial: one symbol, zero-length code (BrotligHuffmanTable.cpp:87-101);
    // simple: 2..4 symbols, k-th symbol in sub-stream k, table filled in STORED order (:102-125)
    const uint32_t nsym = type ? ((hdr >> 2) & 3u) + 1u : 1u;
    const uint32_t tree_select = (hdr >> 4) & 1u;
    if (type && nsym < 2) return kPageErrTable;
    uint32_t sym = 0;
    if (lane < nsym) sym = br_read(rd, in, max_bits);
    if (lane < 4) d->sym[lane] = sym;
    if (lane == 0) {
      d->type = type;
      d->n = nsym;
      d->shape = nsym < 4 ? nsym - 2 : (tree_select ? 3u : 2u);
    }
    return 0;
  }
  if (type != 2) return kPageErrTable;

  // ---- complex (:126-200). 1) code-length code: i-th 5-bit length in sub-stream i, storage order
  //      1,2,3,4,0,5,17,6,16,7,8,9,10,11,12,13,14,15 (lane i reads the length of symbol my_sym)
  const uint32_t ncl = ((hdr >> 2) & 15u) + 4u;
  const uint32_t my_sym = lane < 4u ? lane + 1u : lane == 4u ? 0u : lane == 5u ? 5u : lane == 6u ? 17u : lane == 7u ? 6u :
                          lane == 8u ? 16u : lane - 2u;
  uint32_t* cl_len_by_sym = sm->scratch + kScrClLen;   // [18]
  uint32_t myread = 0;
  if (lane < 18) cl_len_by_sym[lane] = 0;   // the reference leaves these uninitialised when ncl < 18
  if (lane < 16) cnt[lane] = 0;
  __syncwarp();
  if (lane < ncl && lane < 18) {
    myread = br_read(rd, in, 5);
    cl_len_by_sym[my_sym] = myread;
  }
  const uint32_t bad = __ballot_sync(kFull, myread > 9u);
  if (bad) return kPageErrTable;   // reference: out-of-bounds table index
  // the code-length code itself must be a complete prefix code or a single symbol (see build_table)
  const uint32_t cl_used = __ballot_sync(kFull, myread != 0u);
  const uint32_t cl_kraft = __reduce_add_sync(kFull, myread ? (512u >> myread) : 0u);
  if (cl_kraft != 512u && __popc(cl_used) != 1) return kPageErrTable;
  __syncwarp();
  // canonical codes over symbols 0..ncl-1 (GenerateHuffmanTable is called with size = ncl, :145),
  // but the per-length counts come from every length that was read (:135-142)
  const uint32_t ls = (lane < ncl && lane < 18) ? cl_len_by_sym[lane] : 0u;
  const uint32_t msame = __match_any_sync(kFull, ls);
  const uint32_t rank = __popc(msame & ((1u << lane) - 1u));
  uint32_t code = 0, mycode = 0, prevcnt = 0;
#pragma unroll 1
  for (uint32_t L = 1; L <= 9; ++L) {
    code = (code + prevcnt) << 1;
    prevcnt = __popc(__ballot_sync(kFull, myread == L));
    if (ls == L) mycode = code + rank;
  }
  // LUT: zero (sym 0, len 0) wherever no code lands; codes owning >= 32 entries are filled by the whole warp, one code
  // at a time, the others by their own lane (<= 16 entries each, all of them at once)
  reinterpret_cast<uint4*>(cl_lut)[lane] = make_uint4(0u, 0u, 0u, 0u);
  reinterpret_cast<uint4*>(cl_lut)[32u + lane] = make_uint4(0u, 0u, 0u, 0u);
  __syncwarp();
  const uint32_t rev = ls ? (__brev(mycode) >> (32u - ls)) & 511u : 0u;
  const uint16_t entry = (uint16_t)(lane | (ls << 8));
  uint32_t wide = __ballot_sync(kFull, ls != 0u && ls <= 4u);
  while (wide) {
    const int k = __ffs((int)wide) - 1;
    wide &= wide - 1;
    const uint32_t Lk = __shfl_sync(kFull, ls, k);
    lut_fill_coop(cl_lut, 9u, __shfl_sync(kFull, rev, k), Lk, (uint16_t)((uint32_t)k | (Lk << 8)), lane);
  }
  if (ls > 4u) {
#pragma unroll 1
    for (uint32_t j = rev; j < 512u; j += 1u << ls) cl_lut[j] = entry;
  }
  __syncwarp();

  // ---- 2) the code lengths themselves: k-th code-length symbol lives in sub-stream k mod 32. They are not stored as
  //      an array of 728 lengths: every symbol with a non-zero length goes to a compact list (symbol | length << 10,
  //      symbol order), and the lengths are counted on the way.
  uint32_t filled = 0, used = 0, prev_carry = bgx::kInitialRepeatLen;
#pragma unroll 1
  while (filled < alphabet) {
    br_topup(rd, in);
    const uint32_t pk = br_peek(rd);
    const uint32_t e = cl_lut[pk & 511u];
    const uint32_t s = e & 0xffu, l = e >> 8;
    uint32_t run = 1, nb = l;
    if (s == (uint32_t)bgx::kRepeatPrev) { run = 3 + ((pk >> l) & 3u); nb = l + 2; }
    else if (s == (uint32_t)bgx::kRepeatZero) { run = 3 + ((pk >> l) & 7u); nb = l + 3; }
    // "repeat previous" repeats the last explicit length before it (the lanes past the end of the alphabet lie above
    // every lane that counts, so they cannot be that one)
    const uint32_t expl = __ballot_sync(kFull, s < 16u);
    const uint32_t below = expl & ((1u << lane) - 1u);
    const uint32_t pv = __shfl_sync(kFull, s, below ? (31 - __clz((int)below)) : 0);
    const uint32_t prevval = below ? pv : prev_carry;
    const uint32_t val = s == (uint32_t)bgx::kRepeatPrev ? prevval : (s == (uint32_t)bgx::kRepeatZero ? 0u : s);
    // positions: symbols covered | list entries produced << 16, one scan
    const uint32_t incl = warp_incl_scan(run | (val ? run << 16 : 0u), lane);
    const uint32_t start = filled + (incl & 0xffffu) - run;
    const bool active = start < alphabet;   // otherwise this symbol does not exist: consume nothing
    const uint32_t run_c = active ? umin32(run, alphabet - start) : 0u;
    const uint32_t nz = val ? run_c : 0u;
    if (nz) {
      const uint32_t pos = used + (incl >> 16) - run;
      const uint32_t tag = val << 10;
#pragma unroll 1
      for (uint32_t k = 0; k < nz; ++k) list[pos + k] = (uint16_t)((start + k) | tag);
      atomicAdd(&cnt[val], nz);
    }
    if (active) br_skip(rd, in, nb);
    const uint32_t expl_act = expl & __ballot_sync(kFull, active);
    if (expl_act) prev_carry = __shfl_sync(kFull, s, 31 - __clz((int)expl_act));
    else (void)__shfl_sync(kFull, s, 0);
    filled += __shfl_sync(kFull, incl, 31) & 0xffffu;
    used += __reduce_add_sync(kFull, nz);
  }
  if (lane == 0) {
    d->type = 2u;
    d->n = used;
  }
  return 0u;
}

// LUT + canonical arrays of one code from its description.
BGX_DEV uint32_t build_from_desc(WarpSmem* sm, const TableDesc* d, const TableRef& t, const uint16_t* list, const uint32_t* cnt, uint32_t lane) {
  const uint32_t type = d->type, n = d->n, bits = t.bits;
  uint16_t* const lut = t.lut;
  if (type == 2u) return build_table(sm, list, cnt, n, t, lane);
  if (type == 0u) {
    const uint16_t sym = (uint16_t)d->sym[0];
#pragma unroll 1
    for (uint32_t j = lane; j < (1u << bits); j += 32) lut[j] = sym;   // length field 0
  } else if (type == 1u) {
    // codes are consecutive in stored order
    const uint32_t shape = d->shape;
    const uint32_t lens4 = shape == 0 ? 0x0011u : shape == 1 ? 0x0221u : shape == 2 ? 0x2222u : 0x3321u;
    const uint32_t codes4 = shape == 0 ? 0x0010u : shape == 1 ? 0x0320u : shape == 2 ? 0x3210u : 0x7620u;
#pragma unroll 1
    for (uint32_t k = 0; k < n && k < 4u; ++k) {
      const uint32_t Lk = (lens4 >> (4 * k)) & 15u;
      const uint32_t ck = (codes4 >> (4 * k)) & 15u;
      lut_fill_coop(lut, bits, __brev(ck) >> (32 - Lk), Lk, (uint16_t)((d->sym[k] & 0x3ffu) | (Lk << 10)), lane);
    }
  }
  return 0u;
}

// The three prefix codes of a page -- insert&copy (728 symbols), distance (544), literal (256), in stream order
// (PageDecoder.cpp:126-147) -- read and built by ONE out-of-line routine working on copies of the reader state.
// Where a description lives between read_table and build_from_desc:
//   code 0: list in the output ring [0, 1456);   code-length-code LUT (512 x u16) in litq + lut_lit[0, 256)
//   code 1: list in lut_lit[256, 1024) (1536 B >= 1088);   the same code-length-code LUT
//   code 2: list in the output ring [0, 512);   code-length-code LUT in the output ring [1024, 2048)
// (the literal LUT is built last, from the ring, when everything that borrowed its room is dead).
// Measured and dropped: the consumer warp building table k while the producer reads description k + 1 (two CTA
// barriers per table): +-0 on 4 KiB pages, whose table phase is bound by instruction FETCH -- the code of a phase
// that runs once per page is cold every time -- not by the instructions it executes; 16 KiB pages lost 3.6 %.
struct TableIo { BitRd rd; PageIn in; };
BGX_DEV TableDesc* table_desc(WarpSmem* sm, uint32_t k) { return reinterpret_cast<TableDesc*>(sm->scratch + kScrDesc) + k; }
constexpr uint32_t kClLutSpill = kLitQ >= 1024u ? 0u : (1024u - kLitQ) / 2u;   // u16 entries of lut_lit under the code-length-code LUT
BGX_DEV uint16_t* table_list(WarpSmem* sm, uint32_t k) { return k == 1u ? sm->lut_lit + kClLutSpill : reinterpret_cast<uint16_t*>(sm->ring); }
static_assert(offsetof(WarpSmem, lut_lit) == offsetof(WarpSmem, litq) + kLitQ && kLitQ >= 512 && (1 << kLitLutBits) >= kClLutSpill + bgx::kNumDistSymbols,
              "table phase: litq (+ the head of lut_lit) hold a code-length-code LUT, the rest of lut_lit the distance list");

BGX_COLD uint32_t load_tables(WarpSmem* sm, TableIo* io, uint32_t lane) {
  BitRd rd = io->rd;
  PageIn in = io->in;
  uint32_t terr = 0;
#pragma unroll 1
  for (uint32_t k = 0; k < 3u && !terr; ++k) {
    TableDesc* d = table_desc(sm, k);
    uint32_t* cnt = sm->scratch + kScrCntA;
    TableRef t;
    t.lut = k == 0 ? sm->lut_cmd : k == 1 ? sm->lut_dist : sm->lut_lit;
    t.aux = &sm->aux[k];
    t.sorted = k == 0 ? static_cast<void*>(sm->sorted_cmd) : k == 1 ? static_cast<void*>(sm->sorted_dist) : static_cast<void*>(sm->sorted_lit);
    t.bits = k == 0 ? (uint32_t)kCmdLutBits : k == 1 ? (uint32_t)kDistLutBits : (uint32_t)kLitLutBits;
    t.alphabet = k == 0 ? (uint32_t)bgx::kNumCmdSymbols : k == 1 ? (uint32_t)bgx::kNumDistSymbols : (uint32_t)bgx::kNumLitSymbols;
    t.sorted_u8 = k == 2 ? 1u : 0u;
    uint16_t* cl_lut = k < 2u ? reinterpret_cast<uint16_t*>(sm->litq) : reinterpret_cast<uint16_t*>(sm->ring + 1024);
    terr = read_table(sm, rd, in, t.alphabet, d, cl_lut, table_list(sm, k), cnt, lane);
    __syncwarp();
    if (!terr) terr = build_from_desc(sm, d, t, table_list(sm, k), cnt, lane);
    __syncwarp();
  }
  io->rd = rd;
  io->in = in;
  return terr;
}

// ---------------------------------------------------------------------------------------------
struct PageJob {
  const uint8_t* in;        // compressed page (4-byte aligned)
  uint32_t in_size;
  uint32_t in_limit;        // bytes readable from `in` (to the end of the stream buffer incl. slack), >= in_size
  uint8_t* out;             // where this page's bytes go
  uint32_t out_size;
  uint32_t allow_delta;     // stream is preconditioned (page delta flag is honoured, PageDecoder.cpp:87)
};

struct PageResult {
  uint32_t status;          // kPage* bits
  uint32_t is_delta;        // page header flag (&& allow_delta)
};

// read one output byte of the current page at position p (< current write position)
BGX_DEV uint8_t out_byte(const WarpSmem* sm, const uint8_t* out, int32_t ring_lo, uint32_t p) {
  return ((int32_t)p >= ring_lo) ? sm->ring[p & (kRing - 1)] : out[p];
}

// write ring bytes [from, to) to global memory; byte granular, coalesced
BGX_DEV void flush_bytes(const WarpSmem* sm, uint8_t* out, uint32_t from, uint32_t to, uint32_t lane) {
  for (uint32_t p = from + lane; p < to; p += 32) out[p] = sm->ring[p & (kRing - 1)];
}

#ifndef BGX_LIT_TOPUP_MASK
#define BGX_LIT_TOPUP_MASK 6   // staging top-up when (literal index & mask) == 0: 6 = every 4th pair, 0 = every pair
#endif
// Decodes `cnt` literals in this lane (lane-dependent count allowed) into the literal ring at
// page-global literal indices tail + j*32 + lane.
BGX_DEV void decode_literals(WarpSmem* sm, BitRd& rd, PageIn& in, uint32_t tail, uint32_t cnt, uint32_t lane) {
  const saddr_t litq_a = saddr_pinned(sm->litq), lut_a = saddr_pinned(sm->lut_lit);
  uint32_t q = (tail + lane) & (kLitQ - 1);
  // two literals per 32-bit peek (2 x 15 bits at most): one window refill check per pair. An odd count ends
  // with a pair whose second half is neither stored nor consumed.
#pragma unroll 1
  for (uint32_t j = 0; j < cnt; j += 2) {
    // staging top-up every 4th pair: 4 pairs consume at most 120 bits = 4 words, the most one top-up covers. The lanes
    // cross their 16-byte chunks at different pairs, so a top-up per pair runs its body (wait + address + cp.async)
    // nearly every time for three or four lanes; every 4th pair it runs a quarter as often for half the warp.
    if ((j & (uint32_t)BGX_LIT_TOPUP_MASK) == 0u) br_topup1(rd, in);
    const uint32_t pk = br_peek(rd);
    uint32_t len1, len2;
    const uint32_t s1 = huff_decode_at<kLitLutBits>(lut_a, sm->aux[2], sm->sorted_lit, bgx::kNumLitSymbols, pk, len1);
    const uint32_t s2 = huff_decode_at<kLitLutBits>(lut_a, sm->aux[2], sm->sorted_lit, bgx::kNumLitSymbols, pk >> len1, len2);
    const bool two = j + 1u < cnt;
    sts_u8(litq_a + q, s1);
    if (two) sts_u8(litq_a + ((q + 32u) & (kLitQ - 1)), s2);
    q = (q + 64u) & (kLitQ - 1);
    br_skip(rd, in, two ? len1 + len2 : len1);
  }
}

// The page decoder: a CTA of TWO warps decodes one page as a two-stage pipeline over rounds.
//   warp 0, the PRODUCER: owns the 32 bit readers; reads the tables, decodes each round's commands
//           and literals (everything that consumes bits), resolves the distance ring and publishes the
//           round in a RoundBuf;
//   warp 1, the CONSUMER: places literals and match copies in the output ring and streams the page to HBM.
// The two roles are separate loops (each warp only keeps its own state in registers). Rounds travel
// through a ring of kQ RoundBufs guarded by mbarriers in shared memory, so that neither warp executes a CTA-wide
// barrier inside a page and the producer may run up to kQ rounds ahead:
//   full[q]   the producer ARRIVES after publishing round r (q = r mod kQ); the consumer WAITs before reading it
//   empty[q]  the consumer ARRIVES once it no longer needs the round's RoundBuf and literals; the producer
//             WAITs on it before it reuses the slot (and earlier, when the literal ring is short of room)
// A round with long runs (more than kRoundMax bytes, or more literals than the literal ring holds) travels as a
// sequence of "virtual" rounds that each fit (see the producer), so the consumer only ever sees one kind of round.
// The mbarriers are re-initialised at the start of every page (the CTA is persistent and decodes many pages).

// full[] / empty[] are mbarriers in shared memory (no per-SM resource besides 8 bytes each;
// named bar.sync barriers would cap the resident CTAs per SM): one elected lane arrives (release) after a
// __syncwarp, every lane of the waiting warp observes the phase (acquire). Round j uses phase j / kQ of slot
// j mod kQ, so the parity to wait for is (j / kQ) & 1.
BGX_DEV void mbar_init(saddr_t a, uint32_t count) {
#ifdef BGX_EMULATED
  wemu::mbar_init(reinterpret_cast<uint64_t*>(a), count);
#else
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
#endif
}
BGX_DEV void mbar_inval(saddr_t a) {
#ifndef BGX_EMULATED
  asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(a) : "memory");
#endif
}
BGX_DEV void mbar_init_fence() {
#ifndef BGX_EMULATED
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
}
BGX_DEV void mbar_arrive(saddr_t a) {
#ifdef BGX_EMULATED
  wemu::mbar_arrive(reinterpret_cast<uint64_t*>(a));
#else
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(a) : "memory");
#endif
}
// Waits for the phase with this parity to complete. Returns false when it did not within kMaxPolls polls (about a
// second): a hand-over that a valid stream cannot produce. Both warps of the page then leave with kPageErrHang, so a
// corrupt stream can never hang the persistent kernel (the emulator's dead-lock detection plays this role on the CPU).
#ifndef BGX_WAIT_HINT_NS
#define BGX_WAIT_HINT_NS 0x989680
#endif
#ifndef BGX_WAIT_FAST_POLLS
#define BGX_WAIT_FAST_POLLS 1024
#endif
#ifndef BGX_WAIT_SLEEP_NS
#define BGX_WAIT_SLEEP_NS 128
#endif
constexpr uint32_t kFastPolls = BGX_WAIT_FAST_POLLS;   // polls before the waiting warp starts to sleep between polls
constexpr uint32_t kMaxPolls = 1u << 23;               // sleeping polls before a wait is declared hung (about a second)
BGX_DEV bool mbar_wait(saddr_t a, uint32_t parity) {
#ifdef BGX_EMULATED
  wemu::mbar_wait(reinterpret_cast<uint64_t*>(a), parity);
  return true;
#else
  // try_wait comes back after a short time whatever limit it is given. Hand-overs between rounds take a few polls and must be seen at once (sleeping from the 16th poll on costs
  // structured binary 2.5 % and 16 KiB pages 5 %); only a warp that has polled kFastPolls times (~10 us: the consumer
  // during a table phase, or a page that hangs) sleeps between polls. The poll count doubles as the hang guard
  // (kMaxPolls sleeping polls are far beyond any wait of a valid stream).
  // (A try_wait with a time limit is three instructions -- SYNCS.PHASECHK.TRYWAIT, NANOSLEEP.SYNCS, SYNCS.PHASECHK -- and
  // the sleep ends with ANY barrier event on the SM, so a waiting warp polls every ~90 cycles while 16 pages hand rounds
  // over. A tighter poll loop written in PTX, four polls per trip, halved the instructions per poll and lost 1-4 %.)
  uint32_t done, polls = 0;
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(a), "r"(parity), "r"((uint32_t)BGX_WAIT_HINT_NS) : "memory");
    if (done) return true;
    if (++polls > kFastPolls) {
      if (polls > kMaxPolls) return false;
      __nanosleep(BGX_WAIT_SLEEP_NS);
    }
  }
#endif
}
// the whole warp signals: its earlier shared-memory accesses are ordered before the elected lane's arrive
BGX_DEV void warp_arrive(saddr_t a, uint32_t lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(a);
}
BGX_DEV uint32_t ld_volatile_u32(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }

// ------------------------------------------------------------------------------------- PRODUCER
// Hand-over state of the producer. In the round loop it lives in registers (every member function is inlined); the
// rare long-run path (slow_round, out of line so that it stays out of the instruction cache) works on a copy.
struct ProdCtx {
  WarpSmem* sm;
  saddr_t full_a, empty_a;
  uint32_t lane;
  uint32_t rnd;            // rounds published so far
  uint32_t synced;         // empty[] phases taken so far: rounds < synced are known to be consumed
  uint32_t lit_tail;       // literals decoded so far
  uint32_t lit_head_p;     // literals that the rounds published so far consume

  // waits until slot rnd % kQ is free; false = the consumer stopped taking rounds (kPageErrHang is set)
  BGX_DEV bool acquire_slot() {
#ifdef BGX_HANDOVER_SYNCTHREADS
    return true;   // (lock-step verification build: the CTA barrier in publish() is the hand-over)
#endif
    while (synced + kQ <= rnd) {   // the slot still holds round rnd - kQ: wait until the consumer is done with it
      if (!mbar_wait(empty_a + 8u * (synced & (kQ - 1u)), (synced / kQ) & 1u)) { hang(); return false; }
      ++synced;
    }
    return true;
  }
  BGX_DEV void hang() { if (lane == 0) sm->ctl.err = kPageErrHang; }
  // the literal ring must hold `newlits` more literals next to those of every round the consumer may still be
  // working on (rounds >= synced): takes more empty[] phases while that helps; false if they cannot fit at all
  BGX_DEV bool wait_lit_room(uint32_t newlits) {
#ifdef BGX_HANDOVER_SYNCTHREADS
    // lock step: while round rnd is produced the consumer works on round rnd - 1; everything older is consumed
    return (lit_tail - (rnd ? sm->ctl.phead[(rnd - 1u) & (kQ - 1u)] : lit_head_p)) + newlits <= kLitQ;
#endif
    uint32_t head_known = synced == rnd ? lit_head_p : sm->ctl.phead[synced & (kQ - 1u)];
    bool fits = (lit_tail - head_known) + newlits <= kLitQ;
    while (!fits && synced < rnd) {
      if (!mbar_wait(empty_a + 8u * (synced & (kQ - 1u)), (synced / kQ) & 1u)) { hang(); return false; }
      ++synced;
      head_known = synced == rnd ? lit_head_p : sm->ctl.phead[synced & (kQ - 1u)];
      fits = (lit_tail - head_known) + newlits <= kLitQ;
    }
    return fits;
  }
  // publishes round `rnd`: final distance; inclusive sums (output | literals << 16) | flags
  BGX_DEV void publish(uint32_t dxv, uint32_t pkv) {
    const uint32_t q = rnd & (kQ - 1u);
    if (lane == 0) sm->ctl.phead[q] = lit_head_p;   // literal head at the start of this round
    sm->rb[q].cmd[lane] = make_uint2(dxv, pkv);
    hand_over(q);
    ++rnd;
  }
  BGX_DEV void hand_over(uint32_t q) {
#ifdef BGX_HANDOVER_SYNCTHREADS
    (void)q;
    __syncthreads();
#else
    warp_arrive(full_a + 8u * q, lane);
#endif
  }
  // ends the page with an error: the slot of round `rnd` (free: every caller has acquired it) carries the abort flag
  BGX_DEV void publish_abort(uint32_t err) {
    const uint32_t q = rnd & (kQ - 1u);
    if (err && lane == 0) sm->ctl.err = err;
    sm->rb[q].cmd[lane] = make_uint2(0u, kPkAbort);
    hand_over(q);
  }
};

struct SlowRound {           // what slow_round needs of the producer's registers (copied in and out around the call)
  ProdCtx pc;
  BitRd rd;
  PageIn in;
  uint32_t ins, cpy, mine, dx, pdone;
};

// A round with long runs (more output than the ring path takes at once, or more literals than the literal ring
// holds) is handed over as a sequence of VIRTUAL rounds that each fit: a virtual round takes as many whole commands
// as fit, in order, or -- when the next command alone is too big -- a piece of it (first <= kSplitLits of its literals
// at a time, then <= kRoundMax bytes of its copy at a time; a copy split in two is the same byte-serial copy,
// PageDecoder.cpp:222-232). The consumer sees ordinary rounds in which the other lanes carry empty commands; the
// round's literals are decoded row by row (32 at a time, all lanes) as the virtual rounds need them.
// Returns false when the page was aborted (the abort round is then published).
BGX_COLD bool slow_round(SlowRound* a) {
  ProdCtx pc = a->pc;          // (local copies: the argument block lives in local memory)
  BitRd rd = a->rd;
  PageIn in = a->in;
  const uint32_t dx = a->dx;
  const bool pdone = a->pdone != 0;
  struct WriteBack { SlowRound* a; ProdCtx& pc; BitRd& rd; PageIn& in; BGX_DEV ~WriteBack() { a->pc = pc; a->rd = rd; a->in = in; } } wb{a, pc, rd, in};
  const uint32_t lane = pc.lane;
  uint32_t rem_ins = a->ins, rem_cpy = a->cpy;   // what is left of this lane's command
  uint32_t s_mine = a->mine;                     // literals this lane still has to decode in this round
  bool first = true;
  for (;;) {
    const uint32_t work = __ballot_sync(kFull, (rem_ins | rem_cpy) != 0u);
    if (!first && !pc.acquire_slot()) return false;
    first = false;
    const uint32_t a0 = work ? (uint32_t)(__ffs((int)work) - 1) : 32u;   // first command with something left
    uint32_t v_ins = rem_ins, v_cpy = rem_cpy;
    if (lane == a0) {                        // the leading command may have to be clipped
      if (v_ins > kSplitLits) { v_ins = kSplitLits; v_cpy = 0; }
      else if (v_cpy > kRoundMax - v_ins) v_cpy = kRoundMax - v_ins;
    }
    const bool clipped = __ballot_sync(kFull, lane == a0 && (v_ins != rem_ins || v_cpy != rem_cpy)) != 0u;
    uint64_t vtot, vpk;
    if (clipped) {
      // a piece of the leading command is the whole virtual round (the usual case inside a long literal run: every
      // virtual round but the last of a command): its sums are the piece itself from lane a0 on -- no scan
      vtot = __shfl_sync(kFull, ((uint64_t)(v_ins + v_cpy) << 32) | v_ins, (int)a0);
      vpk = lane < a0 ? 0ull : vtot;
      if (lane != a0) { v_ins = 0; v_cpy = 0; }
    } else {
      const uint64_t vincl = warp_incl_scan64(((uint64_t)(v_ins + v_cpy) << 32) | v_ins, lane);
      const bool ok = (uint32_t)(vincl >> 32) <= kRoundMax && (uint32_t)vincl <= kSplitLits;
      const uint32_t notok = ~__ballot_sync(kFull, ok) & ~(0xffffffffu >> (31u - (a0 & 31u)));   // lanes > a0 that do not fit
      const uint32_t b0 = a0 >= 32u ? 32u : (notok ? (uint32_t)(__ffs((int)notok) - 1) : 32u);
      if (lane >= b0) { v_ins = 0; v_cpy = 0; }
      vtot = __shfl_sync(kFull, vincl, (int)((b0 ? b0 : 1u) - 1u));   // sums over the taken commands
      vpk = lane < b0 ? vincl : vtot;
    }
    const uint32_t vr_ins = (uint32_t)vtot;
    rem_ins -= v_ins;
    rem_cpy -= v_cpy;
    const bool last_virtual = __ballot_sync(kFull, (rem_ins | rem_cpy) != 0u) == 0u;
    // literals: whole rows until the virtual round is covered; everything the round still carries with the last one
    const uint32_t have = pc.lit_tail - pc.lit_head_p;
    const uint32_t want = vr_ins > have ? vr_ins - have : 0u;
    const uint32_t rows = (want + 31u) >> 5;
    const uint32_t c = last_virtual ? s_mine : (s_mine < rows ? s_mine : rows);
    const uint32_t total = __reduce_add_sync(kFull, c);
    if (have + total < vr_ins) {   // the stream does not carry the literals it inserts
      pc.publish_abort(kPageErrLiterals);
      return false;
    }
    // (Inside a long literal run the literal ring is the bottleneck: a virtual round's literals only fit once the consumer
    //  has placed the previous round's. Decoding the rows that fit before that wait and the rest after it was measured
    //  and lost: textures 172 -> 165 GB/s, binary -3 % -- a second inlined literal loop costs more than the overlap buys.)
    if (!pc.wait_lit_room(total)) {
      pc.publish_abort(kPageErrLiterals);
      return false;
    }
    decode_literals(pc.sm, rd, in, pc.lit_tail, c, lane);
    s_mine -= c;
    pc.publish(dx, (uint32_t)(vpk >> 32) | ((uint32_t)vpk << 16) | ((last_virtual && pdone) ? kPkLast : 0u));
    pc.lit_tail += total;
    pc.lit_head_p += vr_ins;
    if (last_virtual) return true;
  }
}

BGX_DEV void producer_warp(const PageJob& job, WarpSmem* sm) {
  const uint32_t lane = pinned(lane_id());
  const uint32_t lt_mask = pinned((1u << lane) - 1u);
  PageCtl* ctl = &sm->ctl;
  const uint32_t out_size = job.out_size;

  PageIn in;
  BitRd rd;
  uint32_t npostfix = 0, ndirect = 0;
  if (lane == 0) {
    ctl->is_delta = 0;
  }
  in.base = reinterpret_cast<const uint32_t*>(job.in);
  in.lim = (job.in_limit >> 2) ? (job.in_limit >> 2) - 1 : 0;
  {
    const uintptr_t a = reinterpret_cast<uintptr_t>(job.in);
    in.g16 = reinterpret_cast<const uint4*>(a & ~(uintptr_t)15);
    const uint32_t span = (uint32_t)(a & 15u) + job.in_limit;       // bytes readable from g16
    in.lim16 = (span >> 4) ? (span >> 4) - 1 : 0;
    in.c0 = 0;
    in.stage_a = saddr(&sm->stage[lane][0]);
    in.swz = ((lane >> 1) & 3u) << 4;
    in.issued = 0;
  }
  // ---- length-code tables (RFC 7932 section 5)
  if (lane < 24) {
    sm->lenlut[lane] = bgx::insert_base(lane) | (bgx::insert_extra_bits(lane) << 16);
    sm->lenlut[24 + lane] = bgx::copy_base(lane) | (bgx::copy_extra_bits(lane) << 16);
  }
  // ---- page header + sub-stream size table (PageDecoder.cpp:79-121): every lane parses the header
  //      words it needs itself (they are the first few words of the page: broadcast loads)
  uint32_t sub_off;
  {
    auto hdr_bits = [&](uint32_t hpos, uint32_t n) -> uint32_t {   // n <= 25
      const uint32_t w = hpos >> 5, b = hpos & 31u;
      const uint32_t lo = ld_word(in, w), hi = ld_word(in, w + 1);
      return __funnelshift_r(lo, hi, b) & low_mask(n);
    };
    npostfix = hdr_bits(0, 2);
    ndirect = hdr_bits(2, 4) << npostfix;
    const uint32_t is_delta = (hdr_bits(6, 1) && job.allow_delta) ? 1u : 0u;
    if (lane == 0) ctl->is_delta = is_delta;
    const uint32_t base_bits = bgx::floor_log2((job.in_size + 31u) / 32u) + 1u;
    const uint32_t dbits_bits = bgx::floor_log2(bgx::floor_log2(job.in_size - 1u) + 1u) + 1u;
    const uint32_t base_size = hdr_bits(8, base_bits);
    const uint32_t delta_bits = hdr_bits(8 + base_bits, dbits_bits);
    const uint32_t tbl = 8 + base_bits + dbits_bits;
    const uint32_t hdr_bytes = ((tbl + 32u * delta_bits + 31u) / 32u) * 4u;
    const uint32_t dmine = delta_bits ? hdr_bits(tbl + lane * delta_bits, delta_bits > 25 ? 25 : delta_bits) : 0u;
    const uint32_t mysize = base_size + dmine;
    sub_off = hdr_bytes + warp_incl_scan(mysize, lane) - mysize;
  }
  br_init(rd, in, sub_off);
  __syncwarp();
  // ---- the three prefix codes (load_tables)
  uint32_t terr;
  {
    TableIo io;
    io.rd = rd; io.in = in;
    terr = load_tables(sm, &io, lane);
    rd = io.rd; in = io.in;
  }
  __syncwarp();

  ProdCtx pc;
  pc.sm = sm;
  pc.full_a = saddr_pinned(sm->mbar);
  pc.empty_a = pc.full_a + 8u * kQ;
  pc.lane = lane;
  pc.rnd = 0;
  pc.synced = 0;
  pc.lit_tail = 0;
  pc.lit_head_p = 0;
  uint32_t ringv = (0x100f0b04u >> (8u * (lane & 3u))) & 0xffu;   // distance ring {4, 11, 15, 16} (PageDecoder.cpp:150-153): lane l keeps entry l & 3
  bool pdone = false;
  uint32_t pos_p = 0;          // bytes the rounds published so far produce
  if (terr) {   // a malformed prefix-code description: the page ends here
    pc.publish_abort(terr);
    return;
  }
  const uint32_t postfix_mask = (1u << npostfix) - 1u;
  const saddr_t smem_a = saddr_pinned(sm);   // the page arena as a shared-window address (LUT look-ups: one multiply-add + LDS)
  for (;;) {
    if (!pc.acquire_slot()) break;
    br_topup1(rd, in);
    // ---- one command per lane, speculatively: the lanes behind the sentinel consume nothing. Straight-line code
    //      with two window refills (after the insert&copy part and after the distance part); only 24-bit extra
    //      fields branch off.
    const uint32_t pk = br_peek(rd);
    uint32_t len;
    const uint32_t sym = huff_decode_at<kCmdLutBits>(smem_a + (uint32_t)offsetof(WarpSmem, lut_cmd), sm->aux[0], sm->sorted_cmd, bgx::kNumCmdSymbols, pk, len);
    const uint32_t sent = __ballot_sync(kFull, sym == (uint32_t)bgx::kCmdSentinel);
    const uint32_t n = umin32((uint32_t)__ffs((int)sent) - 1u, 32u);       // commands in this round (no sentinel: ffs = 0)
    pdone = sent != 0;
    const bool act = lane < n;
    const bool has_copy = sym < (uint32_t)bgx::kCmdSentinel;            // else insert-only (PageDecoder.cpp:308-317)
    uint32_t ic = sym - (uint32_t)bgx::kCmdSentinel;
    ic = has_copy ? bgx::icp_insert_code(sym) : (ic > 23u ? 23u : ic);
    const saddr_t lenlut_a = smem_a + (uint32_t)offsetof(WarpSmem, lenlut);
    const uint32_t ei = lds_u32(lenlut_a + 4u * ic);
    const uint32_t ec = has_copy ? lds_u32(lenlut_a + 96u + 4u * bgx::icp_copy_code(sym)) : 0u;
    const uint32_t nbi = ei >> 16, nbc = ec >> 16;
    uint32_t adv = len + nbi + nbc;
    uint32_t ins = ei & 0xffffu, cpy = ec & 0xffffu;
    if (act && adv > 32u) {           // 24-bit extras: field by field (rare, out of line)
      ColdBits cb;
      cb.r = rd; cb.in = in;
      ins += cold_read_fields(&cb, len, nbi, nbc);
      cpy += cb.second;
      rd = cb.r;
      adv = 0;
    } else {                          // symbol + both extra-bit fields out of the one 32-bit peek
      ins += shr32(pk, len) & low_mask(nbi);
      cpy += shr32(pk, len + nbi) & low_mask(nbc);
    }
    if (!act) {
      ins = 0;
      cpy = 0;
      adv = lane == n ? len : 0u;     // the sentinel's code bits are consumed; its lane then continues with literals
    }
    br_skip(rd, in, adv);
    uint32_t dx = 0;                  // explicit distance, or 0x80000000 | short code 0..15
    {
      uint32_t adv2 = 0;
      if (cpy) {
        dx = 0x80000000u;             // implicit "last distance" (symbol < 128, PageDecoder.cpp:305)
        if (sym >= 128u) {
          const uint32_t pk2 = br_peek(rd);
          uint32_t len2;
          const uint32_t dcode = huff_decode_at<kDistLutBits>(smem_a + (uint32_t)offsetof(WarpSmem, lut_dist), sm->aux[1], sm->sorted_dist, bgx::kNumDistSymbols, pk2, len2);
          const bool expl = dcode >= 16u + ndirect;                     // distance with extra bits (PageDecoder.cpp:376-394)
          const uint32_t v = dcode - ndirect - 16u;
          uint32_t nb = 1u + (v >> (npostfix + 1u));
          nb = expl ? (nb > 24u ? 24u : nb) : 0u;
          uint32_t extra;
          if (len2 + nb <= 32u) {
            extra = shr32(pk2, len2) & low_mask(nb);
            adv2 = len2 + nb;
          } else {                    // (rare, out of line)
            ColdBits cb;
            cb.r = rd; cb.in = in;
            extra = cold_read_fields(&cb, len2, nb, 0u);
            rd = cb.r;
          }
          const uint32_t h = v >> npostfix, lo = v & postfix_mask;
          const uint32_t dexp = (((((2u + (h & 1u)) << nb) - 4u + extra) << npostfix) + lo + ndirect + 1u) & 0x7fffffffu;
          // direct codes (PageDecoder.cpp:369-373) / ring codes
          dx = expl ? dexp : (dcode >= 16u ? dcode - 15u : (0x80000000u | dcode));
        }
      }
      br_skip(rd, in, adv2);
    }
    // ---- distance ring, resolved by relaxation (PageDecoder.cpp:345-404)
    {
      const bool has_copy = cpy != 0;
      const uint32_t dcode = (dx >> 31) ? (dx & 0xffu) : 16u;   // 16 = explicit distance
      uint32_t dist = (dx >> 31) ? 0u : dx;
      const uint32_t push = __ballot_sync(kFull, has_copy && dcode != 0);   // commands that enter the ring
      const uint32_t any_short = __ballot_sync(kFull, has_copy && dcode < 16u);
      if (any_short) {
      const uint32_t below = push & lt_mask;
      bool unresolved = has_copy && dcode < 16u;
      uint32_t slot = 0;       // which ring slot (0..3) the short code refers to, and the offset applied to it
      int32_t delta = 0;
      if (unresolved) {
        if (dcode < 4u) slot = dcode;
        else {
          const uint32_t c = dcode - 4u;              // 0..11
          slot = c >= 6u ? 1u : 0u;
          const uint32_t k = c >= 6u ? c - 6u : c;    // 0..5 => -1 +1 -2 +2 -3 +3
          delta = (int32_t)(k >> 1) + 1;
          if (!(k & 1u)) delta = -delta;
        }
      }
      // source: the slot-th most recent pusher below me, else the carried ring
      uint32_t b = below;
      for (uint32_t k = 0; k < slot && b; ++k) b &= ~(1u << (31 - __clz((int)b)));
      const uint32_t npush_below = __popc(below);
      const bool from_carry = slot >= npush_below;
      const uint32_t src_lane = from_carry ? 0u : (uint32_t)(31 - __clz((int)b));
      const uint32_t cslot = slot - (from_carry ? npush_below : 0u);
      const uint32_t carry_val = __shfl_sync(kFull, ringv, (int)cslot);
#ifndef BGX_RING_RELAX
      // Pointer jumping: every unresolved command holds (source lane, accumulated offset); a step either picks up the
      // source's final value or adopts the source's own source, so a chain of length n resolves in ceil(log2 n) + 1
      // steps instead of the n of a plain relaxation (-DBGX_RING_RELAX; record-like data averages 5.6 relaxation
      // steps per round, 2.9 jumps: structured binary 182 -> 192 GB/s).
      if (unresolved && from_carry) {
        dist = (uint32_t)((int32_t)carry_val + delta);
        unresolved = false;
      }
      uint32_t ptr = src_lane;
      int32_t off = delta;
      uint32_t pend = __ballot_sync(kFull, unresolved);
      while (pend) {
        BGX_STAT(emu_stats().ring_iters++);
        const uint32_t t_dist = __shfl_sync(kFull, dist, ptr);
        const uint32_t t_ptr = __shfl_sync(kFull, ptr, ptr);
        const int32_t t_off = __shfl_sync(kFull, off, ptr);
        if (unresolved) {
          if (!((pend >> ptr) & 1u)) {            // the source is final: so am I
            dist = (uint32_t)((int32_t)t_dist + off);
            unresolved = false;
          } else {                                // hop over the source
            off += t_off;
            ptr = t_ptr;
          }
        }
        pend = __ballot_sync(kFull, unresolved);
      }
#else
      uint32_t resolved = __ballot_sync(kFull, !unresolved);
      while (resolved != kFull) {
        BGX_STAT(emu_stats().ring_iters++);
        const uint32_t v = __shfl_sync(kFull, dist, src_lane);
        const bool ready = unresolved && (from_carry || ((resolved >> src_lane) & 1u));
        if (ready) {
          dist = (uint32_t)((int32_t)(from_carry ? carry_val : v) + delta);
          unresolved = false;
        }
        resolved = __ballot_sync(kFull, !unresolved);
      }
#endif
      }
      if (push) {   // new carried ring = the four most recent pushers of this round, then the old ring
        // lane l keeps ring[l & 3]: it wants the (l & 3)-th most recent pusher, or -- past the pushers -- an old entry
        const uint32_t j = lane & 3u;
        uint32_t pbj = push;
#pragma unroll
        for (uint32_t k = 0; k < 3; ++k)
          if (k < j && pbj) pbj &= 0x7fffffffu >> __clz((int)pbj);      // drop the highest set bit
        const uint32_t from_new = __shfl_sync(kFull, dist, pbj ? 31 - __clz((int)pbj) : 0);
        const uint32_t from_old = __shfl_sync(kFull, ringv, (int)((j - (uint32_t)__popc(push)) & 3u));
        ringv = pbj ? from_new : from_old;
      }
      dx = dist;
    }
    // ---- positions: one warp scan gives every command its output and literal offsets. Unless a command carries a
    //      long run, both sums travel in one 32-bit word (bytes produced | literals << 16: 32 x 2046 < 2^16) -- the
    //      very word the consumer gets.
    uint32_t incl_tot, incl_ins, round_ins, round_out, pkv = 0;
    const bool big = __any_sync(kFull, (ins | cpy) > 1023u);
    if (!big) {
      pkv = warp_incl_scan((ins + cpy) | (ins << 16), lane);
      const uint32_t last = __shfl_sync(kFull, pkv, 31);
      incl_tot = pkv & 0xffffu; incl_ins = pkv >> 16;
      round_out = last & 0xffffu; round_ins = last >> 16;
    } else {
      const uint64_t incl_both = warp_incl_scan64(((uint64_t)(ins + cpy) << 32) | ins, lane);
      incl_tot = (uint32_t)(incl_both >> 32); incl_ins = (uint32_t)incl_both;
      round_ins = __shfl_sync(kFull, incl_ins, 31);
      round_out = __shfl_sync(kFull, incl_tot, 31);
    }
    (void)incl_ins;
    // ---- literals of this round (PageDecoder.cpp:196-206)
    const uint32_t avail = pc.lit_tail - pc.lit_head_p;     // decoded ahead of need in earlier rounds (< 32)
    const uint32_t need = round_ins > avail ? round_ins - avail : 0u;
    const uint32_t mult = n ? (n == 32u ? (need + 31u) >> 5 : cold_udiv(need + n - 1u, n)) : 0u;   // (n < 32: last round only)
    const uint32_t rl = n * mult;                     // literals the stream carries for this round
    // literal indices lit_tail + j*32 + lane (lit_tail is a multiple of 32 until the last round)
    uint32_t mine = n == 32u ? mult : (rl > lane ? (rl - lane + 31u) >> 5 : 0u);
    {
      // every command is validated here, so the consumer only ever sees rounds it can execute blindly: the page
      // must hold the round, and a match must start inside the page (PageDecoder.cpp:222-232 trusts both)
      const uint32_t o_cpy_page = pos_p + incl_tot - cpy;   // page offset where this command's copy lands
      const uint32_t baddist = __ballot_sync(kFull, cpy != 0u && (dx == 0u || dx > o_cpy_page));
      pos_p += round_out;
      if (round_out > out_size || pos_p > out_size) { pc.publish_abort(kPageErrOverrun); break; }
      if (baddist) { pc.publish_abort(kPageErrDistance); break; }
    }
    bool fast = !big && round_out <= kRoundMax && rl <= kLitQ;
    if (fast) fast = pc.wait_lit_room(rl);
    BGX_STAT(emu_stats().rounds++; emu_stats().lits += rl; if (!fast) emu_stats().slow_rounds++);
    if (fast) {
      decode_literals(sm, rd, in, pc.lit_tail, mine, lane);
      pc.publish(dx, pkv | (pdone ? kPkLast : 0u));
      pc.lit_tail += rl;
      pc.lit_head_p += round_ins;
    } else {   // a round with long runs: virtual rounds (slow_round)
      SlowRound a;
      a.pc = pc; a.rd = rd; a.in = in;
      a.ins = ins; a.cpy = cpy; a.mine = mine; a.dx = dx; a.pdone = pdone ? 1u : 0u;
      const bool ok = slow_round(&a);
      pc = a.pc; rd = a.rd; in = a.in;
      if (!ok) break;
    }
    if (pdone) break;
  }
}


// ------------------------------------------------------------------------------------- CONSUMER
// rare paths of the consumer, out of line
BGX_COLD void cold_flush_bytes(const WarpSmem* sm, uint8_t* out, uint32_t from, uint32_t to, uint32_t zero_to, uint32_t lane) {
#pragma unroll 1
  for (uint32_t p = from + lane; p < to; p += 32) out[p] = sm->ring[p & (kRing - 1)];
#pragma unroll 1
  for (uint32_t z = to + lane; z < zero_to; z += 32) out[z] = 0;
}
// Positions are kept in "v-space": v = page offset + skew, skew = (output address & 15), so that v % 16 == 0 is
// a 16-byte boundary of global memory whatever the caller's pointer is (word loads of far matches and the vector
// flush are then always aligned). Every round the consumer sees was validated by the producer (it fits the page,
// every match starts inside the page), so nothing here can fail.
#ifndef BGX_PIECE_BATCH
#define BGX_PIECE_BATCH 1
#endif
constexpr int kPieceBatch = BGX_PIECE_BATCH;   // chunks (of 32 pieces) whose source loads are issued back to back (2 was the
                                               // better depth until the kernel shrank; now 1 wins by 1-2 %, 4 loses 3 %)

BGX_DEV void consumer_warp(const PageJob& job, WarpSmem* sm) {
  const uint32_t lane = pinned(lane_id());
  const uint32_t lt_mask = pinned((1u << lane) - 1u);
  const uint32_t le_mask = pinned(0xffffffffu >> (31u - lane));
  const uint32_t skew = (uint32_t)(reinterpret_cast<uintptr_t>(job.out) & 15u);
  uint8_t* const outb = job.out - skew;     // 16-byte aligned: byte v of the page lives at outb[v]
  const uint32_t end_v = skew + job.out_size;
  uint32_t pos = skew;         // v of the next byte to produce
  uint32_t flushed = skew;     // bytes below are in global memory
  uint32_t lit_head = 0;       // page-global literal index of the next literal to place
  const saddr_t sb = saddr_pinned(sm);       // one base register; everything else is a constant offset from it
  const saddr_t full_a = sb + (uint32_t)offsetof(WarpSmem, mbar), empty_a = full_a + 8u * kQ;
  const saddr_t ring_a = sb + (uint32_t)offsetof(WarpSmem, ring), litq_a = sb + (uint32_t)offsetof(WarpSmem, litq);
  const saddr_t tab_a = sb + (uint32_t)offsetof(WarpSmem, scratch), rb_a = sb + (uint32_t)offsetof(WarpSmem, rb);
  const saddr_t tab2_a = tab_a + 128u, tab3_a = tab_a + 640u;
  const uint32_t lane4 = pinned(4u * lane);

  for (uint32_t r = 0;; ++r) {
    const uint32_t q = r & (kQ - 1u);
#ifdef BGX_HANDOVER_SYNCTHREADS
    __syncthreads();
#else
    if (!mbar_wait(full_a + 8u * q, (r / kQ) & 1u)) {   // the producer stopped publishing
      if (lane == 0) sm->ctl.err = kPageErrHang;
      break;
    }
#endif
    const uint2 cw = lds_u32x2(rb_a + (uint32_t)sizeof(RoundBuf) * q + 8u * lane);
    if (cw.y & kPkAbort) break;
    const uint32_t dist = cw.x;                        // resolved and validated by the producer
    const uint32_t pk = cw.y & kPkSums;
    uint32_t pk_prev = __shfl_up_sync(kFull, pk, 1);
    if (lane == 0) pk_prev = 0;
    const uint32_t pk_last = __shfl_sync(kFull, pk, 31);
    const uint32_t lq = pk_prev >> 16;                 // round-local index of this command's first literal
    const uint32_t ins = (pk >> 16) - lq;
    const uint32_t cpy = (pk & 0xffffu) - (pk_prev & 0xffffu) - ins;
    const uint32_t round_out = pk_last & 0xffffu;
    const uint32_t round_ins = pk_last >> 16;
    const uint32_t o_ins = pos + (pk_prev & 0xffffu);  // where this command's literals go
    const uint32_t o_cpy = o_ins + ins;                // where its copy goes
    const uint32_t round_end = pos + round_out;
    const int32_t ring_lo = (int32_t)skew > (int32_t)round_end - (int32_t)kRing ? (int32_t)skew : (int32_t)round_end - (int32_t)kRing;
#ifdef BGX_STATS
    {
      const uint32_t mi = __reduce_max_sync(kFull, ins < kCoopLen ? ins : 0u);
      const uint32_t ncp = __popc(__ballot_sync(kFull, cpy != 0));
      const uint32_t far = __reduce_add_sync(kFull, (cpy && (int32_t)(o_cpy - dist) < ring_lo) ? cpy : 0u);
      const uint32_t ov = __popc(__ballot_sync(kFull, cpy != 0 && dist < cpy));
      BGX_STAT(emu_stats().sum_max_ins += mi; emu_stats().ins_bytes += round_ins; emu_stats().copies += ncp;
               emu_stats().copy_bytes += round_out - round_ins; emu_stats().far_bytes += far; emu_stats().overlap_copies += ov);
    }
#endif
    // ---- inserts, flattened: lane t places literal t of the round (perfectly balanced, any length).
    //      Commands with literals are compacted into tab[]; a per-chunk bit mask of their first
    //      literal index turns "which command owns literal t" into one popc.
    if (round_ins) {
      const uint32_t has = __ballot_sync(kFull, ins != 0);
      if (ins) sts_u32(tab_a + 4u * __popc(has & lt_mask), o_ins - lq);
      __syncwarp();
      const uint32_t icidx = lq >> 5;
      const uint32_t icbit = ins ? (1u << (lq & 31u)) : 0u;
      const uint32_t ilast = __popc(has) - 1u;
      uint32_t before = 0;
#pragma unroll 1
      for (uint32_t c = 0, t = lane; 32u * c < round_ins; ++c, t += 32u) {
        const uint32_t M = __reduce_or_sync(kFull, icidx == c ? icbit : 0u);
        uint32_t ord = before + __popc(M & le_mask) - 1u;
        before += __popc(M);
        ord = ord < ilast ? ord : ilast;
        const uint32_t base = lds_u32(tab_a + 4u * ord);
        if (t < round_ins)   // (lanes past the round's last literal must not touch slots the producer is filling)
          sts_u8(ring_at(ring_a, base + t), lds_u8(litq_at(litq_a, lit_head + t)));
      }
    }
#ifdef BGX_HANDOVER_SYNCTHREADS
    __syncwarp();
#else
    warp_arrive(empty_a + 8u * q, lane);   // the RoundBuf and this round's literals are no longer needed
#endif
    // ---- copies. Wave 1: every copy whose source already is final, i.e. lies below the destination of the first
    //      copy of the round (or is that copy), flattened over 4-byte PIECES: lane t moves piece t of the wave
    //      (perfectly balanced, any length mix). Ready copies are compacted into tab2[]; a per-chunk bit mask of
    //      their first piece index turns "which copy owns piece t" into one popc. A piece is fetched as two aligned
    //      words + a funnel shift, from the ring (near) or from L2 (far matches: everything below `flushed` is in
    //      global memory, and ring_lo + 512 < flushed). The source words of kPieceBatch chunks can be requested
    //      back to back before the first of them is stored (the pieces of a wave are independent: a round then pays
    //      the L2 latency of its far matches once per batch, not once per chunk); the depth is 1 today, see kPieceBatch.
    uint32_t pending = __ballot_sync(kFull, cpy != 0);
    if (pending) {
      const uint32_t first = (uint32_t)__ffs((int)pending) - 1u;
      const uint32_t hwm = __shfl_sync(kFull, o_cpy, (int)first);      // everything below is final
      const bool ready1 = cpy != 0 && dist >= cpy && (lane == first || o_cpy - dist + cpy <= hwm);
      const uint32_t np1 = ready1 ? (cpy + 3u) >> 2 : 0u;
      const uint32_t E1 = warp_incl_scan(np1, lane);
      const uint32_t T1 = __shfl_sync(kFull, E1, 31);
      const uint32_t S1 = E1 - np1;
      const uint32_t m1 = __ballot_sync(kFull, ready1);
      pending &= ~m1;
      // wave-1 entry: (dst, source, end of the copy) relative to 4 * first piece index, so that piece t of the wave only
      // adds 4 t to the first two and subtracts it from the third
      if (ready1) sts_u32x4(tab2_a + 16u * __popc(m1 & lt_mask), o_cpy - 4u * S1, o_cpy - dist - 4u * S1, 4u * S1 + cpy, 0u);
      // the others, in command order: ring offset of dst | ring offset of source << 11 | length << 22 (63: not a
      // single step), and for those a second word: dst - round start (< 1024) | length (<= 1024) << 10 | distance
      // (< 1024) << 21. A copy that is not in wave 1 overlaps itself or reads bytes of this round, so its distance is
      // < kRoundMax and its source lies in the ring.
      else if (cpy) {
        const uint32_t slot = 4u * __popc(pending & lt_mask);
        sts_u32(tab3_a + slot, (o_cpy & (kRing - 1u)) | (((o_cpy - dist) & (kRing - 1u)) << 11) | ((cpy <= 32u && dist >= cpy ? cpy : 63u) << 22));
        sts_u32(tab3_a + 128u + slot, (o_cpy - pos) | (cpy << 10) | (dist << 21));   // for the long / overlapping ones
      }
      __syncwarp();
      const uint32_t cidx = S1 >> 5;
      const uint32_t cbit = ready1 ? (1u << (S1 & 31u)) : 0u;
      const uint32_t last = __popc(m1) - 1u;
      uint32_t before = 0;
#pragma unroll 1
      for (uint32_t c0 = 0; 32u * c0 < T1; c0 += kPieceBatch) {
        uint32_t w0[kPieceBatch], w1[kPieceBatch], dd[kPieceBatch], sr[kPieceBatch];
#pragma unroll
        for (int u = 0; u < kPieceBatch; ++u) {
          const uint32_t c = c0 + (uint32_t)u;
          sr[u] = 0;
          if (32u * c < T1) {                                           // (uniform)
            const uint32_t t4 = 128u * c + lane4;
            const uint32_t M = __reduce_or_sync(kFull, cidx == c ? cbit : 0u);
            uint32_t ord = before + __popc(M & le_mask) - 1u;
            before += __popc(M);
            ord = ord < last ? ord : last;                              // lanes past the end read a valid entry
            const uint4 e = lds_u32x4(tab2_a + 16u * ord);
            dd[u] = e.x + t4;                                           // destination of this piece
            const uint32_t sp = e.y + t4;                               // its source
            int32_t rem = (int32_t)(e.z - t4);                          // bytes of the copy from this piece on (<= 0 past the end)
            rem = rem > 4 ? 4 : rem;
            if (rem > 0) {
              if ((int32_t)sp >= ring_lo) {
#ifdef BGX_STRICT_LOADS   // (verification build: only the bytes of the piece are read, so that racecheck sees no overlap)
                uint32_t v4 = 0;
                for (int k = 0; k < rem; ++k) v4 |= lds_u8(ring_at(ring_a, sp + (uint32_t)k)) << (8 * k);
                w0[u] = v4 << ((sp & 3u) * 8u);
                w1[u] = (sp & 3u) ? v4 >> (32u - (sp & 3u) * 8u) : 0u;
#else
                w0[u] = lds_u32(ring_at(ring_a, sp & ~3u));
                w1[u] = lds_u32(ring_at(ring_a, (sp & ~3u) + 4u));
#endif
              } else {
                const uint32_t* g = reinterpret_cast<const uint32_t*>(outb + (sp & ~3u));
                w0[u] = ldg_u32(g);
                w1[u] = ldg_u32(g + 1);
              }
              sr[u] = ((sp & 3u) << 3) | ((uint32_t)rem << 5);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < kPieceBatch; ++u) {
          const uint32_t m = sr[u];
          const uint32_t v = __funnelshift_r(w0[u], w1[u], m);          // (shift = low five bits = 8 * (source & 3))
          if (m & (7u << 5)) sts_u8(ring_at(ring_a, dd[u]), v);
          if (m & (6u << 5)) sts_u8(ring_at(ring_a, dd[u] + 1u), v >> 8);   // rem >= 2
          if (m >= (3u << 5)) sts_u8(ring_at(ring_a, dd[u] + 2u), v >> 16);
          if (m & (4u << 5)) sts_u8(ring_at(ring_a, dd[u] + 3u), v >> 24);  // rem == 4
        }
      }
      __syncwarp();
      //    The remaining copies (they read bytes of this round's copies, or overlap themselves): in command order,
      //    each by the whole warp (lane j moves byte j; almost always a single step). In-order execution satisfies
      //    every dependency; an overlapping copy (dist < len) repeats its `dist`-byte pattern exactly as the
      //    byte-serial reference loop does (PageDecoder.cpp:222-232).
      const uint32_t npend = __popc(pending);
#pragma unroll 1
      for (uint32_t i = 0; i < npend;) {
        const uint32_t a = lds_u32(tab3_a + 4u * i);
        const uint32_t n_s = a >> 22;
        BGX_STAT(emu_stats().wavefronts++; emu_stats().sum_max_cpy += n_s);
        if (n_s <= 32u) {                                                // the usual case: one step, no overlap
          if (lane < n_s) sts_u8(ring_at(ring_a, a + lane), lds_u8(ring_at(ring_a, (a >> 11) + lane)));
        } else {                                                         // long or overlapping
          const uint32_t e = lds_u32(tab3_a + 128u + 4u * i);
          const uint32_t o_k = pos + (e & 1023u), n_k = (e >> 10) & 2047u, d_k = e >> 21;
#pragma unroll 1
          for (uint32_t j = lane; j < n_k; j += 32) {
            const uint32_t mj = j < d_k ? j : j % d_k;
            sts_u8(ring_at(ring_a, o_k + j), lds_u8(ring_at(ring_a, o_k - d_k + mj)));
          }
        }
        __syncwarp();
        ++i;
      }
    }
    pos = round_end;
    lit_head += round_ins;
    // ---- write-combined flush of complete 512-byte chunks (16 bytes per lane; v % 16 == 0 is a 16-byte boundary)
    if (pos - flushed >= kFlushChunk) {
      if (flushed & 15u) {   // page start of an unaligned output buffer
        const uint32_t to = (flushed + 15u) & ~15u;
        cold_flush_bytes(sm, outb, flushed, to, to, lane);
        flushed = to;
      }
      while (pos - flushed >= kFlushChunk) {
        const uint32_t f = flushed + 16u * lane;
        *reinterpret_cast<uint4*>(outb + f) = lds_u32x4(ring_at(ring_a, f));
        flushed += kFlushChunk;
      }
      __syncwarp();
    }
    if (cw.y & kPkLast) {
      // ---- last round: whatever is still only in the ring, then zero-fill (the reference memsets the page first)
      cold_flush_bytes(sm, outb, flushed, pos, end_v, lane);
      break;
    }
  }
}

// All threads of the CTA call this with identical arguments.
BGX_DEV_NOINLINE PageResult decode_page_cta(const PageJob& job, WarpSmem* sm, bool first_page) {
#ifndef BGX_EMULATED
  __builtin_assume(__isGlobal(job.out));
  __builtin_assume(__isGlobal(job.in));
#endif
  if (warp_index() == 0 && lane_id() == 0) {
    // (the CTA is persistent: from its second page on the words hold the previous page's barriers, possibly mid-phase
    //  after an abort -- invalidate them before they are initialised again)
    for (uint32_t i = 0; i < 2u * kQ; ++i) {
      if (!first_page) mbar_inval(saddr(&sm->mbar[i]));
      mbar_init(saddr(&sm->mbar[i]), 1u);
    }
    mbar_init_fence();
    sm->ctl.err = 0;
  }
  __syncthreads();
  if (warp_index() == 0) producer_warp(job, sm);
  else consumer_warp(job, sm);
  __syncthreads();
  PageResult res;
  res.status = sm->ctl.err;
  res.is_delta = sm->ctl.is_delta;
  __syncthreads();     // everybody has read the result before the page arena is reused
  return res;
}


// ---------------------------------------------------------------------------------------------
// Pre-conditioned streams (BC1..BC5 textures). Pages decode into a scratch buffer that holds the
// *conditioned* layout (one plane per block field); then
//   1. delta_decode_warp  -- per page, in place: byte prefix sums over the colour end-point planes
//                            (PageDecoder::DeltaDecode, PageDecoder.cpp:446-471), restarted per page;
//   2. decondition_block  -- gathers the fields of one block from the planes and writes the block at
//                            its texture address (inverse of PageDecoder::DeconditionBC1_5,
//                            PageDecoder.cpp:406-444, which scatters byte by byte).
struct DeltaPlanes {
  uint32_t count;          // colour planes (0..4)
  uint32_t lo[4], hi[4];   // [lo, hi) of each plane in the conditioned buffer
};

BGX_DEV void delta_decode_warp(uint8_t* page, uint32_t page_start, uint32_t page_size, const DeltaPlanes& dp) {
  const uint32_t lane = lane_id();
  const uint32_t page_end = page_start + page_size;
  for (uint32_t i = 0; i < dp.count; ++i) {
    const uint32_t cs = dp.lo[i], ce = dp.hi[i];
    if (!(cs < page_end && page_start < ce)) continue;
    const uint32_t a = cs > page_start ? cs - page_start : 0u;
    const uint32_t b = ce < page_end ? ce - page_start : page_size;
    uint32_t carry = 0;
    for (uint32_t base = a; base < b; base += 128u) {
      const uint32_t i0 = base + 4u * lane;
      uint32_t y[4];
      uint32_t run = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t x = (i0 + k < b) ? page[i0 + k] : 0u;
        run = (run + x) & 0xffu;
        y[k] = run;
      }
      const uint32_t incl = warp_incl_scan(run, lane);
      const uint32_t add = carry + incl - run;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (i0 + k < b) page[i0 + k] = (uint8_t)(y[k] + add);
      carry = (carry + __shfl_sync(kFull, incl, 31)) & 0xffu;
    }
    __syncwarp();
  }
}

// Texture byte offset of block `t` (index in conditioned plane order over all mips).
BGX_HD uint32_t block_texture_offset(const bgx::PreconLayout& L, uint32_t t) {
  uint32_t mip = 0;
  while (mip + 1 < L.num_mips && t >= L.mip_off_blocks[mip + 1]) ++mip;
  const uint32_t block = t - L.mip_off_blocks[mip];
  const uint32_t W = L.width_blocks[mip], H = L.height_blocks[mip];
  uint32_t row = block / W, col = block - row * W;
  if (L.swizzle && W >= 2 && H >= 2) {
    const uint32_t remW = W & 1u, effW = W - remW, effH = H - (H & 1u);
    if (row < effH && col < effW) {   // undo the 2x2 block-group swizzle
      const uint32_t eff_block = block - row * remW;
      const uint32_t grp = eff_block >> 2, in_grp = eff_block & 3u;
      const uint32_t gpr = effW >> 1;
      row = 2u * (grp / gpr) + (in_grp >> 1);
      col = 2u * (grp % gpr) + (in_grp & 1u);
    }
  }
  return L.mip_off_bytes[mip] + row * L.pitch_bytes[mip] + col * L.block_bytes;
}

// Gathers block `t` from the planes and writes it to the texture, byte by byte (any alignment).
BGX_HD void decondition_block(const bgx::PreconLayout& L, uint32_t t, const uint8_t* planes, uint8_t* tex) {
  uint8_t* dst = tex + block_texture_offset(L, t);
  for (uint32_t sub = 0; sub < L.num_sub; ++sub) {
    const uint32_t sz = L.sub_size[sub];
    const uint8_t* src = planes + L.sub_stream_off[sub] + t * sz;
    for (uint32_t k = 0; k < sz; ++k) dst[L.sub_off[sub] + k] = src[k];
  }
}

// Same, for 4-byte aligned plane buffers and block-aligned texture rows: the field layout of the
// format is a compile-time constant, so a block is assembled in registers from the widest loads each
// field allows (every even-sized field starts on an even offset of an even-based plane) and leaves as
// one 8- or 16-byte store.
template <int FMT> struct BcFields;
//                                 field sizes, one hex digit each, first field in the lowest digit
template <> struct BcFields<1> { static constexpr int n = 3, bytes = 8;  static constexpr uint32_t sizes = 0x422u; };
template <> struct BcFields<2> { static constexpr int n = 4, bytes = 16; static constexpr uint32_t sizes = 0x4228u; };
template <> struct BcFields<3> { static constexpr int n = 6, bytes = 16; static constexpr uint32_t sizes = 0x422611u; };
template <> struct BcFields<4> { static constexpr int n = 3, bytes = 8;  static constexpr uint32_t sizes = 0x611u; };
template <> struct BcFields<5> { static constexpr int n = 6, bytes = 16; static constexpr uint32_t sizes = 0x611611u; };

template <int FMT>
BGX_DEV void decondition_block_fast(const bgx::PreconLayout& L, uint32_t t, const uint8_t* planes, uint8_t* tex) {
  using F = BcFields<FMT>;
  uint32_t w[4] = {0u, 0u, 0u, 0u};
  int pos = 0;
#pragma unroll
  for (int sub = 0; sub < F::n; ++sub) {
    const int sz = (int)((F::sizes >> (4 * sub)) & 15u);
    const uint8_t* src = planes + L.sub_stream_off[sub] + t * (uint32_t)sz;
    if (sz == 1) {
      w[pos >> 2] |= (uint32_t)*src << (8 * (pos & 3));
    } else if (sz == 4 && (pos & 3) == 0) {
      w[pos >> 2] = *reinterpret_cast<const uint32_t*>(src);
    } else if (sz == 8 && (pos & 3) == 0) {
      const uint2 q = *reinterpret_cast<const uint2*>(src);
      w[pos >> 2] = q.x;
      w[(pos >> 2) + 1] = q.y;
    } else {   // even size on an even position: 16-bit granules
#pragma unroll
      for (int k = 0; k < sz; k += 2) {
        const int p = pos + k;
        w[p >> 2] |= (uint32_t)*reinterpret_cast<const uint16_t*>(src + k) << (8 * (p & 3));
      }
    }
    pos += sz;
  }
  uint8_t* dst = tex + block_texture_offset(L, t);
  if (F::bytes == 8) *reinterpret_cast<uint2*>(dst) = make_uint2(w[0], w[1]);
  else *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
}

// ---------------------------------------------------------------------------------------------
// raw page (compressed size == uncompressed size, PageDecoder.cpp:70-76): a straight copy.
// (out of line: ONE copy of the unrolled loops in the kernel, outside the address range of the round loops)
BGX_COLD void copy_page_warp(uint8_t* dst, const uint8_t* src, uint32_t n) {
  const uint32_t lane = lane_id();
  const uintptr_t a = reinterpret_cast<uintptr_t>(dst), b = reinterpret_cast<uintptr_t>(src);
  uint32_t i = 0;
  if (((a | b) & 15u) == 0) {
    const uint4* s = reinterpret_cast<const uint4*>(src);
    uint4* d = reinterpret_cast<uint4*>(dst);
    const uint32_t nv = n >> 4;
    uint32_t v = lane;
    for (; v + 7 * 32 < nv; v += 8 * 32) {
      uint4 t[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) t[u] = s[v + u * 32];
#pragma unroll
      for (int u = 0; u < 8; ++u) d[v + u * 32] = t[u];
    }
#pragma unroll 1
    for (; v < nv; v += 32) d[v] = s[v];
    i = nv << 4;
#if BGX_RAW_PATH >= 1
  } else if (((a & 15u) == 0) && ((b & 7u) == 0)) {
    // typical stream layout: page data sits 8 bytes off a 16-byte boundary (8-byte header + 4n-byte table):
    // two 8-byte loads per 16-byte store, 8 stores in flight per lane
    const uint2* s = reinterpret_cast<const uint2*>(src);
    uint4* d = reinterpret_cast<uint4*>(dst);
    const uint32_t nv = n >> 4;
    uint32_t v = lane;
    for (; v + 7 * 32 < nv; v += 8 * 32) {
      uint2 t[16];
#pragma unroll
      for (int u = 0; u < 8; ++u) { t[2 * u] = s[2 * (v + u * 32)]; t[2 * u + 1] = s[2 * (v + u * 32) + 1]; }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        uint4 w;
        w.x = t[2 * u].x; w.y = t[2 * u].y; w.z = t[2 * u + 1].x; w.w = t[2 * u + 1].y;
        d[v + u * 32] = w;
      }
    }
#pragma unroll 1
    for (; v < nv; v += 32) {
      const uint2 lo = s[2 * v], hi = s[2 * v + 1];
      uint4 w;
      w.x = lo.x; w.y = lo.y; w.z = hi.x; w.w = hi.y;
      d[v] = w;
    }
    i = nv << 4;
#else
  } else if (((a | b) & 7u) == 0) {
    // typical stream layout: page data sits 8 bytes off a 16-byte boundary (8-byte header + 4n-byte table):
    // copy in coalesced 8-byte units, 16 loads in flight per lane
    const uint2* s = reinterpret_cast<const uint2*>(src);
    uint2* d = reinterpret_cast<uint2*>(dst);
    const uint32_t nv = n >> 3;
    uint32_t v = lane;
    for (; v + 15 * 32 < nv; v += 16 * 32) {
      uint2 t[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) t[u] = s[v + u * 32];
#pragma unroll
      for (int u = 0; u < 16; ++u) d[v + u * 32] = t[u];
    }
#pragma unroll 1
    for (; v < nv; v += 32) d[v] = s[v];
    i = nv << 3;
#endif
  } else if (((a | b) & 3u) == 0) {
    const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
    uint32_t* d = reinterpret_cast<uint32_t*>(dst);
    const uint32_t nv = n >> 2;
    uint32_t v = lane;
    for (; v + 3 * 32 < nv; v += 4 * 32) {
      uint32_t t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) t[u] = s[v + u * 32];
#pragma unroll
      for (int u = 0; u < 4; ++u) d[v + u * 32] = t[u];
    }
#pragma unroll 1
    for (; v < nv; v += 32) d[v] = s[v];
    i = nv << 2;
  }
#pragma unroll 1
  for (uint32_t j = i + lane; j < n; j += 32) dst[j] = src[j];
}

// both warps of the page's CTA take half of a raw page each
BGX_DEV void copy_page_cta_ldst(uint8_t* dst, const uint8_t* src, uint32_t n) {
  uint32_t h = ((n >> 1) + 255u) & ~255u;
  if (h > n) h = n;
  if (warp_index() == 0) copy_page_warp(dst, src, h);
  else copy_page_warp(dst + h, src + h, n - h);
}

#if !defined(BGX_EMULATED) && BGX_RAW_PATH == 2
// Raw page through the copy engines: the page is contiguous in the stream and in the output, so it moves in 4 KiB
// chunks global -> shared (cp.async: 16-byte units when the page sits on a 16-byte boundary of the stream, 8-byte units
// when it sits 8 bytes off one -- the usual layout: 8-byte header + 4n-byte page table) and shared -> global as ONE TMA
// bulk store per chunk (cp.async.bulk.global.shared::cta, SASS UBLKCP) issued by one elected thread: no thread touches
// the data, the output leaves as 4 KiB bursts. Three chunks rotate through the page arena (the tables are not in use).
template <int G>
BGX_DEV void copy_page_cta_bulk(uint8_t* dst, const uint8_t* src, uint32_t n, WarpSmem* sm) {
  constexpr uint32_t kChunk = 4096, kStages = 3;
  static_assert(offsetof(WarpSmem, mbar) >= kStages * kChunk + 128, "the staging chunks fit the page arena below its mbarriers");
  const uint32_t tid = threadIdx.x;
  const saddr_t buf = (saddr(sm) + 127u) & ~127u;
  const uint32_t nchunks = n / kChunk;
  auto issue = [&](uint32_t c) {
    const uint8_t* g = src + (size_t)c * kChunk;
    const saddr_t st = buf + (c % kStages) * kChunk;
#pragma unroll
    for (uint32_t i = 0; i < kChunk / (64u * G); ++i) {
      const uint32_t o = (i * 64u + tid) * G;
      if (G == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(st + o), "l"(g + o) : "memory");
      else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(st + o), "l"(g + o) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (uint32_t c = 0; c < kStages - 1 && c < nchunks; ++c) issue(c);
  for (uint32_t c = 0; c < nchunks; ++c) {
    if (c + kStages - 1 < nchunks) {
      // the stage of chunk c + 2 is the stage of chunk c - 1: its bulk store must have finished reading shared memory
      if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncthreads();
      issue(c + kStages - 1);
      asm volatile("cp.async.wait_group %0;" ::"n"(kStages - 1) : "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the cp.async writes become visible to the bulk copy
    __syncthreads();
    if (tid == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + (size_t)c * kChunk),
                   "r"(buf + (c % kStages) * kChunk), "r"(kChunk) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  __syncthreads();
  const uint32_t done = nchunks * kChunk;
  if (done < n) copy_page_cta_ldst(dst + done, src + done, n - done);
}
#endif

BGX_DEV void copy_page_cta(uint8_t* dst, const uint8_t* src, uint32_t n, WarpSmem* sm) {
#if !defined(BGX_EMULATED) && BGX_RAW_PATH == 2
  const uintptr_t a = reinterpret_cast<uintptr_t>(dst), b = reinterpret_cast<uintptr_t>(src);
  if ((a & 15u) == 0 && n >= 8192u) {
    if ((b & 15u) == 0) { copy_page_cta_bulk<16>(dst, src, n, sm); return; }
    if ((b & 7u) == 0) { copy_page_cta_bulk<8>(dst, src, n, sm); return; }
  }
#endif
  (void)sm;
  copy_page_cta_ldst(dst, src, n);
}

}  // namespace bgxk
