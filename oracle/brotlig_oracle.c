/* oracle/brotlig_oracle.c -- TEST INFRASTRUCTURE ONLY (see brotlig_oracle.h).
 *
 * A CPU restatement, in plain C, of the reference Brotli-G decoder. Each function names the
 * reference code it follows. It is written for clarity, not speed, and is validated against the
 * unmodified reference (oracle/_ref) by tests/test_oracle.py -- "parity pinned by reference outputs".
 */
#include "brotlig_oracle.h"

#include <stdlib.h>
#include <string.h>

#define NUM_STREAMS 32
#define TABLE_BITS 15
#define TABLE_SIZE (1u << TABLE_BITS)
#define NUM_CMD_SYMBOLS 728 /* 704 + sentinel + 23 insert-only (BrotligConstants.h:34-40) */
#define NUM_DIST_SYMBOLS 544
#define NUM_LIT_SYMBOLS 256
#define CMD_SENTINEL 704

static bgo_stats g_stats;
void bgo_last_stats(bgo_stats* out) { *out = g_stats; }

/* ---- RFC 7932 section 5 tables (the reference gets them from google/brotli v1.0.9 and from
 *      inc/common/BrotligCommandLut.h:41-747) ---- */
static const uint32_t kInsBase[24] = {0, 1, 2, 3, 4, 5, 6, 8, 10, 14, 18, 26, 34, 50, 66, 98, 130, 194, 322, 578, 1090, 2114, 6210, 22594};
static const uint32_t kInsExtra[24] = {0, 0, 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 7, 8, 9, 10, 12, 14, 24};
static const uint32_t kCopyBase[24] = {2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 14, 18, 22, 30, 38, 54, 70, 102, 134, 198, 326, 582, 1094, 2118};
static const uint32_t kCopyExtra[24] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 7, 8, 9, 10, 24};
/* which third of the insert / copy code range each 64-symbol cell of the alphabet selects */
static const uint8_t kCellInsHi[11] = {0, 0, 0, 0, 1, 1, 0, 2, 1, 2, 2};
static const uint8_t kCellCopyHi[11] = {0, 1, 0, 1, 0, 1, 2, 0, 2, 1, 2};

static uint32_t floor_log2(uint32_t x) { uint32_t r = 0; while (x >>= 1) ++r; return r; }
/* BrotliG::Log2Floor (src/common/BrotligUtils.cpp:49-56) is the bit length */
static uint32_t bit_length(uint32_t x) { uint32_t r = 0; while (x) { x >>= 1; ++r; } return r; }
static uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

/* ---- LSB-first bit cursor; 32 of them form the de-swizzler (inc/common/BrotligDeswizzler.h:43-206).
 *      The reference keeps a 64-bit window with refill thresholds; the observable behaviour is a
 *      plain LSB-first read of up to 32 bits that may run past the sub-stream end. ---- */
typedef struct { const uint8_t* base; uint64_t bitpos; } cursor_t;
typedef struct { cursor_t s[NUM_STREAMS]; uint32_t cur; } deswizzler_t;

static uint32_t cur_peek(const cursor_t* c, uint32_t n) {
  uint64_t w;
  memcpy(&w, c->base + (c->bitpos >> 3), 8); /* little-endian host assumed, like the reference (:139-142) */
  w >>= (c->bitpos & 7);
  return n >= 32 ? (uint32_t)w : (uint32_t)(w & ((1ull << n) - 1ull));
}
static uint32_t ds_peek(deswizzler_t* d, uint32_t n) { return n ? cur_peek(&d->s[d->cur], n) : 0; }
static void ds_consume(deswizzler_t* d, uint32_t n) { d->s[d->cur].bitpos += n; }
static uint32_t ds_read(deswizzler_t* d, uint32_t n) { uint32_t v = ds_peek(d, n); ds_consume(d, n); return v; }
static void ds_switch(deswizzler_t* d) { d->cur = (d->cur + 1) & (NUM_STREAMS - 1); } /* 5-bit wrap, :205 */
static void ds_reset(deswizzler_t* d) { d->cur = 0; }

/* ---- direct 15-bit lookup tables (symbol, code length), indexed by the MSB-first code ---- */
typedef struct { uint16_t sym[TABLE_SIZE]; uint16_t len[TABLE_SIZE]; } hufftable_t;

static uint32_t reverse_bits(uint32_t v, uint32_t n) {
  uint32_t r = 0;
  for (uint32_t i = 0; i < n; ++i) r |= ((v >> i) & 1u) << (n - 1 - i);
  return r;
}

/* GenerateHuffmanTable -- src/decoder/BrotligHuffmanTable.cpp:44-71. Canonical codes: per length in
 * symbol order; a code of length L owns 2^(maxlen-L) consecutive entries of the direct table. */
static void generate_table(const uint16_t* lens, uint32_t size, uint16_t* counts, uint32_t num_lengths,
                           uint16_t* sym, uint16_t* len) {
  uint16_t next_code[16] = {0};
  const uint32_t maxlen = num_lengths - 1;
  counts[0] = 0;
  for (uint32_t i = 1; i < num_lengths; ++i) next_code[i] = (uint16_t)((next_code[i - 1] + counts[i - 1]) << 1);
  for (uint32_t i = 0; i < size; ++i) {
    const uint32_t l = lens[i];
    if (!l) continue;
    const uint32_t left = maxlen - l;
    const uint32_t start = (uint32_t)((uint16_t)(next_code[l]++) << left) & 0xffffu;
    for (uint32_t k = 0; k < (1u << left); ++k) {
      if (start + k < (1u << maxlen)) { sym[start + k] = (uint16_t)i; len[start + k] = (uint16_t)l; }
    }
  }
}

/* LoadHuffmanTable -- src/decoder/BrotligHuffmanTable.cpp:73-205 */
static int load_table(deswizzler_t* r, uint32_t alphabet, hufftable_t* t, int which) {
  static const uint16_t kFixedLens[4][4] = {{1, 1, 0, 0}, {1, 2, 2, 0}, {2, 2, 2, 2}, {1, 2, 3, 3}};
  static const uint16_t kOrder[18] = {1, 2, 3, 4, 0, 5, 17, 6, 16, 7, 8, 9, 10, 11, 12, 13, 14, 15};
  const uint32_t max_bits = bit_length(alphabet - 1);
  const uint32_t type = ds_read(r, 2);
  if (type <= 2) g_stats.table_types[which][type]++;
  switch (type) {
    case 0: { /* trivial: one symbol, zero-length code (:87-101) */
      ds_consume(r, 4);
      const uint16_t s = (uint16_t)ds_read(r, max_bits);
      for (uint32_t i = 0; i < TABLE_SIZE; ++i) { t->sym[i] = s; t->len[i] = 0; }
      ds_reset(r);
      return 0;
    }
    case 1: { /* simple: 2..4 symbols, symbol k in sub-stream k, fixed length shapes (:102-125) */
      const uint32_t nsym = ds_read(r, 2) + 1;
      const uint32_t tree_select = ds_read(r, 1);
      ds_consume(r, 1);
      if (nsym < 2) return -1; /* the reference indexes FixedCodelengths[-1] here */
      const uint32_t shape = nsym < 4 ? nsym - 2 : (tree_select ? 3 : 2);
      uint32_t pos = 0;
      for (uint32_t i = 0; i < nsym; ++i) {
        const uint16_t l = kFixedLens[shape][i];
        const uint16_t s = (uint16_t)ds_read(r, max_bits);
        const uint32_t cnt = 1u << (TABLE_BITS - l);
        for (uint32_t k = 0; k < cnt && pos < TABLE_SIZE; ++k, ++pos) { t->sym[pos] = s; t->len[pos] = l; }
        ds_switch(r);
      }
      ds_reset(r);
      return 0;
    }
    case 2: { /* complex (:126-200) */
      const uint32_t num_len_syms = ds_read(r, 4) + 4;
      uint16_t cl_lens[18] = {0}; /* uninitialised on the reference's stack when num_len_syms < 18 */
      uint16_t cl_counts[16] = {0};
      for (uint32_t i = 0; i < num_len_syms && i < 18; ++i) {
        const uint16_t l = (uint16_t)ds_read(r, 5);
        cl_lens[kOrder[i]] = l;
        if (l > 9) return -1; /* reference: out-of-bounds count/table index */
        cl_counts[l]++;
        ds_switch(r);
      }
      uint16_t cl_sym[512], cl_len[512];
      memset(cl_sym, 0, sizeof cl_sym);
      memset(cl_len, 0, sizeof cl_len);
      generate_table(cl_lens, num_len_syms, cl_counts, 10, cl_sym, cl_len);
      ds_reset(r);

      uint16_t lens[NUM_CMD_SYMBOLS];
      uint16_t counts[16] = {0};
      uint16_t prev = 8; /* BROTLI_INITIAL_REPEATED_CODE_LENGTH (:149) */
      uint32_t filled = 0;
      while (filled < alphabet) {
        const uint32_t code = reverse_bits(ds_peek(r, 9), 9); /* sBrotligReverseBits9 (:165) */
        ds_consume(r, cl_len[code]);
        const uint16_t s = cl_sym[code];
        if (s == 16) { /* repeat last EXPLICIT length, including an explicit 0 (:170-177,186-192) */
          uint32_t reps = ds_read(r, 2) + 3;
          g_stats.rle16++;
          if (reps > alphabet - filled) return -1; /* reference: assert */
          counts[prev] = (uint16_t)(counts[prev] + reps);
          while (reps--) lens[filled++] = prev;
        } else if (s == 17) { /* run of zeros; does not change "prev" (:178-185) */
          uint32_t reps = ds_read(r, 3) + 3;
          g_stats.rle17++;
          if (reps > alphabet - filled) return -1;
          counts[0] = (uint16_t)(counts[0] + reps);
          while (reps--) lens[filled++] = 0;
        } else {
          prev = s;
          counts[s]++;
          lens[filled++] = s;
        }
        ds_switch(r);
      }
      generate_table(lens, alphabet, counts, 16, t->sym, t->len);
      ds_reset(r);
      return 0;
    }
    default:
      return -1; /* reference throws (:203) */
  }
}

typedef struct {
  uint32_t npostfix, ndirect;
  uint32_t ring[4];
  hufftable_t* tab[3]; /* 0 insert&copy, 1 distance, 2 literal (BrotligConstants.h:104-107) */
  deswizzler_t rd;
} pagedec_t;

static uint32_t decode_symbol(pagedec_t* d, int which) {
  const uint32_t bits = reverse_bits(ds_peek(&d->rd, 15), 15); /* sBrotligReverseBits15 */
  ds_consume(&d->rd, d->tab[which]->len[bits]);
  return d->tab[which]->sym[bits];
}

/* PageDecoder::TranslateDistance -- src/decoder/PageDecoder.cpp:345-404 */
static uint32_t translate_distance(pagedec_t* d, uint32_t code) {
  uint32_t dist;
  if (code < 4) dist = d->ring[code];
  else if (code < 10) { const uint32_t k = (code - 4) / 2 + 1; dist = (code & 1) ? d->ring[0] + k : d->ring[0] - k; }
  else if (code < 16) { const uint32_t k = (code - 10) / 2 + 1; dist = (code & 1) ? d->ring[1] + k : d->ring[1] - k; }
  else if (d->ndirect > 0 && code < 16 + d->ndirect) dist = code - 15;
  else {
    const uint32_t v = code - d->ndirect - 16;
    uint32_t nbits = 1 + (v >> (d->npostfix + 1));
    if (nbits > 32) nbits = 32; /* reference: BrotligBitMask[n] out of bounds; unreachable for valid streams */
    const uint32_t extra = ds_read(&d->rd, nbits);
    const uint32_t hcode = v >> d->npostfix;
    const uint32_t lcode = v & ((1u << d->npostfix) - 1u);
    const uint32_t offset = ((2u + (hcode & 1u)) << (nbits & 31)) - 4u;
    dist = ((offset + extra) << d->npostfix) + lcode + d->ndirect + 1;
  }
  if (code > 0) { d->ring[3] = d->ring[2]; d->ring[2] = d->ring[1]; d->ring[1] = d->ring[0]; d->ring[0] = dist; }
  if (code < 16) g_stats.dist_code_hist[code]++;
  return dist;
}

typedef struct { uint32_t insert_len, copy_len, dist; } command_t;

/* PageDecoder::DecodeCommand -- src/decoder/PageDecoder.cpp:290-320. Returns 1 on the sentinel. */
static int decode_command(pagedec_t* d, command_t* c) {
  const uint32_t sym = decode_symbol(d, 0);
  if (sym <= CMD_SENTINEL) {
    if (sym == CMD_SENTINEL) return 1; /* LUT row 704 is (0,0): end of page (BrotligCommandLut.h:746) */
    const uint32_t cell = sym >> 6;
    const uint32_t ic = kCellInsHi[cell] * 8 + ((sym >> 3) & 7);
    const uint32_t cc = kCellCopyHi[cell] * 8 + (sym & 7);
    c->insert_len = kInsBase[ic] + ds_read(&d->rd, kInsExtra[ic]);
    c->copy_len = kCopyBase[cc] + ds_read(&d->rd, kCopyExtra[cc]);
    uint32_t dcode = 0;
    if (sym >= 128) dcode = decode_symbol(d, 1); else g_stats.implicit_dist0++;
    c->dist = translate_distance(d, dcode);
    if (sym < 128) g_stats.dist_code_hist[0]--; /* implicit: counted separately */
  } else { /* insert-only: insert code = sym - 704, no copy, distance untouched (:308-317) */
    const uint32_t ic = sym - CMD_SENTINEL;
    c->insert_len = kInsBase[ic] + ds_read(&d->rd, kInsExtra[ic]);
    c->copy_len = 0;
    g_stats.insert_only++;
  }
  return 0;
}

typedef struct {
  int precondition, swizzle, pitch_aligned, initialized;
  uint32_t format, num_mips, block_bytes, num_sub, num_color_sub, total_blocks;
  uint32_t sub_size[6], sub_off[6], color_sub[6];
  uint32_t width_blocks[33], height_blocks[33], pitch_bytes[33], num_blocks[33];
  uint32_t mip_off_bytes[34], mip_off_blocks[34], sub_stream_off[7];
} precon_t;

/* BrotligDataconditionParams::Initialize -- inc/common/BrotligDataConditioner.h:92-237 */
static int precon_init(precon_t* p, uint32_t in_size) {
  static const uint32_t kBlock[6] = {1, 8, 16, 16, 8, 16};
  static const uint32_t kNumSub[6] = {1, 3, 4, 6, 3, 6};
  static const uint32_t kSizes[6][6] = {{1}, {2, 2, 4}, {8, 2, 2, 4}, {1, 1, 6, 2, 2, 4}, {1, 1, 6}, {1, 1, 6, 1, 1, 6}};
  static const uint32_t kNumColor[6] = {0, 2, 2, 2, 2, 4};
  static const uint32_t kColor[6][4] = {{0}, {0, 1}, {1, 2}, {3, 4}, {0, 1}, {0, 1, 3, 4}};
  const uint32_t f = (p->format >= 1 && p->format <= 5) ? p->format : 0;
  const uint32_t block_px = f ? 4 : 1;
  p->block_bytes = kBlock[f];
  p->num_sub = kNumSub[f];
  for (uint32_t i = 0; i < p->num_sub; ++i) p->sub_size[i] = kSizes[f][i];
  p->num_color_sub = kNumColor[f];
  for (uint32_t i = 0; i < p->num_color_sub; ++i) p->color_sub[i] = kColor[f][i];
  if (p->num_mips == 0) p->num_mips = 1;
  p->total_blocks = p->num_blocks[0] = p->width_blocks[0] * p->height_blocks[0];
  /* pitch_bytes[0] is always non-zero on the decode side (header stores pitch-1) */
  uint32_t wpx = (p->width_blocks[0] * block_px) / 2, hpx = (p->height_blocks[0] * block_px) / 2;
  for (uint32_t mip = 1; mip <= p->num_mips; ++mip) {
    if (mip < p->num_mips) {
      p->width_blocks[mip] = (wpx + block_px - 1) / block_px;
      p->height_blocks[mip] = (hpx + block_px - 1) / block_px;
      p->num_blocks[mip] = p->width_blocks[mip] * p->height_blocks[mip];
      const uint32_t tight = p->width_blocks[mip] * p->block_bytes;
      p->pitch_bytes[mip] = p->pitch_aligned ? ((tight + 255u) / 256u) * 256u : tight;
      p->total_blocks += p->num_blocks[mip];
    }
    p->mip_off_bytes[mip] = p->mip_off_bytes[mip - 1] + p->pitch_bytes[mip - 1] * p->height_blocks[mip - 1];
    p->mip_off_blocks[mip] = p->mip_off_blocks[mip - 1] + p->num_blocks[mip - 1];
    wpx /= 2;
    hpx /= 2;
  }
  if (p->mip_off_bytes[p->num_mips] != in_size) return 0;
  for (uint32_t sub = 1; sub <= p->num_sub; ++sub) {
    if (sub < p->num_sub) p->sub_off[sub] = p->sub_off[sub - 1] + p->sub_size[sub - 1];
    p->sub_stream_off[sub] = p->sub_stream_off[sub - 1];
    for (uint32_t mip = 0; mip < p->num_mips; ++mip) p->sub_stream_off[sub] += p->num_blocks[mip] * p->sub_size[sub - 1];
  }
  if (p->sub_stream_off[p->num_sub] != p->total_blocks * p->block_bytes) return 0;
  p->initialized = 1;
  return 1;
}

/* PageDecoder::DeconditionBC1_5 -- src/decoder/PageDecoder.cpp:406-444 */
static uint32_t decondition_addr(const precon_t* p, uint32_t offset_in_plane, uint32_t sub) {
  uint32_t adj = offset_in_plane, mip = 0;
  while (adj >= p->mip_off_blocks[mip + 1] * p->sub_size[sub]) ++mip;
  adj -= p->mip_off_blocks[mip] * p->sub_size[sub];
  const uint32_t block = adj / p->sub_size[sub];
  uint32_t row = block / p->width_blocks[mip], col = block % p->width_blocks[mip];
  const uint32_t W = p->width_blocks[mip], H = p->height_blocks[mip];
  const int swz = p->swizzle && W >= 2 && H >= 2;
  const uint32_t remW = W % 2, remH = H % 2, effW = W - remW, effH = H - remH;
  if (swz && row < effH && col < effW) {
    const uint32_t eff_block = block - row * remW;
    const uint32_t groups_per_row = effW / 2;
    const uint32_t grp = eff_block / 4, in_grp = eff_block % 4;
    row = 2 * (grp / groups_per_row) + in_grp / 2;
    col = 2 * (grp % groups_per_row) + in_grp % 2;
  }
  return p->mip_off_bytes[mip] + row * p->pitch_bytes[mip] + col * p->block_bytes + p->sub_off[sub] + adj % p->sub_size[sub];
}

/* PageDecoder::DeltaDecode -- src/decoder/PageDecoder.cpp:446-471 */
static void delta_decode(const precon_t* p, uint32_t page_start, uint32_t page_end, uint8_t* data) {
  for (uint32_t i = 0; i < p->num_color_sub; ++i) {
    const uint32_t sub = p->color_sub[i];
    const uint32_t cs = p->sub_stream_off[sub], ce = p->sub_stream_off[sub + 1];
    if (cs < page_end && page_start < ce) {
      const uint32_t a = cs > page_start ? cs - page_start : 0;
      const uint32_t b = ce < page_end ? ce - page_start : page_end - page_start;
      for (uint32_t e = a + 1; e < b; ++e) data[e] = (uint8_t)(data[e] + data[e - 1]);
    }
  }
}

/* PageDecoder::Run -- src/decoder/PageDecoder.cpp:65-268 */
static int decode_page(pagedec_t* d, const precon_t* pc, uint32_t page_size, const uint8_t* in, uint32_t in_size,
                       uint8_t* output, uint32_t out_size, uint32_t out_offset) {
  uint8_t* page_out = output + out_offset;
  uint8_t* temp = NULL;
  g_stats.pages++;
  if (pc->precondition) { temp = (uint8_t*)malloc(out_size ? out_size : 1); page_out = temp; }

  if (out_size == in_size) { /* stored raw (:70-76) */
    memcpy(page_out, in, out_size);
    g_stats.raw_pages++;
  } else {
    /* page header, LSB-first (:79-89) */
    uint64_t hpos = 0;
#define HDR_READ(n, dst) do { uint64_t w_; memcpy(&w_, in + (hpos >> 3), 8); dst = (uint32_t)((w_ >> (hpos & 7)) & ((1ull << (n)) - 1ull)); hpos += (n); } while (0)
    uint32_t ndmsb, is_delta, rsvd, base_size, delta_bits;
    HDR_READ(2, d->npostfix);
    HDR_READ(4, ndmsb);
    d->ndirect = ndmsb << d->npostfix;
    HDR_READ(1, is_delta);
    is_delta = is_delta && pc->precondition;
    HDR_READ(1, rsvd);
    (void)rsvd;
    /* sub-stream size table (:100-121) */
    const uint32_t base_bits = floor_log2((in_size + NUM_STREAMS - 1) / NUM_STREAMS) + 1;
    const uint32_t dbits_bits = floor_log2(floor_log2(in_size - 1) + 1) + 1;
    HDR_READ(base_bits, base_size);
    HDR_READ(dbits_bits, delta_bits);
    uint64_t hdr_bits = 8 + base_bits + dbits_bits + (uint64_t)NUM_STREAMS * delta_bits;
    hdr_bits = ((hdr_bits + 31) / 32) * 32;
    uint64_t idx = hdr_bits / 8;
    for (int i = 0; i < NUM_STREAMS; ++i) {
      uint32_t delta = 0;
      if (delta_bits) HDR_READ(delta_bits, delta);
      d->rd.s[i].base = in + idx;
      d->rd.s[i].bitpos = 0;
      idx += base_size + delta;
    }
#undef HDR_READ
    ds_reset(&d->rd);
    if (load_table(&d->rd, NUM_CMD_SYMBOLS, d->tab[0], 0)) { free(temp); return -1; }
    if (load_table(&d->rd, NUM_DIST_SYMBOLS, d->tab[1], 1)) { free(temp); return -1; }
    if (load_table(&d->rd, NUM_LIT_SYMBOLS, d->tab[2], 2)) { free(temp); return -1; }
    d->ring[0] = 4; d->ring[1] = 11; d->ring[2] = 15; d->ring[3] = 16; /* (:150-153) */
    memset(page_out, 0, out_size);

    uint8_t* litq = (uint8_t*)malloc(page_size + 64 * 1024);
    size_t lq_front = 0, lq_back = 0;
    size_t lq_cap = page_size + 64 * 1024;
    uint32_t wpos = 0, prev_tail = 0;
    command_t last = {0, 0, 0};
    int found_sentinel = 0;
    while (!found_sentinel) { /* one round = up to 32 commands, one per sub-stream (:174-236) */
      command_t cq[NUM_STREAMS];
      uint32_t ncmd = 0, litcount = 0;
      while (ncmd != NUM_STREAMS) {
        command_t c = last; /* the reference reuses one command object: dist survives insert-only commands */
        if (decode_command(d, &c)) { found_sentinel = 1; break; }
        last = c;
        litcount += c.insert_len;
        cq[ncmd++] = c;
        ds_switch(&d->rd);
      }
      ds_reset(&d->rd);
      g_stats.rounds++;
      g_stats.commands += ncmd;
      /* literals of this round (:196-206) */
      const uint32_t need = litcount > prev_tail ? litcount - prev_tail : 0;
      const uint32_t mult = ncmd ? (need + ncmd - 1) / ncmd : 0;
      uint32_t rl = ncmd * mult;
      prev_tail = rl + prev_tail - litcount;
      g_stats.literals_decoded += rl;
      while (rl--) {
        if (lq_back >= lq_cap) { free(litq); free(temp); return -1; }
        litq[lq_back++] = (uint8_t)decode_symbol(d, 2);
        ds_switch(&d->rd);
      }
      /* inserts and copies, in command order; overlapping copies replicate byte by byte (:209-233) */
      for (uint32_t k = 0; k < ncmd; ++k) {
        const command_t* c = &cq[k];
        if ((uint64_t)wpos + c->insert_len + c->copy_len > out_size) { free(litq); free(temp); return -1; }
        if (c->copy_len && (c->dist == 0 || c->dist > wpos + c->insert_len)) { free(litq); free(temp); return -1; }
        memcpy(page_out + wpos, litq + lq_front, c->insert_len);
        wpos += c->insert_len;
        lq_front += c->insert_len;
        g_stats.literals_emitted += c->insert_len;
        if (c->insert_len > g_stats.max_insert_len) g_stats.max_insert_len = c->insert_len;
        if (c->copy_len > g_stats.max_copy_len) g_stats.max_copy_len = c->copy_len;
        if (c->copy_len && c->dist < c->copy_len) g_stats.overlap_copies++;
        for (uint32_t j = 0; j < c->copy_len; ++j, ++wpos) page_out[wpos] = page_out[wpos - c->dist];
      }
    }
    free(litq);
    if (is_delta) { delta_decode(pc, out_offset, out_offset + out_size, page_out); g_stats.delta_pages++; }
  }

  /* BCn de-conditioning scatter (:243-265); bytes of pitch padding are never written */
  if (pc->precondition) {
    const uint32_t tex_size = pc->total_blocks * pc->block_bytes;
    if (out_offset < tex_size) {
      uint32_t sub = 0;
      while (out_offset >= pc->sub_stream_off[sub + 1]) ++sub;
      uint32_t index = 0;
      while (index < out_size) {
        const uint32_t off_in_plane = index + out_offset - pc->sub_stream_off[sub];
        output[decondition_addr(pc, off_in_plane, sub)] = page_out[index++];
        if (out_offset + index >= tex_size) break;
        if (out_offset + index >= pc->sub_stream_off[sub + 1]) ++sub;
      }
    }
    free(temp);
  }
  return 0;
}

uint32_t bgo_decompressed_size(const uint8_t* src) {
  const uint32_t w0 = rd32(src), w1 = rd32(src + 4);
  const uint32_t num_pages = w0 >> 16, page_size = (32u * 1024u) << (w1 & 3u), last = (w1 >> 2) & 0x3ffffu;
  return num_pages * page_size - (last ? page_size - last : 0); /* DataStream.h:60-64 */
}

int bgo_decode(uint32_t input_size, const uint8_t* src, uint32_t* output_size, uint8_t* output) {
  (void)input_size;
  memset(&g_stats, 0, sizeof g_stats);
  const uint32_t w0 = rd32(src), w1 = rd32(src + 4);
  const uint32_t id = w0 & 0xff, magic = (w0 >> 8) & 0xff;
  if (id != (magic ^ 0xffu)) return 14; /* BROTLIG_ERROR_CORRUPT_STREAM (BrotligDecoder.cpp:438-441) */
  if (id != 5) return 15;               /* BROTLIG_ERROR_INCORRECT_STREAM_FORMAT (:443-446) */
  memset(output, 0, *output_size);      /* (:448) */
  const uint32_t num_pages = w0 >> 16;
  const uint32_t page_size = (32u * 1024u) << (w1 & 3u);
  const uint32_t last_page_size = (w1 >> 2) & 0x3ffffu;
  const uint32_t out_total = bgo_decompressed_size(src);
  const uint8_t* p = src + 8;

  precon_t* pc = (precon_t*)calloc(1, sizeof(precon_t));
  pc->precondition = (int)((w1 >> 20) & 1u);
  if (pc->precondition) { /* PreconditionHeader, DataStream.h:89-108; +1s at BrotligDecoder.cpp:470-476 */
    const uint32_t p0 = rd32(p), p1 = rd32(p + 4);
    pc->swizzle = (int)(p0 & 1u);
    pc->pitch_aligned = (int)((p0 >> 1) & 1u);
    pc->width_blocks[0] = ((p0 >> 2) & 0x7fffu) + 1;
    pc->height_blocks[0] = ((p0 >> 17) & 0x7fffu) + 1;
    pc->format = p1 & 0xffu;
    pc->num_mips = ((p1 >> 8) & 0x1fu) + 1;
    pc->pitch_bytes[0] = ((p1 >> 13) & 0x7ffffu) + 1;
    precon_init(pc, *output_size);
    p += 8;
  }
  const uint8_t* table = p; /* page table: BrotligDecoder.cpp:397-399 */
  const uint8_t* pages = p + 4 * (size_t)num_pages;

  pagedec_t* d = (pagedec_t*)calloc(1, sizeof(pagedec_t));
  for (int i = 0; i < 3; ++i) d->tab[i] = (hufftable_t*)calloc(1, sizeof(hufftable_t));
  int rc = 0;
  for (uint32_t i = 0; i < num_pages && rc == 0; ++i) { /* PageDecoderJob, BrotligDecoder.cpp:296-329 */
    const uint32_t in_off = i ? rd32(table + 4 * i) : 0;
    const uint32_t in_size = (i < num_pages - 1) ? rd32(table + 4 * (i + 1)) - in_off : rd32(table);
    const uint32_t out_off = i * page_size;
    const uint32_t out_size = (i == num_pages - 1 && last_page_size) ? last_page_size : page_size;
    if (decode_page(d, pc, page_size, pages + in_off, in_size, output, out_size, out_off)) rc = 14;
  }
  for (int i = 0; i < 3; ++i) free(d->tab[i]);
  free(d);
  free(pc);
  *output_size = out_total;
  return rc;
}
