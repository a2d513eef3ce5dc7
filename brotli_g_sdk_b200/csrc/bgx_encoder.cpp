// bgx_encoder.cpp -- from-scratch CPU encoder for the Brotli-G wire format (see
// include/brotlig_b200_encoder.h for why it exists). Not on the decode hot path.
//
// What it mirrors in the reference (format, not code):
//   page pipeline        /root/reference/src/encoder/PageEncoder.cpp:247-574
//   table storage        /root/reference/src/encoder/BrotligHuffman.cpp:192-363
//   swizzled writer      /root/reference/src/common/BrotligSwizzler.cpp:68-189
//   stream assembly      /root/reference/src/BrotligEncoder.cpp:537-607
//   forward conditioner  /root/reference/src/common/BrotligDataConditioner.cpp:28-133
// The LZ77 parse is our own hash-chain matcher (the reference calls google/brotli's Zopfli
// back-reference search, which is not available offline); any parse is a valid stream.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/brotlig_b200_encoder.h"
#include "bgx_format.h"

namespace {
using namespace bgx;

// ------------------------------------------------------------------ bit writer (LSB-first)
struct BitWriter {
  std::vector<uint8_t> bytes;
  uint64_t acc = 0;
  int nacc = 0;
  size_t total_bits = 0;
  void put(uint32_t nbits, uint64_t v) {
    while (nbits > 0) {
      const uint32_t take = nbits > 32 ? 32 : nbits;
      const uint64_t part = v & ((take == 64) ? ~0ull : ((1ull << take) - 1ull));
      acc |= part << nacc;
      nacc += (int)take;
      total_bits += take;
      while (nacc >= 8) {
        bytes.push_back((uint8_t)acc);
        acc >>= 8;
        nacc -= 8;
      }
      v >>= take;
      nbits -= take;
    }
  }
  size_t byte_size() const { return (total_bits + 7) / 8; }
  void clear() { bytes.clear(); acc = 0; nacc = 0; total_bits = 0; }
  void flush() {
    if (nacc > 0) {
      bytes.push_back((uint8_t)acc);
      acc = 0;
      nacc = 0;
    }
  }
};

// 32 interleaved sub-streams with a round-robin cursor (lane = sub-stream on the GPU).
struct Swizzled {
  BitWriter bs[kNumSubstreams];
  int cur = 0;
  void put(uint32_t n, uint64_t v, bool advance = false) {
    bs[cur].put(n, v);
    if (advance) next();
  }
  void next() { cur = (cur + 1) % kNumSubstreams; }
  void reset() { cur = 0; }
  void clear() { for (auto& b : bs) b.clear(); cur = 0; }
};

// ------------------------------------------------------------------ prefix codes
uint16_t reverse_bits(uint32_t v, int n) {
  uint32_t r = 0;
  for (int i = 0; i < n; ++i) r |= ((v >> i) & 1u) << (n - 1 - i);
  return (uint16_t)r;
}

// Huffman code lengths, limited to max_len by flattening small counts (count = max(count, limit)).
void build_lengths(const uint32_t* hist_in, int n, int max_len, uint8_t* lens) {
  std::vector<uint32_t> hist(hist_in, hist_in + n);
  std::fill(lens, lens + n, 0);
  struct Node { uint64_t w; int left, right; };
  for (uint32_t limit = 1;; limit *= 2) {
    std::vector<Node> nodes;
    std::vector<int> leaves;
    for (int i = 0; i < n; ++i)
      if (hist[i]) { nodes.push_back({hist[i], -1 - i, 0}); leaves.push_back((int)nodes.size() - 1); }
    if (nodes.empty()) return;
    if (nodes.size() == 1) { lens[-1 - nodes[0].left] = 0; return; }
    // two-queue Huffman on sorted leaves (stable on symbol index for determinism)
    std::vector<int> order(leaves);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return nodes[a].w < nodes[b].w; });
    std::vector<int> q2;
    size_t i1 = 0, i2 = 0;
    auto pop = [&]() {
      int r;
      if (i1 < order.size() && (i2 >= q2.size() || nodes[order[i1]].w <= nodes[q2[i2]].w)) r = order[i1++];
      else r = q2[i2++];
      return r;
    };
    const size_t nleaves = order.size();
    for (size_t k = 0; k + 1 < nleaves; ++k) {
      int a = pop(), b = pop();
      nodes.push_back({nodes[a].w + nodes[b].w, a, b});
      q2.push_back((int)nodes.size() - 1);
    }
    // depths
    std::vector<int> depth(nodes.size(), 0);
    int maxd = 0;
    for (int k = (int)nodes.size() - 1; k >= (int)nleaves; --k) {   // internal nodes, parents before children
      depth[nodes[k].left] = depth[k] + 1;
      depth[nodes[k].right] = depth[k] + 1;
    }
    for (size_t k = 0; k < nleaves; ++k) maxd = std::max(maxd, depth[k]);
    if (maxd <= max_len) {
      for (size_t k = 0; k < nleaves; ++k) lens[-1 - nodes[k].left] = (uint8_t)depth[k];
      return;
    }
    for (int i = 0; i < n; ++i)
      if (hist[i]) hist[i] = std::max(hist[i], limit * 2);
  }
}

// Canonical codes: shorter first, then symbol order; returned bit-reversed so that they can be
// written LSB-first (the decoder sees the code MSB-first: BrotligHuffman.cpp:185, PageDecoder.cpp:292).
void assign_codes(const uint8_t* lens, int n, uint16_t* codes) {
  uint32_t count[16] = {0}, next[16] = {0};
  for (int i = 0; i < n; ++i) count[lens[i]]++;
  count[0] = 0;
  for (int l = 1; l < 16; ++l) next[l] = (next[l - 1] + count[l - 1]) << 1;
  for (int i = 0; i < n; ++i) codes[i] = lens[i] ? reverse_bits(next[lens[i]]++, lens[i]) : 0;
}

struct Code {
  std::vector<uint16_t> code;
  std::vector<uint8_t> len;
};

// Run-length code of a code-length vector with symbols 0..15, 16 (repeat previous explicit 3..6)
// and 17 (zeros 3..10). Same shape of output as BrotligUtils.cpp:118-228.
void rle_code_lengths(const uint8_t* lens, int n, int mode, std::vector<uint8_t>& syms, std::vector<uint8_t>& extra) {
  if (mode == 1) {
    for (int i = 0; i < n; ++i) { syms.push_back(lens[i]); extra.push_back(0); }
    return;
  }
  int prev = kInitialRepeatLen;
  int i = 0;
  while (i < n) {
    const int v = lens[i];
    int reps = 1;
    if (i > 0)
      while (i + reps < n && lens[i + reps] == v) ++reps;
    if (i == 0) {
      syms.push_back((uint8_t)v); extra.push_back(0);
    } else if (v == 0) {
      int r = reps;
      if (r == 11) { syms.push_back(0); extra.push_back(0); --r; }
      if (r < 3) {
        while (r--) { syms.push_back(0); extra.push_back(0); }
      } else {
        while (true) {
          const int m = r > 10 ? 10 : r;
          r -= m;
          syms.push_back(kRepeatZero); extra.push_back((uint8_t)(m - 3));
          if (r < 3) break;
        }
        while (r--) { syms.push_back(0); extra.push_back(0); }
      }
    } else {
      int r = reps;
      if (prev != v) { syms.push_back((uint8_t)v); extra.push_back(0); --r; }
      if (r == 7) { syms.push_back((uint8_t)v); extra.push_back(0); --r; }
      if (r < 3) {
        while (r-- > 0) { syms.push_back((uint8_t)v); extra.push_back(0); }
      } else {
        while (true) {
          const int m = r > 6 ? 6 : r;
          r -= m;
          syms.push_back(kRepeatPrev); extra.push_back((uint8_t)(m - 3));
          if (r < 3) break;
        }
        while (r-- > 0) { syms.push_back((uint8_t)v); extra.push_back(0); }
      }
    }
    prev = v;
    i += reps;
  }
}

// Builds the prefix code for one alphabet and writes its description (trivial / simple / complex).
// Returns table type 0/1/2.
int build_and_store_table(const uint32_t* hist, int n, Swizzled& w, Code& out, int rle_mode) {
  out.code.assign(n, 0);
  out.len.assign(n, 0);
  int count = 0;
  int s4[4] = {0, 0, 0, 0};
  for (int i = 0; i < n; ++i)
    if (hist[i]) { if (count < 4) s4[count] = i; ++count; }
  const uint32_t max_bits = bit_length((uint32_t)(n - 1));
  w.reset();
  if (count <= 1) {
    w.put(2, 0);
    w.put(4, 1);            // ignored by the decoder (BrotligHuffmanTable.cpp:89)
    w.put(max_bits, (uint32_t)s4[0]);
    w.reset();
    return 0;
  }
  build_lengths(hist, n, kMaxCodeLen, out.len.data());
  assign_codes(out.len.data(), n, out.code.data());
  if (count <= 4) {
    // stored order = (length, symbol) = canonical order; decoder fills its table in stored order
    std::sort(s4, s4 + count, [&](int a, int b) {
      return out.len[a] != out.len[b] ? out.len[a] < out.len[b] : a < b;
    });
    w.put(2, 1);
    w.put(2, (uint32_t)(count - 1));
    w.put(1, (count == 4 && out.len[s4[0]] == 1) ? 1u : 0u);   // tree select: {1,2,3,3} vs {2,2,2,2}
    w.put(1, 0);
    for (int k = 0; k < count; ++k) w.put(max_bits, (uint32_t)s4[k], true);
    w.reset();
    return 1;
  }
  w.put(2, 2);
  w.put(4, kNumCodeLenCodes - 4);
  std::vector<uint8_t> syms, extra;
  rle_code_lengths(out.len.data(), n, rle_mode, syms, extra);
  uint32_t rhist[kNumCodeLenCodes] = {0};
  for (uint8_t s : syms) rhist[s]++;
  int used = 0;
  for (int i = 0; i < kNumCodeLenCodes; ++i) used += rhist[i] != 0;
  if (used == 1) rhist[rhist[0] ? 1 : 0] = 1;   // a lone symbol would get a 0-bit code; give it a partner
  uint8_t rlen[kNumCodeLenCodes];
  uint16_t rcode[kNumCodeLenCodes];
  build_lengths(rhist, kNumCodeLenCodes, kMaxCodeLenCodeLen, rlen);
  assign_codes(rlen, kNumCodeLenCodes, rcode);
  static const uint8_t kOrder[kNumCodeLenCodes] = {1, 2, 3, 4, 0, 5, 17, 6, 16, 7, 8, 9, 10, 11, 12, 13, 14, 15};
  for (int i = 0; i < kNumCodeLenCodes; ++i) w.put(5, rlen[kOrder[i]], true);
  w.reset();
  for (size_t i = 0; i < syms.size(); ++i) {
    const uint8_t s = syms[i];
    w.put(rlen[s], rcode[s]);
    if (s == kRepeatPrev) w.put(2, extra[i]);
    else if (s == kRepeatZero) w.put(3, extra[i]);
    w.next();
  }
  w.reset();
  return 2;
}

// ------------------------------------------------------------------ LZ77 parse (page-local)
struct Cmd {
  uint32_t insert_len;
  uint32_t copy_len;     // 0 => insert-only
  uint32_t distance;
  // filled in later
  uint16_t prefix;       // insert&copy symbol 0..727
  uint16_t dist_sym;     // distance symbol (valid when has_dist)
  uint8_t has_dist;
  uint8_t dist_nbits;
  uint32_t dist_extra;
};

inline uint32_t hash4(const uint8_t* p) {
  uint32_t v;
  memcpy(&v, p, 4);
  return (v * 0x9E3779B1u) >> 17;   // 15 bits
}
inline uint32_t match_len(const uint8_t* a, const uint8_t* b, uint32_t maxlen) {
  uint32_t l = 0;
  while (l + 8 <= maxlen) {
    uint64_t x, y;
    memcpy(&x, a + l, 8);
    memcpy(&y, b + l, 8);
    if (x != y) return l + (uint32_t)(__builtin_ctzll(x ^ y) >> 3);
    l += 8;
  }
  while (l < maxlen && a[l] == b[l]) ++l;
  return l;
}

// Per-thread buffers reused across pages (fresh allocations per page serialise on the kernel's mm lock).
struct PageScratch {
  std::vector<int32_t> head, prev;
  std::vector<Cmd> cmds, split;
  std::vector<int> ring_code;
  std::vector<uint8_t> lits;
  Swizzled w;
};

void lz77_parse(const uint8_t* in, uint32_t n, const bgxenc_options& opt, PageScratch& S, uint32_t* tail_literals) {
  const int max_chain = opt.max_chain > 0 ? opt.max_chain : 16;
  std::vector<int32_t>& head = S.head;
  std::vector<int32_t>& prev = S.prev;
  std::vector<Cmd>& cmds = S.cmds;
  head.assign(1 << 15, -1);
  prev.resize(n);
  cmds.clear();
  uint32_t last_dist = 0;   // most recent distance (for cheap repeat matches); ring start {4,...} handled below
  uint32_t i = 0, lit_start = 0;
  auto insert_pos = [&](uint32_t p) {
    if (p + 4 <= n) {
      const uint32_t h = hash4(in + p);
      prev[p] = head[h];
      head[h] = (int32_t)p;
    }
  };
  auto find = [&](uint32_t p, uint32_t* best_dist) -> uint32_t {
    uint32_t best = 0;
    const uint32_t maxlen = n - p;
    if (maxlen < 4) {
      // tail: still allow a short repeat of the last distance
      if (last_dist && last_dist <= p && maxlen >= 2) {
        uint32_t l = match_len(in + p, in + p - last_dist, maxlen);
        if (l >= 2) { *best_dist = last_dist; return l; }
      }
      return 0;
    }
    if (last_dist && last_dist <= p) {
      uint32_t l = match_len(in + p, in + p - last_dist, maxlen);
      if (l >= 3) { best = l; *best_dist = last_dist; }
    }
    int32_t c = head[hash4(in + p)];
    int chain = max_chain;
    while (c >= 0 && chain-- > 0) {
      const uint32_t d = p - (uint32_t)c;
      if (best >= maxlen) break;
      if (best < 4 || in[c + best] == in[p + best]) {
        const uint32_t l = match_len(in + p, in + c, maxlen);
        if (l >= 4 && l > best) { best = l; *best_dist = d; }   // ties keep the earlier (repeat / nearer) candidate
      }
      c = prev[c];
    }
    return best;
  };
  while (i < n) {
    uint32_t d = 0;
    uint32_t len = find(i, &d);
    if (len >= 2 && opt.lazy && i + 1 < n && len < 64) {
      // one-step lazy evaluation
      insert_pos(i);
      uint32_t d2 = 0;
      uint32_t len2 = find(i + 1, &d2);
      if (len2 > len + 1) {
        ++i;
        len = len2;
        d = d2;
        // fallthrough with the better match at i (position i-1 stays a literal)
        Cmd c{};
        c.insert_len = i - lit_start;
        c.copy_len = len;
        c.distance = d;
        cmds.push_back(c);
        for (uint32_t k = 0; k < len; ++k) insert_pos(i + k);
        i += len;
        lit_start = i;
        last_dist = d;
        continue;
      }
      Cmd c{};
      c.insert_len = i - lit_start;
      c.copy_len = len;
      c.distance = d;
      cmds.push_back(c);
      for (uint32_t k = 1; k < len; ++k) insert_pos(i + k);
      i += len;
      lit_start = i;
      last_dist = d;
      continue;
    }
    if (len >= 2) {
      Cmd c{};
      c.insert_len = i - lit_start;
      c.copy_len = len;
      c.distance = d;
      cmds.push_back(c);
      for (uint32_t k = 0; k < len; ++k) insert_pos(i + k);
      i += len;
      lit_start = i;
      last_dist = d;
      continue;
    }
    insert_pos(i);
    ++i;
  }
  *tail_literals = n - lit_start;
}

// ------------------------------------------------------------------ symbols
uint32_t insert_code_of(uint32_t len) {
  if (len < 6) return len;
  if (len < 130) {
    const uint32_t nb = floor_log2(len - 2) - 1u;
    return (nb << 1) + ((len - 2) >> nb) + 2;
  }
  if (len < 2114) return floor_log2(len - 66) + 10;
  if (len < 6210) return 21;
  if (len < 22594) return 22;
  return 23;
}
uint32_t copy_code_of(uint32_t len) {
  if (len < 10) return len - 2;
  if (len < 134) {
    const uint32_t nb = floor_log2(len - 6) - 1u;
    return (nb << 1) + ((len - 6) >> nb) + 4;
  }
  if (len < 2118) return floor_log2(len - 70) + 12;
  return 23;
}
// RFC 7932 section 5: (insert code, copy code, implicit distance 0) -> insert&copy symbol
uint32_t combine_codes(uint32_t ic, uint32_t cc, bool implicit_dist0) {
  const uint32_t low = (cc & 7u) | ((ic & 7u) << 3);
  if (implicit_dist0) return (cc < 8) ? low : (low | 64u);
  static const uint16_t cell_base[3][3] = {{128, 192, 384}, {256, 320, 512}, {448, 576, 640}};
  return cell_base[ic >> 3][cc >> 3] | low;
}

void encode_distance(uint32_t dist, uint32_t npostfix, uint32_t ndirect, uint16_t* sym, uint8_t* nbits, uint32_t* extra) {
  if (dist <= ndirect) {           // direct codes: symbol 16 + dist - 1
    *sym = (uint16_t)(15 + dist);
    *nbits = 0;
    *extra = 0;
    return;
  }
  const uint32_t d = (dist - ndirect - 1) + (4u << npostfix);
  const uint32_t bucket = floor_log2(d) - 1;
  const uint32_t postfix = d & ((1u << npostfix) - 1u);
  const uint32_t prefix = (d >> bucket) & 1u;
  const uint32_t offset = (2u + prefix) << bucket;
  const uint32_t nb = bucket - npostfix;
  *sym = (uint16_t)(16 + ndirect + ((2 * (nb - 1) + prefix) << npostfix) + postfix);
  *nbits = (uint8_t)nb;
  *extra = (d - offset) >> npostfix;
}

double entropy_bits(const uint32_t* h, int n) {
  uint64_t tot = 0;
  for (int i = 0; i < n; ++i) tot += h[i];
  if (!tot) return 0;
  double b = 0;
  for (int i = 0; i < n; ++i)
    if (h[i]) b -= (double)h[i] * std::log2((double)h[i] / (double)tot);
  return b;
}

thread_local bgxenc_stats g_stats;

struct PageStats {
  uint64_t commands = 0, literals = 0, ring_hits[16] = {0}, implicit0 = 0, insert_only = 0;
  int table_type[3] = {-1, -1, -1};
  bool raw = false;
};

// ------------------------------------------------------------------ page encode
// Returns compressed bytes written to out (== n means stored raw).
uint32_t encode_page(const uint8_t* in, uint32_t n, bool is_delta, bool is_last_page, const bgxenc_options& opt,
                     std::vector<uint8_t>& out, PageStats& st,
                     const uint8_t* raw_src /* what to store when falling back to raw */, PageScratch& S) {
  auto store_raw = [&]() {
    out.assign(raw_src, raw_src + n);
    st.raw = true;
    return n;
  };
  if (n <= 16) return store_raw();

  std::vector<Cmd>& cmds = S.cmds;
  uint32_t tail = 0;
  lz77_parse(in, n, opt, S, &tail);

  uint64_t num_literals = tail;
  for (auto& c : cmds) num_literals += c.insert_len;
  // Cheap incompressibility test (shape of PageEncoder.cpp:60-85): few matches, nearly all literals,
  // sampled byte entropy close to 8 bits.
  if (opt.allow_raw && cmds.size() < (n >> 8) + 2 && (double)num_literals > 0.99 * n) {
    uint32_t h[256] = {0};
    uint32_t t = 0;
    for (uint32_t p = 0; p < n; p += 13, ++t) h[in[p]]++;
    if (entropy_bits(h, 256) > (double)t * 7.92) return store_raw();
  }

  // optional splitting of long literal runs into insert-only commands (decoder coverage)
  if (opt.split_insert_over > 0) {
    std::vector<Cmd>& s = S.split;
    s.clear();
    for (auto c : cmds) {
      while ((int)c.insert_len > opt.split_insert_over) {
        Cmd io{};
        io.insert_len = (uint32_t)opt.split_insert_over;
        s.push_back(io);
        c.insert_len -= (uint32_t)opt.split_insert_over;
      }
      s.push_back(c);
    }
    cmds.swap(s);
  }
  if (tail) {
    Cmd io{};
    io.insert_len = tail;
    cmds.push_back(io);
  }

  // ---- distance ring codes (PageDecoder.cpp:345-404 is the inverse)
  uint32_t ring[4] = {4, 11, 15, 16};
  std::vector<int>& ring_code = S.ring_code;       // -1 => explicit distance
  ring_code.assign(cmds.size(), -1);
  for (size_t k = 0; k < cmds.size(); ++k) {
    Cmd& c = cmds[k];
    if (!c.copy_len) continue;
    int code = -1;
    if (opt.use_ring_codes) {
      const uint32_t d = c.distance;
      if (d == ring[0]) code = 0;
      else if (d == ring[1]) code = 1;
      else if (d == ring[2]) code = 2;
      else if (d == ring[3]) code = 3;
      else {
        static const int delta[6] = {-1, 1, -2, 2, -3, 3};
        for (int j = 0; j < 6 && code < 0; ++j)
          if ((int64_t)ring[0] + delta[j] == (int64_t)d) code = 4 + j;
        for (int j = 0; j < 6 && code < 0; ++j)
          if ((int64_t)ring[1] + delta[j] == (int64_t)d) code = 10 + j;
      }
    }
    ring_code[k] = code;
    if (code != 0) { ring[3] = ring[2]; ring[2] = ring[1]; ring[1] = ring[0]; ring[0] = c.distance; }
  }

  // ---- choose NPOSTFIX / NDIRECT (PageEncoder.cpp:324-377 searches the same space with brotli's cost model)
  uint32_t best_np = 0, best_nd = 0;
  if (opt.npostfix >= 0 || opt.ndirect_msb >= 0) {
    best_np = (uint32_t)std::max(0, opt.npostfix) & 3u;
    best_nd = ((uint32_t)std::max(0, opt.ndirect_msb) & 15u) << best_np;
  } else {
    double best_cost = 1e300;
    for (uint32_t np = 0; np <= 3; ++np) {
      for (uint32_t msb : {0u, 1u, 2u, 4u, 8u, 15u}) {
        const uint32_t nd = msb << np;
        uint32_t h[kNumDistSymbols] = {0};
        double extra_bits = 0;
        for (size_t k = 0; k < cmds.size(); ++k) {
          if (!cmds[k].copy_len) continue;
          if (ring_code[k] >= 0) { h[ring_code[k]]++; continue; }
          uint16_t s; uint8_t nb; uint32_t ex;
          encode_distance(cmds[k].distance, np, nd, &s, &nb, &ex);
          h[s]++;
          extra_bits += nb;
        }
        const double cost = entropy_bits(h, kNumDistSymbols) + extra_bits;
        if (cost < best_cost) { best_cost = cost; best_np = np; best_nd = nd; }
      }
    }
  }

  // ---- symbols
  uint32_t hc[kNumCmdSymbols] = {0}, hd[kNumDistSymbols] = {0}, hl[kNumLitSymbols] = {0};
  {
    uint32_t pos = 0;
    for (size_t k = 0; k < cmds.size(); ++k) {
      Cmd& c = cmds[k];
      for (uint32_t j = 0; j < c.insert_len; ++j) hl[in[pos + j]]++;
      pos += c.insert_len + c.copy_len;
      const uint32_t ic = insert_code_of(c.insert_len);
      if (!c.copy_len) {
        c.prefix = (uint16_t)(kCmdSentinel + ic);   // 705..727; insert_len >= 1 so ic >= 1
        c.has_dist = 0;
        st.insert_only++;
      } else {
        const uint32_t cc = copy_code_of(c.copy_len);
        const bool implicit0 = ring_code[k] == 0 && ic < 8 && cc < 16;
        c.prefix = (uint16_t)combine_codes(ic, cc, implicit0);
        c.has_dist = !implicit0;
        if (implicit0) st.implicit0++;
        if (c.has_dist) {
          if (ring_code[k] >= 0) { c.dist_sym = (uint16_t)ring_code[k]; c.dist_nbits = 0; c.dist_extra = 0; }
          else encode_distance(c.distance, best_np, best_nd, &c.dist_sym, &c.dist_nbits, &c.dist_extra);
          hd[c.dist_sym]++;
        }
        if (ring_code[k] >= 0) st.ring_hits[ring_code[k]]++;
      }
      hc[c.prefix]++;
      st.commands++;
    }
    hc[kCmdSentinel]++;
    st.literals += num_literals;
  }

  // ---- entropy-coded payload
  Swizzled& w = S.w;
  w.clear();
  Code ccode, dcode, lcode;
  st.table_type[0] = build_and_store_table(hc, kNumCmdSymbols, w, ccode, opt.rle_mode);
  st.table_type[1] = build_and_store_table(hd, kNumDistSymbols, w, dcode, opt.rle_mode);
  st.table_type[2] = build_and_store_table(hl, kNumLitSymbols, w, lcode, opt.rle_mode);

  {
    // literal queue in page order
    std::vector<uint8_t>& lits = S.lits;
    lits.clear();
    lits.reserve(num_literals);
    uint32_t pos = 0;
    for (auto& c : cmds) {
      lits.insert(lits.end(), in + pos, in + pos + c.insert_len);
      pos += c.insert_len + c.copy_len;
    }
    const uint8_t pad_literal = (uint8_t)(std::max_element(hl, hl + kNumLitSymbols) - hl);
    size_t lq = 0, k = 0;
    uint64_t carry = 0;   // literals already emitted ahead of need ("prev_tail", PageDecoder.cpp:196-199)
    bool done = false;
    while (!done) {
      w.reset();
      uint32_t ncmd = 0;
      uint64_t round_ins = 0;
      for (int s = 0; s < kNumSubstreams; ++s) {
        if (k == cmds.size()) {                       // sentinel ends the page, in whatever sub-stream it lands
          w.put(ccode.len[kCmdSentinel], ccode.code[kCmdSentinel]);
          done = true;
          break;
        }
        const Cmd& c = cmds[k++];
        w.put(ccode.len[c.prefix], ccode.code[c.prefix]);
        const uint32_t ic = insert_code_of(c.insert_len);
        w.put(insert_extra_bits(ic), c.insert_len - insert_base(ic));
        if (c.copy_len) {
          const uint32_t cc = copy_code_of(c.copy_len);
          w.put(copy_extra_bits(cc), c.copy_len - copy_base(cc));
          if (c.has_dist) {
            w.put(dcode.len[c.dist_sym], dcode.code[c.dist_sym]);
            w.put(c.dist_nbits, c.dist_extra);
          }
        }
        round_ins += c.insert_len;
        ++ncmd;
        w.next();
      }
      w.reset();
      const uint64_t need = round_ins > carry ? round_ins - carry : 0;
      const uint64_t mult = ncmd ? (need + ncmd - 1) / ncmd : 0;
      uint64_t rl = (uint64_t)ncmd * mult;
      carry = rl + carry - round_ins;
      // Literal g of the page always lands in sub-stream g mod 32. Slots past the last real literal
      // are padded with the most frequent literal in every round but the last, where the decoder
      // never looks at them (PageEncoder.cpp:526-534 pads the last round only on the last page).
      while (rl--) {
        uint8_t b;
        if (lq < lits.size()) b = lits[lq++];
        else if (!done || is_last_page) b = pad_literal;
        else break;
        w.put(lcode.len[b], lcode.code[b], true);
      }
    }
  }

  // ---- page header + sub-stream size table (BrotligSwizzler.cpp:68-142, PageDecoder.cpp:79-121)
  uint32_t sizes[kNumSubstreams], min_size = ~0u, sum = 0;
  for (int s = 0; s < kNumSubstreams; ++s) {
    w.bs[s].flush();
    sizes[s] = (uint32_t)w.bs[s].byte_size();
    min_size = std::min(min_size, sizes[s]);
    sum += sizes[s];
  }
  uint32_t delta_bits = 1;
  for (int s = 0; s < kNumSubstreams; ++s) delta_bits = std::max(delta_bits, bit_length(sizes[s] - min_size));
  const uint32_t payload = (sum + 3u) & ~3u;
  uint32_t total = payload + 4, base_bits = 0, dbits_bits = 0, header_bytes = 0;
  for (int iter = 0; iter < 16; ++iter) {
    base_bits = floor_log2((total + kNumSubstreams - 1) / kNumSubstreams) + 1;
    dbits_bits = floor_log2(floor_log2(total - 1) + 1) + 1;
    const uint32_t hbits = 8 + base_bits + dbits_bits + kNumSubstreams * delta_bits;
    header_bytes = ((hbits + 31u) / 32u) * 4u;
    if (header_bytes + payload == total) break;
    total = header_bytes + payload;
  }
  if (header_bytes + payload != total || total >= n) return store_raw();

  BitWriter hw;
  hw.put(2, best_np);
  hw.put(4, best_nd >> best_np);
  hw.put(1, is_delta ? 1u : 0u);
  hw.put(1, 0);
  hw.put(base_bits, min_size);
  hw.put(dbits_bits, delta_bits);
  for (int s = 0; s < kNumSubstreams; ++s) hw.put(delta_bits, sizes[s] - min_size);
  hw.flush();
  out.assign(total, 0);
  memcpy(out.data(), hw.bytes.data(), hw.bytes.size());
  uint32_t off = header_bytes;
  for (int s = 0; s < kNumSubstreams; ++s) {
    if (sizes[s]) memcpy(out.data() + off, w.bs[s].bytes.data(), sizes[s]);
    off += sizes[s];
  }
  return total;
}

// ------------------------------------------------------------------ forward conditioner
void swizzle_mip(uint8_t* data, uint32_t size, uint32_t block, uint32_t wb, uint32_t hb, uint32_t pitch) {
  if (wb < 2 || hb < 2) return;
  std::vector<uint8_t> tmp(data, data + size);
  const uint32_t ew = wb - (wb % 2), eh = hb - (hb % 2);
  uint32_t orow = 0, ocol = 0;
  for (uint32_t r = 0; r < eh; r += 2)
    for (uint32_t c = 0; c < ew; c += 2)
      for (uint32_t dr = 0; dr < 2; ++dr)
        for (uint32_t dc = 0; dc < 2; ++dc) {
          memcpy(data + orow * pitch + ocol * block, tmp.data() + (r + dr) * pitch + (c + dc) * block, block);
          if (++ocol == ew) { ocol = 0; ++orow; }
        }
}

void condition(const uint8_t* src, uint32_t size, const PreconLayout& L, uint8_t* dst) {
  std::vector<uint8_t> tmp(src, src + size);
  memset(dst, 0, size);
  if (L.swizzle)
    for (uint32_t m = 0; m < L.num_mips; ++m)
      swizzle_mip(tmp.data() + L.mip_off_bytes[m], L.pitch_bytes[m] * L.height_blocks[m], L.block_bytes,
                  L.width_blocks[m], L.height_blocks[m], L.pitch_bytes[m]);
  uint32_t cursor[kMaxSubBlocks];
  for (uint32_t s = 0; s < L.num_sub; ++s) cursor[s] = L.sub_stream_off[s];
  for (uint32_t m = 0; m < L.num_mips; ++m)
    for (uint32_t r = 0; r < L.height_blocks[m]; ++r)
      for (uint32_t c = 0; c < L.width_blocks[m]; ++c) {
        uint32_t p = L.mip_off_bytes[m] + r * L.pitch_bytes[m] + c * L.block_bytes;
        for (uint32_t s = 0; s < L.num_sub; ++s) {
          memcpy(dst + cursor[s], tmp.data() + p, L.sub_size[s]);
          p += L.sub_size[s];
          cursor[s] += L.sub_size[s];
        }
      }
}

// per-page delta coding of the colour end-point planes (PageEncoder.cpp:576-616)
bool delta_encode_page(uint8_t* page, uint32_t page_start, uint32_t page_end, const PreconLayout& L) {
  bool any = false;
  for (uint32_t i = 0; i < L.num_color_sub; ++i) {
    const uint32_t sub = L.color_sub[i];
    const uint32_t cs = L.sub_stream_off[sub], ce = L.sub_stream_off[sub + 1];
    if (cs < page_end && page_start < ce) {
      const uint32_t a = cs > page_start ? cs - page_start : 0;
      const uint32_t b = ce < page_end ? ce - page_start : page_end - page_start;
      uint8_t prevv = page[a];
      for (uint32_t e = a + 1; e < b; ++e) {
        const uint8_t cur = page[e];
        page[e] = (uint8_t)(cur - prevv);
        prevv = cur;
      }
      any = true;
    }
  }
  return any;
}

bool layout_from_options(const bgxenc_options& o, uint32_t size, PreconLayout* L) {
  return precon_layout_init(L, (uint32_t)o.format, o.width_blocks, o.height_blocks, o.pitch_bytes,
                            o.num_mips ? o.num_mips : 1, o.swizzle != 0, o.pitch_aligned != 0, size);
}

}  // namespace

extern "C" {

void bgxenc_default_options(bgxenc_options* o) {
  memset(o, 0, sizeof(*o));
  o->page_size = 65536;
  o->npostfix = -1;
  o->ndirect_msb = -1;
  o->max_chain = 16;
  o->lazy = 1;
  o->use_ring_codes = 1;
  o->allow_raw = 1;
  o->num_mips = 1;
}

uint32_t bgxenc_max_compressed_size(uint32_t input_size, uint32_t page_size, int precondition) {
  if (!page_size) page_size = 65536;
  const uint64_t pages = ((uint64_t)input_size + page_size - 1) / page_size;
  const uint64_t v = kStreamHeaderBytes + (precondition ? kPreconHeaderBytes : 0) + pages * 4 + input_size + 16;
  return v > 0xffffffffull ? 0xffffffffu : (uint32_t)v;
}

int bgxenc_condition(const uint8_t* src, uint32_t size, uint8_t* dst, const bgxenc_options* opt) {
  PreconLayout L;
  if (!layout_from_options(*opt, size, &L)) return kErrGeneric;
  condition(src, size, L, dst);
  return kOk;
}

void bgxenc_last_stats(bgxenc_stats* out) { *out = g_stats; }

int bgxenc_encode(const uint8_t* src, uint32_t size, uint8_t* dst, uint32_t* dst_size, const bgxenc_options* opt_in) {
  bgxenc_options opt;
  if (opt_in) opt = *opt_in; else bgxenc_default_options(&opt);
  if (!opt.page_size) opt.page_size = 65536;
  if (opt.page_size < kMinPageSize) return 2;    // BROTLIG_ERROR_MIN_PAGE_SIZE
  if (opt.page_size > kMaxPageSize) return 3;    // BROTLIG_ERROR_MAX_PAGE_SIZE
  if (opt.page_size & (opt.page_size - 1)) return kErrGeneric;
  const uint32_t page_size = opt.page_size;
  const uint32_t num_pages = size ? (uint32_t)(((uint64_t)size + page_size - 1) / page_size) : 0;
  if (num_pages > kMaxPagesPerStream) return 4;  // BROTLIG_ERROR_MAX_NUM_PAGES
  const uint32_t last_size = num_pages ? size - (num_pages - 1) * page_size : 0;

  PreconLayout L{};
  bool precon = opt.precondition != 0;
  std::vector<uint8_t> conditioned;
  const uint8_t* data = src;
  if (precon) {
    if (!layout_from_options(opt, size, &L)) precon = false;   // reference: warn and encode unconditioned
    else {
      conditioned.resize(size);
      condition(src, size, L, conditioned.data());
      data = conditioned.data();
    }
  }

  std::vector<std::vector<uint8_t>> pages(num_pages);
  std::vector<PageStats> pstats(num_pages);
  std::atomic<uint32_t> next{0};
  auto worker = [&]() {
    std::vector<uint8_t> scratch;
    PageScratch S;
    for (;;) {
      const uint32_t p = next.fetch_add(1);
      if (p >= num_pages) break;
      const uint32_t off = p * page_size;
      const uint32_t n = (p == num_pages - 1) ? last_size : page_size;
      const uint8_t* in = data + off;
      bool is_delta = false;
      if (precon && opt.delta_encode) {
        scratch.assign(in, in + n);
        is_delta = delta_encode_page(scratch.data(), off, off + n, L);
        if (is_delta) in = scratch.data();
      }
      // raw fallback stores the un-delta'd bytes (PageEncoder.cpp:321,568)
      encode_page(in, n, is_delta, p == num_pages - 1, opt, pages[p], pstats[p], data + off, S);
    }
  };
  uint32_t nt = opt.num_threads > 0 ? (uint32_t)opt.num_threads : std::max(1u, std::thread::hardware_concurrency());
  nt = std::min(nt, std::max(1u, num_pages));
  std::vector<std::thread> th;
  for (uint32_t t = 1; t < nt; ++t) th.emplace_back(worker);
  worker();
  for (auto& t : th) t.join();

  // ---- assemble: header, [precondition header], page table, pages
  uint64_t need = kStreamHeaderBytes + (precon ? kPreconHeaderBytes : 0) + (uint64_t)num_pages * 4;
  for (auto& pg : pages) need += pg.size();
  if (need > *dst_size) return kErrGeneric;
  uint8_t* o = dst;
  uint32_t idx = 0;
  for (uint32_t ps = page_size / kMinPageSize; ps > 1; ps >>= 1) ++idx;
  const uint32_t lps = (last_size == page_size) ? 0 : last_size;
  const uint32_t w0 = kStreamId | ((kStreamId ^ 0xffu) << 8) | (num_pages << 16);
  const uint32_t w1 = idx | (lps << 2) | ((precon ? 1u : 0u) << 20);
  memcpy(o, &w0, 4);
  memcpy(o + 4, &w1, 4);
  o += 8;
  if (precon) {
    const uint32_t p0 = (L.swizzle ? 1u : 0u) | ((L.pitch_aligned ? 1u : 0u) << 1) | ((L.width_blocks[0] - 1) << 2) |
                        ((L.height_blocks[0] - 1) << 17);
    const uint32_t p1 = (L.format & 0xffu) | ((L.num_mips - 1) << 8) | ((L.pitch_bytes[0] - 1) << 13);
    memcpy(o, &p0, 4);
    memcpy(o + 4, &p1, 4);
    o += 8;
  }
  uint8_t* table = o;
  o += (size_t)num_pages * 4;
  uint32_t cur = 0;
  for (uint32_t p = 0; p < num_pages; ++p) {
    memcpy(table + 4 * p, &cur, 4);
    memcpy(o + cur, pages[p].data(), pages[p].size());
    cur += (uint32_t)pages[p].size();
  }
  if (num_pages) {
    const uint32_t last_csize = (uint32_t)pages[num_pages - 1].size();   // table[0] = compressed size of the LAST page
    memcpy(table, &last_csize, 4);
  }
  *dst_size = (uint32_t)need;

  g_stats = bgxenc_stats{};
  for (auto& s : pstats) {
    g_stats.pages++;
    g_stats.raw_pages += s.raw;
    g_stats.commands += s.commands;
    g_stats.literals += s.literals;
    g_stats.implicit_dist0 += s.implicit0;
    g_stats.insert_only_cmds += s.insert_only;
    for (int i = 0; i < 16; ++i) g_stats.ring_code_hits[i] += s.ring_hits[i];
    for (int a = 0; a < 3; ++a)
      if (s.table_type[a] >= 0) g_stats.table_types[a][s.table_type[a]]++;
  }
  return kOk;
}

}  // extern "C"
