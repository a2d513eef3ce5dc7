// tests/emul/emul_decode.cpp -- TEST INFRASTRUCTURE ONLY.
// Runs the device code of brotli_g_sdk_b200/csrc/page_decode.cuh on the CPU warp emulator
// (warp_emul.h), one emulated warp per page, so kernel logic can be checked without a GPU.
#include "warp_emul.h"
// clang-format off
#include "../../brotli_g_sdk_b200/csrc/page_decode.cuh"
#include "../../brotli_g_sdk_b200/csrc/host_plan.h"
// clang-format on
#include <vector>

extern "C" {

// Decodes every page of a (non-preconditioned view of a) stream. For preconditioned streams the
// output is the *conditioned* byte plane sequence (before delta decode and de-conditioning).
// status_out[page] receives the kernel status, flags_out[page] the delta flag. Returns a BROTLIG_ERROR.
int emul_decode_stream(const uint8_t* src, uint32_t src_size, uint8_t* dst, uint32_t dst_capacity, uint32_t* status_out,
                       uint32_t* flags_out, uint64_t* collectives_out) {
  // the device path requires 16-byte aligned stream buffers (cp.async staging): give it one
  std::vector<uint8_t> aligned_buf(src_size + 64 + 80);
  uint8_t* al = aligned_buf.data() + ((64 - (reinterpret_cast<uintptr_t>(aligned_buf.data()) & 63)) & 63);
  memcpy(al, src, src_size);
  memset(al + src_size, 0xDB, 64);
  src = al;
  bgx::StreamInfo si;
  const int rc = bgx::parse_stream_header(src, &si);
  if (rc) return rc;
  if (si.uncompressed_size > dst_capacity) return bgx::kErrGeneric;
  const uint8_t* table = src + si.header_bytes;
  const uint8_t* pages = table + 4 * (size_t)si.num_pages;
  // pre-conditioned streams: pages decode into the conditioned planes, then delta + de-condition
  bgx::PreconLayout layout;
  bgxk::DeltaPlanes planes_desc{};
  std::vector<uint8_t> planes;
  uint8_t* const final_dst = dst;
  if (si.preconditioned) {
    const bgx::PreconHeaderFields f = bgx::parse_precon_header(src + 8);
    if (!bgx::precon_layout_init(&layout, f.format, f.width_blocks, f.height_blocks, f.pitch_bytes, f.num_mips,
                                 f.swizzled != 0, f.pitch_aligned != 0, si.uncompressed_size))
      return bgx::kErrCorruptStream;
    planes_desc.count = layout.num_color_sub;
    for (uint32_t c = 0; c < layout.num_color_sub; ++c) {
      planes_desc.lo[c] = layout.sub_stream_off[layout.color_sub[c]];
      planes_desc.hi[c] = layout.sub_stream_off[layout.color_sub[c] + 1];
    }
    planes.assign(si.uncompressed_size + 64, 0xEE);
    dst = planes.data();
  }
  bgxk::WarpSmem* sm = new bgxk::WarpSmem();
  memset(sm, 0xCD, sizeof(*sm));
  uint64_t coll = 0;
  int worst = 0;
  for (uint32_t p = 0; p < si.num_pages; ++p) {
    const bgx::PageExtent e = bgx::page_extent(si, table, p);
    bgxk::PageResult res{0, 0};
    // the same page-extent validation as bgx_decode_pages_kernel (bgx_cuda.cu)
    const uint8_t* src_end = src + ((src_size + 15u) & ~15u);
    const size_t avail = (size_t)(src_end - pages);
    const bool in_bounds = (size_t)e.in_off <= avail && (size_t)e.in_size <= avail - (size_t)e.in_off;
    if (!in_bounds) {
      res.status = bgxk::kPageErrTable;
    } else if (e.in_size == e.out_size) {
      coll += wemu::run_block(2, [&] { bgxk::copy_page_cta(dst + e.out_off, pages + e.in_off, e.out_size, sm); });
    } else if (e.in_size < 8u || (e.in_off & 3u) != 0u) {
      res.status = bgxk::kPageErrTable;
    } else {
      bgxk::PageJob job;
      job.in = pages + e.in_off;
      job.in_size = e.in_size;
      job.in_limit = (uint32_t)(avail - (size_t)e.in_off);
      job.out = dst + e.out_off;
      job.out_size = e.out_size;
      job.allow_delta = si.preconditioned;
      bgxk::PageResult results[64];
      coll += wemu::run_block(2, [&] { results[wemu::thread_id()] = bgxk::decode_page_cta(job, sm, true); });
      res = results[0];
      for (int l = 1; l < 64; ++l)
        if (results[l].status != res.status || results[l].is_delta != res.is_delta) res.status |= 0x80000000u;
      if (!res.status && res.is_delta)
        coll += wemu::run_warp([&] { bgxk::delta_decode_warp(dst + e.out_off, e.out_off, e.out_size, planes_desc); });
    }
    if (status_out) status_out[p] = res.status;
    if (flags_out) flags_out[p] = res.is_delta;
    if (res.status) worst = bgx::kErrCorruptStream;
  }
  delete sm;
  if (si.preconditioned) {
    for (size_t i = si.uncompressed_size; i < planes.size(); ++i)
      if (planes[i] != 0xEE) return bgx::kErrGeneric;   // a page wrote past the scratch planes
    memset(final_dst, 0, si.uncompressed_size);          // pitch padding stays 0 (BrotligDecoder.cpp:448)
    for (uint32_t t = 0; t < layout.total_blocks; ++t) bgxk::decondition_block(layout, t, planes.data(), final_dst);
  }
  if (collectives_out) *collectives_out = coll;
  return worst;
}

#ifdef BGX_STATS
void emul_stats(uint64_t* out14, int reset) {
  memcpy(out14, &bgxk::emu_stats(), sizeof(bgxk::EmuStats));
  if (reset) memset(&bgxk::emu_stats(), 0, sizeof(bgxk::EmuStats));
}
#endif
// build_table on a given set of code lengths (n <= 728): returns the status every lane agreed on (0 = usable)
uint32_t emul_table_status(const uint8_t* lens, uint32_t n) {
  bgxk::WarpSmem* sm = new bgxk::WarpSmem();
  memset(sm, 0, sizeof(*sm));
  // the compact list and the per-length counts load_table hands to build_table
  uint16_t* list = reinterpret_cast<uint16_t*>(sm->ring);
  uint32_t used = 0;
  for (uint32_t s = 0; s < n; ++s)
    if (lens[s]) { list[used++] = (uint16_t)(s | ((uint32_t)lens[s] << 10)); sm->scratch[lens[s] & 15u]++; }
  uint32_t st[32];
  wemu::run_warp([&] {
    bgxk::TableRef t{sm->lut_cmd, &sm->aux[0], sm->sorted_cmd, (uint32_t)bgxk::kCmdLutBits, (uint32_t)n, 0u};
    st[wemu::lane()] = bgxk::build_table(sm, list, sm->scratch, used, t, (uint32_t)wemu::lane());
  });
  uint32_t r = st[0];
  for (int l = 1; l < 32; ++l)
    if (st[l] != r) r = 0xffffffffu;
  delete sm;
  return r;
}

// build_table on a given set of code lengths, returning what it built: the LUT (1 << bits entries), `sorted`
// (n entries, widened to u16), limit[16] and base[16]. kind 0: the 9-bit insert&copy tables (u16 sorted), kind 2: the
// literal tables (10-bit LUT, u8 sorted, n <= 256). Returns the status every lane agreed on.
uint32_t emul_build_table(const uint8_t* lens, uint32_t n, uint32_t kind, uint16_t* lut_out, uint16_t* sorted_out, uint16_t* limit_out,
                          uint16_t* base_out) {
  bgxk::WarpSmem* sm = new bgxk::WarpSmem();
  memset(sm, 0xa5, sizeof(*sm));     // stale bytes everywhere: whatever build_table does not write must not matter
  memset(sm->scratch, 0, sizeof(sm->scratch));
  uint16_t* list = reinterpret_cast<uint16_t*>(sm->ring);
  uint32_t used = 0;
  for (uint32_t s = 0; s < n; ++s)
    if (lens[s]) { list[used++] = (uint16_t)(s | ((uint32_t)lens[s] << 10)); sm->scratch[lens[s] & 15u]++; }
  uint32_t st[32];
  const uint32_t bits = kind == 2 ? (uint32_t)bgxk::kLitLutBits : (uint32_t)bgxk::kCmdLutBits;
  wemu::run_warp([&] {
    bgxk::TableRef t{kind == 2 ? sm->lut_lit : sm->lut_cmd, &sm->aux[kind], kind == 2 ? (void*)sm->sorted_lit : (void*)sm->sorted_cmd, bits, n,
                     kind == 2 ? 1u : 0u};
    st[wemu::lane()] = bgxk::build_table(sm, list, sm->scratch, used, t, (uint32_t)wemu::lane());
  });
  uint32_t r = st[0];
  for (int l = 1; l < 32; ++l)
    if (st[l] != r) r = 0xffffffffu;
  memcpy(lut_out, kind == 2 ? sm->lut_lit : sm->lut_cmd, sizeof(uint16_t) << bits);
  for (uint32_t i = 0; i < n; ++i) sorted_out[i] = kind == 2 ? (uint16_t)sm->sorted_lit[i] : sm->sorted_cmd[i];
  memcpy(limit_out, sm->aux[kind].limit, 32);
  memcpy(base_out, sm->aux[kind].base, 32);
  delete sm;
  return r;
}

uint32_t emul_warp_smem_bytes() { return (uint32_t)sizeof(bgxk::WarpSmem); }

// The host-pointer pipeline's segment planner (host_plan.h), for the CPU unit test. Writes 7 numbers per segment
// (stream, page_begin, page_count, up0, up1, dn0, dn1); returns the number of segments or -1.
int emul_plan_segments(const uint8_t* input, uint32_t input_size, uint64_t target, uint64_t* out7, uint32_t max_segments) {
  bgx::StreamInfo si;
  if (input_size < bgx::kStreamHeaderBytes || bgx::parse_stream_header(input, &si)) return -1;
  std::vector<bgx::HostSegment> seg;
  bgx::plan_stream_segments(0, input, input_size, si, (size_t)target, seg);
  if (seg.size() > max_segments) return -1;
  for (size_t k = 0; k < seg.size(); ++k) {
    const bgx::HostSegment& g = seg[k];
    const uint64_t v[7] = {g.stream, g.page_begin, g.page_count, g.up0, g.up1, g.dn0, g.dn1};
    for (int j = 0; j < 7; ++j) out7[7 * k + j] = v[j];
  }
  return (int)seg.size();
}
}
