// host_plan.h -- host-side parse of a Brotli-G stream into page jobs. Shared by the CUDA launcher
// (bgx_cuda.cu) and by the CPU warp-emulator test driver, so both walk the page table the same way.
//
// Reference semantics restated here:
//   header validation     /root/reference/src/BrotligDecoder.cpp:436-446
//   page table            /root/reference/src/BrotligDecoder.cpp:397-399 (location), :310-314 (meaning):
//                         tbl[i>0] = offset of page i from the end of the table, page 0 at offset 0,
//                         tbl[0] = compressed size of the LAST page.
#pragma once
#include <stdint.h>

#include "bgx_format.h"

namespace bgx {

struct PageExtent {
  uint32_t in_off;     // from the first page byte (end of the page table)
  uint32_t in_size;
  uint32_t out_off;
  uint32_t out_size;
};

// `table` points at the page table (host or device readable by the caller).
BGX_HD PageExtent page_extent(const StreamInfo& si, const uint8_t* table, uint32_t page) {
  PageExtent e;
  e.in_off = page ? load_le32(table + 4 * (size_t)page) : 0u;
  e.in_size = (page + 1 < si.num_pages) ? load_le32(table + 4 * (size_t)(page + 1)) - e.in_off : load_le32(table);
  e.out_off = page * si.page_size;
  e.out_size = (page + 1 == si.num_pages && si.last_page_size) ? si.last_page_size : si.page_size;
  return e;
}

// A SEGMENT is the unit of the host-pointer pipeline (upload -> decode -> download, bgx_decode_batch_host): a whole
// stream, or -- for streams that are large against the batch -- a range of its pages. Pages are independent and
// contiguous in the stream, so a range needs the stream's bytes up to the end of its last page and produces a
// contiguous slice of the output. Offsets come from the (untrusted) page table and are clamped into the stream.
struct HostSegment {
  uint32_t stream, page_begin, page_count;   // page_count 0 = the whole stream
  size_t up0, up1;                           // stream bytes [up0, up1) to upload before this segment can run
  size_t dn0, dn1;                           // output bytes [dn0, dn1) it produces
};

// Appends the segments of stream `i` (about `target` bytes of input + output each) to `seg` (a std::vector-like
// container with push_back).
template <typename Vec>
inline void plan_stream_segments(uint32_t i, const uint8_t* input, uint32_t input_size, const StreamInfo& si, size_t target,
                                 Vec& seg) {
  const size_t bytes = (size_t)input_size + si.uncompressed_size;
  const size_t table_end = (size_t)si.header_bytes + 4ull * si.num_pages;
  size_t parts = (bytes + target - 1) / target;
  if (parts > si.num_pages) parts = si.num_pages;
  if (si.preconditioned || parts < 2 || table_end > input_size) {   // (a truncated table is reported by the plan)
    seg.push_back(HostSegment{i, 0, 0, 0, input_size, 0, si.uncompressed_size});
    return;
  }
  const uint32_t per = (si.num_pages + (uint32_t)parts - 1) / (uint32_t)parts;
  size_t up_prev = 0;
  for (uint32_t pb = 0; pb < si.num_pages; pb += per) {
    const uint32_t pc = per < si.num_pages - pb ? per : si.num_pages - pb;
    const bool last = pb + pc == si.num_pages;
    // end of the range's last page inside the stream: page table entry pb + pc (offset from the end of the table)
    size_t up1 = last ? (size_t)input_size : table_end + load_le32(input + si.header_bytes + 4ull * (pb + pc));
    const size_t lo = up_prev > table_end ? up_prev : table_end;   // corrupt tables stay in bounds and in order
    if (up1 < lo) up1 = lo;
    if (up1 > input_size) up1 = input_size;
    size_t dn0 = (size_t)pb * si.page_size;
    size_t dn1 = last ? (size_t)si.uncompressed_size : (size_t)(pb + pc) * si.page_size;
    if (dn1 > si.uncompressed_size) dn1 = si.uncompressed_size;   // (parse_stream_header rejects sizes that wrap; belt and braces)
    if (dn0 > dn1) dn0 = dn1;
    seg.push_back(HostSegment{i, pb, pc, up_prev, up1, dn0, dn1});
    up_prev = up1;
  }
}

}  // namespace bgx
