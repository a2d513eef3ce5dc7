// host_plan.h -- host-side parse of a Brotli-G stream into page jobs. Shared by the CUDA launcher
// (bgx_api.cu) and by the CPU warp-emulator test driver, so both walk the page table the same way.
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

}  // namespace bgx
