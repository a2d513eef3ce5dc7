/* oracle/shim: Log2FloorNonZero as published in google/brotli v1.0.9 c/enc/fast_log.h (floor(log2(n)), n > 0). */
#ifndef BGX_SHIM_BROTLI_FAST_LOG_H
#define BGX_SHIM_BROTLI_FAST_LOG_H
#include <math.h>
#include "../common/platform.h"
static inline uint32_t Log2FloorNonZero(size_t n) {
  uint32_t r = 0;
  while (n >>= 1) ++r;
  return r;
}
static inline double FastLog2(size_t v) { return log2((double)v); }
#endif
