/* oracle/shim: empty stand-in for google/brotli c/enc/bit_cost.h (encoder-only; unused on the decode path). */
#ifndef BGX_SHIM_BROTLI_ENC_bit_cost_H
#define BGX_SHIM_BROTLI_ENC_bit_cost_H
#include "../common/platform.h"
#endif
