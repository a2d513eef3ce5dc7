/* oracle/shim: empty stand-in for google/brotli c/enc/entropy_encode.h (encoder-only; unused on the decode path). */
#ifndef BGX_SHIM_BROTLI_ENC_entropy_encode_H
#define BGX_SHIM_BROTLI_ENC_entropy_encode_H
#include "../common/platform.h"
#endif
