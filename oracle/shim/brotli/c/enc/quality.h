/* oracle/shim: empty stand-in for google/brotli c/enc/quality.h (encoder-only; unused on the decode path). */
#ifndef BGX_SHIM_BROTLI_ENC_quality_H
#define BGX_SHIM_BROTLI_ENC_quality_H
#include "../common/platform.h"
#endif
