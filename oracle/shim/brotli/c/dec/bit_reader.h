/* oracle/shim: stand-in for google/brotli c/dec/bit_reader.h (nothing from it is used by the reference decode path). */
#ifndef BGX_SHIM_BROTLI_DEC_BIT_READER_H
#define BGX_SHIM_BROTLI_DEC_BIT_READER_H
#include "../common/platform.h"
#endif
