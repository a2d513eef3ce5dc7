/* oracle/shim: stand-in for google/brotli c/dec/huffman.h (nothing from it is used by the reference decode path). */
#ifndef BGX_SHIM_BROTLI_DEC_HUFFMAN_H
#define BGX_SHIM_BROTLI_DEC_HUFFMAN_H
#include "../common/platform.h"
#endif
