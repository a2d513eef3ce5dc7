/* oracle/shim: minimal stand-in for google/brotli v1.0.9 c/include/brotli/types.h (not vendored by the reference). */
#ifndef BGX_SHIM_BROTLI_TYPES_H
#define BGX_SHIM_BROTLI_TYPES_H
#include <stddef.h>
#include <stdint.h>
#define BROTLI_BOOL int
#define BROTLI_TRUE 1
#define BROTLI_FALSE 0
#define TO_BROTLI_BOOL(X) (!!(X) ? BROTLI_TRUE : BROTLI_FALSE)
#define BROTLI_MAKE_UINT64_T(high, low) ((((uint64_t)(high)) << 32) | low)
/* from google/brotli c/include/brotli/encode.h (public API constants) */
#define BROTLI_MIN_WINDOW_BITS 10
#define BROTLI_MAX_WINDOW_BITS 24
#define BROTLI_LARGE_MAX_WINDOW_BITS 30
#define BROTLI_DEFAULT_WINDOW 22
#define BROTLI_MAX_QUALITY 11
#define BROTLI_DEFAULT_QUALITY 11
#endif
