/* oracle/shim: minimal stand-in for google/brotli v1.0.9 c/common/platform.h. */
#ifndef BGX_SHIM_BROTLI_PLATFORM_H
#define BGX_SHIM_BROTLI_PLATFORM_H
#include <string.h>
#include "../include/brotli/types.h"
#define BROTLI_INLINE inline
#define BROTLI_UNUSED(X) (void)(X)
#define BROTLI_MIN(T, A, B) (((A) < (B)) ? (A) : (B))
#define BROTLI_MAX(T, A, B) (((A) > (B)) ? (A) : (B))
#define BROTLI_SWAP(T, A, I, J) { T __brotli_swap_tmp = (A)[(I)]; (A)[(I)] = (A)[(J)]; (A)[(J)] = __brotli_swap_tmp; }
#define BROTLI_DCHECK(x)
#endif
