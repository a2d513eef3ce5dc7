/* oracle/shim: stand-in for google/brotli c/common/context.h; the decode path needs only the type names. */
#ifndef BGX_SHIM_BROTLI_CONTEXT_H
#define BGX_SHIM_BROTLI_CONTEXT_H
#include "platform.h"
typedef enum ContextType { CONTEXT_LSB6 = 0, CONTEXT_MSB6 = 1, CONTEXT_UTF8 = 2, CONTEXT_SIGNED = 3 } ContextType;
typedef const uint8_t* ContextLut;
#endif
