/* oracle/shim: the pieces of google/brotli v1.0.9 c/enc/command.h that the reference decode path touches:
 * struct Command and the RFC 7932 section 5 insert/copy length code tables (base value, extra bits). */
#ifndef BGX_SHIM_BROTLI_ENC_COMMAND_H
#define BGX_SHIM_BROTLI_ENC_COMMAND_H
#include "../common/constants.h"
#include "../common/platform.h"
#include "fast_log.h"
static const uint32_t kBgxShimInsBase[BROTLI_NUM_INS_COPY_CODES] = {
    0, 1, 2, 3, 4, 5, 6, 8, 10, 14, 18, 26, 34, 50, 66, 98, 130, 194, 322, 578, 1090, 2114, 6210, 22594};
static const uint32_t kBgxShimInsExtra[BROTLI_NUM_INS_COPY_CODES] = {
    0, 0, 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 7, 8, 9, 10, 12, 14, 24};
static const uint32_t kBgxShimCopyBase[BROTLI_NUM_INS_COPY_CODES] = {
    2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 14, 18, 22, 30, 38, 54, 70, 102, 134, 198, 326, 582, 1094, 2118};
static const uint32_t kBgxShimCopyExtra[BROTLI_NUM_INS_COPY_CODES] = {
    0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 7, 8, 9, 10, 24};
static inline uint32_t GetInsertBase(uint16_t inscode) { return kBgxShimInsBase[inscode]; }
static inline uint32_t GetInsertExtra(uint16_t inscode) { return kBgxShimInsExtra[inscode]; }
static inline uint32_t GetCopyBase(uint16_t copycode) { return kBgxShimCopyBase[copycode]; }
static inline uint32_t GetCopyExtra(uint16_t copycode) { return kBgxShimCopyExtra[copycode]; }
typedef struct Command {
  uint32_t insert_len_;
  uint32_t copy_len_;
  uint32_t dist_extra_;
  uint16_t cmd_prefix_;
  uint16_t dist_prefix_;
} Command;
#endif
