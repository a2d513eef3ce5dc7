/* oracle/brotlig_oracle.h -- TEST INFRASTRUCTURE ONLY. Never linked into or called by the product path.
 *
 * Plain-C restatement of the reference's CPU decoder (the parity oracle for the CUDA page kernels).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 *
 * PARITY PINNING: the reference ships no tests, golden vectors or sample .brotlig files (SURVEY.md
 * section 4), so this restatement is pinned against OUTPUTS OF THE REFERENCE ITSELF: the unmodified
 * reference decode TUs are compiled from /root/reference into oracle/_ref/libbrotlig_ref.so
 * (oracle/Makefile) and tests/test_oracle.py checks oracle == reference == source bytes on generated
 * streams; the committed fixtures under tests/golden/ carry reference-verified SHA-256 digests.
 */
#ifndef BROTLIG_ORACLE_H
#define BROTLIG_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* BrotliG::DecompressedSize -- /root/reference/src/BrotligDecoder.cpp:34-38 */
uint32_t bgo_decompressed_size(const uint8_t* src);

/* BrotliG::DecodeCPU -- /root/reference/src/BrotligDecoder.cpp:426-519 (single worker).
 * *output_size: in = size of the output buffer (zero-filled, and used for the BCn layout check),
 * out = uncompressed size. Returns a BROTLIG_ERROR value (0 OK, 14 corrupt, 15 wrong format).
 * Like the reference, reads up to 8 bytes past the end of a page's sub-streams. */
int bgo_decode(uint32_t input_size, const uint8_t* src, uint32_t* output_size, uint8_t* output);

/* Per-page statistics of the last bgo_decode call (used by tests to prove corner cases were hit). */
typedef struct bgo_stats {
  uint64_t pages, raw_pages, rounds, commands, literals_emitted, literals_decoded;
  uint64_t dist_code_hist[16];     /* resolved distance symbols 0..15 (explicit symbol, not implicit) */
  uint64_t implicit_dist0;         /* commands with insert&copy symbol < 128 */
  uint64_t insert_only;            /* symbols 705..727 */
  uint64_t overlap_copies;         /* copies with distance < length */
  uint64_t table_types[3][3];      /* [icp,dist,lit][trivial,simple,complex] */
  uint64_t rle16, rle17;           /* code-length repeat symbols seen */
  uint64_t max_insert_len, max_copy_len;
  uint64_t delta_pages;
} bgo_stats;
void bgo_last_stats(bgo_stats* out);

#ifdef __cplusplus
}
#endif
#endif
