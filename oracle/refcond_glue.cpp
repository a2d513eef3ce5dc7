// oracle/refcond_glue.cpp -- TEST INFRASTRUCTURE ONLY.
// Thin C entry point over the reference's own forward pre-conditioner
// (BrotliG::Condition, /root/reference/src/common/BrotligDataConditioner.cpp:121-133, and
// BrotligDataconditionParams::Initialize, inc/common/BrotligDataConditioner.h:92-237) so tests can
// check our encoder-side conditioner byte-for-byte against the reference's.
#include "common/BrotligDataConditioner.h"

extern "C" int refcond_condition(uint32_t size, const uint8_t* in, uint8_t* out, uint32_t format,
                                 uint32_t widthInBlocks, uint32_t heightInBlocks, uint32_t pitchInBytes,
                                 uint32_t numMips, int swizzle, int pitchAligned) {
  BrotliG::BrotligDataconditionParams p = {};
  p.precondition = true;
  p.swizzle = swizzle != 0;
  p.pitchd3d12aligned = pitchAligned != 0;
  p.format = static_cast<BROTLIG_DATA_FORMAT>(format);
  p.widthInBlocks[0] = widthInBlocks;
  p.heightInBlocks[0] = heightInBlocks;
  p.pitchInBytes[0] = pitchInBytes;
  p.numMipLevels = numMips;
  if (!p.Initialize(size)) return -1;
  uint32_t outSize = 0;
  uint8_t* outData = nullptr;
  BrotliG::Condition(size, in, p, outSize, outData);
  if (outSize != size) { delete[] outData; return -2; }
  memcpy(out, outData, size);
  delete[] outData;
  return 0;
}
