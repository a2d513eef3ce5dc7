#!/bin/bash
# Builds the page-kernel experiments that round 1 left emulator-verified but unmeasured and
# prints the one gpurun command that benches them all (about 15 s of GPU time per variant).
set -e
cd "$(dirname "$0")/.."
scripts/variants.sh \
  base "" \
  pj "-DBGX_RING_PJ" \
  insp64 "-DBGX_INS_PIECES=64" \
  insp160 "-DBGX_INS_PIECES=160" \
  split384 "-DBGX_SPLIT_LITS=384" \
  litq1k "-DBGX_Q=2 -DBGX_LITQ=1024 -DBGX_SPLIT_LITS=512" \
  combo "-DBGX_RING_PJ -DBGX_INS_PIECES=160 -DBGX_Q=2 -DBGX_LITQ=1024 -DBGX_SPLIT_LITS=512" > /dev/null
echo "gpurun --timeout 600 -- 'for v in base pj insp64 insp160 split384 litq1k combo base; do BGX_CUDA_LIB=\$PWD/build/variants/libbgx_\$v.so timeout 120 python scripts/gpu_bench_kinds.py 16 32 text,binary,mixed,lowent,texture 2>&1 | tail -1; done | tee gpurun_out/r2_variants.log'"
