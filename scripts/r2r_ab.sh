#!/bin/bash
# Dev aid: GPU parity (default build), then payload kinds + small-page sweep for the default build and every build/variants/libbgx_<name>.so named
# usage: bash scripts/r2r_ab.sh <tag> [variant names...]
TAG=${1:-ab}; shift
O=gpurun_out/${TAG}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > ${O}_pytest.log
cat ${O}_pytest.log
for v in default "$@"; do
  if [ $v = default ]; then unset BGX_CUDA_LIB; else export BGX_CUDA_LIB=$PWD/build/variants/libbgx_$v.so; fi
  echo "== $v" >> ${O}_variants.log
  timeout 300 python scripts/gpu_bench_kinds.py 64 32 text,binary,mixed,lowent,texture 2>&1 | tail -1 >> ${O}_variants.log
  timeout 300 python scripts/page_size_sweep.py 2048 mixed 4096,16384,65536 2>&1 | grep -o "^[0-9]* .*decompressed_GBps': [0-9.]*" | sed "s/{.*decompressed_GBps': / /" | tr '\n' ' ' >> ${O}_variants.log
  echo >> ${O}_variants.log
done
cat ${O}_variants.log
