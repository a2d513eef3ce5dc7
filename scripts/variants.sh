#!/bin/bash
# Dev aid: builds alternative copies of the CUDA library with extra -D flags into build/variants/ (git-ignored,
# shipped to the GPU box by gpurun) so that one GPU call can A/B several kernel variants:
#   scripts/variants.sh name1 "-DFOO=1" name2 "-DBAR=2 -DBAZ" ...
# Select one at run time with BGX_CUDA_LIB=build/variants/libbgx_<name>.so
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -shared $flags \
    brotli_g_sdk_b200/csrc/bgx_cuda.cu brotli_g_sdk_b200/csrc/brotlig_api.cpp -o build/variants/libbgx_$name.so &
done
wait
ls -la build/variants
