for v in default nosplit pb1 pb1split default pb1; do
  if [ $v = default ]; then unset BGX_CUDA_LIB; else export BGX_CUDA_LIB=$PWD/build/variants/libbgx_$v.so; fi
  timeout 300 python scripts/gpu_bench_kinds.py 64 32 text,binary,mixed,lowent,texture 2>&1 | tail -1
done
