"""brotli_g_sdk_b200 -- a B200-native Brotli-G decompressor behind the reference SDK's decode API.

Hot path (CUDA, sm_100a):   brotli_g_sdk_b200/csrc/page_decode.cuh, bgx_cuda.cu  -> libbrotlig_b200.so
Drop-in boundary (C ABI):   include/brotlig_b200.h, include/brotlig_b200/BrotliG.h
Host mirror (this package): decoder.DecompressedSize / DecodeGPU / DecodeCPU / BrotligDecoder
CPU-side stream encoder:    encoder.Encode / MaxCompressedSize (test + benchmark input generator)
"""
from .decoder import (BROTLIG_ERROR_CORRUPT_STREAM, BROTLIG_ERROR_INCORRECT_STREAM_FORMAT, BROTLIG_OK, BrotligDecoder,
                      BrotligError, DecodeCPU, DecodeGPU, DecompressedSize)
from .encoder import Condition, DataconditionParams, Encode, MaxCompressedSize

__all__ = [
    "BROTLIG_OK", "BROTLIG_ERROR_CORRUPT_STREAM", "BROTLIG_ERROR_INCORRECT_STREAM_FORMAT", "BrotligDecoder",
    "BrotligError", "DecodeCPU", "DecodeGPU", "DecompressedSize", "Condition", "DataconditionParams", "Encode",
    "MaxCompressedSize",
]
