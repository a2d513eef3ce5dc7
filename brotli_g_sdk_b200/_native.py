"""ctypes bindings of the two native libraries. The CUDA library is loaded lazily and loudly:
there is no Python or CPU fallback for decoding -- if libbrotlig_b200.so is missing the import of
the decode API raises."""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# BGX_CUDA_LIB lets kernel experiments load an alternative build of the same library (scripts/variants.sh)
CUDA_LIB_PATH = os.environ.get("BGX_CUDA_LIB") or os.path.join(HERE, "libbrotlig_b200.so")
ENC_LIB_PATH = os.path.join(HERE, "libbrotlig_b200_enc.so")

_cuda = None
_enc = None

c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_u32p = ctypes.POINTER(ctypes.c_uint32)


PROGRESS_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32)   # bgx_progress_fn


class BgxStream(ctypes.Structure):
    """mirror of bgx_stream (include/brotlig_b200.h)"""
    _fields_ = [
        ("d_src", ctypes.c_void_p), ("src_size", ctypes.c_uint32), ("src_capacity", ctypes.c_uint32),
        ("d_dst", ctypes.c_void_p), ("dst_capacity", ctypes.c_uint32),
        ("page_begin", ctypes.c_uint32), ("page_count", ctypes.c_uint32),
        ("header", ctypes.c_uint8 * 16),
    ]


class BgxPlanInfo(ctypes.Structure):
    _fields_ = [
        ("pages", ctypes.c_uint64), ("raw_pages_unknown", ctypes.c_uint64),
        ("compressed_bytes", ctypes.c_uint64), ("uncompressed_bytes", ctypes.c_uint64),
        ("kernels_per_launch", ctypes.c_uint32), ("grid_blocks", ctypes.c_uint32), ("block_threads", ctypes.c_uint32),
        ("smem_bytes_per_block", ctypes.c_uint32), ("sm_count", ctypes.c_uint32),
    ]


class EncOptions(ctypes.Structure):
    """mirror of bgxenc_options (include/brotlig_b200_encoder.h)"""
    _fields_ = [
        ("page_size", ctypes.c_uint32), ("npostfix", ctypes.c_int32), ("ndirect_msb", ctypes.c_int32),
        ("max_chain", ctypes.c_int32), ("lazy", ctypes.c_int32), ("use_ring_codes", ctypes.c_int32),
        ("rle_mode", ctypes.c_int32), ("split_insert_over", ctypes.c_int32), ("allow_raw", ctypes.c_int32),
        ("num_threads", ctypes.c_int32), ("precondition", ctypes.c_int32), ("format", ctypes.c_int32),
        ("width_blocks", ctypes.c_uint32), ("height_blocks", ctypes.c_uint32), ("pitch_bytes", ctypes.c_uint32),
        ("num_mips", ctypes.c_uint32), ("swizzle", ctypes.c_int32), ("pitch_aligned", ctypes.c_int32),
        ("delta_encode", ctypes.c_int32),
    ]


class EncStats(ctypes.Structure):
    _fields_ = [
        ("pages", ctypes.c_uint64), ("raw_pages", ctypes.c_uint64), ("commands", ctypes.c_uint64),
        ("literals", ctypes.c_uint64), ("ring_code_hits", ctypes.c_uint64 * 16), ("implicit_dist0", ctypes.c_uint64),
        ("insert_only_cmds", ctypes.c_uint64), ("table_types", (ctypes.c_uint64 * 3) * 3),
    ]


def cuda_lib() -> ctypes.CDLL:
    """The CUDA decoder library. Raises (never falls back) when it has not been built."""
    global _cuda
    if _cuda is None:
        if not os.path.exists(CUDA_LIB_PATH):
            raise RuntimeError(
                f"{CUDA_LIB_PATH} is missing: build it with `python -m brotli_g_sdk_b200.build` "
                "(nvcc, sm_100a). brotli_g_sdk_b200 has no CPU decode fallback.")
        lib = ctypes.CDLL(CUDA_LIB_PATH)
        vp = ctypes.c_void_p
        lib.bgx_create.argtypes = [ctypes.POINTER(vp), ctypes.c_int]
        lib.bgx_create.restype = ctypes.c_int
        lib.bgx_destroy.argtypes = [vp]
        lib.bgx_destroy.restype = None
        lib.bgx_last_error.argtypes = [vp]
        lib.bgx_last_error.restype = ctypes.c_char_p
        lib.bgx_decompressed_size.argtypes = [vp]
        lib.bgx_decompressed_size.restype = ctypes.c_uint32
        lib.bgx_decode_host.argtypes = [vp, ctypes.c_uint32, vp, c_u32p, vp, ctypes.POINTER(ctypes.c_double)]
        lib.bgx_decode_host.restype = ctypes.c_int
        lib.bgx_decode_batch_host.argtypes = [vp, ctypes.c_uint32, ctypes.POINTER(vp), c_u32p, ctypes.POINTER(vp), c_u32p,
                                              ctypes.POINTER(ctypes.c_double)]
        lib.bgx_decode_batch_host.restype = ctypes.c_int
        lib.bgx_decode_batch_host_multi.argtypes = [ctypes.POINTER(vp), ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(vp), c_u32p,
                                                    ctypes.POINTER(vp), c_u32p, ctypes.POINTER(ctypes.c_double)]
        lib.bgx_decode_batch_host_multi.restype = ctypes.c_int
        lib.bgx_decode_host_progress.argtypes = [vp, ctypes.c_uint32, vp, c_u32p, vp, ctypes.POINTER(ctypes.c_double), PROGRESS_FN, vp,
                                                 ctypes.c_uint32]
        lib.bgx_decode_host_progress.restype = ctypes.c_int
        lib.bgx_plan_create.argtypes = [vp, ctypes.POINTER(BgxStream), ctypes.c_uint32, ctypes.POINTER(vp)]
        lib.bgx_plan_create.restype = ctypes.c_int
        lib.bgx_plan_launch.argtypes = [vp, vp, vp]
        lib.bgx_plan_launch.restype = ctypes.c_int
        lib.bgx_plan_finish.argtypes = [vp, vp, c_u32p]
        lib.bgx_plan_finish.restype = ctypes.c_int
        lib.bgx_plan_destroy.argtypes = [vp]
        lib.bgx_plan_destroy.restype = None
        lib.bgx_plan_get_info.argtypes = [vp, ctypes.POINTER(BgxPlanInfo)]
        lib.bgx_plan_get_info.restype = None
        _cuda = lib
    return _cuda


def enc_lib() -> ctypes.CDLL:
    global _enc
    if _enc is None:
        if not os.path.exists(ENC_LIB_PATH):
            raise RuntimeError(f"{ENC_LIB_PATH} is missing: build it with `python -m brotli_g_sdk_b200.build`")
        lib = ctypes.CDLL(ENC_LIB_PATH)
        vp = ctypes.c_void_p
        lib.bgxenc_default_options.argtypes = [ctypes.POINTER(EncOptions)]
        lib.bgxenc_default_options.restype = None
        lib.bgxenc_max_compressed_size.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int]
        lib.bgxenc_max_compressed_size.restype = ctypes.c_uint32
        lib.bgxenc_encode.argtypes = [vp, ctypes.c_uint32, vp, c_u32p, ctypes.POINTER(EncOptions)]
        lib.bgxenc_encode.restype = ctypes.c_int
        lib.bgxenc_condition.argtypes = [vp, ctypes.c_uint32, vp, ctypes.POINTER(EncOptions)]
        lib.bgxenc_condition.restype = ctypes.c_int
        lib.bgxenc_last_stats.argtypes = [ctypes.POINTER(EncStats)]
        lib.bgxenc_last_stats.restype = None
        _enc = lib
    return _enc
