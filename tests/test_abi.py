"""CPU tests (-m "not gpu"): the C-ABI libraries load and export every function their headers declare;
host-only calls work without a device; device calls fail loudly (no fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = re.sub(r"//[^\n]*", "", txt)
    return sorted(set(re.findall(r"\b(bgx\w*|bgxenc_\w+)\s*\(", txt)))


@pytest.fixture(scope="module")
def built():
    from brotli_g_sdk_b200 import build
    build.build_all()
    return build


def test_cuda_library_exports_the_declared_abi(built):
    lib = ctypes.CDLL(built.CUDA_LIB)
    names = [n for n in _declared_functions("brotlig_b200.h") if not n.endswith("_context") and n not in ("bgx_plan", "bgx_stream", "bgx_plan_info")]
    assert "bgx_decode_host" in names and "bgx_plan_launch" in names
    for n in names:
        assert hasattr(lib, n), f"libbrotlig_b200.so does not export {n}"
    # the reference-named entry points (include/brotlig_b200/BrotliG.h): unmangled like the reference's
    for n in ("DecompressedSize", "DecodeCPU"):
        assert hasattr(lib, n), n
    assert hasattr(lib, "_Z9DecodeGPUbjPKhPjPhRd"), "C++-linkage DecodeGPU(bool, uint32_t, const uint8_t*, uint32_t*, uint8_t*, double&)"


def test_encoder_library_exports_the_declared_abi(built):
    lib = ctypes.CDLL(built.ENC_LIB)
    names = [n for n in _declared_functions("brotlig_b200_encoder.h") if n not in ("bgxenc_options", "bgxenc_stats")]
    assert len(names) >= 5
    for n in names:
        assert hasattr(lib, n), f"libbrotlig_b200_enc.so does not export {n}"


def test_decompressed_size_is_host_only(built, sdk):
    data = np.zeros(2 * 65536 + 1000, np.uint8)
    s = sdk.Encode(data)
    assert sdk.DecompressedSize(s) == len(data)


def test_cuda_kernels_are_built_for_sm_100a(built):
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "--list-elf", built.CUDA_LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_no_device_means_a_loud_failure_not_a_fallback(built, sdk):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    s = sdk.Encode(np.arange(1000, dtype=np.uint8))
    with pytest.raises(sdk.BrotligError):
        sdk.DecodeGPU(False, s)
