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


def test_host_pipeline_segments_tile_the_stream(sdk, emulator):
    """bgx_decode_batch_host pipelines over page ranges of large streams (host_plan.h): the ranges must tile the
    pages, the uploads must tile the stream and reach the end of each range's last page, the downloads must tile
    the output -- and a corrupt page table must not push any range outside the buffers"""
    import numpy as np
    from brotli_g_sdk_b200 import datagen
    data = np.concatenate([datagen.text_like(400000, seed=71), datagen.random_bytes(200000, seed=72), datagen.low_entropy(123456, seed=73)])
    s = sdk.Encode(data)
    n_pages = int(s[2]) | (int(s[3]) << 8)
    table = np.frombuffer(s[8: 8 + 4 * n_pages].tobytes(), dtype="<u4")
    table_end = 8 + 4 * n_pages
    page_end = lambda p: table_end + (int(table[p]) if p < n_pages else len(s) - table_end)   # end of page p - 1
    assert emulator.plan_segments(s, 1 << 30) == [(0, 0, 0, len(s), 0, len(data))]      # small against the target: whole
    for target in (64 << 10, 200 << 10, 300000):
        seg = emulator.plan_segments(s, target)
        assert len(seg) >= 2
        pages = up = dn = 0
        for pb, pc, up0, up1, dn0, dn1 in seg:
            assert pb == pages and pc > 0 and up0 == up and dn0 == dn and up1 >= up0 and dn1 > dn0
            pages += pc
            assert up1 >= (len(s) if pages == n_pages else page_end(pages)), "range's last page not uploaded"
            assert dn1 == min(pages * 65536, len(data))
            up, dn = up1, dn1
        assert pages == n_pages and up == len(s) and dn == len(data)
    rng = np.random.default_rng(9)
    for trial in range(50):
        bad = s.copy()
        k = 8 + 4 * int(rng.integers(0, n_pages))
        bad[k: k + 4] = rng.integers(0, 256, 4, dtype=np.uint8)
        up = 0
        for pb, pc, up0, up1, dn0, dn1 in emulator.plan_segments(bad, 100 << 10):
            assert up0 == up and up0 <= up1 <= len(bad) and dn0 < dn1 <= len(data)
            up = up1
        assert up == len(bad)
