"""The drop-in boundary from the C++ side: a translation unit written against include/brotlig_b200/BrotliG.h (the
reference's names: DecompressedSize / DecodeCPU / DecodeGPU, /root/reference/inc/BrotligDecoder.h:32-33,
sample/BrotligGPUDecoder.h:24) compiles, links libbrotlig_b200.so and -- on the GPU box -- decodes."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "consumer.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "consumer")


def _build():
    from brotli_g_sdk_b200 import build
    build.build_all()
    libdir = os.path.dirname(build.CUDA_LIB)
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(SRC), os.path.getmtime(build.CUDA_LIB)):
        subprocess.run(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), SRC, "-L" + libdir, "-lbrotlig_b200",
                        "-Wl,-rpath," + libdir, "-o", EXE], check=True)
    return EXE


def test_cpp_consumer_compiles_and_links():
    exe = _build()
    out = subprocess.run(["nm", "-u", "--demangle", exe], capture_output=True, text=True).stdout
    for sym in ("DecompressedSize", "DecodeCPU", "DecodeGPU(bool, unsigned int, unsigned char const*, unsigned int*, unsigned char*, double&)"):
        assert sym in out, f"{sym} is not resolved from the library:\n{out}"


@pytest.mark.gpu
def test_cpp_consumer_decodes_and_time_accumulates(tmp_path):
    sys.path.insert(0, ROOT)
    import brotli_g_sdk_b200 as b
    from brotli_g_sdk_b200 import datagen
    exe = _build()
    data = np.concatenate([datagen.text_like(40 * 65536 + 777, seed=91), datagen.random_bytes(3 * 65536, seed=92)])
    s = b.Encode(data)
    sp, dp = str(tmp_path / "s.brotlig"), str(tmp_path / "d.bin")
    s.tofile(sp)
    data.tofile(dp)
    r = subprocess.run([exe, sp, dp], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "CONSUMER OK" in r.stdout, r.stdout + r.stderr
    assert "feedback calls 44" in r.stdout, r.stdout          # one call per page
    # the callback returns true after 5 pages: decoding stops at a page-group boundary, the rest of the output is zero
    r = subprocess.run([exe, sp, dp, "5"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "CONSUMER OK" in r.stdout, r.stdout + r.stderr
