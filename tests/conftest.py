"""pytest configuration: the `gpu` marker and shared helpers.

CPU-side checkers used here (TEST INFRASTRUCTURE, never on the product path):
  oracle/_build/libbrotlig_oracle.so   our plain-C restatement of the reference decoder
  oracle/_ref/libbrotlig_ref.so        the UNMODIFIED reference decoder built from /root/reference (optional:
                                       only where it was built; it travels to the GPU box as a prebuilt file)
  tests/emul/libbgx_emul.so            the CUDA device code compiled for the CPU warp emulator
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _run(cmd, cwd=None):
    subprocess.run(cmd, check=True, cwd=cwd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)


class Oracle:
    """ctypes face of oracle/brotlig_oracle.c"""

    class Stats(ctypes.Structure):
        _fields_ = [(n, ctypes.c_uint64) for n in ("pages", "raw_pages", "rounds", "commands", "literals_emitted", "literals_decoded")] + [
            ("dist_code_hist", ctypes.c_uint64 * 16), ("implicit_dist0", ctypes.c_uint64), ("insert_only", ctypes.c_uint64),
            ("overlap_copies", ctypes.c_uint64), ("table_types", (ctypes.c_uint64 * 3) * 3), ("rle16", ctypes.c_uint64),
            ("rle17", ctypes.c_uint64), ("max_insert_len", ctypes.c_uint64), ("max_copy_len", ctypes.c_uint64),
            ("delta_pages", ctypes.c_uint64)]

    def __init__(self):
        so = os.path.join(ROOT, "oracle", "_build", "libbrotlig_oracle.so")
        src = os.path.join(ROOT, "oracle", "brotlig_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            _run(["make", "oracle"], cwd=os.path.join(ROOT, "oracle"))
        self.lib = ctypes.CDLL(so)
        self.lib.bgo_decompressed_size.restype = ctypes.c_uint32
        self.lib.bgo_decompressed_size.argtypes = [ctypes.c_void_p]
        self.lib.bgo_decode.restype = ctypes.c_int
        self.lib.bgo_decode.argtypes = [ctypes.c_uint32, ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint32), ctypes.c_void_p]

    def size(self, stream: np.ndarray) -> int:
        return int(self.lib.bgo_decompressed_size(stream.ctypes.data))

    def decode(self, stream: np.ndarray, expect_rc: int = 0) -> np.ndarray:
        s = np.concatenate([np.ascontiguousarray(stream, dtype=np.uint8), np.zeros(16, np.uint8)])   # over-read slack
        n = self.size(s)
        out = np.full(n + 16, 0xA5, dtype=np.uint8)
        osz = ctypes.c_uint32(n)
        rc = self.lib.bgo_decode(len(stream), s.ctypes.data, ctypes.byref(osz), out.ctypes.data)
        assert rc == expect_rc, f"oracle rc {rc}"
        assert (out[n:] == 0xA5).all(), "oracle wrote past the output"
        return out[: osz.value]

    def stats(self) -> "Oracle.Stats":
        st = Oracle.Stats()
        self.lib.bgo_last_stats(ctypes.byref(st))
        return st


class Reference:
    """the unmodified reference DecodeCPU (oracle/_ref), loaded RTLD_LOCAL (its symbols clash with ours)"""

    def __init__(self, path):
        self.lib = ctypes.CDLL(path, mode=ctypes.RTLD_LOCAL)
        self.lib.DecompressedSize.restype = ctypes.c_uint32
        self.lib.DecompressedSize.argtypes = [ctypes.c_void_p]
        self.lib.DecodeCPU.restype = ctypes.c_int
        self.lib.DecodeCPU.argtypes = [ctypes.c_uint32, ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint32), ctypes.c_void_p, ctypes.c_void_p]

    def decode(self, stream: np.ndarray) -> np.ndarray:
        s = np.concatenate([np.ascontiguousarray(stream, dtype=np.uint8), np.zeros(16, np.uint8)])
        n = int(self.lib.DecompressedSize(s.ctypes.data))
        out = np.full(n + 16, 0xA5, dtype=np.uint8)
        osz = ctypes.c_uint32(n)
        rc = self.lib.DecodeCPU(len(stream), s.ctypes.data, ctypes.byref(osz), out.ctypes.data, None)
        assert rc == 0, f"reference rc {rc}"
        return out[: osz.value]


class Emulator:
    """brotli_g_sdk_b200/csrc/page_decode.cuh compiled with g++ against tests/emul/warp_emul.h"""

    def __init__(self):
        d = os.path.join(ROOT, "tests", "emul")
        extra = os.environ.get("BGX_EMUL_FLAGS", "").split()     # e.g. -DBGX_PIECE_PREFETCH: emulate a kernel variant
        so = os.path.join(d, "libbgx_emul%s.so" % ("_" + "".join(c for c in "".join(extra) if c.isalnum()) if extra else ""))
        deps = [os.path.join(d, "emul_decode.cpp"), os.path.join(d, "warp_emul.h"),
                os.path.join(ROOT, "brotli_g_sdk_b200", "csrc", "page_decode.cuh"),
                os.path.join(ROOT, "brotli_g_sdk_b200", "csrc", "bgx_format.h"),
                os.path.join(ROOT, "brotli_g_sdk_b200", "csrc", "host_plan.h")]
        if not os.path.exists(so) or any(os.path.getmtime(x) > os.path.getmtime(so) for x in deps):
            _run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-w", "-I" + d] + extra + [deps[0], "-o", so])
        self.lib = ctypes.CDLL(so)
        self.lib.emul_decode_stream.restype = ctypes.c_int
        self.lib.emul_plan_segments.restype = ctypes.c_int

    def plan_segments(self, stream: np.ndarray, target: int):
        """host_plan.h plan_stream_segments for one stream: list of (page_begin, page_count, up0, up1, dn0, dn1)"""
        s = np.ascontiguousarray(stream, dtype=np.uint8)
        out = (ctypes.c_uint64 * (7 * 4096))()
        n = self.lib.emul_plan_segments(ctypes.c_void_p(s.ctypes.data), ctypes.c_uint32(len(s)), ctypes.c_uint64(target), out, 4096)
        assert n >= 0
        return [tuple(int(out[7 * k + j]) for j in range(1, 7)) for k in range(n)]

    def decode(self, stream: np.ndarray, expect_rc: int = 0, dst_offset: int = 0):
        s = np.ascontiguousarray(stream, dtype=np.uint8)
        n = (int(s[2]) | (int(s[3]) << 8))
        w1 = int(s[4]) | (int(s[5]) << 8) | (int(s[6]) << 16) | (int(s[7]) << 24)
        ps = 32768 << (w1 & 3)
        last = (w1 >> 2) & 0x3FFFF
        size = n * ps - ((ps - last) if last else 0)
        base = np.full(size + 64 + 32, 0xEE, dtype=np.uint8)
        skew = (-base.ctypes.data) % 16 + dst_offset          # output starts `dst_offset` bytes past a 16-byte boundary
        out = base[skew: skew + size + 64]
        st = (ctypes.c_uint32 * max(n, 1))()
        fl = (ctypes.c_uint32 * max(n, 1))()
        coll = ctypes.c_uint64(0)
        rc = self.lib.emul_decode_stream(ctypes.c_void_p(s.ctypes.data), ctypes.c_uint32(len(s)), ctypes.c_void_p(out.ctypes.data),
                                         ctypes.c_uint32(size), st, fl, ctypes.byref(coll))
        assert rc == expect_rc, f"emulator rc {rc} status {list(st)[:8]}"
        assert (out[size:] == 0xEE).all(), "emulated kernel wrote past the page"
        return out[:size], list(st), list(fl)


@pytest.fixture(scope="session")
def oracle():
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    p = os.path.join(ROOT, "oracle", "_ref", "libbrotlig_ref.so")
    if not os.path.exists(p):
        if os.path.isdir("/root/reference"):
            _run(["make", "ref"], cwd=os.path.join(ROOT, "oracle"))
        else:
            pytest.skip("oracle/_ref not built (needs /root/reference)")
    return Reference(p)


@pytest.fixture(scope="session")
def emulator():
    return Emulator()


@pytest.fixture(scope="session")
def sdk():
    from brotli_g_sdk_b200 import build
    build.build_encoder()
    import brotli_g_sdk_b200 as b
    return b


def sha256(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
