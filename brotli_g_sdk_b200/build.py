"""Builds the native libraries of brotli_g_sdk_b200 in-tree.

  libbrotlig_b200.so      CUDA kernels (sm_100a) + C ABI (include/brotlig_b200.h) + C++ shim (BrotliG::*)
  libbrotlig_b200_enc.so  CPU-side stream encoder (include/brotlig_b200_encoder.h); no CUDA dependency

nvcc cross-compiles without a GPU, so this runs on the CPU-only development box; the built .so files
travel to the GPU box with the repository snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
CUDA_LIB = os.path.join(HERE, "libbrotlig_b200.so")
ENC_LIB = os.path.join(HERE, "libbrotlig_b200_enc.so")

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC", "-shared",
]


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _sources(names: list[str]) -> list[str]:
    return [os.path.join(CSRC, n) for n in names]


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found; the CUDA library cannot be built")
    return p


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    srcs = _sources(["bgx_cuda.cu", "brotlig_api.cpp"])
    deps = srcs + _sources(["page_decode.cuh", "bgx_format.h", "host_plan.h"]) + [
        os.path.join(HERE, "..", "include", "brotlig_b200.h"),
        os.path.join(HERE, "..", "include", "brotlig_b200", "BrotliG.h"),
    ]
    if not force and _newer(CUDA_LIB, deps):
        return CUDA_LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + srcs + ["-o", CUDA_LIB]
    subprocess.run(cmd, check=True)
    return CUDA_LIB


def build_encoder(force: bool = False) -> str:
    srcs = _sources(["bgx_encoder.cpp"])
    deps = srcs + _sources(["bgx_format.h"]) + [os.path.join(HERE, "..", "include", "brotlig_b200_encoder.h")]
    if not force and _newer(ENC_LIB, deps):
        return ENC_LIB
    cxx = shutil.which("g++") or "g++"
    subprocess.run([cxx, "-std=c++17", "-O2", "-fPIC", "-shared", "-pthread"] + srcs + ["-o", ENC_LIB], check=True)
    return ENC_LIB


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_encoder(force)
    build_cuda(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built:", CUDA_LIB, ENC_LIB)
