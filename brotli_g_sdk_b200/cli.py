"""Command line twin of the reference sample (`sample/brotlig_cli.cpp:174-222, 367-650`):

    python -m brotli_g_sdk_b200.cli [options] <file>

  <file>.brotlig given  -> decompress on the GPU (the reference's `-gpu` path; there is no CPU decode here)
  any other file        -> compress with the CPU-side encoder into <file>.brotlig

Options (reference names where they exist):
  -pagesize <bytes>     32768 | 65536 | 131072          (reference: -pagesize)
  -num-repeat <n>       repeat the (de)compression n times and report the average  (reference: -num-repeat)
  -output <path>        output file
  -verbose              print per-stream details
  -precondition -swizzle -delta-encode -data-format <1..5> -texture-width <px> -texture-height <px>
  -num-mip-levels <n> -texture-pitchd3d12aligned        texture pre-conditioning (reference: same names)

Like the reference CLI, the decompression bandwidth it prints is INPUT (compressed) bytes per second of
kernel time (`brotlig_cli.cpp:626-636`); the decompressed GB/s is printed next to it.
"""
from __future__ import annotations

import argparse
import sys
import time

import numpy as np

EXT = ".brotlig"


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="brotlig_b200", add_help=True)
    ap.add_argument("file")
    ap.add_argument("-pagesize", type=int, default=65536)
    ap.add_argument("-num-repeat", dest="num_repeat", type=int, default=1)
    ap.add_argument("-output", default="")
    ap.add_argument("-verbose", action="store_true")
    ap.add_argument("-gpu", action="store_true", help="accepted for compatibility; decompression always runs on the GPU")
    ap.add_argument("-precondition", action="store_true")
    ap.add_argument("-swizzle", action="store_true")
    ap.add_argument("-delta-encode", dest="delta", action="store_true")
    ap.add_argument("-data-format", dest="fmt", type=int, default=0)
    ap.add_argument("-texture-width", dest="width", type=int, default=0)
    ap.add_argument("-texture-height", dest="height", type=int, default=0)
    ap.add_argument("-num-mip-levels", dest="mips", type=int, default=1)
    ap.add_argument("-texture-pitchd3d12aligned", dest="aligned", action="store_true")
    a = ap.parse_args(argv)

    import brotli_g_sdk_b200 as bg
    src = np.fromfile(a.file, dtype=np.uint8)
    if a.file.endswith(EXT):
        dst_path = a.output or a.file[: -len(EXT)]
        n = bg.DecompressedSize(src)
        out = np.empty(n, dtype=np.uint8)
        kernel_ms = 0.0
        t0 = time.perf_counter()
        for rep in range(a.num_repeat):
            print(f"Round {rep + 1} of {a.num_repeat}")
            out, ms = bg.DecodeGPU(False, src, out)
            kernel_ms += ms
        wall = (time.perf_counter() - t0) / a.num_repeat
        kernel_ms /= a.num_repeat
        out.tofile(dst_path)
        print(f"Saving decompressed file {dst_path}")
        print(f"BrotliG GPU decompressor: {len(src)} -> {n} bytes")
        print(f"Processed in {kernel_ms:.3f} ms (kernel), {wall * 1e3:.3f} ms (wall incl. PCIe)")
        if kernel_ms > 0:
            print(f"Bandwidth {len(src) / kernel_ms / 2**30 * 1e3:.3f} GiB/s of input, {n / kernel_ms / 1e6:.3f} GB/s decompressed")
        return 0

    dst_path = a.output or a.file + EXT
    dc = None
    if a.precondition:
        dc = bg.DataconditionParams(precondition=True, swizzle=a.swizzle, delta_encode=a.delta, format=a.fmt,
                                    width_blocks=(a.width + 3) // 4, height_blocks=(a.height + 3) // 4, num_mips=a.mips,
                                    pitch_aligned=a.aligned)
    t0 = time.perf_counter()
    for rep in range(a.num_repeat):
        print(f"Round {rep + 1} of {a.num_repeat}")
        stream = bg.Encode(src, page_size=a.pagesize, dcParams=dc)
    dt = (time.perf_counter() - t0) / a.num_repeat
    stream.tofile(dst_path)
    print(f"Saving compressed file {dst_path}")
    print(f"BrotliG CPU compressor: {len(src)} -> {len(stream)} bytes, ratio {len(src) / max(1, len(stream)):.3f}")
    print(f"Processed in {dt * 1e3:.3f} ms, {len(src) / dt / 2**30:.3f} GiB/s")
    if a.verbose:
        from brotli_g_sdk_b200 import encoder
        print(encoder.last_stats())
    return 0


if __name__ == "__main__":
    sys.exit(main())
