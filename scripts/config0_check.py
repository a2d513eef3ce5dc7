"""BASELINE.json configs[0]: a single 64 KiB page and a 1 MiB buffer (16 pages) of the low-entropy source, decoded by the
reference's CPU DecodeCPU (oracle/_ref, or the oracle port) and by the 1-GPU kernel; asserts both equal the source."""
import ctypes, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import brotli_g_sdk_b200 as b
from brotli_g_sdk_b200 import datagen
sys.path.insert(0, ROOT)
import bench
dec = b.BrotligDecoder(0)
for n in (65536, 1 << 20):
    d = datagen.low_entropy(n, seed=datagen.SEED_CONFIG1)
    s = b.Encode(d, page_size=65536)
    cpu = bench.CpuDecoder([s], [len(d)])
    t_cpu = min(cpu.run(1) for _ in range(5))
    assert np.array_equal(cpu.outs[0][:n], d), "DecodeCPU != source"
    out, ms = dec.decode_host(s)
    best = ms
    for _ in range(5):
        out, ms = dec.decode_host(s)
        best = min(best, ms)
    assert np.array_equal(out, d), "GPU != source"
    assert np.array_equal(out, cpu.outs[0][:n]), "GPU != DecodeCPU"
    print(f"configs[0] {n} bytes ({(n + 65535) // 65536} pages, ratio {n / len(s):.2f}): bit-exact; {cpu.kind} DecodeCPU {t_cpu * 1e3:.3f} ms "
          f"({n / t_cpu / 1e9:.3f} GB/s, {cpu.cores} thread(s)); GPU kernel {best:.4f} ms ({n / best / 1e6:.2f} GB/s)")
