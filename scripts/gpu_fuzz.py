"""Corruption fuzz of the page kernel ON THE GPU (run by tests/test_gpu_fuzz.py under a host watchdog):
corrupts payload bytes, runs of bytes and page-table entries of corpus streams, decodes them device-resident in batches
and requires: the launch ends (the kernel's hang guard + the caller's time-out), guard bands around every output buffer
and after every input buffer are intact, and every stream ends either bit-exact or with a page status.
usage: gpu_fuzz.py <trials> <seed>"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import brotli_g_sdk_b200 as b  # noqa: E402
from corpus import corner_cases  # noqa: E402

trials, seed = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(seed)
GUARD = 512
names = ["text", "structured_binary", "mixed", "sixteen_pages_lowent", "ring_codes", "long_insert_long_copy", "big_alphabet", "period3",
         "run_structure_fuzz1", "page_32k", "insert_only_split", "four_symbols_skew", "big_inserts_small_copies", "partial_last_page"]
cc = corner_cases()
base = []
for nme in names:
    d, kw = cc[nme]
    d = d[:200000]
    s = b.Encode(d, **kw)
    n = int(s[2]) | int(s[3]) << 8
    if len(s) >= 8 + 4 * n + 24:
        base.append((nme, s, d, n))
dec = b.BrotligDecoder(0)
done = clean = failed_pages = 0
t0 = time.time()
BATCH = 64
while done < trials:
    keep, descs, meta = [], [], []
    for _ in range(min(BATCH, trials - done)):
        nme, s, d, n = base[int(rng.integers(0, len(base)))]
        bad = s.copy()
        mode = int(rng.integers(0, 3))
        if mode == 0:     # flip a few payload bytes
            for _ in range(int(rng.integers(1, 4))):
                bad[int(rng.integers(8 + 4 * n, len(bad)))] ^= int(rng.integers(1, 256))
        elif mode == 1:   # clobber a run
            k = int(rng.integers(8 + 4 * n, len(bad) - 8))
            L = int(rng.integers(1, 64))
            bad[k:k + L] = rng.integers(0, 256, len(bad[k:k + L]), dtype=np.uint8)
        else:             # corrupt a page-table entry
            k = 8 + 4 * int(rng.integers(0, n))
            bad[k:k + 4] = rng.integers(0, 256, 4, dtype=np.uint8)
        cap = ((len(bad) + 15) // 16) * 16
        t_in = torch.full((cap + GUARD,), 0xEE, dtype=torch.uint8, device="cuda")
        t_in[: len(bad)] = torch.from_numpy(bad).cuda()
        t_out = torch.full((len(d) + 2 * GUARD,), 0xEE, dtype=torch.uint8, device="cuda")
        keep.append((t_in, t_out))
        descs.append(dict(d_src=t_in.data_ptr(), src_size=len(bad), src_capacity=cap, d_dst=t_out.data_ptr() + GUARD, dst_capacity=len(d),
                          header=bytes(bad[:16])))
        meta.append((nme, d, cap))
    plan = dec.plan(descs)
    plan.launch()
    bad_pages = plan.finish()
    failed_pages += bad_pages
    for (t_in, t_out), (nme, d, cap) in zip(keep, meta):
        o = t_out.cpu().numpy()
        assert (o[:GUARD] == 0xEE).all() and (o[GUARD + len(d):] == 0xEE).all(), f"{nme}: wrote outside the output buffer"
        assert (t_in[cap:].cpu().numpy() == 0xEE).all(), f"{nme}: input slack was written"
        if np.array_equal(o[GUARD: GUARD + len(d)], d):
            clean += 1
    plan.close()
    done += len(descs)
print(f"GPU FUZZ OK trials {done} bit-exact anyway {clean} failed pages {failed_pages} seconds {time.time() - t0:.1f}", flush=True)
