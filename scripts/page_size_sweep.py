"""BASELINE configs[4]: page-size sweep on the mixed-entropy payload, one GPU.
32/64/128 KiB via the stream header's PageSizeIdx; 4/8/16 KiB "pages" as single-page streams
(NumPages = 1, LastPageSize = n) batched by the thousands. Writes gpurun_out/page_size_sweep.json.
usage: page_size_sweep.py [total MiB per size=256] [kind=mixed] [sizes=4096,...,131072]
SWEEP_SINGLE=1: exactly one launch per size and no timing loop (for `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum
-k regex:bgx_decode_pages`: launch i of the capture is size i of the list)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import brotli_g_sdk_b200 as bg
from brotli_g_sdk_b200 import datagen

total = int(sys.argv[1]) << 20 if len(sys.argv) > 1 else 256 << 20
dec = bg.BrotligDecoder(0)
kind = sys.argv[2] if len(sys.argv) > 2 else "mixed"
sizes = tuple(int(x) for x in sys.argv[3].split(",")) if len(sys.argv) > 3 else (4096, 8192, 16384, 32768, 65536, 131072)
gen = {"mixed": datagen.mixed, "text": datagen.text_like, "binary": datagen.structured_binary, "lowent": datagen.low_entropy}[kind]
data = gen(64 << 20, seed=datagen.SEED_CONFIG4)
res = {}
for ps in sizes:
    if ps >= 32768:
        streams = [bg.Encode(data, page_size=ps)]
        usizes = [len(data)]
    else:
        n = (16 << 20) // ps     # 16 MiB of unique payload cut into single-page streams
        streams = [bg.Encode(data[i * ps:(i + 1) * ps], page_size=32768) for i in range(n)]
        usizes = [ps] * n
    reps = max(1, total // sum(usizes))
    blob_in = np.concatenate([np.concatenate([s, np.zeros((-len(s)) % 256 + 256, np.uint8)]) for s in streams])
    offs = np.cumsum([0] + [len(s) + ((-len(s)) % 256 + 256) for s in streams])
    descs, keep = [], []
    for r in range(reps):
        t_in = torch.from_numpy(blob_in).cuda()
        t_out = torch.empty(sum(usizes), dtype=torch.uint8, device="cuda")
        keep.append((t_in, t_out))
        o = 0
        for i, (s, u) in enumerate(zip(streams, usizes)):
            descs.append(dict(d_src=t_in.data_ptr() + int(offs[i]), src_size=len(s), src_capacity=int(offs[i + 1] - offs[i]),
                              d_dst=t_out.data_ptr() + o, dst_capacity=u, header=bytes(s[:16])))
            o += u
    plan = dec.plan(descs)
    ts = torch.cuda.Stream()
    torch.cuda.synchronize()
    single = os.environ.get("SWEEP_SINGLE") == "1"
    for _ in range(0 if single else 3):
        plan.launch(ts.cuda_stream)
    assert plan.finish() == 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 1 if single else 5
    e0.record(ts)
    for _ in range(K):
        plan.launch(ts.cuda_stream)
    e1.record(ts)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    out = keep[0][1].cpu().numpy()
    ok = bool(np.array_equal(out, data[: len(out)]))
    cbytes = sum(len(s) for s in streams) * reps
    ubytes = sum(usizes) * reps
    res[str(ps)] = {"page_bytes": ps, "streams": len(descs), "decompressed_GBps": ubytes / ms / 1e6, "ratio": ubytes / cbytes,
                    "algorithmic_GBps": (ubytes + cbytes) / ms / 1e6, "ms": ms, "bit_exact": ok}
    print(ps, res[str(ps)], flush=True)
    del keep, plan
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/page_size_sweep%s.json" % ("_single" if os.environ.get("SWEEP_SINGLE") == "1" else ""), "w"), indent=1)
