"""One device-resident decode launch of a chosen payload (for ncu captures).
usage: gpu_prof_one.py <kind> <MiB per stream> <copies> [launches]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import brotli_g_sdk_b200 as b
from brotli_g_sdk_b200 import datagen
kind, mib, copies = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
launches = int(sys.argv[4]) if len(sys.argv) > 4 else 1
gen = {"text": datagen.text_like, "lowent": datagen.low_entropy, "random": datagen.random_bytes, "texture": None,
       "binary": datagen.structured_binary, "mixed": datagen.mixed}[kind]
if kind == "texture":     # a 16 MiB BC3 texture with swizzle + delta pre-conditioning, whatever the size argument says
    from brotli_g_sdk_b200.encoder import DataconditionParams
    data = datagen.bc_texture(1024, 1024, 3, seed=21)
    s = b.Encode(data, dcParams=DataconditionParams(precondition=True, swizzle=True, delta_encode=True, format=3,
                                                    width_blocks=1024, height_blocks=1024))
else:
    data = gen(mib << 20, seed=21)
    s = b.Encode(data)
dec = b.BrotligDecoder(0)
sd, keep = [], []
for c in range(copies):
    t_in = torch.empty(len(s) + 64, dtype=torch.uint8, device="cuda"); t_in[: len(s)] = torch.from_numpy(s).cuda()
    t_out = torch.empty(len(data), dtype=torch.uint8, device="cuda")
    keep.append((t_in, t_out))
    sd.append(dict(d_src=t_in.data_ptr(), src_size=len(s), src_capacity=len(s) + 64, d_dst=t_out.data_ptr(), dst_capacity=len(data), header=bytes(s[:16])))
plan = dec.plan(sd)
ts = torch.cuda.Stream()
torch.cuda.synchronize()
for _ in range(launches):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(ts); plan.launch(ts.cuda_stream); e1.record(ts); torch.cuda.synchronize()
    print(kind, "ms", e0.elapsed_time(e1), "GB/s out", len(data) * copies / e0.elapsed_time(e1) / 1e6)
assert plan.finish() == 0
print("algorithmic_bytes", (len(s) + len(data)) * copies, "compressed", len(s), "uncompressed", len(data), "copies", copies)
print("verify", bool(np.array_equal(keep[-1][1].cpu().numpy(), data)))
