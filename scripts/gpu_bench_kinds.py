"""Device-resident decode throughput of several payload kinds in one process (kernel experiments).
usage: gpu_bench_kinds.py [MiB per stream=16] [copies=32] [kinds=text,binary,mixed,lowent]
       ("texture" is a 16 MiB BC3 texture with swizzle + delta pre-conditioning, whatever the size argument says)
Prints one line: <lib> kind=GB/s ... (median of 5 launches, outputs verified against the source)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import brotli_g_sdk_b200 as b
from brotli_g_sdk_b200 import datagen
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 16
copies = int(sys.argv[2]) if len(sys.argv) > 2 else 32
kinds = (sys.argv[3] if len(sys.argv) > 3 else "text,binary,mixed,lowent").split(",")
gen = {"text": datagen.text_like, "lowent": datagen.low_entropy, "random": datagen.random_bytes,
       "binary": datagen.structured_binary, "mixed": datagen.mixed}
dec = b.BrotligDecoder(0)
res = []
for kind in kinds:
    if kind == "texture":
        from brotli_g_sdk_b200.encoder import DataconditionParams
        data = datagen.bc_texture(1024, 1024, 3, seed=21)
        s = b.Encode(data, dcParams=DataconditionParams(precondition=True, swizzle=True, delta_encode=True, format=3,
                                                         width_blocks=1024, height_blocks=1024))
    else:
        data = gen[kind](mib << 20, seed=21)
        s = b.Encode(data)
    sd, keep = [], []
    for c in range(copies):
        t_in = torch.empty(len(s) + 64, dtype=torch.uint8, device="cuda"); t_in[: len(s)] = torch.from_numpy(s).cuda()
        t_out = torch.empty(len(data), dtype=torch.uint8, device="cuda")
        keep.append((t_in, t_out))
        sd.append(dict(d_src=t_in.data_ptr(), src_size=len(s), src_capacity=len(s) + 64, d_dst=t_out.data_ptr(), dst_capacity=len(data), header=bytes(s[:16])))
    plan = dec.plan(sd)
    ts = torch.cuda.Stream()
    torch.cuda.synchronize()
    times = []
    for i in range(7):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(ts); plan.launch(ts.cuda_stream); e1.record(ts); torch.cuda.synchronize()
        if i >= 2: times.append(e0.elapsed_time(e1))
    assert plan.finish() == 0
    ok = bool(np.array_equal(keep[-1][1].cpu().numpy(), data)) and bool(np.array_equal(keep[0][1].cpu().numpy(), data))
    res.append(f"{kind}={len(data) * copies / float(np.median(times)) / 1e6:.1f}{'' if ok else '(WRONG)'}")
    del plan, keep, sd
print(os.path.basename(os.environ.get("BGX_CUDA_LIB", "default")), " ".join(res), flush=True)
