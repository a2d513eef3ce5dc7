"""First-contact GPU check (run under gpurun): correctness on a spread of streams through the C ABI,
then rough device-resident throughput. Writes gpurun_out/quick.json."""
import ctypes, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import brotli_g_sdk_b200 as b
from brotli_g_sdk_b200 import datagen

os.makedirs("gpurun_out", exist_ok=True)
res = {"device": torch.cuda.get_device_name(0), "host_cpus": os.cpu_count()}
dec = b.BrotligDecoder(0)
rng = np.random.default_rng(3)
cases = {
    "lowent_1page": datagen.low_entropy(65536, seed=1),
    "lowent_1MiB": datagen.low_entropy(1 << 20, seed=2),
    "text_3MiB": datagen.text_like(3 << 20, seed=3),
    "binary_2MiB": datagen.structured_binary(2 << 20, seed=4),
    "random_1MiB": datagen.random_bytes(1 << 20, seed=5),
    "const": np.full(300000, 7, np.uint8),
    "period3": np.tile(np.array([1, 2, 3], np.uint8), 50000),
    "tiny": datagen.low_entropy(1000, seed=6),
    "partial_last": datagen.low_entropy(65536 * 2 + 12345, seed=7),
}
ok_all = True
for name, data in cases.items():
    for ps in (32768, 65536, 131072):
        s = b.Encode(data, page_size=ps)
        out, ms = dec.decode_host(s)
        ok = bool(np.array_equal(out, data))
        ok_all &= ok
        res[f"{name}_ps{ps}"] = {"ok": ok, "ratio": round(len(data) / len(s), 3), "kernel_ms": round(ms, 4)}
        if not ok:
            bad = np.nonzero(out != data)[0]
            res[f"{name}_ps{ps}"]["first_bad"] = int(bad[0]) if len(bad) else -1
            res[f"{name}_ps{ps}"]["nbad"] = int(len(bad))
res["all_ok"] = ok_all
print("correctness:", ok_all, flush=True)

def bench_resident(name, unique_streams, copies, iters=5):
    """streams resident in HBM, `copies` distinct copies of each unique stream, kernel-only timing"""
    plans = []
    total_out = 0
    total_in = 0
    sd = []
    keep = []
    for s, usize in unique_streams:
        for c in range(copies):
            t_in = torch.empty(len(s) + 64, dtype=torch.uint8, device="cuda")
            t_in[: len(s)] = torch.from_numpy(s).cuda()
            t_out = torch.empty(usize, dtype=torch.uint8, device="cuda")
            keep.append((t_in, t_out))
            sd.append(dict(d_src=t_in.data_ptr(), src_size=len(s), src_capacity=len(s) + 64, d_dst=t_out.data_ptr(),
                           dst_capacity=usize, header=bytes(s[:16])))
            total_out += usize
            total_in += len(s)
    plan = dec.plan(sd)
    ts = torch.cuda.Stream()
    st = ts.cuda_stream
    torch.cuda.synchronize()
    for _ in range(2):
        plan.launch(st)
    assert plan.finish() == 0
    times = []
    for _ in range(iters):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(ts); plan.launch(st); e1.record(ts); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    assert plan.finish() == 0
    # verify one copy of each
    t = float(np.median(times))
    res[name] = {"out_GB": total_out / 1e9, "in_GB": total_in / 1e9, "ms": round(t, 3),
                 "decompressed_GBps": round(total_out / t / 1e6, 1), "algorithmic_GBps": round((total_in + total_out) / t / 1e6, 1),
                 "info": plan.info, "times": [round(x, 3) for x in times]}
    print(name, res[name], flush=True)
    return keep

t0 = time.time()
txt = datagen.text_like(32 << 20, seed=11)
s_txt = b.Encode(txt)
res["encode_text_MBps"] = round(len(txt) / (time.time() - t0) / 1e6, 1)
keep = bench_resident("text_32MiBx16", [(s_txt, len(txt))], 16)
o = keep[0][1].cpu().numpy(); res["text_verify"] = bool(np.array_equal(o, txt)); del keep
le = datagen.low_entropy(32 << 20, seed=12); s_le = b.Encode(le)
keep = bench_resident("lowent_32MiBx16", [(s_le, len(le))], 16)
res["lowent_verify"] = bool(np.array_equal(keep[3][1].cpu().numpy(), le)); del keep
rb = datagen.random_bytes(64 << 20, seed=13); s_rb = b.Encode(rb)
keep = bench_resident("random_64MiBx32", [(s_rb, len(rb))], 32)
res["random_verify"] = bool(np.array_equal(keep[5][1].cpu().numpy(), rb)); del keep
sb = datagen.structured_binary(32 << 20, seed=14); s_sb = b.Encode(sb)
keep = bench_resident("binary_32MiBx16", [(s_sb, len(sb))], 16)
res["binary_verify"] = bool(np.array_equal(keep[1][1].cpu().numpy(), sb)); del keep
json.dump(res, open("gpurun_out/quick.json", "w"), indent=1)
print("done")
