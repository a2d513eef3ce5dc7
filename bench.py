#!/usr/bin/env python
"""bench.py -- decompressed GB/s (bit-exact) on 64 KiB-page Brotli-G streams, % of HBM roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload mixed|random|texture|...]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the decode hot path over one batch of synthetic streams. The default workload is
BASELINE.json configs[3]: a 16 GiB mixed-entropy batch (50 % text, 25 % structured binary, 25 % random),
page_size 65536, 256 streams x 64 MiB (16 unique streams x 16 replicas in distinct HBM buffers), whose
streams are partitioned over the N GPUs of the run (STRONG scaling: the batch is fixed, every rank decodes
its share, no data-path collective). At N = 1 it is the largest single-GPU configuration of BASELINE.json.

  value      whole-job decompressed GB/s, streams already resident in HBM, CUDA events around K launches, max over ranks
  e2e        the same metric through the C-ABI host-pointer call (bgx_decode_batch_host): pinned host buffers,
             H2D + decode + D2H inside the timed region; pcie_ceiling = the same bytes moved by plain
             cudaMemcpyAsync in both directions at once (what the bus of this box allows)
  roofline   algorithmic bytes (compressed read + decompressed written) / kernel time vs the measured HBM peak
  cpu_baseline  the reference's own DecodeCPU (oracle/_ref, unmodified, built from /root/reference) on the box's host
             cores over a bounded sample of the same streams
  secondary  (N = 1) BASELINE.json configs[1]: 4 GiB of random bytes -- every page is stored raw, so this is a memcpy
             upper bound of the page kernel, not a decoder number; secondary_texture: configs[2], 4 GiB of pre-conditioned
             BC3 textures (page kernel + de-conditioning kernel)
  sharded    (N > 1) SURVEY section 8e level 2: single 64 MiB streams living on one rank, replicated with ONE NCCL
             broadcast and decoded by page range on every rank; reported with and without the broadcast time
--impl reference times the reference CPU decoder alone on the same config (rank 0 only).
"""
from __future__ import annotations

import argparse
import concurrent.futures
import ctypes
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

STREAM_BYTES = 64 << 20
PAGE = 65536
CACHE_DIR = os.environ.get("BGX_BENCH_CACHE", "/tmp/bgx_bench_cache")
METRIC = "decompressed GB/s (bit-exact) on 64 KiB-page streams; % of HBM roofline"


# ------------------------------------------------------------------------------------------- helpers
def measured_peak_gbs() -> tuple[float, str]:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """samples SM clock / throttle reasons through NVML (a few hundred Hz) while the timed region runs;
    falls back to polling nvidia-smi when pynvml is unavailable"""

    def __init__(self, index: int):
        self.index = index
        self.samples: list[tuple[float, int]] = []
        self.max_mhz = None
        self.power = []
        self._stop = threading.Event()
        self._t = None
        self._nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None
        self._t = threading.Thread(target=self._run_nvml if self._nvml else self._run_smi, daemon=True)
        self._t.start()

    def _run_nvml(self):
        n = self._nvml
        while not self._stop.is_set():
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM))
                try:
                    reasons = int(n.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                except Exception:
                    reasons = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                self.samples.append((mhz, reasons))
                self.power.append(n.nvmlDeviceGetPowerUsage(self._h) / 1000.0)
            except Exception:
                pass
            time.sleep(0.002)

    def _run_smi(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active"
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=2).stdout.strip().split(",")
                self.samples.append((float(out[0]), int(out[2].strip(), 16)))
                self.max_mhz = float(out[1])
            except Exception:
                pass

    def stop(self) -> dict:
        self._stop.set()
        if self._t:
            self._t.join(timeout=3)
        bits = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        sm = [m for m, _ in self.samples]
        reasons = sorted({name for _, r in self.samples for bit, name in bits.items() if r & bit})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(sm), "power_w_max": max(self.power) if self.power else None,
                "source": "nvml" if self._nvml else "nvidia-smi"}


# ------------------------------------------------------------------------------------------- workloads
WORKLOADS = {
    "mixed": "configs[3]: {gib} GiB mixed-entropy (50% text, 25% structured binary, 25% random) multi-stream batch, "
             "{n} streams x 64 MiB, page_size=65536, streams partitioned over the GPUs",
    "random": "configs[1]: {gib} GiB random-byte buffer, page_size=65536 ({n} streams x 64 MiB, all pages raw: memcpy upper bound)",
    "texture": "configs[2]: {gib} GiB of BC3 textures, 1024x1024 blocks (16 MiB) each, pre-conditioned (swizzle + delta), page_size=65536",
    "text": "{gib} GiB text-like (Zipf word model), {n} streams x 64 MiB, page_size=65536",
    "binary": "{gib} GiB structured binary (fixed-width records), {n} streams x 64 MiB, page_size=65536",
    "lowent": "{gib} GiB low-entropy (configs[0] source model), {n} streams x 64 MiB, page_size=65536",
}
DEFAULT_GIB = {"mixed": 16, "random": 4, "texture": 4, "text": 4, "binary": 4, "lowent": 4}


def _code_tag() -> str:
    """the cache is only valid for the generator + encoder sources it was made with"""
    h = hashlib.sha256()
    for rel in ("brotli_g_sdk_b200/datagen.py", "brotli_g_sdk_b200/csrc/bgx_encoder.cpp", "brotli_g_sdk_b200/encoder.py"):
        try:
            h.update(open(os.path.join(ROOT, rel), "rb").read())
        except OSError:
            pass
    return h.hexdigest()[:10]


def _gen_one(kind: str, seed: int, sdk, threads: int):
    """one unique stream of the workload: (stream bytes, source bytes); cached on disk (the reference arm and the
    GPU arm of one driver run, and the ranks of one torchrun, share the encoder's work)"""
    from brotli_g_sdk_b200 import datagen
    tag = f"{kind}_{seed:x}_{STREAM_BYTES}_{_code_tag()}"
    ps, pd = os.path.join(CACHE_DIR, tag + ".brotlig"), os.path.join(CACHE_DIR, tag + ".src")
    try:
        if kind != "random" and os.path.exists(ps) and os.path.exists(pd):
            s, d = np.fromfile(ps, dtype=np.uint8), np.fromfile(pd, dtype=np.uint8)
            if len(d) and len(s) > 16:
                return s, d
    except Exception:
        pass
    if kind == "texture":
        from brotli_g_sdk_b200.encoder import DataconditionParams
        p = DataconditionParams(precondition=True, swizzle=True, delta_encode=True, format=3, width_blocks=1024, height_blocks=1024)
        d = datagen.bc_texture(1024, 1024, 3, seed=seed)
        s = sdk.Encode(d, page_size=PAGE, dcParams=p, num_threads=threads)
    else:
        gen = {"mixed": datagen.mixed, "text": datagen.text_like, "binary": datagen.structured_binary, "lowent": datagen.low_entropy,
               "random": datagen.random_bytes}[kind]
        d = gen(STREAM_BYTES, seed=seed)
        s = sdk.Encode(d, page_size=PAGE, num_threads=threads)
    try:
        if kind == "random":
            return s, d
        os.makedirs(CACHE_DIR, exist_ok=True)
        for path, arr in ((ps, s), (pd, d)):
            tmp = f"{path}.{os.getpid()}.tmp"
            arr.tofile(tmp)
            os.replace(tmp, path)
    except Exception:
        pass
    return s, d


def unique_streams(kind: str, n_unique: int, sdk, seed0: int):
    """the unique streams of a workload, generated and encoded by a few threads (numpy and the encoder release the GIL)"""
    cores = os.cpu_count() or 8
    workers = max(1, min(n_unique, cores // 2))
    with concurrent.futures.ThreadPoolExecutor(workers) as ex:
        res = list(ex.map(lambda i: _gen_one(kind, seed0 + i, sdk, max(1, cores // workers)), range(n_unique)))
    return [r[0] for r in res], [r[1] for r in res]


def workload_plan(kind: str, total_bytes: int, n_unique: int):
    """(number of streams, unique streams, uncompressed bytes per stream)"""
    per = 1024 * 1024 * 16 if kind == "texture" else STREAM_BYTES
    n = max(1, total_bytes // per)
    return n, min(n, n_unique), per


def workload_config(args, world: int):
    n, uniq, per = workload_plan(args.workload, args.size_gib << 30, args.unique)
    return {"workload": WORKLOADS[args.workload].format(gib=args.size_gib, n=n), "total_bytes": args.size_gib << 30,
            "page_size": PAGE, "streams": n, "stream_bytes": per, "unique_streams": uniq, "replicas": (n + uniq - 1) // uniq,
            "l2_policy": "inputs+outputs per step (>= 2x payload) are far larger than the 126 MB L2; no flush needed",
            "parallelism": f"streams partitioned over {world} GPU(s) (size balanced), one process per GPU, no data-path collective"}


# ------------------------------------------------------------------------------------------- CPU reference
def load_reference_cpu():
    p = os.path.join(ROOT, "oracle", "_ref", "libbrotlig_ref.so")
    if os.path.exists(p):
        lib = ctypes.CDLL(p, mode=ctypes.RTLD_LOCAL)
        lib.DecodeCPU.restype = ctypes.c_int
        lib.DecodeCPU.argtypes = [ctypes.c_uint32, ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint32), ctypes.c_void_p, ctypes.c_void_p]
        return lib, "reference"
    # oracle port (single-threaded restatement) -- only when the reference could not be built
    so = os.path.join(ROOT, "oracle", "_build", "libbrotlig_oracle.so")
    if not os.path.exists(so):
        subprocess.run(["make", "oracle"], cwd=os.path.join(ROOT, "oracle"), check=True, stdout=subprocess.DEVNULL)
    lib = ctypes.CDLL(so)
    lib.bgo_decode.restype = ctypes.c_int
    return lib, "port"


class CpuDecoder:
    """the reference DecodeCPU (all the host threads it chooses to use) over a list of streams"""

    def __init__(self, streams, usizes):
        self.lib, self.kind = load_reference_cpu()
        self.streams, self.usizes = streams, usizes
        self.padded = [np.concatenate([s, np.zeros(16, np.uint8)]) for s in streams]
        self.outs = [np.zeros(u + 16, np.uint8) for u in usizes]      # pre-faulted
        hw = os.cpu_count() or 1
        workers = min(128, hw)
        self.cores = workers if (self.kind == "reference" and (usizes[0] // PAGE) > 2 * workers) else 1
        self.hw = hw

    def run(self, k: int) -> float:
        t0 = time.perf_counter()
        for i in range(k):
            n = ctypes.c_uint32(self.usizes[i])
            if self.kind == "reference":
                rc = self.lib.DecodeCPU(len(self.streams[i]), self.padded[i].ctypes.data, ctypes.byref(n), self.outs[i].ctypes.data, None)
            else:
                rc = self.lib.bgo_decode(ctypes.c_uint32(len(self.streams[i])), ctypes.c_void_p(self.padded[i].ctypes.data), ctypes.byref(n),
                                         ctypes.c_void_p(self.outs[i].ctypes.data))
            assert rc == 0
        return time.perf_counter() - t0

    def sample(self, budget_s: float, repeats: int = 3) -> dict:
        t1 = self.run(1)
        k = int(max(1, min(len(self.streams), budget_s / repeats / max(t1, 1e-6))))
        t = statistics.median([self.run(k) for _ in range(repeats)])
        nbytes = sum(self.usizes[:k])
        return {"value": nbytes / t / 1e9, "unit": "GB/s decompressed", "cores": self.cores, "kind": self.kind,
                "sample": f"{k} of the {len(self.streams)} unique streams ({nbytes / 2**20:.0f} MiB, {nbytes // PAGE} pages), median of {repeats}",
                "host_threads_available": self.hw, "k": k}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    import brotli_g_sdk_b200 as sdk
    from brotli_g_sdk_b200 import build, datagen
    build.build_encoder()
    _, uniq, per = workload_plan(args.workload, args.size_gib << 30, args.unique)
    streams, sources = unique_streams(args.workload, uniq, sdk, datagen.SEED_CONFIG4)
    cpu = CpuDecoder(streams, [len(d) for d in sources])
    # each step decodes a bounded sample of the workload: k of its unique streams
    t1 = cpu.run(1)
    budget = max(0.5, min(8.0, 120.0 / max(1, args.steps + args.warmup)))
    k = int(max(1, min(len(streams), budget / max(t1, 1e-6))))
    for _ in range(args.warmup):
        cpu.run(k)
    times = [cpu.run(k) for _ in range(args.steps)]
    assert np.array_equal(cpu.outs[0][: len(sources[0])], sources[0]), "reference output != source"
    nbytes = sum(cpu.usizes[:k])
    t = statistics.median(times)
    v = nbytes / t / 1e9
    sample = f"each step = {k} of the {len(streams)} unique streams ({nbytes / 2**20:.0f} MiB, {nbytes // PAGE} pages) through DecodeCPU, median of {args.steps}"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": v, "unit": "GB/s", "cores": cpu.cores, "kind": cpu.kind, "sample": sample, "host_threads_available": cpu.hw},
        "e2e": {"value": v, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; brotli_g_sdk_b200 has no CPU decode path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    # Everything that is not the JSON line goes to stderr: NCCL prints its version banner on stdout when the
    # first communicator is created, and the contract is ONE JSON line on stdout.
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import brotli_g_sdk_b200 as sdk
    from brotli_g_sdk_b200 import build, datagen
    from brotli_g_sdk_b200.multi_gpu import cuda_decode_fn, decode_sharded_stream, partition_streams
    if rank == 0:
        build.build_all()
    if world > 1:
        dist.barrier()
    dec = sdk.BrotligDecoder(local)
    dev = torch.device("cuda", local)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def get_unique(kind, total_bytes, n_unique, seed0):
        """the unique streams of a workload (rank 0 encodes, the others then read the disk cache)"""
        _, uniq, _ = workload_plan(kind, total_bytes, n_unique)
        if rank == 0:
            out = unique_streams(kind, uniq, sdk, seed0)
        if world > 1:
            dist.barrier()
        if rank != 0:
            out = unique_streams(kind, uniq, sdk, seed0)
        return out

    def share_of(kind, total_bytes, streams):
        """this rank's share of a batch of total_bytes: indices into the unique streams (stream i of the batch is a
        replica of unique stream i % unique; whole streams per rank, balanced by compressed size)"""
        n, _, _ = workload_plan(kind, total_bytes, len(streams))
        which = [i % len(streams) for i in range(n)]
        mine = partition_streams([len(streams[w]) for w in which], world)[rank]
        return [which[i] for i in mine]

    def resident(streams, sources, share):
        keep, descs = [], []
        for w in share:
            s = streams[w]
            t_in = torch.empty(len(s) + 64, dtype=torch.uint8, device=dev)
            t_in[: len(s)].copy_(torch.from_numpy(s))
            t_in[len(s):].zero_()
            t_out = torch.empty(len(sources[w]), dtype=torch.uint8, device=dev)
            keep.append((t_in, t_out))
            descs.append(dict(d_src=t_in.data_ptr(), src_size=len(s), src_capacity=len(s) + 64, d_dst=t_out.data_ptr(),
                              dst_capacity=len(sources[w]), header=bytes(s[:16])))
        torch.cuda.synchronize(dev)
        return keep, dec.plan(descs)

    def timed(plan, steps, warmup, sampler=None):
        ts = torch.cuda.Stream(dev)
        for _ in range(warmup):
            plan.launch(ts.cuda_stream)
        ts.synchronize()
        assert plan.finish() == 0
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        if sampler:
            sampler.start()
            time.sleep(0.02)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ts)
        for _ in range(steps):
            plan.launch(ts.cuda_stream)
        e1.record(ts)
        ts.synchronize()
        torch.cuda.synchronize(dev)
        clocks = sampler.stop() if sampler else None
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        assert plan.finish() == 0
        return max_over_ranks(ms), clocks

    def verify(keep, share, sources, what):
        for j in sorted({0, len(share) - 1}):
            assert np.array_equal(keep[j][1].cpu().numpy(), sources[share[j]]), f"GPU output != source bytes ({what})"

    def ncu_traffic(kind, algorithmic_bytes):
        # DRAM bytes per launch from the committed ncu capture of this kernel on this kind of payload
        # (profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum / algorithmic bytes of that capture)
        try:
            return float(json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[kind]["ratio_to_algorithmic"]) * algorithmic_bytes
        except Exception:
            return None

    peak, peak_src = measured_peak_gbs()

    # ---------------- primary workload: device-resident, the streams of the batch partitioned over the ranks
    total = args.size_gib << 30
    streams, sources = get_unique(args.workload, total, args.unique, datagen.SEED_CONFIG4)
    share = share_of(args.workload, total, streams)
    keep, plan = resident(streams, sources, share)
    sampler = ClockSampler(local) if rank == 0 else None
    ms, clocks = timed(plan, args.steps, args.warmup, sampler)
    verify(keep, share, sources, args.workload)
    my_out = float(sum(len(sources[w]) for w in share))
    my_in = float(sum(len(streams[w]) for w in share))
    all_out, all_in = sum_over_ranks(my_out), sum_over_ranks(my_in)
    ms_step = ms / args.steps
    value = all_out / (ms_step / 1e3) / 1e9
    achieved = (all_in + all_out) / world / (ms_step / 1e3) / 1e9          # per GPU (ranks hold equal shares)
    launches_per_step = plan.info["kernels_per_launch"]
    grid = plan.info["grid_blocks"]
    del keep, plan
    torch.cuda.empty_cache()

    # ---------------- end to end through the host-pointer C-ABI call (pinned host buffers), a sample of the same batch
    e2e = None
    if not args.no_e2e:
        e2e_total = min(total, args.e2e_gib << 30)
        e_share = share_of(args.workload, e2e_total, streams)
        uniq_used = sorted(set(e_share))
        pin_in = {w: torch.from_numpy(streams[w]).pin_memory() for w in uniq_used}      # replicas share the pinned source
        pin_out = [torch.empty(len(sources[w]), dtype=torch.uint8).pin_memory() for w in e_share]
        np_in = [pin_in[w].numpy() for w in e_share]
        np_out = [t.numpy() for t in pin_out]
        e_out = float(sum(len(sources[w]) for w in e_share))
        e_in = float(sum(len(streams[w]) for w in e_share))
        for _ in range(2):
            dec.decode_batch_host(np_in, np_out)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            dec.decode_batch_host(np_in, np_out)
        el = max_over_ranks(time.perf_counter() - t0)
        for j in sorted({0, len(e_share) - 1}):
            assert np.array_equal(np_out[j], sources[e_share[j]]), "e2e output != source bytes"
        # what the bus allows: the same bytes as plain copies, both directions at once, no kernel
        d_in = torch.empty(int(max(len(streams[w]) for w in uniq_used)) + 64, dtype=torch.uint8, device=dev)
        d_out = torch.empty(int(max(len(sources[w]) for w in uniq_used)), dtype=torch.uint8, device=dev)
        s_up, s_dn = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

        def copies():
            with torch.cuda.stream(s_up):
                for w in e_share:
                    d_in[: len(streams[w])].copy_(pin_in[w], non_blocking=True)
            with torch.cuda.stream(s_dn):
                for j, w in enumerate(e_share):
                    pin_out[j].copy_(d_out[: len(sources[w])], non_blocking=True)
            s_up.synchronize()
            s_dn.synchronize()
        copies()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(max(2, args.steps // 2)):
            copies()
        el_c = max_over_ranks(time.perf_counter() - t0) / max(2, args.steps // 2)
        all_e_out, all_e_in = sum_over_ranks(e_out), sum_over_ranks(e_in)
        e2e = {"value": all_e_out * args.steps / el / 1e9, "unit": "GB/s", "h2d_bytes_per_step": int(all_e_in), "d2h_bytes_per_step": int(all_e_out),
               "call": "bgx_decode_batch_host (C ABI), pinned host buffers, one call per rank and step",
               "sample": f"{int(all_e_out) >> 30} GiB of the batch per step ({len(e_share)} streams on rank 0)",
               "pcie_ceiling": {"value": all_e_out / el_c / 1e9, "unit": "GB/s decompressed",
                                "how": "same H2D + D2H bytes as cudaMemcpyAsync on two streams from the same pinned buffers, no kernel"}}
        e2e["frac_of_pcie_ceiling"] = e2e["value"] / e2e["pcie_ceiling"]["value"]
        del pin_in, pin_out, d_in, d_out
        torch.cuda.empty_cache()

    # ---------------- secondary (N = 1): the raw-page copy path, configs[1]
    secondary = None
    if world == 1 and not args.no_secondary and args.workload != "random":
        s2, d2 = get_unique("random", 4 << 30, 8, datagen.SEED_CONFIG2)      # (content is irrelevant to a copy: 8 unique x 8)
        sh2 = share_of("random", 4 << 30, s2)
        keep2, plan2 = resident(s2, d2, sh2)
        st2 = max(3, args.steps // 2)
        ms2, _ = timed(plan2, st2, 3)
        ms2 /= st2
        verify(keep2, sh2, d2, "random")
        o2, i2 = float(sum(len(d2[w]) for w in sh2)), float(sum(len(s2[w]) for w in sh2))
        secondary = {"workload": WORKLOADS["random"].format(gib=4, n=64), "note": "raw pages only: a memcpy upper bound of the page kernel, not a decoder figure",
                     "value": o2 / (ms2 / 1e3) / 1e9, "unit": "GB/s", "ms_per_step": ms2,
                     "roofline": {"bound": "hbm", "achieved": (i2 + o2) / (ms2 / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                                  "frac": (i2 + o2) / (ms2 / 1e3) / 1e9 / peak, "traffic": ncu_traffic("random", i2 + o2)}}
        del keep2, plan2, s2, d2
        torch.cuda.empty_cache()

    # ---------------- second secondary (N = 1): configs[2], 4 GiB of pre-conditioned BC3 textures (page kernel + de-conditioning)
    secondary_texture = None
    if world == 1 and not args.no_secondary and args.workload != "texture":
        s3, d3 = get_unique("texture", 4 << 30, 8, datagen.SEED_CONFIG3)
        sh3 = share_of("texture", 4 << 30, s3)
        keep3, plan3 = resident(s3, d3, sh3)
        st3 = max(3, args.steps // 2)
        ms3, _ = timed(plan3, st3, 3)
        ms3 /= st3
        verify(keep3, sh3, d3, "texture")
        o3, i3 = float(sum(len(d3[w]) for w in sh3)), float(sum(len(s3[w]) for w in sh3))
        secondary_texture = {"workload": WORKLOADS["texture"].format(gib=4, n=len(sh3)), "value": o3 / (ms3 / 1e3) / 1e9, "unit": "GB/s",
                             "ms_per_step": ms3, "kernels_per_step": plan3.info["kernels_per_launch"], "compression_ratio": o3 / i3,
                             "roofline": {"bound": "hbm", "achieved": (i3 + o3) / (ms3 / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                                          "frac": (i3 + o3) / (ms3 / 1e3) / 1e9 / peak, "traffic": ncu_traffic("texture", i3 + o3)}}
        del keep3, plan3, s3, d3
        torch.cuda.empty_cache()

    # ---------------- sharded streams (N > 1): one NCCL broadcast per stream, page ranges per rank
    sharded = None
    if world > 1 and not args.no_sharded and args.workload != "texture":
        n_sh = min(4, len(streams))
        t_bc = t_dec = 0.0
        bc_bytes = out_bytes = 0
        for rep in range(2):                                  # first pass warms NCCL and the kernels up
            t_bc = t_dec = 0.0
            bc_bytes = out_bytes = 0
            for i in range(n_sh):
                owner = i % world
                src = torch.from_numpy(streams[i]).to(dev) if rank == owner else None
                tm = {}
                torch.cuda.synchronize(dev)
                dist.barrier()
                shard, (lo, hi), nbytes = decode_sharded_stream(src, owner, cuda_decode_fn(dec, tm), device=dev,
                                                                capacity=len(streams[i]), timing=tm)   # (size known from the request metadata)
                t_bc += max_over_ranks(tm.get("broadcast_ms", 0.0))
                t_dec += max_over_ranks(tm.get("decode_ms", 0.0))
                bc_bytes += int(tm.get("broadcast_bytes", 0))
                out_bytes += len(sources[i])
                if rep == 1:
                    want = sources[i][lo * PAGE: lo * PAGE + int(shard.numel())]
                    assert np.array_equal(shard.cpu().numpy(), want), "sharded decode != source bytes"
        sharded = {"streams": n_sh, "stream_bytes": STREAM_BYTES, "collective": "one ncclBroadcast of [size | stream] per stream (dist.broadcast, NCCL over NVLink)",
                   "decode_gbs": out_bytes / (t_dec / 1e3) / 1e9, "with_broadcast_gbs": out_bytes / ((t_dec + t_bc) / 1e3) / 1e9,
                   "broadcast_gbs": bc_bytes / (t_bc / 1e3) / 1e9 if t_bc else None, "broadcast_bytes": bc_bytes,
                   "nvlink_peer_copy_gbs_reference": 770.0, "unit": "GB/s decompressed (broadcast_gbs: compressed bytes / broadcast time)"}

    # ---------------- CPU baseline (rank 0, N = 1 only): reference DecodeCPU on a bounded sample of the same streams
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        c = CpuDecoder(streams, [len(d) for d in sources]).sample(budget_s=15.0)
        cpu = {k: c[k] for k in ("value", "unit", "cores", "kind", "sample", "host_threads_available")}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args, world),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(args.workload, (all_in + all_out) / world), "peak_source": peak_src, "kernel": "bgx_decode_pages_kernel",
                         "algorithmic_bytes_per_launch": (all_in + all_out) / world, "per": "GPU", "grid_blocks": grid,
                         "compression_ratio": all_out / all_in},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches_per_step * args.steps, "clocks": clocks,
            "secondary": secondary, "secondary_texture": secondary_texture, "sharded": sharded, "bit_exact": True,
        }
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="mixed", choices=sorted(WORKLOADS))
    ap.add_argument("--size-gib", type=int, default=0, help="total batch size (default: 16 for mixed, 4 otherwise)")
    ap.add_argument("--unique", type=int, default=16, help="unique streams encoded (the batch replicates them into distinct buffers)")
    ap.add_argument("--e2e-gib", type=int, default=4, help="size of the batch sample the host-pointer (e2e) leg moves per step")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-sharded", action="store_true")
    args = ap.parse_args()
    if args.size_gib <= 0:
        args.size_gib = DEFAULT_GIB[args.workload]
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
