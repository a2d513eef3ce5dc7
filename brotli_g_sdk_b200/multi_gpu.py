"""Multi-GPU launcher: one process per GPU (torch.distributed), pages are the unit of parallelism.

Replaces the single-device D3D12 host of the reference (sample/BrotligGPUDecoder.cpp:260-748), which has
no multi-GPU story. Brotli-G pages -- and therefore streams -- are independent (no cross-page LZ77
references, per-page prefix codes and distance ring: PageDecoder.cpp:126-153), so:

  level 1  whole streams are assigned to ranks, size balanced: NO communication on the data path;
  level 2  one big stream (fewer streams than ranks): every rank decodes a contiguous page range
           [lo, hi) of it into its own output shard. The compressed bytes live on one rank, so the
           stream is replicated with exactly ONE collective -- dist.broadcast of [size | stream] over
           NCCL (NVLink 5 / NVSwitch) -- and nothing else ever crosses the fabric; the output stays
           sharded. (Texture streams are not split: their owner decodes them whole.)

The decode itself is always the CUDA path (BrotligDecoder.plan). `decode_fn` exists so that the
plumbing (partitioning, broadcast, shard geometry) can be tested on CPU with the gloo backend.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Sequence

import numpy as np


def partition_streams(sizes: Sequence[int], world: int) -> list[list[int]]:
    """Longest-processing-time-first assignment of stream indices to ranks, balanced by bytes."""
    order = sorted(range(len(sizes)), key=lambda i: -sizes[i])
    load = [0] * world
    out: list[list[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += sizes[i]
    for lst in out:
        lst.sort()
    return out


def shard_pages(num_pages: int, world: int) -> list[tuple[int, int]]:
    """Contiguous page ranges [lo, hi) per rank ([r*P/R, (r+1)*P/R))."""
    return [(r * num_pages // world, (r + 1) * num_pages // world) for r in range(world)]


@dataclass
class StreamGeometry:
    num_pages: int
    page_size: int
    last_page_size: int
    uncompressed_size: int

    @staticmethod
    def parse(header: bytes) -> "StreamGeometry":
        w0 = int.from_bytes(header[0:4], "little")
        w1 = int.from_bytes(header[4:8], "little")
        n = w0 >> 16
        ps = 32768 << (w1 & 3)
        last = (w1 >> 2) & 0x3FFFF
        return StreamGeometry(n, ps, last, n * ps - ((ps - last) if last else 0))

    def range_bytes(self, lo: int, hi: int) -> int:
        if hi <= lo:
            return 0
        n = (hi - lo) * self.page_size
        if hi == self.num_pages and self.last_page_size:
            n -= self.page_size - self.last_page_size
        return n


SHARD_HEADER_BYTES = 16   # the broadcast buffer starts with the stream size (u64 LE) + 8 bytes of padding: payload stays 16-byte aligned


def decode_sharded_stream(stream_on_owner, owner: int, decode_fn: Callable, device=None, group=None, capacity: int | None = None,
                          timing: dict | None = None):
    """Level-2 decode of ONE stream across all ranks of `group`.

    stream_on_owner: uint8 torch tensor holding the stream on rank `owner` (ignored elsewhere).
    capacity:        an upper bound of the stream size that every rank knows (e.g. the largest stream the service
                     accepts, or the size from the request metadata). The receive buffers are allocated from it and the
                     stream size rides in the first 8 bytes of the broadcast buffer, so that EXACTLY ONE collective --
                     the broadcast of [size | stream bytes] over NCCL (NVLink 5 / NVSwitch) -- is issued. Without it the
                     size travels in a second, 8-byte control broadcast first.
    decode_fn(buf, n, geometry, lo, hi) -> output shard (bytes of pages [lo, hi)); buf[:n] is the stream.
    timing:          optional dict; receives "broadcast_ms" (CUDA events around the collective) on CUDA devices.
    Returns (shard, (lo, hi), stream_bytes). Pre-conditioned (texture) streams are not split -- a page range of
    a texture scatters into the whole texture -- so the owner decodes all pages and the other ranks get an empty shard.
    """
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = device if device is not None else (stream_on_owner.device if rank == owner else torch.device("cpu"))
    if capacity is None:   # control-plane pre-broadcast of the size (8 bytes), then the data-path broadcast
        meta = torch.tensor([int(stream_on_owner.numel()) if rank == owner else 0], dtype=torch.int64, device=dev)
        dist.broadcast(meta, src=owner, group=group)
        capacity = int(meta.item())
    cap16 = ((capacity + 15) // 16) * 16
    buf = torch.zeros(SHARD_HEADER_BYTES + cap16 + 64, dtype=torch.uint8, device=dev)
    if rank == owner:
        n = int(stream_on_owner.numel())
        if n > capacity:
            raise ValueError(f"stream of {n} bytes exceeds the agreed capacity {capacity}")
        buf[:8] = torch.from_numpy(np.frombuffer(int(n).to_bytes(8, "little"), dtype=np.uint8).copy()).to(dev)
        buf[SHARD_HEADER_BYTES: SHARD_HEADER_BYTES + n].copy_(stream_on_owner)
    on_cuda = torch.device(dev).type == "cuda"
    if on_cuda and timing is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    dist.broadcast(buf[: SHARD_HEADER_BYTES + cap16], src=owner, group=group)          # <- the one data-path collective
    if on_cuda and timing is not None:
        e1.record()
        e1.synchronize()
        timing["broadcast_ms"] = e0.elapsed_time(e1)
        timing["broadcast_bytes"] = SHARD_HEADER_BYTES + cap16
    head = bytes(buf[: SHARD_HEADER_BYTES + 16].cpu().numpy())
    n = int.from_bytes(head[:8], "little")
    geo = StreamGeometry.parse(head[SHARD_HEADER_BYTES:])
    preconditioned = (int.from_bytes(head[SHARD_HEADER_BYTES + 4: SHARD_HEADER_BYTES + 8], "little") >> 20) & 1
    if preconditioned:
        lo, hi = (0, geo.num_pages) if rank == owner else (0, 0)
    else:
        lo, hi = shard_pages(geo.num_pages, world)[rank]
    payload = buf[SHARD_HEADER_BYTES:]            # 16-byte aligned (torch allocations are 256-byte aligned)
    if hi > lo:
        shard = decode_fn(payload, n, geo, lo, hi)
    else:
        shard = torch.empty(0, dtype=torch.uint8, device=dev)
    return shard, (lo, hi), n


def cuda_decode_fn(decoder, timing: dict | None = None):
    """decode_fn for decode_sharded_stream backed by the CUDA plan interface. `timing`, if given, receives
    "decode_ms": the CUDA-event duration of the launch."""
    import torch

    def fn(buf, n, geo: StreamGeometry, lo: int, hi: int):
        nbytes = geo.range_bytes(lo, hi)
        out = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=buf.device)
        if nbytes == 0:
            return out[:0]
        need = ((n + 15) // 16) * 16
        if int(buf.numel()) < need or buf.data_ptr() % 16:      # the kernel stages input in aligned 16-byte chunks
            padded = torch.zeros(need + 64, dtype=torch.uint8, device=buf.device)
            padded[:n].copy_(buf[:n])
            buf = padded
        plan = decoder.plan([dict(d_src=buf.data_ptr(), src_size=n, src_capacity=int(buf.numel()), d_dst=out.data_ptr(),
                                  dst_capacity=nbytes, header=bytes(buf[:16].cpu().numpy()), page_begin=lo, page_count=hi - lo)])
        # run on a side stream ordered after whatever produced `buf` (e.g. the NCCL broadcast)
        side = torch.cuda.Stream(buf.device)
        side.wait_stream(torch.cuda.current_stream(buf.device))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(side)
        plan.launch(side.cuda_stream)
        e1.record(side)
        bad = plan.finish()
        if timing is not None:
            timing["decode_ms"] = e0.elapsed_time(e1)
        plan.close()
        if bad:
            raise RuntimeError(f"{bad} page(s) failed to decode")
        return out[:nbytes]

    return fn
