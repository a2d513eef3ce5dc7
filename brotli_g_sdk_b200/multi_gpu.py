"""Multi-GPU launcher: one process per GPU (torch.distributed), pages are the unit of parallelism.

Replaces the single-device D3D12 host of the reference (sample/BrotligGPUDecoder.cpp:260-748), which has
no multi-GPU story. Brotli-G pages -- and therefore streams -- are independent (no cross-page LZ77
references, per-page prefix codes and distance ring: PageDecoder.cpp:126-153), so:

  level 1  whole streams are assigned to ranks, size balanced: NO communication on the data path;
  level 2  one big stream (fewer streams than ranks): every rank decodes a contiguous page range
           [lo, hi) of it into its own output shard. The compressed bytes live on one rank, so the
           stream is replicated with exactly ONE collective -- dist.broadcast over NCCL (NVLink 5 /
           NVSwitch) -- and nothing else ever crosses the fabric; the output stays sharded.

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


def decode_sharded_stream(stream_on_owner, owner: int, decode_fn: Callable, device=None, group=None):
    """Level-2 decode of ONE stream across all ranks of `group`.

    stream_on_owner: uint8 torch tensor holding the stream on rank `owner` (ignored elsewhere).
    decode_fn(stream_tensor, geometry, lo, hi) -> output shard (bytes of pages [lo, hi)).
    Returns (shard, (lo, hi), broadcast_bytes). Exactly one collective (the broadcast) is issued.
    """
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = device if device is not None else (stream_on_owner.device if rank == owner else torch.device("cpu"))
    # 1) size, then the bytes: a single broadcast carries [size (8 B) | stream] so there is one collective
    if rank == owner:
        n = int(stream_on_owner.numel())
        meta = torch.tensor([n], dtype=torch.int64, device=dev)
    else:
        meta = torch.zeros(1, dtype=torch.int64, device=dev)
    # the size is needed to allocate the receive buffer; it rides in an 8-byte pre-broadcast that is part of
    # the control plane (not the data path). Callers that know the size can skip it via broadcast_stream().
    dist.broadcast(meta, src=owner, group=group)
    n = int(meta.item())
    # receive buffer: 16-byte aligned (torch allocations are) and padded to the 16-byte staging granularity
    buf = stream_on_owner if rank == owner else torch.zeros(((n + 15) // 16) * 16 + 64, dtype=torch.uint8, device=dev)
    payload = buf[:n]
    dist.broadcast(payload, src=owner, group=group)          # <- the one data-path collective
    geo = StreamGeometry.parse(bytes(payload[:16].cpu().numpy()))
    lo, hi = shard_pages(geo.num_pages, world)[rank]
    shard = decode_fn(buf, n, geo, lo, hi)
    return shard, (lo, hi), n


def cuda_decode_fn(decoder):
    """decode_fn for decode_sharded_stream backed by the CUDA plan interface."""
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
        plan.launch(side.cuda_stream)
        bad = plan.finish()
        plan.close()
        if bad:
            raise RuntimeError(f"{bad} page(s) failed to decode")
        return out[:nbytes]

    return fn
