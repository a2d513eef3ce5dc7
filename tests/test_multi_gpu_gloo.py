"""CPU tests (-m "not gpu"): the N>1 plumbing with world_size 2 on the gloo backend -- stream
partitioning, page-range sharding, and the single broadcast of a sharded stream. The decode step is
injected (oracle based) because there is no GPU here; on the GPU box the same code runs with
cuda_decode_fn (see test_multi_gpu_nccl in test_gpu_multi.py)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from brotli_g_sdk_b200.multi_gpu import StreamGeometry, decode_sharded_stream, partition_streams, shard_pages  # noqa: E402


def test_partition_is_balanced_and_complete():
    sizes = [64, 64, 64, 1, 1, 1, 200, 7, 7, 30]
    parts = partition_streams(sizes, 4)
    assert sorted(i for p in parts for i in p) == list(range(len(sizes)))
    loads = [sum(sizes[i] for i in p) for p in parts]
    assert max(loads) <= 200 and max(loads) - min(loads) <= 200
    assert partition_streams([5, 5], 8).count([]) == 6


def test_page_ranges_tile_the_stream():
    for n in (1, 2, 7, 16, 1000, 65535):
        for w in (1, 2, 3, 8):
            r = shard_pages(n, w)
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
    g = StreamGeometry.parse(bytes([0x05, 0xFA, 0x03, 0x00, 0xA1, 0x0F, 0x00, 0x00]))
    assert (g.num_pages, g.page_size, g.last_page_size, g.uncompressed_size) == (3, 65536, 1000, 2 * 65536 + 1000)
    assert g.range_bytes(0, 2) == 131072 and g.range_bytes(2, 3) == 1000 and g.range_bytes(1, 3) == 65536 + 1000


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, stream_np, expected_np, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conftest import Oracle
    orc = Oracle()

    def oracle_decode_fn(buf, n, geo, lo, hi):
        full = orc.decode(buf[:n].numpy())
        return torch.from_numpy(full[lo * geo.page_size: lo * geo.page_size + geo.range_bytes(lo, hi)].copy())

    owner = 1
    src = torch.from_numpy(stream_np.copy()) if rank == owner else None
    geo = StreamGeometry.parse(bytes(stream_np[:16]))
    ok = True
    # once with the size riding inside the one broadcast (every rank knows an upper bound), once with the 8-byte
    # control pre-broadcast of the size
    for capacity in (len(stream_np) + 1000, None):
        shard, (lo, hi), nbytes = decode_sharded_stream(src, owner, oracle_decode_fn, device=torch.device("cpu"), capacity=capacity)
        want = expected_np[lo * geo.page_size: lo * geo.page_size + geo.range_bytes(lo, hi)]
        ok = ok and bool(np.array_equal(shard.numpy(), want)) and nbytes == len(stream_np)
    gathered = [None] * world
    dist.all_gather_object(gathered, (rank, lo, hi, ok))
    if rank == 0:
        q.put(gathered)
    dist.destroy_process_group()


def test_sharded_stream_broadcast_world2():
    import brotli_g_sdk_b200 as b
    from brotli_g_sdk_b200 import datagen
    data = datagen.mixed(7 * 65536 + 999, seed=77)
    stream = b.Encode(data)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, stream, data, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[3] for r in res), res
    ranges = sorted((r[1], r[2]) for r in res)
    assert ranges[0][0] == 0 and ranges[0][1] == ranges[1][0] and ranges[1][1] == 8
