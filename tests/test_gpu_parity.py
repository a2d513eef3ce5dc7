"""GPU tests (-m gpu): the CUDA path, called through the C ABI (ctypes -> libbrotlig_b200.so), against the
oracle on the same seeded inputs, against the committed golden fixtures, and -- at BASELINE.json sizes --
through size-independent properties (round trip to the source bytes, checksum of checksums).
Bit-exact is the bar: this is byte/integer work."""
import ctypes
import json
import os

import numpy as np
import pytest

from conftest import sha256
from corpus import corner_cases, texture_cases

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def dec(sdk):
    from brotli_g_sdk_b200 import build
    build.build_cuda()
    d = sdk.BrotligDecoder(0)
    yield d
    d.close()


def test_corner_cases_match_oracle(sdk, oracle, dec):
    for name, (data, kw) in corner_cases().items():
        s = sdk.Encode(data, **kw)
        want = oracle.decode(s)
        got, ms = dec.decode_host(s)
        assert np.array_equal(got, want), f"{name}: GPU != oracle (first diff {np.nonzero(got != want)[0][:4]})"
        assert np.array_equal(got, data), name


def test_textures_match_oracle(sdk, oracle, dec):
    for name, (data, p) in texture_cases().items():
        s = sdk.Encode(data, dcParams=p)
        want = oracle.decode(s)
        got, _ = dec.decode_host(s)
        assert np.array_equal(got, want), f"{name}: GPU texture path != oracle"


def test_golden_fixtures(dec):
    idx = json.load(open(os.path.join(GOLDEN, "index.json")))
    for name, meta in idx.items():
        s = np.fromfile(os.path.join(GOLDEN, name + ".brotlig"), dtype=np.uint8)
        out, _ = dec.decode_host(s)
        assert len(out) == meta["size"] and sha256(out) == meta["output_sha256"], name


def test_batch_of_streams_one_launch(sdk, oracle, dec):
    items = list(corner_cases().items())[:14]
    streams = [sdk.Encode(d, **kw) for _, (d, kw) in items]
    outs, ms = dec.decode_batch_host(streams)
    assert ms > 0
    for (name, (d, _)), o in zip(items, outs):
        assert np.array_equal(o, d), name


def test_header_errors_and_output_size_contract(sdk, dec):
    s = sdk.Encode(np.arange(70000, dtype=np.uint32).view(np.uint8))
    bad = s.copy(); bad[1] ^= 0x20
    with pytest.raises(sdk.BrotligError) as e:
        dec.decode_host(bad)
    assert e.value.code == 14       # BROTLIG_ERROR_CORRUPT_STREAM (BrotligDecoder.cpp:438-441)
    bad = s.copy(); bad[0] = 9; bad[1] = 9 ^ 0xFF
    with pytest.raises(sdk.BrotligError) as e:
        dec.decode_host(bad)
    assert e.value.code == 15       # BROTLIG_ERROR_INCORRECT_STREAM_FORMAT (:443-446)
    # *output_size is in/out: a larger buffer is fine and the decoded size comes back
    big = np.full(sdk.DecompressedSize(s) + 100, 0x77, np.uint8)
    out, _ = dec.decode_host(s, big)
    assert len(out) == sdk.DecompressedSize(s)
    assert (big[len(out):] == 0x77).all()


def test_corrupt_payload_is_reported_not_crashed(sdk, dec):
    rng = np.random.default_rng(9)
    data = np.tile(rng.integers(0, 256, 997, dtype=np.uint8), 80)
    s = sdk.Encode(data)
    for _ in range(8):
        bad = s.copy()
        k = int(rng.integers(20, len(bad) - 4))
        bad[k: k + 4] ^= rng.integers(1, 256, 4, dtype=np.uint8)
        try:
            out, _ = dec.decode_host(bad)
        except sdk.BrotligError as e:
            assert e.code == 14
    out, _ = dec.decode_host(s)      # the context still works afterwards
    assert np.array_equal(out, data)


def test_reference_named_c_entry_points(sdk, dec):
    """DecodeCPU / DecompressedSize exported with the reference's names and argument order"""
    from brotli_g_sdk_b200 import build
    lib = ctypes.CDLL(build.CUDA_LIB)
    data = np.tile(np.arange(251, dtype=np.uint8), 1000)
    s = sdk.Encode(data)
    lib.DecompressedSize.restype = ctypes.c_uint32
    n = lib.DecompressedSize(ctypes.c_void_p(s.ctypes.data))
    assert n == len(data)
    out = np.zeros(n, np.uint8)
    osz = ctypes.c_uint32(n)
    lib.DecodeCPU.restype = ctypes.c_int
    rc = lib.DecodeCPU(ctypes.c_uint32(len(s)), ctypes.c_void_p(s.ctypes.data), ctypes.byref(osz), ctypes.c_void_p(out.ctypes.data), None)
    assert rc == 0 and osz.value == n and np.array_equal(out, data)


def test_device_resident_plan_with_page_ranges(sdk, dec):
    """the multi-GPU sharding primitive: a rank decodes pages [begin, begin+count) into its own shard"""
    import torch
    data = corner_cases()["mixed"][0]
    s = sdk.Encode(data)
    npages = (len(data) + 65535) // 65536
    t_in = torch.zeros(len(s) + 64, dtype=torch.uint8, device="cuda")
    t_in[: len(s)] = torch.from_numpy(s).cuda()
    cut = npages // 2
    shards = []
    for begin, count in ((0, cut), (cut, npages - cut)):
        nbytes = min(len(data), (begin + count) * 65536) - begin * 65536
        t_out = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
        plan = dec.plan([dict(d_src=t_in.data_ptr(), src_size=len(s), src_capacity=len(s) + 64, d_dst=t_out.data_ptr(),
                              dst_capacity=nbytes, header=bytes(s[:16]), page_begin=begin, page_count=count)])
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        plan.launch(side.cuda_stream)
        assert plan.finish() == 0
        shards.append(t_out.cpu().numpy())
    assert np.array_equal(np.concatenate(shards), data)


def test_device_output_at_any_alignment(sdk, oracle, dec):
    """the output pointer of a device-resident decode may have any alignment (byte-granular flush and
    far-match paths of the page kernel; raw pages fall back to narrower copies)"""
    import torch
    from brotli_g_sdk_b200 import datagen
    data = np.concatenate([datagen.text_like(200000, seed=41), datagen.random_bytes(70000, seed=42),
                           datagen.structured_binary(100000, seed=43), np.zeros(4000, np.uint8)])
    s = sdk.Encode(data)
    want = oracle.decode(s)
    t_in = torch.zeros(len(s) + 64, dtype=torch.uint8, device="cuda")
    t_in[: len(s)] = torch.from_numpy(s).cuda()
    for off in (1, 2, 3, 4, 8):
        t_out = torch.full((len(data) + 64,), 0xEE, dtype=torch.uint8, device="cuda")
        plan = dec.plan([dict(d_src=t_in.data_ptr(), src_size=len(s), src_capacity=len(s) + 64, d_dst=t_out.data_ptr() + off,
                              dst_capacity=len(data), header=bytes(s[:16]))])
        plan.launch(torch.cuda.current_stream().cuda_stream)
        assert plan.finish() == 0
        got = t_out.cpu().numpy()
        assert np.array_equal(got[off: off + len(data)], want), f"offset {off}"
        assert (got[:off] == 0xEE).all() and (got[off + len(data):] == 0xEE).all(), f"offset {off}: wrote outside the buffer"


def test_full_size_random_buffer_checksum_of_checksums(sdk, dec):
    """BASELINE config 2 shape (64 MiB streams of raw pages), 1 GiB here: per-page CRCs of the decoded
    buffer equal those of the source (checked on the device side via torch, no oracle at this size)."""
    import torch
    import zlib
    from brotli_g_sdk_b200 import datagen
    streams, sources = [], []
    for i in range(4):
        d = datagen.random_bytes(64 << 20, seed=datagen.SEED_CONFIG2 + i)
        sources.append(d)
        streams.append(sdk.Encode(d))
    outs, _ = dec.decode_batch_host(streams)
    for d, o in zip(sources, outs):
        assert zlib.crc32(o.tobytes()) == zlib.crc32(d.tobytes())


def test_full_size_mixed_entropy_batch_is_bit_exact(sdk, dec):
    """BASELINE config 4 shape (64 MiB mixed-entropy streams: compressed and raw pages side by side), 256 MiB
    here, decoded in one device-resident launch: every stream equals its source byte for byte (the oracle
    is too slow at this size; the source is the ground truth the oracle is pinned to)"""
    import torch
    from brotli_g_sdk_b200 import datagen
    keep, descs, sources = [], [], []
    for i in range(4):
        d = datagen.mixed(64 << 20, seed=datagen.SEED_CONFIG4 + 7 * i)
        s = sdk.Encode(d)
        assert len(s) < len(d)
        t_in = torch.zeros(len(s) + 64, dtype=torch.uint8, device="cuda")
        t_in[: len(s)] = torch.from_numpy(s).cuda()
        t_out = torch.full((len(d),), 0xEE, dtype=torch.uint8, device="cuda")
        keep.append((t_in, t_out))
        sources.append(d)
        descs.append(dict(d_src=t_in.data_ptr(), src_size=len(s), src_capacity=len(s) + 64, d_dst=t_out.data_ptr(),
                          dst_capacity=len(d), header=bytes(s[:16])))
    plan = dec.plan(descs)
    for _ in range(2):      # a second launch over the same plan: the persistent CTAs leave no state behind
        plan.launch(torch.cuda.current_stream().cuda_stream)
        assert plan.finish() == 0
    for (t_in, t_out), d in zip(keep, sources):
        assert torch.equal(t_out.cpu(), torch.from_numpy(d)), "GPU output != source bytes"


def test_single_large_stream_through_the_host_call(sdk, dec):
    """DecodeGPU of ONE 64 MiB mixed-entropy stream: the host-pointer path splits it into page ranges that are
    uploaded, decoded and downloaded as a pipeline; the result must not depend on that"""
    from brotli_g_sdk_b200 import datagen
    d = datagen.mixed(64 << 20, seed=datagen.SEED_CONFIG4 + 99)
    s = sdk.Encode(d)
    out, ms = dec.decode_host(s)
    assert ms > 0 and np.array_equal(out, d), "host-pointer decode of a large stream != source bytes"


def test_page_size_sweep_small_single_page_streams(sdk, oracle, dec):
    """config 5: 4/8/16 KiB "pages" are single-page streams (NumPages = 1, LastPageSize = n)"""
    from brotli_g_sdk_b200 import datagen
    streams, sources = [], []
    for i, n in enumerate([4096, 8192, 16384] * 20):
        d = datagen.text_like(n, seed=100 + i)
        sources.append(d)
        streams.append(sdk.Encode(d))
    outs, _ = dec.decode_batch_host(streams)
    for d, o, s in zip(sources, outs, streams):
        assert np.array_equal(o, d)
    assert np.array_equal(oracle.decode(streams[7]), sources[7])


def test_thousands_of_ragged_streams_in_one_launch(sdk, dec):
    """More streams than resident CTAs, with uneven page counts: every CTA walks the flat page queue across many
    streams, so the kernel's stream look-up runs its proportional guess, its neighbour check and its fallback search."""
    from brotli_g_sdk_b200 import datagen
    rng = np.random.default_rng(77)
    uniq = []
    for i, n in enumerate([1, 300, 1500, 4096, 9000, 16384, 40000, 65536, 65537, 150000, 200000, 333]):
        d = datagen.mixed(n, seed=300 + i) if n > 2000 else datagen.text_like(n, seed=300 + i)
        uniq.append((d, sdk.Encode(d)))
    # long stretches of one-page streams with a few many-page streams in between, then a ragged tail
    order = [0, 1, 2, 3] * 900 + [9, 10] * 8 + [4, 5] * 700 + list(rng.integers(0, len(uniq), size=1500))
    streams = [uniq[k][1] for k in order]
    outs, ms = dec.decode_batch_host(streams)
    assert ms > 0 and len(outs) == len(order)
    for j, k in enumerate(order):
        assert np.array_equal(outs[j], uniq[k][0]), f"stream {j} (kind {k}) != source bytes"


def test_plan_rejects_bad_buffers(sdk, dec):
    """device-resident contract: 16-byte aligned stream pointer, capacity covering the stream rounded up to 16"""
    import torch
    s = sdk.Encode(np.arange(100000, dtype=np.uint8))
    t_in = torch.zeros(len(s) + 80, dtype=torch.uint8, device="cuda")
    t_in[1: 1 + len(s)] = torch.from_numpy(s).cuda()
    t_out = torch.zeros(100000, dtype=torch.uint8, device="cuda")
    with pytest.raises(sdk.BrotligError) as e:
        dec.plan([dict(d_src=t_in.data_ptr() + 1, src_size=len(s), src_capacity=len(s) + 64, d_dst=t_out.data_ptr(), dst_capacity=100000,
                       header=bytes(s[:16]))])
    assert e.value.code == 16 and "aligned" in str(e.value)
    with pytest.raises(sdk.BrotligError) as e:
        dec.plan([dict(d_src=t_in.data_ptr(), src_size=len(s), src_capacity=len(s), d_dst=t_out.data_ptr(), dst_capacity=100000,
                       header=bytes(s[:16]))] if len(s) % 16 else [dict(d_src=t_in.data_ptr(), src_size=len(s), src_capacity=len(s) - 1,
                                                                        d_dst=t_out.data_ptr(), dst_capacity=100000, header=bytes(s[:16]))])
    assert e.value.code in (14, 16)
    with pytest.raises(sdk.BrotligError):
        dec.plan([dict(d_src=t_in.data_ptr(), src_size=len(s), src_capacity=len(s) + 64, d_dst=t_out.data_ptr(), dst_capacity=999,
                       header=bytes(s[:16]))])


def test_cli_round_trip(tmp_path, sdk):
    """python -m brotli_g_sdk_b200.cli: compress a file, decompress the .brotlig on the GPU (twin of brotlig_cli)"""
    import subprocess
    import sys
    from brotli_g_sdk_b200 import datagen
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "payload.bin"
    data = datagen.mixed(700000, seed=3)
    data.tofile(src)
    env = dict(os.environ, PYTHONPATH=root)
    r = subprocess.run([sys.executable, "-m", "brotli_g_sdk_b200.cli", "-pagesize", "65536", str(src)], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    comp = tmp_path / "payload.bin.brotlig"
    assert comp.exists() and comp.stat().st_size < len(data)
    out = tmp_path / "restored.bin"
    r = subprocess.run([sys.executable, "-m", "brotli_g_sdk_b200.cli", "-num-repeat", "2", "-output", str(out), str(comp)],
                       capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    assert "GB/s decompressed" in r.stdout
    assert np.array_equal(np.fromfile(out, dtype=np.uint8), data)


def test_progress_callback_and_abort_through_the_c_abi(sdk, oracle, dec):
    """bgx_decode_host_progress: one report per page, in order; a true return stops the decode between two page groups and
    leaves zeros behind (BrotligDecoder.cpp:318-325,448)"""
    from brotli_g_sdk_b200 import datagen
    data = np.concatenate([datagen.text_like(21 * 65536, seed=71), datagen.structured_binary(11 * 65536 + 500, seed=72)])
    s = sdk.Encode(data)
    want = oracle.decode(s)
    seen = []
    out, ms = dec.decode_host_progress(s, lambda p, n: seen.append((p, n)) or False, pages_per_group=8)
    assert np.array_equal(out, want) and ms > 0
    assert seen == [(p, 33) for p in range(33)]
    seen.clear()
    out, _ = dec.decode_host_progress(s, lambda p, n: seen.append(p) or p >= 9, output=np.full(len(data), 0xA5, np.uint8), pages_per_group=8)
    assert seen == list(range(10))                       # stopped inside the second group of 8 pages
    assert np.array_equal(out[: 16 * 65536], want[: 16 * 65536]) and not out[16 * 65536:].any()


def test_batch_over_several_contexts(sdk, oracle):
    """bgx_decode_batch_host_multi: whole streams split over the contexts (one per device; on a one-GPU box two contexts of
    the same device), outputs identical to the oracle"""
    import torch
    from brotli_g_sdk_b200 import datagen
    from brotli_g_sdk_b200.decoder import decode_batch_host_multi
    ndev = max(1, torch.cuda.device_count())
    decs = [sdk.BrotligDecoder(i % ndev) for i in range(max(2, min(ndev, 4)))]
    datas = [datagen.mixed(n, seed=80 + i) for i, n in enumerate((300000, 65536, 1 << 20, 70000, 5, 2 * 65536 + 1, 650000))]
    streams = [sdk.Encode(d) for d in datas]
    outs, ms = decode_batch_host_multi(decs, streams)
    for d, s, o in zip(datas, streams, outs):
        assert np.array_equal(o, oracle.decode(s)) and np.array_equal(o, d)
    assert ms > 0
    for d in decs:
        d.close()
