"""Seeded test payloads that push the decoder through the corner cases SURVEY.md section 8c lists.
Each entry: name -> (data, encoder kwargs). Shared by the oracle, emulator and GPU parity tests."""
from __future__ import annotations

import functools

import numpy as np

from brotli_g_sdk_b200 import datagen


def _rng(seed):
    return np.random.Generator(np.random.Philox(seed))


def ring_code_workout(n: int, seed: int) -> np.ndarray:
    """records whose matches reuse the last / second-last distance and +-1..3 around them"""
    r = _rng(seed)
    base = r.integers(0, 256, size=257, dtype=np.uint8)
    out = [base]
    total = len(base)
    dists = [17, 64, 129, 255]
    buf = base[-300:]                      # the last <= 300 bytes produced so far
    while total < n:
        d = dists[int(r.integers(0, 4))] + int(r.integers(-3, 4))
        l = int(r.integers(3, 40))
        d = max(1, min(d, len(buf)))
        seg = np.array([buf[len(buf) - d + (k % d)] for k in range(l)], dtype=np.uint8)
        lit = r.integers(0, 256, size=int(r.integers(0, 4)), dtype=np.uint8)
        out += [seg, lit]
        total += l + len(lit)
        buf = np.concatenate([buf, seg, lit])[-300:]
    return np.concatenate(out)[:n].copy()


def run_structure_fuzz(n: int, seed: int) -> np.ndarray:
    """literal runs and copies whose lengths sit on and around the page kernel's hand-over limits (256 literals per
    virtual round, 1024 bytes per round, 512-byte literal ring, 32-byte rows), near and far, some overlapping"""
    r = _rng(seed)
    edges = np.array([1, 2, 3, 31, 32, 33, 63, 64, 65, 255, 256, 257, 287, 288, 289, 511, 512, 513, 767, 768, 1023, 1024,
                      1025, 1279, 1280, 2047, 2048, 2049, 3000, 5000])
    out = np.empty(n + 6000, np.uint8)
    pos = 0
    while pos < n:
        ins = int(r.choice(edges)) if r.random() < 0.5 else int(r.integers(0, 12))
        out[pos: pos + ins] = r.integers(0, 256, ins, dtype=np.uint8)
        pos += ins
        if pos == 0:
            continue
        cpy = int(r.choice(edges)) if r.random() < 0.4 else int(r.integers(2, 40))
        dist = int(r.choice([1, 2, 3, 7, 64, 300, 2047, 2048, 2049, 5000, 40000])) if r.random() < 0.7 else int(r.integers(1, pos + 1))
        dist = min(dist, pos)
        out[pos: pos + cpy] = np.resize(out[pos - dist: pos], cpy)   # byte-serial copy: an overlap replicates its pattern
        pos += cpy
    return out[:n].copy()


@functools.lru_cache(maxsize=1)
def corner_cases() -> dict:
    r = _rng(0xC0FFEE)
    c = {}
    c["single_page_lowent"] = (datagen.low_entropy(65536, seed=1), {})
    c["sixteen_pages_lowent"] = (datagen.low_entropy(1 << 20, seed=2)[: 5 * 65536], {})
    c["partial_last_page"] = (datagen.low_entropy(2 * 65536 + 12345, seed=3), {})
    c["tiny_4k"] = (datagen.low_entropy(4096, seed=4), {})
    c["tiny_33"] = (datagen.low_entropy(33, seed=5), {})
    c["tiny_17_raw"] = (datagen.low_entropy(17, seed=6), {})
    c["one_byte"] = (np.array([42], np.uint8), {})
    c["const_run"] = (np.full(100000, 7, np.uint8), {})                       # trivial literal table, dist 1, 24-bit copy length
    c["period3"] = (np.tile(np.array([1, 2, 3], np.uint8), 30000), {})        # overlapping copy, dist 3
    c["period2_two_syms"] = (np.tile(np.array([9, 200], np.uint8), 40000), {})
    c["two_symbols"] = (r.integers(0, 2, 50000).astype(np.uint8), {})          # simple table {1,1}
    c["three_symbols"] = (r.integers(0, 3, 50000).astype(np.uint8), {})        # simple table {1,2,2}
    c["four_symbols_flat"] = (r.integers(0, 4, 50000).astype(np.uint8), {})    # simple table {2,2,2,2}
    c["four_symbols_skew"] = (r.choice(4, 50000, p=[.7, .2, .05, .05]).astype(np.uint8), {})   # {1,2,3,3}
    c["random_raw"] = (datagen.random_bytes(140000, seed=7), {})               # raw pages + ragged raw last page
    c["page_32k"] = (datagen.low_entropy(100000, seed=8), dict(page_size=32768))
    c["page_128k"] = (datagen.text_like(300000, seed=9), dict(page_size=131072))
    c["no_rle_tables"] = (datagen.text_like(100000, seed=10), dict(rle_mode=1))
    c["insert_only_split"] = (datagen.text_like(100000, seed=11), dict(split_insert_over=5))
    c["big_alphabet"] = (_rng(12).integers(0, 250, 100000).astype(np.uint8) // 3 * 3, {})
    long_lit = np.concatenate([datagen.random_bytes(30000, seed=13), np.zeros(20000, np.uint8), datagen.random_bytes(3000, seed=14)])
    c["long_insert_long_copy"] = (np.concatenate([long_lit, long_lit[:40000]]), {})   # insert >= 22594, copy >= 2118
    c["period700"] = (np.tile(datagen.random_bytes(700, seed=15), 200), {})
    c["ring_codes"] = (ring_code_workout(150000, 16), {})
    c["text"] = (datagen.text_like(200000, seed=17), {})
    c["structured_binary"] = (datagen.structured_binary(200000, seed=18), {})
    c["mixed"] = (datagen.mixed(600000, seed=19), {})
    for np_ in range(4):
        for nd in (0, 15):
            c[f"npostfix{np_}_ndirect{nd}"] = (ring_code_workout(70000, 20 + np_), dict(npostfix=np_, ndirect_msb=nd))
    # rounds made of many big commands: every command inserts hundreds of literals and copies a little (the page kernel
    # splits such rounds into virtual rounds: whole commands, then pieces of one command)
    rr = _rng(0xB16)
    key = rr.integers(0, 256, 64, dtype=np.uint8)
    parts = []
    for i in range(90):
        parts += [rr.integers(0, 256, int(rr.integers(300, 1500)), dtype=np.uint8), key[: int(rr.integers(8, 64))]]
    c["big_inserts_small_copies"] = (np.concatenate(parts), {})
    parts_big = parts
    # ... and the mirror image: a few literals, then kilobytes copied from far back (copies longer than a round takes)
    base = rr.integers(0, 256, 9000, dtype=np.uint8)
    parts = [base]
    for i in range(40):
        o = int(rr.integers(0, 4000))
        parts += [rr.integers(0, 256, int(rr.integers(1, 6)), dtype=np.uint8), base[o: o + int(rr.integers(1100, 4800))]]
    c["small_inserts_big_copies"] = (np.concatenate(parts), {})
    for sd in (1, 2, 3):
        c[f"run_structure_fuzz{sd}"] = (run_structure_fuzz(140000, 0xF00 + sd), {})
    c["run_structure_fuzz_small"] = (run_structure_fuzz(45000, 0xF10), {})      # small enough for a golden fixture
    c["big_inserts_small"] = (np.concatenate(parts_big[:40]), {})
    c["no_ring_codes"] = (datagen.text_like(80000, seed=30), dict(use_ring_codes=0))
    c["greedy_short_chain"] = (datagen.text_like(80000, seed=31), dict(lazy=0, max_chain=1))
    return c


def texture_cases() -> dict:
    """name -> (data, DataconditionParams kwargs)"""
    c = {}
    from brotli_g_sdk_b200.encoder import DataconditionParams as P
    def tex(fmt, w, h, mips=1, swizzle=False, delta=False, aligned=False, seed=1):
        p = P(precondition=True, swizzle=swizzle, delta_encode=delta, format=fmt, width_blocks=w, height_blocks=h, num_mips=mips, pitch_aligned=aligned)
        size = p.texture_size()
        r = _rng(seed)
        if mips == 1 and not aligned and fmt in (1, 3):
            data = datagen.bc_texture(w, h, fmt, seed=seed)
        else:
            data = (r.integers(0, 256, size=size, dtype=np.uint8) & r.choice(np.array([0xFF, 0x0F, 0x3C, 0x81], np.uint8), size=size))
        assert len(data) == size, (len(data), size)
        return data, p
    c["bc1_64x64"] = tex(1, 64, 64)
    c["bc1_swz_delta"] = tex(1, 128, 96, swizzle=True, delta=True, seed=2)
    c["bc2_swz"] = tex(2, 33, 17, swizzle=True, seed=3)
    c["bc3_256x256_swz_delta"] = tex(3, 256, 256, swizzle=True, delta=True, seed=4)
    c["bc3_odd_mips"] = tex(3, 37, 23, mips=3, swizzle=True, delta=True, seed=5)
    c["bc4_aligned_pitch"] = tex(4, 50, 20, aligned=True, delta=True, seed=6)
    c["bc5_mips4_aligned_swz"] = tex(5, 64, 48, mips=4, aligned=True, swizzle=True, delta=True, seed=7)
    c["bc1_1x1"] = tex(1, 1, 1, seed=8)
    c["bc3_big_delta_noswz"] = tex(3, 200, 180, delta=True, seed=9)
    return c
