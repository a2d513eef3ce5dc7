"""Seeded synthetic payloads for the BASELINE.json configs. numpy only.

  low_entropy        config 1: 16-symbol skewed alphabet (P ~ 2^-k) + injected back-references
  random_bytes       config 2: incompressible bytes (every page is stored raw)
  bc_texture         config 3: BC1/BC3-like blocks: smooth end-point fields + noisy index bits
  text_like          config 4: Zipf word model
  structured_binary  config 4: fixed-width records with counters / floats
  mixed              config 4/5: 50 % text, 25 % structured binary, 25 % random
"""
from __future__ import annotations

import numpy as np

SEED_CONFIG1 = 0x5EED0001
SEED_CONFIG2 = 0x5EED0002
SEED_CONFIG3 = 0x5EED0003
SEED_CONFIG4 = 0x5EED0004


def random_bytes(n: int, seed: int = SEED_CONFIG2) -> np.ndarray:
    rng = np.random.Generator(np.random.Philox(seed))
    return rng.integers(0, 256, size=n, dtype=np.uint8)


def low_entropy(n: int, seed: int = SEED_CONFIG1) -> np.ndarray:
    rng = np.random.Generator(np.random.Philox(seed))
    p = 2.0 ** -np.arange(1, 17)
    p /= p.sum()
    a = (rng.choice(16, size=n, p=p).astype(np.uint8) + 65)
    # back-references: length 4..64, distance 1..250, roughly every 64 bytes (vectorised per batch of sites)
    if n > 400:
        sites = np.arange(300, n - 70, 64) + rng.integers(0, 32, size=len(np.arange(300, n - 70, 64)))
        lens = rng.integers(4, 65, size=len(sites))
        dists = rng.integers(1, 251, size=len(sites))
        for s, l, d in zip(sites.tolist(), lens.tolist(), dists.tolist()):
            if d >= l:
                a[s:s + l] = a[s - d:s - d + l]
            else:
                for k in range(l):
                    a[s + k] = a[s + k - d]
    return a


def text_like(n: int, seed: int = SEED_CONFIG4) -> np.ndarray:
    rng = np.random.Generator(np.random.Philox(seed))
    vocab = 4096
    wl = rng.integers(2, 11, size=vocab)
    letters = rng.choice(26, size=int(wl.sum()), p=_letter_p()).astype(np.uint8) + 97
    starts = np.concatenate([[0], np.cumsum(wl)[:-1]])
    ranks = rng.zipf(1.25, size=n // 4 + 16) % vocab
    seps = rng.choice(np.frombuffer(b"     ,.\n", dtype=np.uint8), size=len(ranks))
    # token i = word ranks[i] + one separator; tokens are laid end to end until n bytes are covered (vectorised:
    # every output byte looks up its token and its offset inside it)
    tok_len = wl[ranks] + 1
    ends = np.cumsum(tok_len)
    k = min(int(np.searchsorted(ends, n)) + 1, len(ranks))
    tok_len, ends = tok_len[:k], ends[:k]
    total = int(ends[-1])
    tok = np.repeat(np.arange(k, dtype=np.int64), tok_len)
    off = np.arange(total, dtype=np.int64) - (ends - tok_len)[tok]
    is_sep = off == wl[ranks[:k]][tok]
    out = np.where(is_sep, seps[:k][tok], letters[np.minimum(starts[ranks[:k]][tok] + off, len(letters) - 1)]).astype(np.uint8)
    if len(out) < n:
        out = np.concatenate([out, np.full(n - len(out), 32, np.uint8)])
    return out[:n].copy()


def _letter_p():
    p = np.array([8.2, 1.5, 2.8, 4.3, 12.7, 2.2, 2.0, 6.1, 7.0, 0.2, 0.8, 4.0, 2.4, 6.7, 7.5, 1.9, 0.1, 6.0, 6.3, 9.1,
                  2.8, 1.0, 2.4, 0.2, 2.0, 0.1])
    return p / p.sum()


def structured_binary(n: int, seed: int = SEED_CONFIG4 + 1) -> np.ndarray:
    rng = np.random.Generator(np.random.Philox(seed))
    rec = 32
    m = n // rec + 1
    out = np.zeros((m, rec), dtype=np.uint8)
    out[:, 0:4] = np.arange(m, dtype=np.uint32).view(np.uint8).reshape(m, 4)                 # counter
    out[:, 4:8] = (rng.normal(100.0, 5.0, size=m).astype(np.float32)).view(np.uint8).reshape(m, 4)   # float field
    out[:, 8:12] = rng.integers(0, 16, size=(m, 4), dtype=np.uint8)                          # small enums
    out[:, 12:16] = np.cumsum(rng.integers(0, 3, size=m)).astype(np.uint32).view(np.uint8).reshape(m, 4)
    out[:, 16:24] = np.frombuffer(b"RECORD__", dtype=np.uint8)                                # tag
    out[:, 24:28] = rng.integers(0, 256, size=(m, 4), dtype=np.uint8)                        # noise
    out[:, 28:32] = 0
    return out.reshape(-1)[:n].copy()


def mixed(n: int, seed: int = SEED_CONFIG4) -> np.ndarray:
    """50 % text, 25 % structured binary, 25 % random, interleaved in 256 KiB slabs"""
    slab = 256 * 1024
    nslabs = (n + slab - 1) // slab
    kinds = [0, 1, 0, 2]
    parts = []
    for i in range(nslabs):
        k = kinds[i % 4]
        if k == 0:
            parts.append(text_like(slab, seed + 17 * i))
        elif k == 1:
            parts.append(structured_binary(slab, seed + 17 * i))
        else:
            parts.append(random_bytes(slab, seed + 17 * i))
    return np.concatenate(parts)[:n].copy()


def bc_texture(width_blocks: int, height_blocks: int, fmt: int = 3, seed: int = SEED_CONFIG3) -> np.ndarray:
    """BC1 (fmt 1, 8-byte blocks) or BC3 (fmt 3, 16-byte blocks) like data, tight pitch, one mip."""
    rng = np.random.Generator(np.random.Philox(seed))
    h, w = height_blocks, width_blocks
    yy, xx = np.mgrid[0:h, 0:w]
    def smooth565(phase):
        r = ((np.sin(xx / 37.0 + phase) + 1) * 15.5).astype(np.uint16) & 31
        g = ((np.cos(yy / 29.0 + phase) + 1) * 31.5).astype(np.uint16) & 63
        b_ = ((np.sin((xx + yy) / 53.0 + phase) + 1) * 15.5).astype(np.uint16) & 31
        return (r << 11) | (g << 5) | b_
    c0 = smooth565(0.0)
    c1 = smooth565(0.7)
    idx = rng.integers(0, 1 << 32, size=(h, w), dtype=np.uint64).astype(np.uint32)
    idx &= rng.choice(np.array([0xFFFFFFFF, 0x0F0F0F0F, 0x00FF00FF, 0x55555555], dtype=np.uint32), size=(h, w))
    color = np.zeros((h, w, 8), dtype=np.uint8)
    color[..., 0:2] = c0.astype("<u2").view(np.uint8).reshape(h, w, 2)
    color[..., 2:4] = c1.astype("<u2").view(np.uint8).reshape(h, w, 2)
    color[..., 4:8] = idx.astype("<u4").view(np.uint8).reshape(h, w, 4)
    if fmt == 1:
        return color.reshape(-1).copy()
    a0 = ((np.sin(xx / 41.0) + 1) * 127.5).astype(np.uint8)
    a1 = ((np.cos(yy / 31.0) + 1) * 127.5).astype(np.uint8)
    aidx = rng.integers(0, 256, size=(h, w, 6), dtype=np.uint8) & 0x3F
    blk = np.zeros((h, w, 16), dtype=np.uint8)
    blk[..., 0] = a0
    blk[..., 1] = a1
    blk[..., 2:8] = aidx
    blk[..., 8:16] = color
    return blk.reshape(-1).copy()
