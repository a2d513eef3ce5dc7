"""Python face of the CPU-side Brotli-G encoder (libbrotlig_b200_enc.so).

Mirrors the reference's encode entry points (`BrotliG::MaxCompressedSize`, `BrotliG::Encode`,
/root/reference/inc/BrotligEncoder.h:34-37) closely enough that tests and benchmarks read like the
reference's CLI round trip; the texture parameters mirror `BrotligDataconditionParams`
(/root/reference/inc/common/BrotligDataConditioner.h:29-62).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field

import numpy as np

from ._native import EncOptions, EncStats, enc_lib

BROTLIG_DATA_FORMAT_BC1 = 1
BROTLIG_DATA_FORMAT_BC2 = 2
BROTLIG_DATA_FORMAT_BC3 = 3
BROTLIG_DATA_FORMAT_BC4 = 4
BROTLIG_DATA_FORMAT_BC5 = 5
BLOCK_BYTES = {1: 8, 2: 16, 3: 16, 4: 8, 5: 16}


@dataclass
class DataconditionParams:
    """decode-relevant subset of BrotligDataconditionParams"""
    precondition: bool = False
    swizzle: bool = False
    delta_encode: bool = False
    format: int = 0
    width_blocks: int = 0
    height_blocks: int = 0
    pitch_bytes: int = 0          # 0 => tight (or 256-aligned with pitch_aligned)
    num_mips: int = 1
    pitch_aligned: bool = False

    def texture_size(self) -> int:
        """total bytes of all mips (what `Initialize(inSize)` checks against)"""
        bb = BLOCK_BYTES[self.format]
        w, h = self.width_blocks, self.height_blocks
        wpx, hpx = (w * 4) // 2, (h * 4) // 2
        total = 0
        for mip in range(self.num_mips):
            if mip == 0:
                pitch = self.pitch_bytes or (-(-(w * bb) // 256) * 256 if self.pitch_aligned else w * bb)
            else:
                w, h = (wpx + 3) // 4, (hpx + 3) // 4
                wpx //= 2
                hpx //= 2
                pitch = -(-(w * bb) // 256) * 256 if self.pitch_aligned else w * bb
            total += pitch * h
        return total


def _as_u8(data) -> np.ndarray:
    a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    a = np.ascontiguousarray(a.view(np.uint8).reshape(-1))
    return a


def _options(page_size: int, dc: DataconditionParams | None, **kw) -> EncOptions:
    o = EncOptions()
    enc_lib().bgxenc_default_options(ctypes.byref(o))
    o.page_size = page_size
    if dc is not None and dc.precondition:
        o.precondition = 1
        o.format = dc.format
        o.width_blocks = dc.width_blocks
        o.height_blocks = dc.height_blocks
        o.pitch_bytes = dc.pitch_bytes
        o.num_mips = dc.num_mips
        o.swizzle = int(dc.swizzle)
        o.pitch_aligned = int(dc.pitch_aligned)
        o.delta_encode = int(dc.delta_encode)
    for k, v in kw.items():
        if not hasattr(o, k):
            raise TypeError(f"unknown encoder option {k}")
        setattr(o, k, v)
    return o


def MaxCompressedSize(input_size: int, page_size: int = 65536, precondition: bool = False) -> int:
    return int(enc_lib().bgxenc_max_compressed_size(input_size, page_size, int(precondition)))


def Encode(data, page_size: int = 65536, dcParams: DataconditionParams | None = None, **options) -> np.ndarray:
    """Returns the Brotli-G stream for `data` as a uint8 array (with 16 bytes of zero slack after it,
    not counted in the length: use `stream[:n]`/`len`). Raises ValueError on BROTLIG_ERROR != OK."""
    src = _as_u8(data)
    o = _options(page_size, dcParams, **options)
    cap = MaxCompressedSize(src.size, page_size, bool(o.precondition))
    dst = np.zeros(cap + 16, dtype=np.uint8)
    n = ctypes.c_uint32(cap)
    rc = enc_lib().bgxenc_encode(src.ctypes.data, src.size, dst.ctypes.data, ctypes.byref(n), ctypes.byref(o))
    if rc != 0:
        raise ValueError(f"bgxenc_encode failed with BROTLIG_ERROR {rc}")
    return dst[: n.value].copy()


def Condition(data, dcParams: DataconditionParams) -> np.ndarray:
    """Forward BCn pre-conditioning alone (twin of BrotliG::Condition)."""
    src = _as_u8(data)
    o = _options(65536, dcParams)
    dst = np.zeros(src.size, dtype=np.uint8)
    rc = enc_lib().bgxenc_condition(src.ctypes.data, src.size, dst.ctypes.data, ctypes.byref(o))
    if rc != 0:
        raise ValueError(f"bgxenc_condition failed with {rc}")
    return dst


def last_stats() -> dict:
    s = EncStats()
    enc_lib().bgxenc_last_stats(ctypes.byref(s))
    return {
        "pages": s.pages, "raw_pages": s.raw_pages, "commands": s.commands, "literals": s.literals,
        "ring_code_hits": list(s.ring_code_hits), "implicit_dist0": s.implicit_dist0,
        "insert_only_cmds": s.insert_only_cmds,
        "table_types": [[s.table_types[a][t] for t in range(3)] for a in range(3)],
    }
