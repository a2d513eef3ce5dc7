"""CPU tests (-m "not gpu"): the oracle restatement is pinned against the unmodified reference decoder
and against the committed golden fixtures; the corpus is shown to reach the corner cases it claims."""
import json
import os

import numpy as np
import pytest

from conftest import sha256
from corpus import corner_cases, texture_cases

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_oracle_matches_reference_and_source(sdk, oracle, reference):
    for name, (data, kw) in corner_cases().items():
        s = sdk.Encode(data, **kw)
        ref = reference.decode(s)
        orc = oracle.decode(s)
        assert np.array_equal(ref, data), f"{name}: reference != source"
        assert np.array_equal(orc, ref), f"{name}: oracle != reference"


def test_oracle_matches_reference_on_textures(sdk, oracle, reference):
    for name, (data, p) in texture_cases().items():
        s = sdk.Encode(data, dcParams=p)
        assert (s[6] >> 4) & 1 == 1, f"{name}: stream is not flagged preconditioned"
        ref = reference.decode(s)
        orc = oracle.decode(s)
        # bytes of pitch padding are not part of the texture: both decoders leave them 0
        assert np.array_equal(orc, ref), f"{name}: oracle != reference"
        assert _texture_equal(ref, data, p), f"{name}: reference != source on block bytes"


def _texture_equal(out, data, p):
    """compare only block bytes (rows are width_blocks*block_bytes wide inside pitch_bytes)"""
    from brotli_g_sdk_b200.encoder import BLOCK_BYTES
    bb = BLOCK_BYTES[p.format]
    w, h = p.width_blocks, p.height_blocks
    wpx, hpx = (w * 4) // 2, (h * 4) // 2
    off = 0
    ok = True
    for mip in range(p.num_mips):
        if mip:
            w, h = (wpx + 3) // 4, (hpx + 3) // 4
            wpx //= 2
            hpx //= 2
        tight = w * bb
        pitch = (p.pitch_bytes if (mip == 0 and p.pitch_bytes) else (-(-tight // 256) * 256 if p.pitch_aligned else tight))
        a = out[off: off + pitch * h].reshape(h, pitch)
        b = data[off: off + pitch * h].reshape(h, pitch)
        ok &= bool(np.array_equal(a[:, :tight], b[:, :tight]))
        ok &= bool((a[:, tight:] == 0).all())
        off += pitch * h
    return ok


def test_corpus_reaches_the_corner_cases(sdk, oracle):
    """the generator must hit these densely; the oracle's counters prove it."""
    tot = {"dist": np.zeros(16, np.int64), "implicit": 0, "insert_only": 0, "overlap": 0, "types": np.zeros((3, 3), np.int64),
           "rle16": 0, "rle17": 0, "max_ins": 0, "max_copy": 0, "raw": 0, "rounds": 0, "carry": 0}
    for name, (data, kw) in corner_cases().items():
        s = sdk.Encode(data, **kw)
        oracle.decode(s)
        st = oracle.stats()
        tot["dist"] += np.array(list(st.dist_code_hist), dtype=np.int64)
        tot["implicit"] += st.implicit_dist0
        tot["insert_only"] += st.insert_only
        tot["overlap"] += st.overlap_copies
        tot["types"] += np.array([[st.table_types[a][t] for t in range(3)] for a in range(3)], dtype=np.int64)
        tot["rle16"] += st.rle16
        tot["rle17"] += st.rle17
        tot["max_ins"] = max(tot["max_ins"], st.max_insert_len)
        tot["max_copy"] = max(tot["max_copy"], st.max_copy_len)
        tot["raw"] += st.raw_pages
        tot["rounds"] += st.rounds
        tot["carry"] += st.literals_decoded - st.literals_emitted
    assert (tot["dist"] >= 20).all(), f"every distance short code 0..15 must be hit densely: {tot['dist']}"
    assert tot["implicit"] > 1000 and tot["insert_only"] > 1000 and tot["overlap"] > 100
    assert tot["types"][2][0] > 0 and tot["types"][2][1] >= 4 and tot["types"][2][2] > 0, "literal table: trivial, simple, complex"
    assert tot["types"][1][0] > 0 and tot["types"][1][2] > 0, "distance table: trivial and complex"
    assert tot["types"][0][1] > 0 and tot["types"][0][2] > 0, "command table: simple and complex"
    assert tot["rle16"] > 0 and tot["rle17"] > 0
    assert tot["max_ins"] >= 22594, "24-extra-bit insert length"
    assert tot["max_copy"] >= 2118, "24-extra-bit copy length"
    assert tot["raw"] > 0 and tot["rounds"] > 1000 and tot["carry"] > 0


def test_header_errors(oracle, sdk):
    s = sdk.Encode(np.arange(5000, dtype=np.uint8) % 7)
    bad = s.copy(); bad[1] ^= 0x10
    out = np.zeros(5000 + 16, np.uint8)
    import ctypes
    n = ctypes.c_uint32(5000)
    assert oracle.lib.bgo_decode(len(bad), bad.ctypes.data, ctypes.byref(n), out.ctypes.data) == 14   # BROTLIG_ERROR_CORRUPT_STREAM
    bad = s.copy(); bad[0] = 6; bad[1] = 6 ^ 0xFF
    assert oracle.lib.bgo_decode(len(bad), bad.ctypes.data, ctypes.byref(n), out.ctypes.data) == 15   # ..._INCORRECT_STREAM_FORMAT


def test_stream_header_bytes_match_the_survey_probe(sdk):
    """3 pages, last page 1000 bytes, 64 KiB pages -> 05 fa 03 00 a1 0f 00 00"""
    s = sdk.Encode(np.zeros(2 * 65536 + 1000, np.uint8))
    assert bytes(s[:8]) == bytes([0x05, 0xFA, 0x03, 0x00, 0xA1, 0x0F, 0x00, 0x00])


def test_golden_fixtures(oracle):
    """tests/golden/*.brotlig were produced by make_golden.py; their outputs were verified with the
    unmodified reference decoder at generation time and are pinned here by SHA-256."""
    idx = json.load(open(os.path.join(GOLDEN, "index.json")))
    assert len(idx) >= 10
    for name, meta in idx.items():
        s = np.fromfile(os.path.join(GOLDEN, name + ".brotlig"), dtype=np.uint8)
        assert sha256(s) == meta["stream_sha256"]
        out = oracle.decode(s)
        assert len(out) == meta["size"] and sha256(out) == meta["output_sha256"], name


def test_forward_conditioner_matches_reference(sdk):
    import ctypes
    p = os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "_ref", "libbrotlig_refcond.so")
    if not os.path.exists(p):
        pytest.skip("oracle/_ref/libbrotlig_refcond.so not built")
    lib = ctypes.CDLL(p, mode=ctypes.RTLD_LOCAL)
    for name, (data, prm) in texture_cases().items():
        mine = sdk.Condition(data, prm)
        ref = np.zeros_like(data)
        from brotli_g_sdk_b200.encoder import BLOCK_BYTES
        tight = prm.width_blocks * BLOCK_BYTES[prm.format]
        pitch0 = prm.pitch_bytes or (-(-tight // 256) * 256 if prm.pitch_aligned else tight)
        rc = lib.refcond_condition(len(data), data.ctypes.data, ref.ctypes.data, prm.format, prm.width_blocks, prm.height_blocks,
                                   pitch0, prm.num_mips, int(prm.swizzle), int(prm.pitch_aligned))
        assert rc == 0, name
        assert np.array_equal(mine, ref), f"{name}: forward conditioner differs from BrotliG::Condition"
