"""CPU tests (-m "not gpu"): the CUDA device code (page_decode.cuh), compiled for the CPU warp
emulator, against the oracle. This is how kernel logic is checked where there is no GPU; the GPU
parity tests (test_gpu_parity.py) run the real kernels through the C ABI."""
import numpy as np

from corpus import corner_cases, texture_cases


def test_emulated_kernel_matches_oracle(sdk, oracle, emulator):
    for name, (data, kw) in corner_cases().items():
        if len(data) > 320000:
            data = data[:320000]
        s = sdk.Encode(data, **kw)
        want = oracle.decode(s)
        got, status, _ = emulator.decode(s)
        assert all(x == 0 for x in status), f"{name}: page status {status}"
        assert np.array_equal(got, want), f"{name}: emulated kernel != oracle (first diff at {np.nonzero(got != want)[0][:4]})"


def test_emulated_kernel_unaligned_output(sdk, oracle, emulator):
    """the caller's buffer may start anywhere: byte-granular flush and far-match paths"""
    from brotli_g_sdk_b200 import datagen
    data = np.concatenate([datagen.text_like(150000, seed=31), np.zeros(5000, np.uint8), datagen.structured_binary(60000, seed=32)])
    s = sdk.Encode(data)
    want = oracle.decode(s)
    for off in (1, 2, 3, 4, 8):
        got, status, _ = emulator.decode(s, dst_offset=off)
        assert all(x == 0 for x in status), f"offset {off}: page status {status}"
        assert np.array_equal(got, want), f"offset {off}: emulated kernel != oracle"


def test_emulated_kernel_textures(sdk, oracle, emulator):
    for name, (data, p) in texture_cases().items():
        s = sdk.Encode(data, dcParams=p)
        want = oracle.decode(s)
        got, status, flags = emulator.decode(s)
        assert all(x == 0 for x in status), f"{name}: page status {status}"
        assert np.array_equal(got, want), f"{name}: emulated texture path != oracle"
        if p.delta_encode:
            assert any(flags), f"{name}: no page took the delta path"


def test_emulated_kernel_rejects_garbage_without_writing_out_of_bounds(sdk, emulator):
    """corrupt payloads must end in a page status, never in an out-of-bounds access (the guard bytes
    around the output are checked inside Emulator.decode)"""
    rng = np.random.default_rng(5)
    data = np.tile(rng.integers(0, 256, 997, dtype=np.uint8), 80)
    s = sdk.Encode(data)
    assert len(s) < len(data) // 2
    n_pages = int(s[2]) | (int(s[3]) << 8)
    for trial in range(18):
        bad = s.copy()
        if trial % 3 == 2:      # a corrupt page-table entry: offsets/sizes that point anywhere
            k = 8 + 4 * int(rng.integers(0, n_pages))
        else:                   # corrupt payload bytes
            k = int(rng.integers(8 + 4 * n_pages, len(bad) - 4))
        bad[k: k + 4] ^= rng.integers(1, 256, 4, dtype=np.uint8)
        try:
            emulator.decode(bad, expect_rc=0)
        except AssertionError as e:
            assert "rc 14" in str(e) or "status" in str(e), str(e)


def test_prefix_code_descriptions_must_be_complete(emulator):
    """Kraft sum of the code lengths: complete codes and single-symbol codes are accepted, over- and under-subscribed
    ones end in kPageErrTable (the reference trusts them: BrotligHuffmanTable.cpp:44-71,135-145)"""
    import ctypes

    def status(lens):
        a = np.zeros(728, np.uint8)
        a[: len(lens)] = lens
        emulator.lib.emul_table_status.restype = ctypes.c_uint32
        return int(emulator.lib.emul_table_status(ctypes.c_void_p(a.ctypes.data), ctypes.c_uint32(728)))

    assert status([1, 1]) == 0
    assert status([1, 2, 3, 3]) == 0
    assert status([2, 2, 2, 2]) == 0
    assert status([15] * 2 + [14] + [13] + [12] + [11] + [10] + [9] + [8] + [7] + [6] + [5] + [4] + [3] + [2] + [1]) == 0
    assert status([3]) == 0                      # a single symbol
    assert status([1, 1, 1]) == 8                # over-subscribed
    assert status([1, 2, 2, 2]) == 8
    assert status([2, 2, 2]) == 8                # incomplete
    assert status([1, 3]) == 8


def test_texture_header_with_a_lying_pitch_is_rejected(sdk, emulator):
    """PreconditionHeader fields are untrusted: a pitch smaller than a row of blocks (or sizes that do not add up) must
    not reach the de-conditioning scatter. 16-byte header: BC1, 64 x 64 blocks, pitch 8, 512 bytes, one raw page."""
    import ctypes
    w0 = 5 | (0xFA << 8) | (1 << 16)                                   # id, magic, 1 page
    w1 = 1 | (512 << 2) | (1 << 20)                                    # 64 KiB pages, last page 512 bytes, preconditioned
    p0 = (63 << 2) | (63 << 17)                                        # 64 x 64 blocks
    p1 = 1 | (0 << 8) | ((8 - 1) << 13)                                # BC1, 1 mip, pitch 8 bytes (a row needs 512)
    hdr = np.array([w0, w1, p0, p1], dtype="<u4").view(np.uint8)
    stream = np.concatenate([hdr, np.array([512], dtype="<u4").view(np.uint8), np.zeros(512, np.uint8)])
    out = np.zeros(512 + 64, np.uint8)
    st = (ctypes.c_uint32 * 1)()
    fl = (ctypes.c_uint32 * 1)()
    coll = ctypes.c_uint64(0)
    rc = emulator.lib.emul_decode_stream(ctypes.c_void_p(stream.ctypes.data), ctypes.c_uint32(len(stream)), ctypes.c_void_p(out.ctypes.data),
                                         ctypes.c_uint32(512), st, fl, ctypes.byref(coll))
    assert rc == 14, rc            # BROTLIG_ERROR_CORRUPT_STREAM
    assert not out.any()


def _random_complete_code(rng, n_used, max_len=15):
    """code lengths of a complete prefix code with n_used symbols: split leaves of a binary tree at random"""
    lens = [0]
    while len(lens) < n_used:
        cand = [i for i, l in enumerate(lens) if l < max_len]
        if not cand:
            break
        i = cand[int(rng.integers(0, len(cand)))]
        l = lens.pop(i)
        lens += [l + 1, l + 1]
    return lens


def test_build_table_against_a_canonical_reference(emulator):
    """LUT, sorted[], limit[] and base[] of build_table for random complete prefix codes of every shape (few / many
    symbols, all lengths up to 15, gaps in the alphabet), against a plain canonical-code construction
    (GenerateHuffmanTable, BrotligHuffmanTable.cpp:44-71: codes per length in symbol order)."""
    import ctypes
    rng = np.random.default_rng(2024)
    f = emulator.lib.emul_build_table
    f.restype = ctypes.c_uint32
    cases = []
    for kind, alphabet, bits in ((0, 728, 9), (2, 256, 10)):
        for n_used in (2, 3, 5, 17, 64, 200, alphabet):
            for rep in range(3):
                cases.append((kind, alphabet, bits, _random_complete_code(rng, n_used)))
        cases.append((kind, alphabet, bits, [8] * 256))                  # flat
        cases.append((kind, alphabet, bits, [1] + list(range(2, 16)) + [15]))   # one code of every length
        cases.append((kind, alphabet, bits, [15] * 2 + list(range(14, 0, -1))))
    for kind, alphabet, bits, code in cases:
        lens = np.zeros(alphabet, np.uint8)
        where = np.sort(rng.choice(alphabet, size=len(code), replace=False))
        lens[where] = rng.permutation(code)
        lut = np.zeros(1 << bits, np.uint16)
        srt = np.zeros(alphabet, np.uint16)
        limit = np.zeros(16, np.uint16)
        base = np.zeros(16, np.uint16)
        st = f(ctypes.c_void_p(lens.ctypes.data), alphabet, kind, ctypes.c_void_p(lut.ctypes.data), ctypes.c_void_p(srt.ctypes.data),
               ctypes.c_void_p(limit.ctypes.data), ctypes.c_void_p(base.ctypes.data))
        assert st == 0, (kind, code)
        # reference construction
        order = sorted((int(l), int(s)) for s, l in enumerate(lens) if l)
        cnt = [0] * 16
        for l, _ in order:
            cnt[l] += 1
        first, off, c, o = [0] * 16, [0] * 16, 0, 0
        for L in range(1, 16):
            c = (c + cnt[L - 1]) << 1
            first[L], off[L] = c, o
            o += cnt[L]
        assert [s for _, s in order] == list(srt[: len(order)]), "sorted[]"
        for L in range(1, 16):
            assert int(limit[L]) == min(0x8000, (first[L] + cnt[L]) << (15 - L)), ("limit", L)
            assert int(base[L]) == (off[L] - first[L]) & 0xffff, ("base", L)
        want = np.zeros(1 << bits, np.uint16)
        nxt = list(first)
        for l, s in order:
            codeword = nxt[l]
            nxt[l] += 1
            if l <= bits:
                rev = int(format(codeword, "0%db" % l)[::-1], 2)
                want[rev:: 1 << l] = s | (l << 10)
            else:
                rev = int(format(codeword >> (l - bits), "0%db" % bits)[::-1], 2)
                want[rev] = 0xffff
        assert np.array_equal(lut, want), ("lut", kind, sorted(code))
