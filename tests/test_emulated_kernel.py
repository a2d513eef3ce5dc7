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
