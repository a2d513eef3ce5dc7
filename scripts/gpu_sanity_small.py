"""Tiny decode through the C ABI without torch (for compute-sanitizer runs)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import brotli_g_sdk_b200 as b
from brotli_g_sdk_b200 import datagen
dec = b.BrotligDecoder(0)
ok = True
for name, d in [("low", datagen.low_entropy(200000, seed=1)), ("text", datagen.text_like(150000, seed=2)),
                ("const", np.full(70000, 9, np.uint8)), ("rand", datagen.random_bytes(70001, seed=3)),
                ("bin", datagen.structured_binary(131072 + 77, seed=4))]:
    s = b.Encode(d)
    out, ms = dec.decode_host(s)
    good = bool(np.array_equal(out, d))
    ok &= good
    print(name, len(d), len(s), good, ms)
print("SANITY", "OK" if ok else "FAIL")
sys.exit(0 if ok else 1)
