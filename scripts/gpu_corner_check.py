"""Which corpus cases decode bit-exact with the library selected by BGX_CUDA_LIB (kernel experiments)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import brotli_g_sdk_b200 as b
from corpus import corner_cases
dec = b.BrotligDecoder(0)
bad = []
for name, (data, kw) in corner_cases().items():
    s = b.Encode(data, **kw)
    try:
        out, _ = dec.decode_host(s)
        ok = bool(np.array_equal(out, data))
        if not ok:
            d = np.nonzero(out != data)[0]
            bad.append(f"{name}: {len(d)} bytes differ, first at {d[:3].tolist()} of {len(data)}")
    except Exception as e:
        bad.append(f"{name}: {e}")
print(os.path.basename(os.environ.get("BGX_CUDA_LIB", "default")), "FAILED:" if bad else "ALL OK", "; ".join(bad))
