"""Generates the committed golden fixtures (run in the development container, where /root/reference exists):

    python tests/golden/make_golden.py

For each small case of tests/corpus.py it writes <name>.brotlig (the stream, from our encoder) and
records in index.json the SHA-256 of the stream and of the output that the UNMODIFIED reference
DecodeCPU (oracle/_ref/libbrotlig_ref.so, built from /root/reference) produces for it -- checked equal
to the source bytes. The fixtures then pin the oracle (and the GPU path) on machines without the reference."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from conftest import Reference, sha256  # noqa: E402
from corpus import corner_cases, texture_cases  # noqa: E402
import brotli_g_sdk_b200 as b  # noqa: E402

ref = Reference(os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref", "libbrotlig_ref.so"))
index = {}
MAX_STREAM = 48 * 1024
cases = {}
for name, (data, kw) in corner_cases().items():
    cases[name] = (b.Encode(data, **kw), data, None)
for name, (data, p) in texture_cases().items():
    cases["tex_" + name] = (b.Encode(data, dcParams=p), data, p)
for name, (s, data, p) in sorted(cases.items()):
    if len(s) > MAX_STREAM:
        continue
    out = ref.decode(s)
    if p is None:
        assert np.array_equal(out, data), name
    s.tofile(os.path.join(HERE, name + ".brotlig"))
    index[name] = {"size": int(len(out)), "stream_sha256": sha256(s), "output_sha256": sha256(out), "preconditioned": p is not None}
json.dump(index, open(os.path.join(HERE, "index.json"), "w"), indent=1, sort_keys=True)
print(len(index), "fixtures,", sum(os.path.getsize(os.path.join(HERE, n + ".brotlig")) for n in index), "bytes")
