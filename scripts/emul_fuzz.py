"""Corruption fuzz of the page kernel on the CPU warp emulator (no GPU needed): takes streams of tests/corpus.py,
corrupts payload bytes, runs of bytes or page-table entries, decodes them with tests/emul/libbgx_emul.so and
requires: no crash, no dead-lock (the emulator aborts on one), nothing written past the output.
usage: emul_fuzz.py <corpus names, comma separated> <seed> <trials per name>   (build the emulator first:
       python -m pytest tests/test_emulated_kernel.py)"""
import sys, os, time, ctypes, subprocess
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,os.path.join(ROOT,'tests')); sys.path.insert(0,ROOT)
import numpy as np
import brotli_g_sdk_b200 as b
from corpus import corner_cases
lib=ctypes.CDLL(os.path.join(ROOT,'tests','emul','libbgx_emul.so'))
lib.emul_decode_stream.restype=ctypes.c_int
cc=corner_cases()
names=sys.argv[1].split(',')
seed=int(sys.argv[2]); trials=int(sys.argv[3])
rng=np.random.default_rng(seed)
stats={}
for name in names:
    d,kw=cc[name]; d=d[:200000]
    s=b.Encode(d,**kw)
    n=int(s[2])|int(s[3])<<8
    if len(s) < 8+4*n+24: continue      # too small to corrupt meaningfully
    for t in range(trials):
        bad=s.copy()
        mode=int(rng.integers(0,3))
        if mode==0:   # flip a few bytes in the payload
            for _ in range(int(rng.integers(1,4))):
                k=int(rng.integers(8+4*n, len(bad)))
                bad[k]^=int(rng.integers(1,256))
        elif mode==1: # clobber a run
            k=int(rng.integers(8+4*n, len(bad)-8)); L=int(rng.integers(1,64))
            bad[k:k+L]=rng.integers(0,256,len(bad[k:k+L]),dtype=np.uint8)
        else:         # corrupt a page-table entry
            k=8+4*int(rng.integers(0,n)); bad[k:k+4]=rng.integers(0,256,4,dtype=np.uint8)
        size=len(d)
        out=np.full(size+64,0xEE,np.uint8); st=(ctypes.c_uint32*max(n,1))(); fl=(ctypes.c_uint32*max(n,1))(); coll=ctypes.c_uint64()
        print(name,t,mode,flush=True)
        rc=lib.emul_decode_stream(ctypes.c_void_p(bad.ctypes.data), ctypes.c_uint32(len(bad)), ctypes.c_void_p(out.ctypes.data), ctypes.c_uint32(size), st, fl, ctypes.byref(coll))
        assert (out[size:]==0xEE).all(), "wrote past the output"
        stats[rc]=stats.get(rc,0)+1
print("DONE",stats)
