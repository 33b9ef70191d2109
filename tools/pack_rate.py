#!/usr/bin/env python
"""Host packer rate (dcb_pack_words, GB/s of text) per unit and store mode; each variant in a process of its own because the
unit is chosen when the library loads.  usage: python tools/pack_rate.py"""
import os
import subprocess
import sys

CHILD = r'''
import numpy as np, time, os, sys, ctypes
sys.path.insert(0, %r)
from decombinator_b200 import _lib
rng = np.random.default_rng(1)
L, n = 250, 4_000_000
buf = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=L * n)].copy()
out = _lib.PinnedBytes(np.zeros(n * 64, dtype=np.uint8))
clean = ctypes.c_int(0)
for nt in (1, 8, 16):
    best = 0
    for _ in range(4):
        t0 = time.perf_counter()
        _lib.lib().dcb_pack_words(buf.ctypes.data, None, None, 0, n, L, 1, 16, out.a.ctypes.data, nt, ctypes.byref(clean))
        best = max(best, L * n / (time.perf_counter() - t0) / 1e9)
    print("   threads %%2d: %%6.1f GB/s of text (clean %%d)" %% (nt, best, clean.value), flush=True)
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

for name, env in (("AVX2", {"DCB_NO_AVX512": "1"}), ("AVX-512, direct stores", {}), ("AVX-512, local buffer", {"DCB_PACK_STORE": "1"}),
                  ("AVX-512, streaming stores", {"DCB_PACK_STORE": "2"})):
    print(name, flush=True)
    e = dict(os.environ)
    e.update(env)
    subprocess.run([sys.executable, "-c", CHILD], env=e, check=False)
