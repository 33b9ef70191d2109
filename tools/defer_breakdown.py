"""Diagnostic: which reads the flat kernel's per-read logic defers, by reason (host build of csrc/dcr_core.cuh)."""
import collections
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np
import simlib
from decombinator_b200 import _lib, tags
from helpers import synth_batch

chain = sys.argv[1] if len(sys.argv) > 1 else "b"
sub = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
nrate = float(sys.argv[3]) if len(sys.argv) > 3 else 0.001
info = tags.load("human", "extended", chain)
vt, jt = info.tables()
n, L = 200000, 250
r1, off, ln = synth_batch(info, n, L, sub, nrate, 0.0, seed=20260003)
packed = _lib.pack_arrays(r1, off, ln, revcomp=True)
simlib.sim_decombine(packed, vt, jt, use_q=True)     # builds / loads the sim
sim = simlib._sim
cls = np.zeros(n, dtype=np.uint8)
u = _lib.union_index(vt, jt)
sim.sim_defer_classes.argtypes = [ctypes.POINTER(_lib.CPacked)] + [ctypes.c_void_p] * 3 + [ctypes.c_int] * 3 + [ctypes.c_void_p]
sim.sim_defer_classes(packed.c, vt.blob(1).ctypes.data, jt.blob(1).ctypes.data, u.ctypes.data, 0, 0, 130, cls.ctypes.data)
c = collections.Counter(cls.tolist())
names = {1: "noV", 2: "noJ", 4: "flagged", 8: "multi", 16: "bothfound"}
print("deferred %.2f %%" % (100.0 * (cls != 0).mean()))
for k, v in sorted(c.items(), key=lambda kv: -kv[1]):
    print("%6.2f %%  %s" % (100.0 * v / n, "+".join(nm for b, nm in names.items() if k & b) or "done"))
