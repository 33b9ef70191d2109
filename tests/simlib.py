"""Test helper: run the kernels' device code compiled for the host (tests/sim) -- TEST SCAFFOLDING."""
import ctypes
import os
import subprocess

import numpy as np

from decombinator_b200 import _lib

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "sim", "sim_decombine.cpp")
_SO = os.path.join(_HERE, "sim", "libdcbsim.so")
_ROOT = os.path.dirname(_HERE)


def _build():
    deps = [_SRC, os.path.join(_ROOT, "decombinator_b200", "csrc", "dcr_core.cuh"),
            os.path.join(_ROOT, "decombinator_b200", "csrc", "dcb_tables.h"), os.path.join(_ROOT, "include", "dcb.h")]
    if not os.path.exists(_SO) or any(os.path.getmtime(d) > os.path.getmtime(_SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", "-I", os.path.join(_ROOT, "include"),
                               _SRC, "-o", _SO])
    return _SO


_sim = None


def sim_decombine(packed, vt, jt, both_frames=False, allow_ns=False, lenthreshold=130, general_only=False,
                  use_union=None, use_q=False, use_marks=True, use_half=False, want_deferred2=False):
    """-> (results, counters, n_deferred) using dcr_exact_read/dcr_general_read on the host.

    use_union: None = what dcb_ctx_create does (union index when V and J share the seed geometry).
    use_q: search through the flat kernel's tables (byte filter + offset table of the union index).
    use_half: reads the exact path defers go through the half-tag path (dcr_half_read) first, the general path takes
    what that passes on; want_deferred2 adds the number of reads that reached the general path to the result.
    use_marks: the general path marks candidate keyword positions with the union suffix filter first (what the
    general kernel does); False = its scans visit every position."""
    global _sim
    if _sim is None:
        _sim = ctypes.CDLL(_build())
        _sim.sim_decombine.argtypes = [ctypes.POINTER(_lib.CPacked)] + [ctypes.c_void_p] * 6 + [ctypes.c_int] * 4 + \
                                      [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64), ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64)]
    res = np.zeros(packed.n_reads, dtype=_lib.RESULT_DTYPE)
    cnt = np.zeros(_lib.NCOUNTERS, dtype=np.uint64)
    nd, nd2 = ctypes.c_uint64(), ctypes.c_uint64()
    half = _lib.half_index(vt, jt) if use_half else None
    if use_half and half is None:
        raise ValueError("the chain has no half-tag index")
    union = _lib.union_index(vt, jt) if use_union in (None, True) else None
    if (use_union or use_q) and union is None:
        raise ValueError("V and J do not share a seed geometry")
    sfilt = _lib.suffix_filter(vt, jt) if use_marks else None
    blobs = [vt.blob(0), jt.blob(0), vt.blob(1), jt.blob(1)] + ([union, None] if union is not None else [vt.blob(2), jt.blob(2)])
    rc = _sim.sim_decombine(packed.c, *[b.ctypes.data if b is not None else None for b in blobs], int(both_frames), int(allow_ns),
                            int(lenthreshold), 2 if use_q else int(general_only), res.ctypes.data, cnt.ctypes.data, ctypes.byref(nd),
                            sfilt.ctypes.data if sfilt is not None else None,
                            half.ctypes.data if half is not None else None, ctypes.byref(nd2))
    assert rc == 0
    if want_deferred2:
        return res, cnt, nd.value, nd2.value
    return res, cnt, nd.value


_LEV_SRC = os.path.join(_HERE, "sim", "sim_lev.cpp")
_LEV_SO = os.path.join(_HERE, "sim", "liblevsim.so")
_lev = None


def lev_sim():
    """ctypes handle on csrc/lev_core.cuh compiled for the host (sim_umi_distance, sim_umi_may_be_within, sim_seq_distance)."""
    global _lev
    if _lev is None:
        deps = [_LEV_SRC, os.path.join(_ROOT, "decombinator_b200", "csrc", "lev_core.cuh")]
        if not os.path.exists(_LEV_SO) or any(os.path.getmtime(d) > os.path.getmtime(_LEV_SO) for d in deps):
            subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", _LEV_SRC, "-o", _LEV_SO])
        _lev = ctypes.CDLL(_LEV_SO)
        _lev.sim_umi_distance.argtypes = [ctypes.c_uint64, ctypes.c_uint64]
        _lev.sim_umi_may_be_within.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int]
        _lev.sim_seq_distance.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        _lev.sim_seq_distance_bytes.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    return _lev
