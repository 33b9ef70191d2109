#!/usr/bin/env python
"""A few passes of the decombine kernels over ONE workload block of bench.py (cfg2 / cfg4, one chain), batch resident in
HBM -- the target of `ncu -k regex:dcb_halftag ...` when a kernel of a workload block is profiled (GPU box).

    python tools/profile_workload.py --cfg 2 --chain a [--reads 4000000] [--passes 3]

Prints the per-kernel device times (CUDA events of the library) and how many reads each tier saw."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

SHAPES = {2: ("human", "extended", ("a", "b"), 20260003, 0.01, 0.001),
          4: ("mouse", "original", ("g", "d"), 20260005, 0.005, 0.0)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", type=int, default=2, choices=sorted(SHAPES))
    ap.add_argument("--chain", default=None)
    ap.add_argument("--reads", type=int, default=4_000_000)
    ap.add_argument("--passes", type=int, default=3)
    ap.add_argument("--force-general", type=int, default=0)
    args = ap.parse_args()
    from decombinator_b200 import _lib, tags
    species, tagset, chains, seed, sub, nrate = SHAPES[args.cfg]
    chain = args.chain or chains[0]
    infos = [tags.load(species, tagset, c) for c in chains]
    n, L = args.reads, 250
    syn = _lib.Synth([(i.v_regions, i.j_regions) for i in infos], seed, L, 0, sub, nrate, 0.0)
    r1, _ = syn.reads(0, n, n_threads=os.cpu_count() or 8)
    off = np.arange(n, dtype=np.uint64) * L
    ln = np.full(n, L, dtype=np.uint32)
    packed = _lib.pack_arrays(r1, off, ln, revcomp=True, n_threads=os.cpu_count() or 8)
    vt, jt = infos[chains.index(chain)].tables()
    ctx = _lib.Context(vt, jt, device=0, force_general=args.force_general)
    ctx.upload(packed)
    ctx.run_resident()
    ctx.timing_enable(True); ctx.timing_reset()
    for _ in range(args.passes):
        ctx.run_resident()
    res, cnt = ctx.download()
    kms, kl = ctx.timing_get()
    print(json.dumps({"cfg": args.cfg, "chain": chain, "reads": n,
                      "kernels_ms": {ctx.exact_kernel_name(): kms[0] / max(1, kl[0]), "dcb_halftag_kernel": kms[2] / max(1, kl[2]),
                                     "dcb_general_kernel": kms[1] / max(1, kl[1])},
                      "queued": ctx.last_deferred() / n, "general": ctx.last_general() / n,
                      "decombined": float(res["status"].mean())}))


if __name__ == "__main__":
    main()
