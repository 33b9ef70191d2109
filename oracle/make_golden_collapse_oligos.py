#!/usr/bin/env python
"""Record the UNMODIFIED reference's collapse stage for the oligo designs beside M13 -- I8, I8_single, NEBIO, TAKARA
(collapse.py:172-189, 367-479) -- as tests/golden/collapse_cases_oligos.json.gz.  Same recipe as make_golden_collapse.py
(build container only; the reference imported as-is with the stand-ins): read 2 of the synthetic molecules is re-laid
out for each design, so that exact and fuzzy spacers, short / long N1, Ns, low-quality barcodes and collisions all occur
for every one of them.  Each case also records the reference's collapse counters.

usage: python oracle/make_golden_collapse_oligos.py
"""
import gzip
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import refenv  # noqa: E402
import make_golden_collapse as G  # noqa: E402

# the generator writes read 2 as: M13 spacer (22), N6, I8 spacer (8), N6, the rest
LAYOUTS = {
    # oligo: (read 2 of that design from the generator's read 2, where its first hexamer starts)
    "I8": (lambda b: "GTCGTGAT" + b[22:] + b[:14][::-1], 8),
    "I8_single": (lambda b: b[22:28] + "ATCACGAC" + b[36:] + b[:22][::-1], 0),
    "NEBIO": (lambda b: b[22:28] + b[36:42] + b[42:47] + "A" + "TACGGG" + b[47:] + b[:23][::-1], 0),
    "TAKARA": (lambda b: b[22:28] + b[36:42] + "GTACGGG" + b[42:] + b[:23][::-1], 0),
}


def main():
    ref = refenv.load()
    C = ref["collapse"]
    cases = []
    tmp = tempfile.mkdtemp(prefix="dcbgoldo")
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        spec = [
            # oligo, species, tags, chain, n reads, pool, L, sub1, sub2, extra collapse args
            ("I8", "human", "extended", "b", 1200, 100, 250, 0.004, 0.01, {}),
            ("I8", "human", "original", "a", 900, 60, 250, 0.006, 0.02, {"bcthreshold": 1, "allowNs": True, "minbcQ": 10}),
            ("I8_single", "human", "extended", "b", 900, 80, 250, 0.004, 0.01, {}),
            ("NEBIO", "human", "extended", "a", 800, 60, 250, 0.004, 0.01, {"bcthreshold": 3}),
            ("TAKARA", "mouse", "original", "b", 800, 60, 250, 0.004, 0.01, {"percentlevdist": 5}),
        ]
        for ci, (oligo, species, tagset, chain, n, pool, L, sub1, sub2, extra) in enumerate(spec):
            layout, n1_at = LAYOUTS[oligo]
            fix = lambda b, f=layout, L=L: f(b)[:L]          # noqa: E731
            rows, args = G.make_rows(ref, species, tagset, chain, n, pool, L, sub1, sub2, 20260400 + ci, oligo, r2_layout=fix, n1_at=n1_at)
            cases.append(G.record_case(C, rows, args, extra, (ci, oligo, species, tagset, chain)))
    finally:
        os.chdir(cwd)
    out = os.path.join(ROOT, "tests", "golden", "collapse_cases_oligos.json.gz")
    with gzip.GzipFile(out, "wb", mtime=0) as fh:
        fh.write(json.dumps({"cases": cases}).encode())
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
