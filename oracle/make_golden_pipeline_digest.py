#!/usr/bin/env python
"""Record digests of what the UNMODIFIED reference produces for decombine + collapse at a size the fixtures cannot hold row by
row (BASELINE configs[3] shape: reads that are copies of UMI-tagged molecules), as tests/golden/pipeline_digest.json.

Build container only (the reference imported as-is with the stand-ins of oracle/standins/).  The paired FASTQ is
regenerated from its seed wherever the test runs (oracle/synth_pairs.py), so only the SHA-256 of the .n12 text, of the .freq
rows and a few counters are committed.

usage: python oracle/make_golden_pipeline_digest.py
"""
import hashlib
import json
import os
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import refenv  # noqa: E402
import synth_pairs  # noqa: E402

SPECS = [
    # name, species, tags, chain, n reads, molecules, L, sub1, sub2, seed, oligo
    ("human_b_120k", "human", "extended", "b", 120_000, 6_000, 250, 0.004, 0.01, 20260500, "M13"),
    ("mouse_a_60k", "mouse", "original", "a", 60_000, 2_000, 250, 0.006, 0.015, 20260501, "M13"),
]


def n12_text(rows):
    return "".join(", ".join(map(str, r[:10])) + "\n" for r in rows)


def freq_text(rows):
    return "".join(", ".join(map(str, r)) + "\n" for r in rows)


def main():
    ref = refenv.load()
    out = {"how": "oracle/make_golden_pipeline_digest.py: unmodified reference + stand-ins, decombinator() then collapsinator(data=rows)",
           "cases": []}
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp(prefix="dcbgoldp"))
    try:
        for name, species, tagset, chain, n, pool, L, sub1, sub2, seed, oligo in SPECS:
            synth_pairs.write_pairs("s_1.fq", "s_2.fq", species, tagset, chain, n, pool, L, sub1, sub2, seed)
            args = ref["io"].create_args_dict(infile="s_1.fq", chain=chain, bc_read="R2", dontcount=True, suppresssummary=True,
                                              dontcheck=True, tagfastadir=refenv.REF_TAGDIR, outpath="", species=species, tags=tagset,
                                              oligo=oligo, command="pipeline")
            t0 = time.time()
            rows = ref["decombine"].decombinator(dict(args))
            t1 = time.time()
            freq = ref["collapse"].collapsinator(dict(args), data=[list(r) for r in rows])
            t2 = time.time()
            case = {"name": name, "spec": [species, tagset, chain, n, pool, L, sub1, sub2, seed, oligo],
                    "n12_rows": len(rows), "n12_sha256": hashlib.sha256(n12_text(rows).encode()).hexdigest(),
                    "freq_rows": len(freq), "freq_sha256": hashlib.sha256(freq_text(freq).encode()).hexdigest(),
                    "freq_head": [list(map(str, r)) for r in freq[:3]],
                    "reference_seconds": {"decombine": round(t1 - t0, 1), "collapse": round(t2 - t1, 1)}}
            print(json.dumps(case)[:400])
            out["cases"].append(case)
    finally:
        os.chdir(cwd)
    path = os.path.join(ROOT, "tests", "golden", "pipeline_digest.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
