#!/usr/bin/env python
"""Record what the UNMODIFIED reference's collapse stage returns, as tests/golden/collapse_cases.json.gz.

Runs only in the build container (needs /root/reference, imported as-is through oracle/refenv.py with the stand-ins
of oracle/standins/).  Inputs are .n12 rows produced by the reference's own decombinator() from synthetic paired
FASTQ in which reads are copies of a pool of molecules (shared UMI + rearrangement) with sequencing errors in both
reads, so that exact-barcode grouping, proto-sequence changes, multi-TCR barcodes, fuzzy spacers, short/long N1,
low-quality barcodes and UMI merges all occur.

usage: python oracle/make_golden_collapse.py
"""
import collections as coll
import gzip
import json
import os
import random
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import refenv  # noqa: E402
import synth_pairs  # noqa: E402
from decombinator_b200 import _lib, tags as dtags  # noqa: E402


def make_rows(ref, species, tagset, chain, n, pool, L, sub1, sub2, seed, oligo, r2_layout=None, n1_at=22):
    """r2_layout(read 2 as the generator wrote it: M13 spacer, N6, I8 spacer, N6, ...) -> read 2 of another oligo design
    (make_golden_collapse_oligos.py); n1_at: where that design's first random hexamer starts."""
    synth_pairs.write_pairs("s_1.fq", "s_2.fq", species, tagset, chain, n, pool, L, sub1, sub2, seed, r2_layout, n1_at)
    args = ref["io"].create_args_dict(infile="s_1.fq", chain=chain, bc_read="R2", dontcount=True, suppresssummary=True,
                                      dontcheck=True, tagfastadir=refenv.REF_TAGDIR, outpath="", species=species, tags=tagset,
                                      oligo=oligo, command="pipeline")
    rows = ref["decombine"].decombinator(args)
    return rows, args


def record_case(C, rows, args, extra, label):
    """One recorded run of the reference's collapse stage over `rows`: every intermediate and the .freq rows."""
    args = dict(args)
    args.update(extra)
    # the stages separately (what the reference's own unit tests call) ...
    C.counts = coll.Counter()
    qp = [args["minbcQ"], args["bcQbelowmin"], args["avgQthreshold"]]
    frac = args["percentlevdist"] / 100
    groups = C.read_in_data([list(r) for r in rows], args, qp, frac, True, open)
    read_counts = {k: v for k, v in sorted(C.counts.items()) if not k.startswith("time_") and isinstance(v, int)}
    group_keys = list(groups.keys())
    group_sizes = [len(v) for v in groups.values()]
    _, blist, umi_proto = C.create_clustering_objs(groups)
    matches = C.make_merge_groups(umi_proto, args["bcthreshold"], True)
    pairs = [[int(i), int(j)] for i, j in zip(matches.row, matches.col)]
    protos = [k.split("|")[2] for k in group_keys]
    verdicts = [bool(C.are_seqs_equivalent(protos[i], protos[j], frac)) for i, j in pairs]
    clusters = C.make_clusters(matches, blist, frac)
    cluster_keys = list(clusters.keys())
    cluster_sizes = [len(v) for v in clusters.values()]
    # ... and the whole stage
    out = C.collapsinator(dict(args), data=[list(r) for r in rows])
    print("case", *label, "rows", len(rows), "groups", len(group_keys), "pairs", len(pairs),
          "merged", sum(verdicts), "clusters", len(cluster_keys), "out", len(out))
    return {"args": {k: v for k, v in args.items() if k != "tagfastadir"}, "rows": rows,
            "group_keys": group_keys, "group_sizes": group_sizes, "umis": [u for u, _ in umi_proto],
            "pairs": pairs, "verdicts": verdicts, "cluster_keys": cluster_keys, "cluster_sizes": cluster_sizes,
            "freq": out, "counts_read_in_data": read_counts}


def main():
    ref = refenv.load()
    C = ref["collapse"]
    cases = []
    tmp = tempfile.mkdtemp(prefix="dcbgoldc")
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        spec = [
            # species, tags, chain, n reads, pool, L, sub1, sub2, extra collapse args
            ("human", "extended", "b", 1500, 120, 250, 0.004, 0.01, {}),
            ("human", "extended", "a", 1200, 60, 250, 0.006, 0.02, {"bcthreshold": 1, "percentlevdist": 5}),
            ("human", "original", "b", 1000, 200, 250, 0.002, 0.005, {"bcthreshold": 3, "minbcQ": 10, "allowNs": True}),
            ("mouse", "original", "a", 800, 40, 250, 0.01, 0.02, {"lenthreshold": 100}),
        ]
        for ci, (species, tagset, chain, n, pool, L, sub1, sub2, extra) in enumerate(spec):
            rows, args = make_rows(ref, species, tagset, chain, n, pool, L, sub1, sub2, 20260300 + ci, "M13")
            cases.append(record_case(C, rows, args, extra, (ci, species, tagset, chain)))
    finally:
        os.chdir(cwd)
    # distance known answers: random sequence pairs with the stand-in polyleven (checked against the reference's
    # own known answers by its tests) -- lengths straddling the 64-symbol word boundaries of the GPU verifier
    rng = random.Random(7)
    import polyleven
    dist = []
    for L in (0, 1, 5, 12, 31, 32, 33, 63, 64, 65, 100, 127, 128, 129, 130, 200, 256, 300):
        for _ in range(6):
            a = "".join(rng.choice("ACGT") for _ in range(L))
            b = list(a)
            for _ in range(rng.randrange(0, 1 + max(1, L // 8))):
                op = rng.randrange(3)
                if op == 0 and b:
                    b[rng.randrange(len(b))] = rng.choice("ACGTN")
                elif op == 1 and b:
                    del b[rng.randrange(len(b))]
                else:
                    b.insert(rng.randrange(len(b) + 1), rng.choice("ACGT"))
            b = "".join(b)
            dist.append([a, b, int(polyleven.levenshtein(a, b))])
    out = os.path.join(ROOT, "tests", "golden", "collapse_cases.json.gz")
    with gzip.GzipFile(out, "wb", mtime=0) as fh:
        fh.write(json.dumps({"cases": cases, "distances": dist}).encode())
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
