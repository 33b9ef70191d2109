#!/usr/bin/env python
"""Record what the UNMODIFIED reference returns, as fixtures under tests/golden/.

Runs only in the build container (needs /root/reference); the GPU box uses the committed
fixtures.  The reference source is imported as-is through oracle/refenv.py (stand-ins for the
five absent third-party wheels, see oracle/standins/README.md).

Outputs
  tests/golden/dcr_cases.json.gz      per-read dcr() results + counter deltas for synthetic and
                                      fuzzed reads over many tag sets / orientations
  tests/golden/decombinator_runs.json.gz  whole-file decombinator() runs (rows + counters) on small
                                      FASTQ pairs written from the same generator

usage: python oracle/make_golden.py
"""
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

import numpy as np  # noqa: E402

import refenv  # noqa: E402
from decombinator_b200 import _lib, tags as dtags  # noqa: E402

COUNTERS = [
    "verr1", "verr2", "jerr1", "jerr2",
    "dcrfilter_intertagN", "dcrfilter_toolong_intertag", "dcrfilter_imposs_deletion", "dcrfilter_tag_overlap",
    "multiple_v_matches", "v_del_failed_tag_at_end", "v_del_failed", "foundv1notv2", "foundv2notv1",
    "no_vtags_found", "multiple_j_matches", "j_del_failed", "foundj1notj2", "foundj2notj1",
    "no_j_assigned", "VJ_assignment_failed",
]

IUPAC = "RYKMSWBDHVUacgtn"


def synth_reads(species, tagset, chain, n, L, sub, nrate, junk, seed, L2=0):
    info = dtags.load(species, tagset, chain)
    syn = _lib.Synth([(info.v_regions, info.j_regions)], seed, L, L2, sub, nrate, junk)
    r1, r2 = syn.reads(0, n, want_r2=bool(L2))
    a = [bytes(r1[i * L:(i + 1) * L]).decode() for i in range(n)]
    b = [bytes(r2[i * L2:(i + 1) * L2]).decode() for i in range(n)] if L2 else None
    return a, b, info


def fuzz(reads, info, rng):
    """Edge cases: truncations, tags at the read ends, non-ACGT symbols, chimeras, tiny reads."""
    out = []
    comp = str.maketrans("ACGT", "TGCA")
    for r in reads:
        op = rng.randrange(13)
        if op == 0:      # cut from the left (oriented right end): V tag / J tag near an end
            out.append(r[rng.randrange(0, len(r)):])
        elif op == 1:    # cut from the right
            out.append(r[:rng.randrange(0, len(r) + 1)])
        elif op == 2:    # both
            a = rng.randrange(0, len(r)); b = rng.randrange(a, len(r) + 1)
            out.append(r[a:b])
        elif op == 3:    # sprinkle N
            s = list(r)
            for _ in range(rng.randrange(1, 6)):
                s[rng.randrange(len(s))] = "N"
            out.append("".join(s))
        elif op == 4:    # sprinkle IUPAC / lower case / U
            s = list(r)
            for _ in range(rng.randrange(1, 6)):
                s[rng.randrange(len(s))] = rng.choice(IUPAC)
            out.append("".join(s))
        elif op == 5:    # chimera of two reads (multiple tag matches)
            o = rng.choice(reads)
            out.append(r[:rng.randrange(len(r))] + o[rng.randrange(len(o)):])
        elif op == 6:    # tiny
            out.append(r[:rng.randrange(0, 30)])
        elif op == 7:    # read = revcomp of (k junk + V region prefix ... ) so the V region starts at oriented pos < 10
            v = rng.choice(info.v_regions)
            j = rng.choice(info.j_regions)
            k = rng.randrange(0, 12)
            mol = "".join(rng.choice("ACGT") for _ in range(k)) + v[:len(v) - rng.randrange(0, 12)] + \
                  "".join(rng.choice("ACGT") for _ in range(rng.randrange(0, 8))) + j[rng.randrange(0, 8):] + \
                  "".join(rng.choice("ACGT") for _ in range(rng.randrange(0, 40)))
            s = list(mol)
            for _ in range(rng.randrange(0, 25)):   # mutations that break the 10-mer walk
                s[rng.randrange(len(s))] = rng.choice("ACGT")
            out.append("".join(s).translate(comp)[::-1])
        elif op == 8:    # oriented read ends exactly at / just after the germline V end, or right after the J tag
            v = rng.choice(info.v_regions)
            mol = v[max(0, len(v) - rng.randrange(60, 200)):] + ("" if rng.random() < 0.5 else rng.choice(info.j_regions)[:rng.randrange(0, 40)])
            out.append(mol.translate(comp)[::-1])
        elif op == 9:    # dense substitutions
            s = list(r)
            for _ in range(rng.randrange(5, 30)):
                s[rng.randrange(len(s))] = rng.choice("ACGT")
            out.append("".join(s))
        elif op == 10:   # J region directly after V tag (deletion skipping / overlap filters)
            v = rng.choice(info.v_regions); j = rng.choice(info.j_regions)
            mol = v[:len(v) - rng.randrange(0, 45)] + j[rng.randrange(0, 30):] + "".join(rng.choice("ACGT") for _ in range(rng.randrange(0, 30)))
            out.append(mol.translate(comp)[::-1])
        elif op == 11:   # J before V.  (This cannot reach dcrfilter_tag_overlap or, with a positive threshold,
            # dcrfilter_toolong_intertag -- nothing can: once vdel <= jump_v - len(vtag) and jdel <= jump_j have passed
            # (dcrfilter_imposs_deletion is checked first, decombine.py:561-565), end_of_v >= v_seq_start + len(vtag) and the
            # J walk only accepts positions >= end_of_v with pos <= jump_j, so the J tag starts at or behind the V tag's end:
            # v_seq_start + len(vtag) <= j_seq_end < j_seq_end + len(jtag).  Both filters are dead code)
            v = rng.choice(info.v_regions); j = rng.choice(info.j_regions)
            mol = j + "".join(rng.choice("ACGT") for _ in range(rng.randrange(0, 20))) + v + \
                  "".join(rng.choice("ACGT") for _ in range(rng.randrange(12, 40)))
            out.append(mol.translate(comp)[::-1])
        else:
            out.append(r)
    return out


def run_reads(ref, inputargs, reads):
    """The per-read part of the reference's main loop (decombine.py:999-1010) with dcr() itself."""
    d = ref["decombine"]
    d.import_tcr_info(inputargs)
    results, deltas = [], []
    prev = dict(d.counts)
    for vdj in reads:
        orientation = inputargs["orientation"]
        if orientation == "reverse":
            recom, frame = d.dcr(d.revcomp(vdj), inputargs), 0
        elif orientation == "forward":
            recom, frame = d.dcr(vdj, inputargs), 1
        else:
            recom, frame = d.dcr(d.revcomp(vdj), inputargs), 0
            if not recom:
                recom, frame = d.dcr(vdj, inputargs), 1
        now = dict(d.counts)
        delta = {k: now.get(k, 0) - prev.get(k, 0) for k in COUNTERS if now.get(k, 0) != prev.get(k, 0)}
        prev = now
        results.append(None if not recom else [int(recom[0]), int(recom[1]), int(recom[2]), int(recom[3]), str(recom[4]),
                                               int(recom[5]), int(recom[6]), frame])
        deltas.append(delta)
    totals = {k: int(d.counts.get(k, 0)) for k in COUNTERS}
    return results, deltas, totals


def base_args(ref, **kw):
    args = ref["io"].create_args_dict(infile="x_1.fq", chain=kw.pop("chain"), bc_read="R2", dontcount=True,
                                      suppresssummary=True, dontcheck=True, tagfastadir=refenv.REF_TAGDIR, outpath="")
    args.update(kw)
    return args


def write_fastq(path, names, seqs, quals):
    with open(path, "wt") as fh:
        for n, s, q in zip(names, seqs, quals):
            fh.write("@%s\n%s\n+\n%s\n" % (n, s, q))


def main():
    ref = refenv.load()
    rng = random.Random(20260101)
    groups = []
    spec = [
        # species, tags, chain, orientation, L, sub, N, junk, n_syn, n_fuzz, allowNs
        ("human", "extended", "b", "reverse", 250, 0.0, 0.0, 0.02, 300, 500, False),
        ("human", "extended", "b", "reverse", 250, 0.01, 0.001, 0.05, 400, 500, False),
        ("human", "extended", "a", "reverse", 150, 0.01, 0.001, 0.05, 400, 500, False),
        ("human", "extended", "a", "both", 250, 0.02, 0.002, 0.05, 300, 400, True),
        ("human", "original", "b", "reverse", 250, 0.02, 0.001, 0.05, 400, 500, False),
        ("human", "original", "a", "both", 200, 0.02, 0.001, 0.05, 300, 400, False),
        ("human", "original", "g", "reverse", 250, 0.01, 0.001, 0.05, 300, 300, False),
        ("human", "original", "d", "forward", 250, 0.01, 0.001, 0.05, 200, 300, False),
        ("mouse", "original", "a", "reverse", 250, 0.02, 0.001, 0.05, 400, 600, False),
        ("mouse", "original", "b", "reverse", 300, 0.01, 0.001, 0.05, 300, 400, False),
        ("mouse", "original", "g", "reverse", 250, 0.005, 0.0, 0.05, 300, 300, False),
        ("mouse", "original", "d", "both", 250, 0.005, 0.001, 0.05, 300, 300, False),
    ]
    for gi, (species, tagset, chain, orient, L, sub, nrate, junk, n_syn, n_fuzz, allowNs) in enumerate(spec):
        reads, _, info = synth_reads(species, tagset, chain, n_syn + n_fuzz, L, sub, nrate, junk, 20260100 + gi)
        if orient == "forward":  # the generator emits reverse-strand reads; flip them for a forward-only run
            reads = [ref["decombine"].revcomp(r) for r in reads]
        fz = fuzz(reads[n_syn:], info, rng)
        if orient == "forward":
            pass
        allreads = reads[:n_syn] + fz
        args = base_args(ref, chain=chain, species=species, tags=tagset, orientation=orient, allowNs=allowNs)
        results, deltas, totals = run_reads(ref, args, allreads)
        n_ok = sum(1 for r in results if r)
        print("group", gi, species, tagset, chain, orient, "reads", len(allreads), "decombined", n_ok,
              {k: v for k, v in totals.items() if v})
        groups.append({"species": species, "tags": tagset, "chain": chain, "orientation": orient, "allowNs": allowNs,
                       "lenthreshold": 130, "reads": allreads, "results": results, "deltas": deltas, "totals": totals})
    # a lenthreshold that can actually trigger, and allowNs variants, on one group
    reads = groups[1]["reads"]
    for allowNs, lenthr in ((True, 130), (False, -85), (False, -100)):
        args = base_args(ref, chain="b", species="human", tags="extended", orientation="reverse", allowNs=allowNs,
                         lenthreshold=lenthr)
        results, deltas, totals = run_reads(ref, args, reads)
        print("variant allowNs", allowNs, "lenthreshold", lenthr, {k: v for k, v in totals.items() if v})
        groups.append({"species": "human", "tags": "extended", "chain": "b", "orientation": "reverse", "allowNs": allowNs,
                       "lenthreshold": lenthr, "reads": reads, "results": results, "deltas": deltas, "totals": totals})
    out = os.path.join(ROOT, "tests", "golden", "dcr_cases.json.gz")
    with gzip.GzipFile(out, "wb", mtime=0) as fh:
        fh.write(json.dumps({"counters": COUNTERS, "groups": groups}).encode())
    print("wrote", out, os.path.getsize(out), "bytes")

    # ---- whole-file runs through the reference's decombinator() ------------------------------------
    runs = []
    tmp = tempfile.mkdtemp(prefix="dcbgold")
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        file_spec = [
            # species, tags, chain, orientation, bc_read, bclength, L, sub, N, junk, n, extra
            ("human", "extended", "b", "reverse", "R2", 42, 250, 0.01, 0.002, 0.1, 300, {}),
            ("human", "extended", "a", "both", "R2", 42, 150, 0.01, 0.002, 0.1, 300, {}),
            ("human", "original", "b", "reverse", "R1", 12, 250, 0.01, 0.002, 0.1, 300, {}),
            ("mouse", "original", "a", "reverse", "R2", 30, 250, 0.01, 0.002, 0.1, 300, {"allowNs": True}),
            ("human", "extended", "b", "reverse", "R2", 42, 250, 0.01, 0.002, 0.1, 120, {"sampling_analysis": True}),
            ("human", "extended", "b", "forward", "R2", 42, 250, 0.0, 0.0, 0.1, 100, {}),
        ]
        for fi, (species, tagset, chain, orient, bc_read, bl, L, sub, nrate, junk, n, extra) in enumerate(file_spec):
            r1, r2, info = synth_reads(species, tagset, chain, n, L, sub, nrate, junk, 20260200 + fi, L2=L)
            qrng = random.Random(fi)
            q1 = ["".join(qrng.choice("FFFFFFF:,#") for _ in range(L)) for _ in range(n)]
            q2 = ["".join(qrng.choice("FFFFFFF:,#") for _ in range(L)) for _ in range(n)]
            if bc_read == "R1":  # barcode is the prefix of R1
                r1 = [b[:bl] + a for a, b in zip(r1, r2)]
                q1 = [qb[:bl] + qa for qa, qb in zip(q1, q2)]
            # put an N into some barcodes
            r2 = [(b[:5] + "N" + b[6:]) if i % 17 == 0 else b for i, b in enumerate(r2)]
            if orient == "forward":
                r1 = [ref["decombine"].revcomp(a) for a in r1]
            names = ["SYN:%d:%d extra words" % (fi, i) for i in range(n)]
            f1 = "sample%d_1.fq" % fi
            write_fastq(f1, names, r1, q1)
            write_fastq(f1.replace("1.f", "2.f"), names, r2, q2)
            args = ref["io"].create_args_dict(infile=f1, chain=chain, bc_read=bc_read, dontcount=True, suppresssummary=True,
                                              dontcheck=True, tagfastadir=refenv.REF_TAGDIR, outpath="", species=species,
                                              tags=tagset, orientation=orient, bclength=bl, **extra)
            rows = ref["decombine"].decombinator(args)
            counts = {k: int(v) for k, v in ref["decombine"].counts.items() if k not in ("start_time", "end_time")}
            # the same run once more with the summary switched on: the text of Logs/..._Decombinator_Summary.csv, with the
            # values of the four lines that depend on where and when it ran replaced (so the fixture regenerates identically)
            sargs = dict(args, suppresssummary=False, dontcheck=False)
            ref["decombine"].decombinator(sargs)
            logs = sorted(os.listdir("Logs"))
            assert len(logs) == fi + 1, logs
            newest = max((os.path.join("Logs", n) for n in logs), key=os.path.getmtime)
            with open(newest) as fh:
                summary = fh.read()
            summary = "\n".join(ln.split(",")[0] + ",<run>" if ln.split(",")[0] in ("Directory", "DateFinished", "TimeFinished", "TimeTaken(Seconds)")
                                else ln for ln in summary.split("\n"))
            summary_name = os.path.basename(newest).split("_", 3)[3]        # without the date prefix YYYY_MM_DD_
            with open(f1) as fh:
                t1 = fh.read()
            with open(f1.replace("1.f", "2.f")) as fh:
                t2 = fh.read()
            print("file run", fi, species, tagset, chain, orient, bc_read, "rows", len(rows))
            runs.append({"args": {k: v for k, v in args.items() if k != "tagfastadir"}, "fastq1": t1, "fastq2": t2,
                         "rows": rows, "counts": counts, "summary": summary, "summary_name": summary_name})
    finally:
        os.chdir(cwd)
    out = os.path.join(ROOT, "tests", "golden", "decombinator_runs.json.gz")
    with gzip.GzipFile(out, "wb", mtime=0) as fh:
        fh.write(json.dumps({"runs": runs}).encode())
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
