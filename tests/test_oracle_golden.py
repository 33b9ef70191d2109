"""The oracle (oracle/dcr_oracle.c) is pinned against the reference's own golden files and against
fixtures recorded from the unmodified reference source (oracle/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

import decombine_oracle as O
from helpers import record_to_list


@pytest.mark.parametrize("chain,name", [("a", "alpha"), ("b", "beta")])
def test_tiny_n12_golden(golden_dir, chain, name):
    # reference tests/test_pipeline.py:63-84 compares exactly these files byte for byte
    rows, counts = O.decombinator_rows(os.path.join(golden_dir, "TINY_1.fq"), chain)
    with open(os.path.join(golden_dir, "dcr_TINY_1_%s.n12" % name)) as fh:
        assert O.n12_text(rows) == fh.read()


def test_tiny_counters_match_shipped_logs(golden_dir):
    # examples/Logs/2025_08_29_{alpha,beta}__Decombinator_Summary.csv:21-49 of the reference
    _, a = O.decombinator_rows(os.path.join(golden_dir, "TINY_1.fq"), "a")
    assert (a["read_count"], a["vj_count"], a["verr2"], a["foundv1notv2"], a["foundv2notv1"], a["no_vtags_found"]) == \
        (106, 35, 1, 7, 9, 55)
    _, b = O.decombinator_rows(os.path.join(golden_dir, "TINY_1.fq"), "b")
    assert (b["read_count"], b["vj_count"], b["verr1"], b["verr2"], b["foundv1notv2"], b["foundv2notv1"],
            b["no_vtags_found"]) == (106, 48, 1, 1, 1, 2, 55)


def test_tiny_original_tagset_known_answers(golden_dir):
    # recorded from the reference with -tg original
    rows, a = O.decombinator_rows(os.path.join(golden_dir, "TINY_1.fq"), "a", tags="original")
    assert len(rows) == 35 and (a["no_vtags_found"], a["foundv2notv1"], a["foundv1notv2"], a["verr2"]) == (54, 10, 7, 1)
    rows, b = O.decombinator_rows(os.path.join(golden_dir, "TINY_1.fq"), "b", tags="original")
    assert len(rows) == 48
    assert (b["verr1"], b["verr2"], b["jerr1"], b["foundv1notv2"], b["foundv2notv1"], b["no_vtags_found"]) == \
        (1, 3, 1, 5, 1, 52)


def test_dcr_cases_per_read(dcr_cases):
    """Every recorded dcr() call: same return value AND same counter increments."""
    for gi, g in enumerate(dcr_cases["groups"]):
        orc = O.Oracle(O.TagSet(g["species"], g["tags"], g["chain"]), g["allowNs"], g["lenthreshold"])
        for i, (read, exp, delta) in enumerate(zip(g["reads"], g["results"], g["deltas"])):
            before = orc.counts.copy()
            rec = orc.decombine_reads([read], g["orientation"])[0]
            got = record_to_list(read, rec, g["orientation"])
            assert got == exp, (gi, i, read)
            dd = {n: int(a - b) for n, a, b in zip(O.COUNTER_NAMES, orc.counts, before) if a != b}
            assert dd == delta, (gi, i, read)
        assert orc.counts_dict() == g["totals"]


def test_dcr_cases_threaded_batch(dcr_cases):
    """The pthread batch entry point gives the same records and counter totals as one read at a time."""
    g = dcr_cases["groups"][1]
    orc = O.Oracle(O.TagSet(g["species"], g["tags"], g["chain"]), g["allowNs"], g["lenthreshold"])
    res = orc.decombine_reads(g["reads"], g["orientation"], nthreads=4)
    assert [record_to_list(r, rec, g["orientation"]) for r, rec in zip(g["reads"], res)] == g["results"]
    assert orc.counts_dict() == g["totals"]


def test_decombinator_file_runs(decombinator_runs, tmp_path):
    """readfq + barcode slicing + row assembly restated in the oracle == the reference's decombinator()."""
    for ri, run in enumerate(decombinator_runs["runs"]):
        a = run["args"]
        if a.get("sampling_analysis"):
            continue  # the oracle's row builder does not restate -sa; the product test covers it
        f1 = tmp_path / ("run%d_1.fq" % ri)
        f1.write_text(run["fastq1"])
        (tmp_path / ("run%d_2.fq" % ri)).write_text(run["fastq2"])
        rows, counts = O.decombinator_rows(str(f1), a["chain"], a["bc_read"], a["bclength"], a["orientation"], a["tags"],
                                           a["species"], a["allowNs"], a["lenthreshold"])
        assert rows == run["rows"], ri
        for k, v in run["counts"].items():
            if k in counts:
                assert counts[k] == v, (ri, k)


def test_revcomp_matches_bio_table():
    assert O.revcomp("ACGTNacgtnRYKMUu") == "aAKMRYnacgtNACGT"


def test_findall_order_end_position_then_longest():
    # human TRBV half2: CCTGTATCTC is inside GCCCTGTATCTCTGT -> the shorter one ends first
    orc = O.Oracle(O.TagSet("human", "extended", "b"))
    read = "AAAAGCCCTGTATCTCTGTAAAA"
    hits = orc.findall(2, read)
    ends = [s + len(orc.ts.v_seqs[k][10:]) for k, s in hits]
    assert ends == sorted(ends)
    ids = [k for k, _ in hits]
    assert 28 in ids and 24 in ids and ids.index(28) < ids.index(24)
