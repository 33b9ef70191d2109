"""CPU-only: (1) the collapse oracle against fixtures recorded from the unmodified reference, (2) the bit-parallel
Levenshtein code the kernels run (csrc/lev_core.cuh, compiled for the host by tests/sim) against the oracle, (3) the
HOST logic of decombinator_b200.collapse -- grouping order, component order, counting -- with the distances supplied
by the oracle in place of the GPU context.  The GPU suite (test_gpu_collapse.py) repeats (3) on the CUDA kernels."""
import collections as coll
import gzip
import json
import os
import random

import numpy as np
import pytest

import collapse_checks
import collapse_oracle as CO
import simlib
from decombinator_b200 import _lib, collapse


@pytest.fixture(scope="module")
def collapse_cases(golden_dir):
    with gzip.open(os.path.join(golden_dir, "collapse_cases.json.gz"), "rt") as fh:
        return json.load(fh)


@pytest.fixture
def oracle_distances(monkeypatch):
    """Install the oracle where decombinator_b200.collapse expects its GPU context (tests only)."""
    monkeypatch.setattr(collapse, "_dist", CO.OracleDist())


def test_collapse_oracle_matches_recorded_distances(collapse_cases):
    for a, b, d in collapse_cases["distances"]:
        assert CO.levenshtein(a, b) == d


def test_collapse_oracle_matches_recorded_pairs_and_verdicts(collapse_cases):
    for case in collapse_cases["cases"]:
        row, col = CO.umi_pairs(case["umis"], case["args"]["bcthreshold"])
        assert [[int(i), int(j)] for i, j in zip(row, col)] == case["pairs"]
        protos = [k.split("|")[2] for k in case["group_keys"]]
        frac = case["args"]["percentlevdist"] / 100
        assert [CO.seqs_equivalent(protos[i], protos[j], frac) for i, j in case["pairs"]] == case["verdicts"]


def test_device_levenshtein_code_matches_oracle(collapse_cases):
    sim = simlib.lev_sim()
    for a, b, d in collapse_cases["distances"]:
        sa, _, _ = _lib.encode_seqs([a, b])
        x, y = np.ascontiguousarray(sa[:len(a)]), np.ascontiguousarray(sa[len(a):])
        assert sim.sim_seq_distance(x.ctypes.data, len(a), y.ctypes.data, len(b)) == d, (len(a), len(b))
    rng = random.Random(3)
    for _ in range(3000):
        L = rng.choice((4, 8, 11, 12, 13, 17, 19))
        a = "".join(rng.choice("ACGTNSL") for _ in range(L))
        b = list(a)
        for _ in range(rng.randrange(0, 4)):
            op = rng.randrange(3)
            if op == 0 and b:
                b[rng.randrange(len(b))] = rng.choice("ACGT")
            elif op == 1 and b:
                del b[rng.randrange(len(b))]
            elif len(b) < 19:
                b.insert(rng.randrange(len(b) + 1), rng.choice("ACGT"))
        b = "".join(b)
        ca, cb = (int(x) for x in _lib.encode_umis([a, b]))
        d = CO.levenshtein(a, b)
        assert sim.sim_umi_distance(ca, cb) == d
        for k in (0, 1, 2, 3):
            if d <= k:  # the prefilter may never reject a pair that is within k edits
                assert sim.sim_umi_may_be_within(ca, cb, k) == 1, (a, b, k)


def test_device_levenshtein_code_on_arbitrary_bytes():
    """The eight-bit-plane form (batches with more than eight distinct characters: lower case, IUPAC) against the oracle."""
    sim = simlib.lev_sim()
    rng = random.Random(5)
    alphabet = "ACGTNRYKMSWBDHVacgtn"
    for _ in range(400):
        L = rng.choice((1, 20, 64, 65, 130, 200, 300))
        a = "".join(rng.choice(alphabet) for _ in range(L))
        b = list(a)
        for _ in range(rng.randrange(0, 12)):
            op = rng.randrange(3)
            if op == 0 and b:
                b[rng.randrange(len(b))] = rng.choice(alphabet)
            elif op == 1 and b:
                del b[rng.randrange(len(b))]
            else:
                b.insert(rng.randrange(len(b) + 1), rng.choice(alphabet))
        b = "".join(b)
        x = np.frombuffer(a.encode(), dtype=np.uint8).copy()
        y = np.frombuffer(b.encode(), dtype=np.uint8).copy()
        assert sim.sim_seq_distance_bytes(x.ctypes.data, len(a), y.ctypes.data, len(b)) == CO.levenshtein(a, b)
    sym, off, ln = _lib.encode_seqs(["ACGTNRYKMSWacgt", "ACGT"])          # more than eight symbols: the characters themselves
    assert sym.max() > 7 and bytes(sym[:4]) == b"ACGT"
    sym, off, ln = _lib.encode_seqs(["ACGTN", "ACGT"])
    assert sym.max() <= 7


@pytest.fixture(scope="module")
def oligo_cases(golden_dir):
    """Runs of the unmodified reference for the oligo designs beside M13 (oracle/make_golden_collapse_oligos.py)."""
    with gzip.open(os.path.join(golden_dir, "collapse_cases_oligos.json.gz"), "rt") as fh:
        return json.load(fh)


def test_host_logic_matches_reference_runs_other_oligos(oligo_cases, oracle_distances):
    """I8 (two cases), I8_single, NEBIO, TAKARA: groups, UMI pairs, clusters, .freq rows and the counters of read_in_data."""
    assert [c["args"]["oligo"] for c in oligo_cases["cases"]] == ["I8", "I8", "I8_single", "NEBIO", "TAKARA"]
    for case in oligo_cases["cases"]:
        collapse_checks.check_case(case)


def test_host_logic_matches_reference_runs(collapse_cases, oracle_distances):
    for case in collapse_cases["cases"]:
        collapse_checks.check_case(case)


@pytest.mark.parametrize("chain,name", [("a", "alpha"), ("b", "beta")])
def test_host_logic_reproduces_golden_freq(golden_dir, tmp_path, oracle_distances, chain, name):
    collapse_checks.check_tiny_freq(golden_dir, tmp_path, chain, name)


def test_reference_unit_answers(oracle_distances):
    collapse_checks.check_reference_unit_answers()


def test_barcode_positions_known_answers():
    """reference tests/test_collapse.py:55-195"""
    c = coll.Counter()
    m13, i8 = "GTCGTGACTGGGAAAACCCTGG", "GTCGTGAT"
    f = collapse.get_barcode_positions
    assert f("GTCGTGACTGGGAAAACCCTGGTTTCCGGTCGTGATAAAGTG", {"oligo": "m13", "allowNs": False}, c) == \
        [len(m13), len(m13) + 6, len(m13) + 6 + len(i8), len(m13) + 6 + len(i8) + 6]
    assert f("GTCGTGATTTTCCGGTCGTGATAAAGTG", {"oligo": "i8", "allowNs": False}, c) == [8, 14, 22, 28]
    assert f("GAAGCTATCACGACATCACTAC", {"oligo": "i8_single", "allowNs": False}, c) == [0, 6, 14, 20]
    assert f("CGGGCTTGGTATCGGCCGATCTACGGG", {"oligo": "nebio", "allowNs": False}, c) == [0, 17]
    assert f("CTCGTTAGGTTCGTACGGGGATTGCA", {"oligo": "takara", "allowNs": False}, c) == [0, 12]
    assert f("GTCGTGACTGGGAAAACCCTGGTTNCCGGTCGTGATAAAGTG", {"oligo": "m13", "allowNs": False}, c) is None
    assert c["getbarcode_fail_N"] == 1
    with pytest.raises(ValueError):
        f("ACGT", {"oligo": "nope", "allowNs": False}, c)
    for spacer, seq, lo, hi in (("GTCGTGACTGGGAAAACCCTGG", "GTCGTGACTGGGAAAACCCTGGTTTCCGGTCGTGATAAAGTG", 0, 32),
                                ("GTCGTGAT", "GTCGTGATTTTCCGGTCGTGATAAAGTG", 0, 18), ("ATCACGAC", "GAAGCTATCACGACATCACTAC", 0, 18),
                                ("TACGGG", "CGGGCTTGGTATCGGCCGATCTACGGG", 18, 28), ("GTACGGG", "CTCGTTAGGTTCGTACGGGGATTGCA", 0, 19)):
        assert collapse.findFirstSpacer({"spcr1": spacer}, seq, lo, hi) == [spacer]


def test_n12_index_and_collapse_lines_equal_the_split_lines(golden_dir, tmp_path):
    """dcb_n12_index / dcb_n12_collapse_rows (host code) against the reference's own parsing of an .n12 file
    (collapse.py:523-530: line.rstrip("\\n").split(", "); :565-590: str(row[:5]), "|".join(...))."""
    for name in ("alpha", "beta"):
        raw = open(os.path.join(golden_dir, "dcr_TINY_1_%s.n12" % name), "rb").read()
        rows = [line.rstrip("\n").split(", ") for line in raw.decode().splitlines(True)]
        cols = collapse.N12Columns.from_file(os.path.join(golden_dir, "dcr_TINY_1_%s.n12" % name), open)
        assert len(cols) == len(rows)
        keep = np.arange(len(rows)) % 3 != 1
        assert cols.subset_rows(keep) == [r for r, k in zip(rows, keep) if k]
        seqs, dcrs, etcs = cols.collapse_lines(keep)
        want = [r for r, k in zip(rows, keep) if k]
        assert seqs == [r[6] for r in want] and dcrs == [str(r[:5]) for r in want]
        assert etcs == ["|".join((str(r[:5]), r[6], r[7], r[5])) for r in want]
        # many rows: every thread's share, cut at row starts
        big = raw * 400
        text = np.frombuffer(big, dtype=np.uint8)
        off, ln = _lib.n12_index(text, n_threads=7)
        assert len(off) == 400 * len(rows)
        flat = [big[int(o):int(o) + int(l)].decode() for o, l in zip(off[-len(rows):].ravel(), ln[-len(rows):].ravel())]
        assert flat == [f for r in rows for f in r]
    # texts the index declines (the caller then reads the lines as the reference does)
    one = b"1, 2, 3, 4, ACGT, id, ACGT, IIII, ACGT, IIII\n"
    assert _lib.n12_index(np.frombuffer(one[:-1], dtype=np.uint8)) is None                       # no final newline
    assert _lib.n12_index(np.frombuffer(one + b"1, 2, 3\n", dtype=np.uint8)) is None              # a short row
    assert _lib.n12_index(np.frombuffer(one.replace(b"\n", b"\r\n"), dtype=np.uint8)) is None    # carriage returns
    quoted = one.replace(b"ACGT, id", b"AC'T, id")
    off, ln = _lib.n12_index(np.frombuffer(quoted, dtype=np.uint8))
    assert _lib.n12_collapse_rows(np.frombuffer(quoted, dtype=np.uint8), off, ln, np.ones(1, dtype=bool)) is None
    p = tmp_path / "q.n12"
    p.write_bytes(quoted)
    seqs, dcrs, etcs = collapse.N12Columns.from_file(str(p), open).collapse_lines(np.ones(1, dtype=bool))
    assert dcrs == [str(["1", "2", "3", "4", "AC'T"])]


def test_native_grouping_equals_the_python_state_machines(oracle_distances, monkeypatch):
    """dcb_group (csrc/group.cpp) against _BarcodeMachine -- the restatement of collapse.py:595-682 the recorded runs of the
    reference pin -- on rows with many copies per barcode, variants one or two edits apart (so that proto-sequences change),
    unrelated sequences on a used barcode (dead barcodes) and S / L / N symbols in barcodes."""
    rng = random.Random(17)
    pool = ["".join(rng.choice("ACGT") for _ in range(rng.randrange(60, 90))) for _ in range(40)]
    barcodes = ["".join(rng.choice("ACGTACGTACGTNSL") for _ in range(12)) for _ in range(150)]
    home = {b: rng.choice(pool) for b in barcodes}
    kept = []
    for i in range(6000):
        b = rng.choice(barcodes)
        seq = list(home[b])
        r = rng.random()
        if r < 0.35:
            for _ in range(rng.randrange(1, 3)):
                seq[rng.randrange(len(seq))] = rng.choice("ACGT")
        elif r < 0.40:
            del seq[rng.randrange(len(seq))]
        elif r < 0.42:
            seq = list(rng.choice(pool))                       # another molecule on this barcode
        seq = "".join(seq)
        kept.append((3 * i + 1, b, seq, "dcr|%s|q|id%d" % (seq, i)))
    calls = []
    real = _lib.Grouping.run
    monkeypatch.setattr(_lib.Grouping, "run", lambda self, v: (calls.append(1), real(self, v))[1])
    for frac in (0.03, 0.1):
        native = collapse._group_rows(kept, frac)
        assert calls, "the library's grouping did not run"
        monkeypatch.setenv("DCB_GROUP_NATIVE", "0")
        n_calls = len(calls)
        python = collapse._group_rows(kept, frac)
        assert len(calls) == n_calls
        monkeypatch.delenv("DCB_GROUP_NATIVE")
        assert sorted(native[0]) == sorted(python[0]) and native[1:] == python[1:]
        assert [g[0] for g in native[0]] == sorted(g[0] for g in python[0])          # the library hands them over in dict order
        assert native[2] > 3 and native[1] > native[2]                              # dead barcodes, and rows dropped behind them
    # rows it does not take: a barcode of another length, a symbol outside ACGTNSL
    assert collapse._group_rows_native(kept[:100] + [(99999, "ACGT", "ACGT", "x")], 0.1) is None
    assert collapse._group_rows_native(kept[:100] + [(99999, "ACGTACGTACGR", "ACGT", "x")], 0.1) is None


def test_empty_n12_raises(tmp_path):
    p = tmp_path / "empty.n12"
    p.write_text("")
    with pytest.raises(ValueError):
        collapse.check_dcr_file(str(p), open)


def test_distances_fail_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.DcbError):
        _lib.Dist(0)
