"""Parity tests of the collapse distance kernels (csrc/collapse.cu) through the C ABI: dcb_umi_pairs and dcb_lev_leq
against the oracle and against fixtures recorded from the unmodified reference; the collapse stage end to end with the
distances on the GPU against the recorded runs and the reference's golden .freq files.  Bit-exact everywhere."""
import collections as coll
import gzip
import json
import os
import random

import numpy as np
import pytest

import collapse_checks
import collapse_oracle as CO
from decombinator_b200 import _lib, collapse

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def collapse_cases(golden_dir):
    with gzip.open(os.path.join(golden_dir, "collapse_cases.json.gz"), "rt") as fh:
        return json.load(fh)


@pytest.fixture(scope="module")
def dist():
    d = _lib.Dist(0)
    yield d
    d.close()


@pytest.fixture(autouse=True)
def real_gpu_context():
    collapse._dist = None   # whatever an earlier (CPU) test installed: the product path builds its own dcb_dist
    yield


def test_lev_leq_matches_recorded_distances(dist, collapse_cases):
    pairs = collapse_cases["distances"]
    seqs = [p[0] for p in pairs] + [p[1] for p in pairs]
    sym, off, ln = _lib.encode_seqs(seqs)
    n = len(pairs)
    a, b = np.arange(n, dtype=np.uint32), np.arange(n, 2 * n, dtype=np.uint32)
    for frac in (0.0, 0.05, 0.1, 0.13, 0.25, 1.0):
        got = dist.lev_leq(sym, off, ln, a, b, frac)
        want = [d <= min(len(x), len(y)) * frac for x, y, d in pairs]
        assert got.tolist() == want, frac


def test_lev_leq_random_pairs_vs_oracle(dist):
    rng = random.Random(11)
    seqs = []
    for _ in range(600):
        L = rng.choice((0, 1, 20, 63, 64, 65, 90, 128, 129, 130, 191, 193, 256, 257, 400, 512))
        a = [rng.choice("ACGTN") for _ in range(L)]
        b = list(a)
        for _ in range(rng.randrange(0, 30)):
            op = rng.randrange(3)
            if op == 0 and b:
                b[rng.randrange(len(b))] = rng.choice("ACGTRY")
            elif op == 1 and b:
                del b[rng.randrange(len(b))]
            elif len(b) < 512:
                b.insert(rng.randrange(len(b) + 1), rng.choice("ACGT"))
        seqs += ["".join(a), "".join(b)]
    sym, off, ln = _lib.encode_seqs(seqs)
    a = np.arange(0, len(seqs), 2, dtype=np.uint32)
    b = a + 1
    for frac in (0.02, 0.1):
        got = dist.lev_leq(sym, off, ln, a, b, frac)
        want = [CO.seqs_equivalent(seqs[i], seqs[j], frac) for i, j in zip(a, b)]
        assert got.tolist() == want


def test_lev_leq_on_more_than_eight_symbols(dist):
    """Lower case and IUPAC symbols in the inter-tag sequences (the reference compares whatever characters it is given):
    the batch goes through dcb_lev_leq_bytes."""
    rng = random.Random(21)
    alphabet = "ACGTNRYKMSWBDHVacgtn"
    seqs = []
    for _ in range(500):
        L = rng.choice((20, 64, 65, 100, 130, 257))
        a = [rng.choice(alphabet) for _ in range(L)]
        b = list(a)
        for _ in range(rng.randrange(0, 14)):
            b[rng.randrange(len(b))] = rng.choice(alphabet)
        seqs += ["".join(a), "".join(b)]
    sym, off, ln = _lib.encode_seqs(seqs)
    assert sym.max() > 7
    a = np.arange(0, len(seqs), 2, dtype=np.uint32)
    got = dist.lev_leq(sym, off, ln, a, a + 1, 0.1)
    assert got.tolist() == [CO.seqs_equivalent(seqs[i], seqs[i + 1], 0.1) for i in a]
    assert collapse.are_seqs_equivalent("ACGTNRYKMSWacgt", "ACGTNRYKMSWacgA", 0.1) is True


def test_lev_leq_rejects_bad_input(dist):
    sym, off, ln = _lib.encode_seqs(["A" * 513, "ACGT"])
    with pytest.raises(_lib.DcbError):
        dist.lev_leq(sym, off, ln, [0], [1], 0.1)
    with pytest.raises(_lib.DcbError):
        _lib.encode_umis(["A" * 20])


@pytest.mark.parametrize("n,alphabet,length,k", [(0, "ACGT", 12, 2), (1, "ACGT", 12, 2), (700, "ACGT", 6, 2), (800, "AC", 12, 2),
                                                  (700, "ACGTNSL", 12, 1), (600, "ACG", 17, 3), (700, "ACGT", 5, 0)])
def test_umi_pairs_vs_oracle(dist, n, alphabet, length, k):
    rng = random.Random(n + k)
    umis = []
    for _ in range(n):
        L = length if rng.random() < 0.8 else max(1, length - rng.randrange(0, 3))
        umis.append("".join(rng.choice(alphabet) for _ in range(L)))
    row, col = dist.umi_pairs(_lib.encode_umis(umis), k)
    wrow, wcol = CO.umi_pairs(umis, k)
    assert np.array_equal(row, wrow) and np.array_equal(col, wcol)
    assert np.all(row < col)
    key = row * (1 << 32) + col
    assert np.all(np.diff(key) > 0)   # sorted row-major, no duplicates


def test_umi_pairs_matches_recorded_runs(dist, collapse_cases, monkeypatch):
    """The pair lists symdel returned inside the unmodified reference (collapse.py:735-742), through both forms of the
    search: all pairs, and deletion neighbourhoods forced for these short lists."""
    for force in ("1000000000", "0"):
        monkeypatch.setenv("DCB_UMI_SYMDEL_MIN", force)
        for case in collapse_cases["cases"]:
            k = case["args"]["bcthreshold"]
            row, col = dist.umi_pairs(_lib.encode_umis(case["umis"]), k)
            assert [[int(i), int(j)] for i, j in zip(row, col)] == case["pairs"]
            if len(case["umis"]) > 1:
                assert dist.last_method() == ("deletion neighbourhoods" if force == "0" and 1 <= k <= 2 else "all pairs")


@pytest.mark.parametrize("n,alphabet,length,k", [(3000, "ACGT", 6, 2), (3000, "AC", 12, 2), (2500, "ACGTNSL", 5, 1), (2500, "ACGTNSL", 12, 2), (2000, "ACG", 17, 2),
                                                  (4000, "ACGT", 8, 1), (40000, "ACGT", 12, 2)])
def test_umi_pairs_deletion_neighbourhoods_equal_all_pairs(dist, monkeypatch, n, alphabet, length, k):
    """Both forms on the same lists (mixed lengths, the padded-barcode alphabet, dense and sparse): identical output."""
    rng = random.Random(1000 * n + k)
    seen, umis = set(), []
    while len(umis) < n:
        L = length if rng.random() < 0.8 else max(1, length - rng.randrange(0, 3))
        u = "".join(rng.choice(alphabet) for _ in range(L))
        if u not in seen or rng.random() < 0.02:        # a few duplicates too (distance 0)
            seen.add(u); umis.append(u)
    codes = _lib.encode_umis(umis)
    monkeypatch.setenv("DCB_UMI_SYMDEL_MIN", "1000000000")
    r0, c0 = dist.umi_pairs(codes, k)
    assert dist.last_method() == "all pairs"
    monkeypatch.setenv("DCB_UMI_SYMDEL_MIN", "0")
    r1, c1 = dist.umi_pairs(codes, k)
    assert dist.last_method() == "deletion neighbourhoods"
    assert (len(r0) > 0 or alphabet == "ACGTNSL") and np.array_equal(r0, r1) and np.array_equal(c0, c1)


@pytest.mark.parametrize("force", ["0", "1000000000"])
def test_umi_pairs_split_over_parts(dist, monkeypatch, force):
    """dcb_umi_pairs_part: the union of the shares of a search split over 1, 2, 3 and 8 GPUs is the whole list (both forms)."""
    monkeypatch.setenv("DCB_UMI_SYMDEL_MIN", force)
    rng = random.Random(99)
    umis = list({"".join(rng.choice("ACGT") for _ in range(rng.choice((8, 8, 8, 7)))) for _ in range(9000)})
    codes = _lib.encode_umis(umis)
    row, col = dist.umi_pairs(codes, 2)
    whole = row * (1 << 32) + col
    assert len(whole) > 1000
    for n_parts in (2, 3, 8):
        got = []
        for part in range(n_parts):
            r, c = dist.umi_pairs(codes, 2, part=part, n_parts=n_parts)
            assert len(r) < len(row)
            got.append(r * (1 << 32) + c)
        assert np.array_equal(np.unique(np.concatenate(got)), whole)


def test_umi_pairs_two_million(dist):
    """BASELINE configs[3] scale: 2 M distinct random 12-nt UMIs, two edits -- ~10^8 pairs.  Every sampled pair is real,
    planted neighbours (substitution, deletion + insertion, two substitutions) are all found, the list is sorted and unique."""
    rng = np.random.default_rng(20260004)
    n = 2_000_000
    vals = np.unique(rng.integers(0, 4 ** 12, size=int(n * 1.1), dtype=np.uint64))[:n]
    rng.shuffle(vals)
    sym = np.stack([(vals >> np.uint64(2 * k)) & np.uint64(3) for k in range(12)], axis=1)
    codes = np.full(n, 12 << 58, dtype=np.uint64)
    for k in range(12):
        codes |= sym[:, k] << np.uint64(3 * k)
    row, col = dist.umi_pairs(codes, 2)
    assert dist.last_method() == "deletion neighbourhoods"
    assert len(row) > 10 * n and np.all(row < col)
    key = row.astype(np.uint64) * np.uint64(1 << 32) + col.astype(np.uint64)
    assert np.all(key[1:] > key[:-1])
    letters = np.array(list("ACGT"))

    def s(i):
        return "".join(letters[sym[i].astype(int)])
    for t in rng.integers(0, len(row), size=300):
        assert CO.levenshtein(s(int(row[t])), s(int(col[t]))) <= 2
    # exhaustive for a few UMIs: their neighbours by brute force over the whole list
    where = {int(v): i for i, v in enumerate(vals.tolist())}
    for i in rng.integers(0, n, size=3):
        a = s(int(i))
        want = set()
        # every string within two edits of a that is in the list (generate by two rounds of single edits)
        def edits(x):
            out = set()
            for p in range(len(x) + 1):
                for ch in "ACGT":
                    out.add(x[:p] + ch + x[p:])
                if p < len(x):
                    out.add(x[:p] + x[p + 1:])
                    for ch in "ACGT":
                        out.add(x[:p] + ch + x[p + 1:])
            return out
        near = set()
        for y in edits(a):
            near |= edits(y)
        for y in near:
            if len(y) == 12 and y != a:
                v = sum("ACGT".index(ch) << (2 * k) for k, ch in enumerate(y))
                if v in where:
                    want.add(where[v])
        lo, hi = np.searchsorted(row, i), np.searchsorted(row, i, side="right")
        got = set(col[lo:hi].tolist()) | set(row[col == i].tolist())
        assert got == want, (int(i), len(got), len(want))


def test_umi_pairs_large_properties(dist):
    """200 k UMIs (2e10 pair tests on the GPU): size-independent properties instead of the brute-force oracle."""
    rng = np.random.default_rng(5)
    n = 200_000
    base = rng.integers(0, 4, size=(n, 12), dtype=np.uint64)
    codes = np.full(n, 12 << 58, dtype=np.uint64)
    for k in range(12):
        codes |= base[:, k] << np.uint64(3 * k)
    row, col = dist.umi_pairs(codes, 1)
    assert np.all(row < col)
    # every reported pair is really within one edit (sampled), and planted neighbours are all found
    letters = np.array(list("ACGT"))
    def s(i):
        return "".join(letters[base[i].astype(int)])
    for t in rng.integers(0, len(row), size=300):
        assert CO.levenshtein(s(int(row[t])), s(int(col[t]))) <= 1
    planted = base[:1000].copy()
    planted[:, 5] = (planted[:, 5] + 1) % 4
    codes2 = codes.copy()
    extra = np.full(1000, 12 << 58, dtype=np.uint64)
    for k in range(12):
        extra |= planted[:, k] << np.uint64(3 * k)
    row2, col2 = dist.umi_pairs(np.concatenate([codes2, extra]), 1)
    found = set(zip(row2.tolist(), col2.tolist()))
    assert all((i, n + i) in found for i in range(1000))


@pytest.fixture(scope="module")
def oligo_cases(golden_dir):
    """Runs of the unmodified reference for the oligo designs beside M13 (oracle/make_golden_collapse_oligos.py)."""
    with gzip.open(os.path.join(golden_dir, "collapse_cases_oligos.json.gz"), "rt") as fh:
        return json.load(fh)


def test_collapse_stage_matches_reference_runs_other_oligos(oligo_cases):
    """I8 (two cases), I8_single, NEBIO, TAKARA: groups, UMI pairs, clusters, .freq rows and the counters of read_in_data."""
    assert [c["args"]["oligo"] for c in oligo_cases["cases"]] == ["I8", "I8", "I8_single", "NEBIO", "TAKARA"]
    for case in oligo_cases["cases"]:
        collapse_checks.check_case(case)


def test_collapse_stage_matches_reference_runs(collapse_cases):
    for case in collapse_cases["cases"]:
        collapse_checks.check_case(case)


@pytest.mark.parametrize("chain,name", [("a", "alpha"), ("b", "beta")])
def test_collapse_reproduces_golden_freq(golden_dir, tmp_path, chain, name):
    collapse_checks.check_tiny_freq(golden_dir, tmp_path, chain, name)


def test_reference_unit_answers_on_gpu():
    collapse_checks.check_reference_unit_answers()


def _host_barcode(bc, q, args, qp):
    """What collapse.py does for one row on the host (the reference's code path): -> (status, n1len, barcode)."""
    c = coll.Counter()
    saved = collapse.counts
    collapse.counts = coll.Counter()
    try:
        locs = collapse.get_barcode_positions(bc, args, c)
        if not locs:
            for key, st in (("getbarcode_fail_N", _lib.BC_FAIL_N), ("getbarcode_fail_nospacerfound", _lib.BC_FAIL_NOSPACER),
                            ("getbarcode_fail_not2spacersfound", _lib.BC_FAIL_NOT2), ("getbarcode_fail_n1tooshort", _lib.BC_FAIL_N1SHORT),
                            ("getbarcode_fail_n1toolong", _lib.BC_FAIL_N1LONG), ("getbarcode_fail_n2pastend", _lib.BC_FAIL_N2END)):
                if c[key]:
                    return st, 0, None, c
            raise AssertionError(c)
        fields = [None] * 8 + [bc, q]
        barcode, bq = collapse.set_barcode(fields, locs, args)
        bad = collapse.check_umi_quality(bq, qp)
        return (_lib.BC_FAIL_QUALITY if bad else _lib.BC_OK), locs[1] - locs[0], barcode, c
    finally:
        collapse.counts = saved


@pytest.mark.parametrize("oligo", ["M13", "I8"])
@pytest.mark.parametrize("allow_ns", [False, True])
def test_barcode_kernel_equals_the_host_path(dist, oligo, allow_ns):
    """dcb_barcodes row by row against get_barcode_positions + set_barcode + check_umi_quality (collapse.py:281-479): every
    row the kernel decides must carry the host path's verdict, N1 length and barcode; rows it hands back (BC_HOST) are
    exactly those whose spacers are not found by an exact search (or that hold odd symbols)."""
    rng = random.Random(7 if oligo == "M13" else 8)
    sp1, sp2 = ("GTCGTGACTGGGAAAACCCTGG", "GTCGTGAT") if oligo == "M13" else ("GTCGTGAT", "GTCGTGAT")
    bcs, quals = [], []
    for i in range(6000):
        n1 = rng.choice((6, 6, 6, 6, 5, 4, 7, 8, 3, 9, 2, 10))
        s = "".join(rng.choice("ACGT") for _ in range(rng.choice((0, 0, 0, 1, 2)))) + sp1 + "".join(rng.choice("ACGT") for _ in range(n1)) + sp2 + \
            "".join(rng.choice("ACGT") for _ in range(rng.choice((6, 6, 6, 8, 5, 3))))
        s = list(s)
        r = rng.random()
        if r < 0.15:                       # substitutions: some fall into a spacer (fuzzy search on the host), some into N1 / N2
            for _ in range(rng.randrange(1, 3)):
                s[rng.randrange(len(s))] = rng.choice("ACGTN")
        elif r < 0.2 and s:
            del s[rng.randrange(len(s))]
        elif r < 0.25:
            s.insert(rng.randrange(len(s) + 1), rng.choice("ACGT"))
        elif r < 0.27:
            s[rng.randrange(len(s))] = rng.choice("acgtRY")
        elif r < 0.3:                      # a second copy of a spacer
            s += list(sp2)
        s = "".join(s)[:rng.choice((42, 42, 42, 60, 30))]
        q = "".join(rng.choice("IIIIIIIIIIFFF:5,#!") for _ in range(len(s)))
        if i % 50 == 0:
            q = q[:-1]                     # a quality string of another length: host
        bcs.append(s); quals.append(q)
    bcs += ["", "A", sp1, sp1 + "ACGTAC" + sp2, sp1 + "ACGTAC" + sp2 + "ACGTA"]
    quals += ["", "I", "I" * len(sp1), "I" * (len(sp1) + 14), "I" * (len(sp1) + 19)]
    args = {"oligo": oligo, "allowNs": allow_ns, "sampling_analysis": False}
    for qp in ([20, 1, 30], [30, 0, 35.5], [2, 5, 0]):
        status, n1, code = dist.barcodes(bcs, quals, _lib.OLIGOS_ON_DEVICE[oligo.lower()], allow_ns, *qp)
        decided = 0
        for i, (bc, q) in enumerate(zip(bcs, quals)):
            if status[i] == _lib.BC_HOST:
                # handed back: no exact hit of one of the spacers in its window, odd symbols, or a ragged quality string
                lo_hi = bc[:10 + len(sp1)]
                assert (sp1 not in lo_hi) or (sp2 not in bc[len(sp1):]) or set(bc) - set("ACGTN") or len(q) != len(bc), (i, bc)
                continue
            decided += 1
            st, hn1, barcode, _ = _host_barcode(bc, q, args, qp)
            assert status[i] == st, (i, bc, q, status[i], st)
            if st in (_lib.BC_OK, _lib.BC_FAIL_QUALITY):
                assert n1[i] == hn1
                got = "".join("ACGTNSL"[(int(code[i]) >> (3 * k)) & 7] for k in range(int(code[i]) >> 58))
                assert got == barcode, (i, bc, got, barcode)
        assert decided > 0.6 * len(bcs)
