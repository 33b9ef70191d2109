"""CPU restatement of the distance arithmetic of the collapse step -- TEST INFRASTRUCTURE (the checker), never
imported by the product.

Follows /root/reference/src/decombinator/collapse.py:
  * ``levenshtein``        unit-cost edit distance = polyleven.levenshtein (collapse.py:360, 364) and the
                           Levenshtein.distance inside pyrepseq.nn.symdel (third-party, pyrepseq==1.5 /
                           Levenshtein==0.25.1 / polyleven==0.8, absent here: the published algorithm is restated)
  * ``seqs_equivalent``    are_seqs_equivalent (collapse.py:355-360)
  * ``umi_pairs``          symdel(..., max_edits=k) + sparse.triu + sum_duplicates (collapse.py:735-742): every pair
                           row < col within k edits, ascending (row, col)

Pinned by tests/test_oracle_golden.py::test_collapse_oracle_* against fixtures recorded from the unmodified
reference (tests/golden/collapse_cases.json.gz) and against the reference's own known answers
(tests/test_collapse.py:43-52, 278-314 of the reference).
"""
import numpy as np


def levenshtein(a, b):
    """Textbook DP, one row at a time."""
    if a == b:
        return 0
    if len(a) < len(b):
        a, b = b, a
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


def seqs_equivalent(seq1, seq2, lev_threshold_fraction):
    threshold = len(min(seq1, seq2, key=len)) * lev_threshold_fraction
    return levenshtein(seq1, seq2) <= threshold


def umi_pairs(umis, max_edits):
    """-> (row, col) int64 arrays."""
    rows, cols = [], []
    n = len(umis)
    for i in range(n):
        a = umis[i]
        for j in range(i + 1, n):
            b = umis[j]
            if abs(len(a) - len(b)) <= max_edits and levenshtein(a, b) <= max_edits:
                rows.append(i)
                cols.append(j)
    return np.array(rows, dtype=np.int64), np.array(cols, dtype=np.int64)


class OracleDist:
    """Same interface as decombinator_b200._lib.Dist, computed on the CPU by the functions above.  Tests install it
    in place of the GPU context to check the HOST logic of decombinator_b200.collapse without a GPU."""

    ALPHABET = "ACGTNSL"

    def umi_pairs(self, codes, max_edits, part=0, n_parts=1):
        """(a share of a split search: the pairs whose row is `part` modulo n_parts -- any split whose union is whole)"""
        umis = []
        for c in np.asarray(codes, dtype=np.uint64).tolist():
            n = c >> 58
            umis.append("".join(chr(ord("a") + ((c >> (3 * k)) & 7)) for k in range(n)))
        row, col = umi_pairs(umis, max_edits)
        mine = row % n_parts == part
        return row[mine], col[mine]

    def lev_leq(self, symbols, off, length, a, b, frac):
        seqs = [bytes(symbols[int(o):int(o) + int(l)]) for o, l in zip(off, length)]
        out = np.zeros(len(a), dtype=bool)
        for t, (i, j) in enumerate(zip(a, b)):
            out[t] = seqs_equivalent(seqs[int(i)], seqs[int(j)], frac)
        return out

    def barcodes(self, bcs, quals, oligo, allow_ns, min_q, max_below, avg_q):
        """No device here: every row is handed back (status 255), i.e. the host runs the reference's own barcode code
        (get_barcode_positions / set_barcode / check_umi_quality) on all of them."""
        n = len(bcs)
        return np.full(n, 255, dtype=np.uint8), np.zeros(n, dtype=np.uint8), np.zeros(n, dtype=np.uint64)

    def barcodes_arrays(self, bbuf, bo, bl, qbuf, qo, ql, oligo, allow_ns, min_q, max_below, avg_q):
        """The columnar form of the same: every row handed back."""
        n = len(bo)
        return np.full(n, 255, dtype=np.uint8), np.zeros(n, dtype=np.uint8), np.zeros(n, dtype=np.uint64)
