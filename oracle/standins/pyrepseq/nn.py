"""symdel(seqs, max_edits, ...) -> scipy COO matrix with an entry (i, j, d) for every
ordered pair i != j whose Levenshtein distance d <= max_edits (pyrepseq docs)."""
import numpy as np
from scipy import sparse
from Levenshtein import distance as _d


def symdel(seqs, seqs2=None, max_edits=1, max_returns=None, n_cpu=1, custom_distance=None,
           max_custom_distance=float("inf"), output_type="triplets", single_presorted=False,
           progress=False, **kw):
    seqs = list(seqs)
    n = len(seqs)
    rows, cols, vals = [], [], []
    # deletion-neighbourhood bucketing keeps this usable for a few thousand UMIs
    from itertools import combinations
    buckets = {}
    for idx, s in enumerate(seqs):
        keys = {s}
        for k in range(1, max_edits + 1):
            if k > len(s):
                break
            for pos in combinations(range(len(s)), k):
                keys.add("".join(c for i, c in enumerate(s) if i not in pos))
        for key in keys:
            buckets.setdefault(key, []).append(idx)
    seen = set()
    for members in buckets.values():
        if len(members) < 2:
            continue
        for a in range(len(members)):
            for b in range(a + 1, len(members)):
                i, j = members[a], members[b]
                if (i, j) in seen:
                    continue
                seen.add((i, j))
                d = _d(seqs[i], seqs[j])
                if d <= max_edits:
                    rows += [i, j]
                    cols += [j, i]
                    vals += [d, d]
    if output_type == "triplets":
        return list(zip(rows, cols, vals))
    return sparse.coo_matrix((np.array(vals, dtype=np.int64), (np.array(rows, dtype=np.int64), np.array(cols, dtype=np.int64))), shape=(n, n))
