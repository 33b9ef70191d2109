"""Stand-in for polyleven==0.8: levenshtein(a, b[, k]) unit-cost edit distance (collapse.py:360,364)."""
from Levenshtein import distance as _d


def levenshtein(a, b, k=-1):
    d = _d(a, b)
    if k is not None and k >= 0 and d > k:
        return k + 1
    return d
