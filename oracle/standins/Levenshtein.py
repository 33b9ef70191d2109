"""Stand-in for Levenshtein==0.25.1: hamming() and distance() (decombine.py:108,309..)."""


def hamming(a, b):
    if len(a) != len(b):
        raise ValueError("Sequences are not the same length.")
    return sum(1 for x, y in zip(a, b) if x != y)


def distance(a, b, score_cutoff=None):
    if len(a) < len(b):
        a, b = b, a
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    d = prev[-1]
    if score_cutoff is not None and d > score_cutoff:
        return score_cutoff + 1
    return d
