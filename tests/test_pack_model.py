"""CPU-only: dcb_pack_reads (csrc/pack.cpp, incl. its eight-bases-at-a-time path) against a plain Python model of the
packing contract: 2 bits per base in the orientation analysed (reverse complement through Bio.Seq's table,
decombine.py:182-184), every non-ACGT symbol as base 0 + an entry of the sparse exception list."""
import numpy as np
import pytest

from decombinator_b200 import _lib

_COMP = str.maketrans("ACGTUMRWSYKVHDBNacgtumrwsykvhdbn", "TGCAAKYWSRMBDHVNtgcaakywsrmbdhvn")


@pytest.mark.parametrize("revcomp", [False, True])
def test_packer_matches_model(revcomp):
    rng = np.random.default_rng(5)
    # letters next to A/C/G/T in ASCII ('@', 'B', 'D', 'F', 'H', 'S', 'U', '`') catch sloppy range tests
    alphabet = list("ACGT") * 12 + list("NacgtURY@BDFHSU`")
    reads = []
    for _ in range(3000):
        L = int(rng.integers(0, 300))
        pool = list("ACGT") if rng.random() < 0.5 else alphabet
        reads.append("".join(rng.choice(pool, L)))
    reads += ["", "A", "ACGTACGT", "ACGTACGTN", "NACGTACGT", "ACGTACGTACGTACGT", "ACGTACGTACGTACGTA", "U" * 9, "acgtacgtACGTACGT"]
    P = _lib.pack_strings(reads, revcomp=revcomp)
    c = P.c.contents if hasattr(P.c, "contents") else P.c
    sw, nexc = c.slot_words, c.n_exc
    words = np.ctypeslib.as_array(c.words, shape=(len(reads) * sw,)).reshape(len(reads), sw)
    got_exc = sorted(zip(np.ctypeslib.as_array(c.exc_read, shape=(max(nexc, 1),))[:nexc].tolist(),
                         np.ctypeslib.as_array(c.exc_pos, shape=(max(nexc, 1),))[:nexc].tolist(),
                         np.ctypeslib.as_array(c.exc_kind, shape=(max(nexc, 1),))[:nexc].tolist()))
    flags = np.ctypeslib.as_array(c.flags, shape=((len(reads) + 31) // 32,))
    want_exc = []
    for r, s in enumerate(reads):
        o = s.translate(_COMP)[::-1] if revcomp else s
        w = [0] * sw
        flagged = False
        for i, ch in enumerate(o):
            code = "ACGT".find(ch)
            if code < 0:
                want_exc.append((r, i, 1 if ch == "N" else 2)); code = 0; flagged = True
            elif revcomp and s[len(s) - 1 - i] == "U":
                want_exc.append((r, i, 3)); flagged = True        # a real base here, not in the other frame
            w[i >> 4] |= code << (2 * (i & 15))
        assert words[r].tolist() == w, (r, s)
        assert bool((int(flags[r >> 5]) >> (r & 31)) & 1) == flagged, (r, s)
    assert got_exc == sorted(want_exc)
    P.free()
