

def test_pack_words_equals_pack_reads_on_clean_reads():
    """dcb_pack_words (the host's share of dcb_decombine_ascii: AVX2, A / C / G / T only) writes the words dcb_pack_reads
    writes, for both strands, ragged and uniform lengths, and says so when another symbol turns up."""
    import numpy as np
    from decombinator_b200 import _lib
    rng = np.random.default_rng(11)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    lens = np.concatenate([rng.integers(0, 321, size=3000), [0, 1, 15, 16, 17, 31, 32, 33, 63, 64, 65, 250, 320]]).astype(np.uint32)
    off = np.zeros(len(lens), dtype=np.uint64)
    off[1:] = np.cumsum(lens[:-1], dtype=np.uint64)
    buf = acgt[rng.integers(0, 4, size=int(lens.sum()))].copy()
    for rc in (False, True):
        ref = _lib.pack_arrays(buf, off, lens, revcomp=rc)
        sw = ref.slot_words
        want = ref.arrays()["words"].reshape(len(lens), sw).copy()
        got, clean = _lib.pack_words(buf, off, lens, rc, sw, n_threads=3)
        if not clean:
            ref.free()
            return                        # a CPU without AVX2: nothing to compare
        assert np.array_equal(got.reshape(len(lens), sw), want)
        part, clean = _lib.pack_words(buf, off, lens, rc, sw, first=1000, count=500)
        assert clean and np.array_equal(part.reshape(500, sw), want[1000:1500])
        ref.free()
    L = 250
    uni = acgt[rng.integers(0, 4, size=L * 5000)].copy()
    ref = _lib.pack_arrays(uni, np.arange(5000, dtype=np.uint64) * L, np.full(5000, L, dtype=np.uint32), revcomp=True)
    got, clean = _lib.pack_words(uni, None, None, True, ref.slot_words, uniform_len=L)
    assert clean and np.array_equal(got, ref.arrays()["words"][:5000 * ref.slot_words])
    ref.free()
    for bad in (b"N", b"a", b"U", b"\xc1"):
        dirty = uni.copy()
        dirty[L * 4321 + 77] = bad[0]
        _, clean = _lib.pack_words(dirty, None, None, True, 16, uniform_len=L)
        assert not clean
