

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


def _model_words(buf, off, lens, rc, sw):
    """2-bit packing restated with numpy (A0 C1 G2 T3, base i of the read to analyse in bits 2 (i % 16) of word i // 16;
    reverse strand: the read backwards, codes complemented) -- no SIMD unit involved."""
    import numpy as np
    code = np.zeros(256, dtype=np.uint32)
    for k, ch in enumerate(b"ACGT"):
        code[ch] = k
    out = np.zeros((len(lens), sw), dtype=np.uint32)
    for r, (o, L) in enumerate(zip(off.tolist(), lens.tolist())):
        c = code[buf[o:o + L]]
        if rc:
            c = 3 - c[::-1]
        pad = np.zeros(sw * 16, dtype=np.uint32)
        pad[:L] = c
        out[r] = (pad.reshape(sw, 16) << (2 * np.arange(16, dtype=np.uint32))[None, :]).sum(axis=1, dtype=np.uint64).astype(np.uint32)
    return out


def test_simd_packers_equal_the_numpy_model():
    """Every length from 0 to 330 on both strands: the AVX-512 unit (reads of 64 bases and more), the AVX2 unit behind it and
    the byte loop write the words of the model; dcb_pack_words and dcb_pack_reads alike."""
    import numpy as np
    from decombinator_b200 import _lib
    rng = np.random.default_rng(5)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    lens = np.concatenate([np.arange(0, 331), rng.integers(0, 331, size=400)]).astype(np.uint32)
    off = np.zeros(len(lens), dtype=np.uint64)
    off[1:] = np.cumsum(lens[:-1], dtype=np.uint64)
    buf = acgt[rng.integers(0, 4, size=int(lens.sum()) + 64)].copy()
    for rc in (False, True):
        ref = _lib.pack_arrays(buf, off, lens, revcomp=rc)
        sw = ref.slot_words
        want = _model_words(buf, off, lens, rc, sw)
        assert np.array_equal(ref.arrays()["words"].reshape(len(lens), sw), want)
        ref.free()
        got, clean = _lib.pack_words(buf, off, lens, rc, sw, n_threads=2)
        if clean:
            assert np.array_equal(got.reshape(len(lens), sw), want)
        # a symbol beyond the four anywhere in a read is noticed by whichever unit packs it
        for L, at in ((250, 0), (250, 63), (250, 64), (250, 191), (250, 249), (64, 63), (130, 129), (40, 39)):
            one = acgt[rng.integers(0, 4, size=L)].copy()
            one[at] = ord("N")
            _, clean = _lib.pack_words(one, None, None, rc, (L + 63) // 64 * 4, uniform_len=L)
            assert not clean
