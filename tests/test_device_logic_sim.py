"""CPU-only: the kernels' per-read device code (csrc/dcr_core.cuh), compiled for the host by tests/sim, against
the reference fixtures and the oracle.  This checks the LOGIC the GPU runs on a box without a GPU; the real
parity tests are the -m gpu ones, which call the CUDA kernels through the C ABI."""
import numpy as np
import pytest

import decombine_oracle as O
import simlib
from decombinator_b200 import _lib, tags
from helpers import assert_records_equal, record_to_list, synth_batch


@pytest.mark.parametrize("general_only,use_union,use_q", [(False, None, False), (False, False, False), (True, None, False),
                                                          (False, None, True)])
def test_sim_matches_reference_fixtures(dcr_cases, general_only, use_union, use_q):
    names = dcr_cases["counters"]
    for gi, g in enumerate(dcr_cases["groups"]):
        info = tags.load(g["species"], g["tags"], g["chain"])
        vt, jt = info.tables()
        if use_q and _lib.union_index(vt, jt) is None:
            continue   # the flat kernel's tables exist for chains whose V and J tags share the seed geometry
        packed = _lib.pack_strings(g["reads"], revcomp=(g["orientation"] != "forward"))
        res, cnt, _ = simlib.sim_decombine(packed, vt, jt, both_frames=(g["orientation"] == "both"), allow_ns=g["allowNs"],
                                           lenthreshold=g["lenthreshold"], general_only=general_only, use_union=use_union,
                                           use_q=use_q)
        got = [record_to_list(r, rec, g["orientation"]) for r, rec in zip(g["reads"], res)]
        bad = [i for i, (a, b) in enumerate(zip(got, g["results"])) if a != b]
        assert not bad, (gi, bad[:5])
        assert {n: int(c) for n, c in zip(names, cnt)} == g["totals"], gi
        packed.free()


@pytest.mark.parametrize("species,tagset,chain,orientation,L,sub,nrate", [
    ("human", "extended", "b", "reverse", 250, 0.0, 0.0),
    ("human", "extended", "b", "reverse", 250, 0.01, 0.001),
    ("human", "extended", "a", "both", 150, 0.01, 0.001),
    ("human", "original", "b", "reverse", 300, 0.02, 0.002),
    ("mouse", "original", "g", "reverse", 250, 0.005, 0.0),
    ("mouse", "original", "d", "both", 250, 0.005, 0.001),
])
def test_sim_matches_oracle_on_synthetic(species, tagset, chain, orientation, L, sub, nrate):
    info = tags.load(species, tagset, chain)
    vt, jt = info.tables()
    n = 30000
    r1, off, ln = synth_batch(info, n, L, sub, nrate, 0.05, seed=11)
    orc = O.Oracle(O.TagSet(species, tagset, chain))
    want = orc.decombine_arrays(r1, off, ln, orientation, nthreads=4)
    packed = _lib.pack_arrays(r1, off, ln, revcomp=(orientation != "forward"))
    res, cnt, deferred = simlib.sim_decombine(packed, vt, jt, both_frames=(orientation == "both"))
    assert_records_equal(res, want, orientation)
    assert np.array_equal(cnt, orc.counts)
    if sub == 0.0:
        assert deferred < 0.1 * n  # the exact-tag path must carry clean data on its own
    if _lib.union_index(vt, jt) is not None:   # the flat kernel's tables find exactly the same tags
        res_q, cnt_q, deferred_q = simlib.sim_decombine(packed, vt, jt, both_frames=(orientation == "both"), use_q=True)
        assert np.array_equal(res_q, res) and np.array_equal(cnt_q, cnt)
        assert deferred_q <= deferred          # the flat kernel also finishes reads whose non-ACGT symbols lie outside the V-J span
        if nrate > 0:
            assert deferred_q < deferred
    packed.free()


def test_sim_half_tag_path_matches_reference_fixtures(dcr_cases):
    """The half-tag path (dcr_half_read: what dcb_halftag_kernel runs per read) between the flat kernel's search and the
    general path, on the fixtures recorded from the reference (fuzzed edge cases: it passes the reads outside its
    interior case on, but must decide most of what the exact search queues)."""
    names = dcr_cases["counters"]
    ran = 0
    for gi, g in enumerate(dcr_cases["groups"]):
        info = tags.load(g["species"], g["tags"], g["chain"])
        vt, jt = info.tables()
        if g["orientation"] == "both" or _lib.union_index(vt, jt) is None or _lib.half_index(vt, jt) is None:
            continue
        packed = _lib.pack_strings(g["reads"], revcomp=(g["orientation"] != "forward"))
        res, cnt, nd, nd2 = simlib.sim_decombine(packed, vt, jt, allow_ns=g["allowNs"], lenthreshold=g["lenthreshold"],
                                                 use_q=True, use_half=True, want_deferred2=True)
        got = [record_to_list(r, rec, g["orientation"]) for r, rec in zip(g["reads"], res)]
        bad = [i for i, (a, b) in enumerate(zip(got, g["results"])) if a != b]
        assert not bad, (gi, bad[:5])
        assert {n: int(c) for n, c in zip(names, cnt)} == g["totals"], gi
        assert nd > 0 and nd2 <= 0.5 * nd
        ran += 1
        packed.free()
    assert ran >= 5


@pytest.mark.parametrize("chain,L,sub,nrate", [("b", 250, 0.01, 0.001), ("a", 250, 0.01, 0.001), ("b", 300, 0.03, 0.005),
                                               ("a", 100, 0.02, 0.01)])
def test_sim_half_tag_path_matches_oracle(chain, L, sub, nrate):
    info = tags.load("human", "extended", chain)
    vt, jt = info.tables()
    n = 40000
    r1, off, ln = synth_batch(info, n, L, sub, nrate, 0.05, seed=77)
    orc = O.Oracle(O.TagSet("human", "extended", chain))
    want = orc.decombine_arrays(r1, off, ln, "reverse", nthreads=4)
    packed = _lib.pack_arrays(r1, off, ln, revcomp=True)
    res, cnt, nd, nd2 = simlib.sim_decombine(packed, vt, jt, use_q=True, use_half=True, want_deferred2=True)
    assert_records_equal(res, want, "reverse")
    assert np.array_equal(cnt, orc.counts)
    assert nd > 0.2 * n and nd2 < 0.05 * n          # the half-tag path decides nearly all the exact search queues
    packed.free()


def test_half_index_covers_j_only_with_long_half_tags():
    """The sampled index needs half tags of >= 10 bases: every chain has one for its V side (the V split is 10); the J
    side is in it only for the human alpha / beta `extended` sets.  A 6-base J split leaves it out (j_ok = 0); those J
    halves get the direct 6-mer table instead (j_short = 1)."""
    for sp, ts, ch, j_ok in (("human", "extended", "a", 1), ("human", "extended", "b", 1), ("human", "original", "a", 0),
                             ("human", "original", "b", 0), ("mouse", "original", "g", 0), ("mouse", "original", "d", 0),
                             ("human", "extended", "g", 0), ("mouse", "extended", "a", 0)):
        vt, jt = tags.load(sp, ts, ch).tables()
        hx = _lib.half_index(vt, jt)
        assert hx is not None
        assert int(hx[17]) == j_ok, (sp, ts, ch, hx[:20])       # DcbHalfIndex.j_ok
        assert int(hx[18]) == 1 - j_ok and (int(hx[19]) > 0) == (j_ok == 0)      # .j_short, .jt_off


@pytest.mark.parametrize("species,tagset,chain,L,sub,nrate", [("mouse", "original", "g", 250, 0.005, 0.0), ("mouse", "original", "d", 250, 0.005, 0.001),
                                                              ("human", "original", "b", 250, 0.01, 0.001), ("human", "original", "a", 150, 0.01, 0.0),
                                                              ("mouse", "extended", "a", 250, 0.01, 0.0), ("human", "extended", "d", 200, 0.02, 0.0),
                                                              ("human", "original", "a", 250, 0.01, 0.002)])   # flat kernel + N: the scan sees the invalid-base column
def test_sim_half_tag_path_short_j_halves(species, tagset, chain, L, sub, nrate):
    """Chains with a 6-base J split (12-nt J tags): their J halves -- and, behind an exact-tag kernel that only searches J
    in reads with one full V tag, the full J tags -- are found with the 6-mer table at every base of the reads whose V is
    assigned.  Mixed with reads of another chain (no V tag at all), as in a file analysed once per chain."""
    info = tags.load(species, tagset, chain)
    other = tags.load(species, tagset, {"g": "d", "d": "g", "a": "b", "b": "a"}[chain])
    vt, jt = info.tables()
    n = 30000
    r1, off, ln = synth_batch(info, n, L, sub, nrate, 0.02, seed=31, sets=[(info.v_regions, info.j_regions), (other.v_regions, other.j_regions)])
    orc = O.Oracle(O.TagSet(species, tagset, chain))
    want = orc.decombine_arrays(r1, off, ln, "reverse", nthreads=4)
    packed = _lib.pack_arrays(r1, off, ln, revcomp=True)
    use_q = _lib.union_index(vt, jt) is not None
    res, cnt, nd, nd2 = simlib.sim_decombine(packed, vt, jt, use_q=use_q, use_half=True, want_deferred2=True)
    assert_records_equal(res, want, "reverse")
    assert np.array_equal(cnt, orc.counts)
    assert nd > 0.5 * n
    if L >= 200:
        # next to nothing is left for the general path -- reads with non-ACGT symbols included: behind the bit-filter exact
        # kernels (which do not search them) both genes are handed over "unknown" and the half-tag path compares the full tags
        assert nd2 < 0.05 * n, (nd, nd2)
    else:
        assert nd2 < 0.5 * nd                        # short reads cut tags off
    packed.free()


def sweep_reads(info, L, seed=5):
    """Every tag of the chain at every start position modulo the seed stride and at both read ends: one V tag and one
    J tag per read (in the oriented frame), random bases elsewhere.  Exercises every sampling phase of the seed index,
    tags cut off by the read ends, and windows that reach into the padding around a packed read."""
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    reads = []
    vs, js = list(info.v_seqs), list(info.j_seqs)
    for k, vtag in enumerate(vs):
        for pv in list(range(0, 17)) + [L // 2 - 40 + d for d in range(9)]:
            jtag = js[(k + pv) % len(js)]
            for pj in (pv + len(vtag) + 30 + (pv % 9), L - len(jtag) - (pv % 4), L - len(jtag) + 1 + (pv % 3)):
                body = bytearray(acgt[rng.integers(0, 4, L)].tobytes())
                body[pv:pv + len(vtag)] = vtag.encode()
                if pj + len(jtag) <= L:
                    body[pj:pj + len(jtag)] = jtag.encode()
                else:                                  # J tag cut off by the read end
                    body[pj:L] = jtag.encode()[:L - pj]
                reads.append(bytes(body[:L]).translate(comp)[::-1].decode())   # stored as the reverse complement
    return reads


@pytest.mark.parametrize("tagset,chain,L", [("extended", "b", 250), ("extended", "a", 150), ("original", "a", 128)])
def test_sim_tag_position_sweep(tagset, chain, L):
    info = tags.load("human", tagset, chain)
    vt, jt = info.tables()
    reads = sweep_reads(info, L)
    orc = O.Oracle(O.TagSet("human", tagset, chain))
    want = orc.decombine_reads(reads, "reverse")
    assert want["ok"].sum() > 0.3 * len(reads)
    packed = _lib.pack_strings(reads, revcomp=True)
    for use_q in (False, True):
        res, cnt, _ = simlib.sim_decombine(packed, vt, jt, use_q=use_q)
        assert_records_equal(res, want, "reverse", "sweep use_q=%s" % use_q)
        assert np.array_equal(cnt, orc.counts)
    packed.free()


@pytest.mark.parametrize("species,tagset,chain,orientation", [("human", "extended", "b", "reverse"), ("human", "original", "b", "both"),
                                                              ("mouse", "original", "g", "reverse")])
def test_sim_marks_and_hit_list_change_nothing(species, tagset, chain, orientation):
    """The general path with the union suffix filter (candidate marks + hit list) against the same code scanning every
    position: identical records and counters on reads with substitutions, N and junk -- the marks may only skip
    positions where no keyword of any of the six sets can end (6-mer keywords of the `original` J sets included)."""
    info = tags.load(species, tagset, chain)
    vt, jt = info.tables()
    r1, off, ln = synth_batch(info, 6000, 200, 0.02, 0.004, 0.1, seed=23)
    packed = _lib.pack_arrays(r1, off, ln, revcomp=(orientation != "forward"))
    both = orientation == "both"
    a = simlib.sim_decombine(packed, vt, jt, both_frames=both, general_only=True, use_marks=True)
    b = simlib.sim_decombine(packed, vt, jt, both_frames=both, general_only=True, use_marks=False)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    sf = _lib.suffix_filter(vt, jt)
    assert sf[0] == min(min(len(t), len(t) - s, s) for ts, s in ((info.v_seqs, info.v_half_split), (info.j_seqs, info.j_half_split))
                        for t in ts)      # kq = the shortest keyword of the six sets
    packed.free()
