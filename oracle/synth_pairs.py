"""Synthetic paired FASTQ for the collapse fixtures -- TEST INFRASTRUCTURE (shared by the scripts that record the
unmodified reference in the build container and by the tests that regenerate the same files on the GPU box; needs
nothing but the generator library libdcbsynth.so and the bundled tag sets).

Reads are copies of a pool of molecules (shared UMI + rearrangement) with sequencing errors in both reads, shuffled; the
barcode read then gets the edge cases the collapse stage has branches for: N1 one base short / long, an N, low
quality, a second molecule on a used barcode."""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def build_pairs(species, tagset, chain, n, pool, L, sub1, sub2, seed, r2_layout=None, n1_at=22):
    """-> (names, read 1, quality 1, read 2, quality 2), lists of str.  r2_layout(read 2 as the generator wrote it: M13
    spacer, N6, I8 spacer, N6, ...) -> read 2 of another oligo design; n1_at: where that design's first hexamer starts."""
    from decombinator_b200 import _lib, tags as dtags
    info = dtags.load(species, tagset, chain)
    syn = _lib.Synth([(info.v_regions, info.j_regions)], seed, L, L, sub1, 0.0, 0.02, umi_pool=pool, sub_rate2=sub2)
    r1, r2 = syn.reads(0, n, want_r2=True)
    a = [bytes(r1[i * L:(i + 1) * L]).decode() for i in range(n)]
    b = [bytes(r2[i * L:(i + 1) * L]).decode() for i in range(n)]
    if r2_layout is not None:
        b = [r2_layout(x) for x in b]
    rng = random.Random(seed)
    order = list(range(n))
    rng.shuffle(order)
    a = [a[i] for i in order]
    b = [b[i] for i in order]
    q2 = []
    for i in range(n):
        q = ["I"] * L
        r = rng.random()
        if r < 0.03:
            b[i] = b[i][:n1_at + 2] + b[i][n1_at + 3:] + "A"          # N1 of 5 bases
        elif r < 0.06:
            b[i] = b[i][:n1_at + 2] + "C" + b[i][n1_at + 2:-1]        # N1 of 7 bases
        elif r < 0.08:
            b[i] = b[i][:n1_at + 8] + "N" + b[i][n1_at + 9:]
        elif r < 0.12:
            for k in rng.sample(range(n1_at, n1_at + 20), 3):
                q[k] = "#"
        elif r < 0.15 and i > 10:
            b[i] = b[rng.randrange(0, i)]               # barcode collision with an unrelated molecule
        q2.append("".join(q))
    q1 = ["I" * L for _ in range(n)]
    names = ["SYN:%d" % i for i in range(n)]
    return names, a, q1, b, q2


def write_pairs(path1, path2, *spec, **kw):
    names, a, q1, b, q2 = build_pairs(*spec, **kw)
    with open(path1, "wt") as fh:
        for nm, s, q in zip(names, a, q1):
            fh.write("@%s\n%s\n+\n%s\n" % (nm, s, q))
    with open(path2, "wt") as fh:
        for nm, s, q in zip(names, b, q2):
            fh.write("@%s\n%s\n+\n%s\n" % (nm, s, q))
