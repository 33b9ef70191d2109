"""CPU-only: the native FASTQ record index (csrc/fastq.cpp, dcb_fastq_index_build) against the general parser, which
keeps the reference's readfq semantics (decombine.py:228-265): same records on strict four-line files, and a clean
hand-over ("not strict") for everything whose handling depends on readfq's quirks."""
import gzip
import os

import numpy as np
import pytest

from decombinator_b200 import _lib, fastq


def _args(path, bc_read="R2", bclength=42, sampling=False, python_fastq=False):
    return {"infile": str(path), "bc_read": bc_read, "bclength": bclength, "sampling_analysis": sampling,
            "python_fastq": python_fastq}


def _write_pair(tmp_path, recs1, recs2, gz=False, tail="\n"):
    def text(recs):
        return "".join("@%s\n%s\n+\n%s%s" % (n, s, q, "\n") for n, s, q in recs)[:-1] + tail
    p1, p2 = tmp_path / ("s_1.fq" + (".gz" if gz else "")), tmp_path / ("s_2.fq" + (".gz" if gz else ""))
    op = gzip.open if gz else open
    with op(p1, "wt") as fh:
        fh.write(text(recs1))
    with op(p2, "wt") as fh:
        fh.write(text(recs2))
    return p1


def _columns(batch):
    return {k: list(getattr(batch, k)) for k in ("ids", "vdj", "vdjqual", "bc", "bcq", "v_tail")}


def _random_records(rng, n, lo=30, hi=80, names=None):
    out = []
    for i in range(n):
        L = int(rng.integers(lo, hi))
        seq = "".join(rng.choice(list("ACGTN"), L, p=[0.24, 0.24, 0.24, 0.24, 0.04]))
        qual = "".join(rng.choice(list("#,:F@+>I"), L))          # quality lines may start with '@', '+', '>'
        name = (names[i] if names else "SYN:%d extra words 1:N:0" % i)
        out.append((name, seq, qual))
    return out


@pytest.mark.parametrize("bc_read,gz,sampling", [("R2", False, False), ("R2", True, True), ("R1", False, True), ("R1", True, False)])
def test_native_index_matches_general_parser(tmp_path, bc_read, gz, sampling):
    rng = np.random.default_rng(3)
    recs1 = _random_records(rng, 501)
    recs2 = _random_records(rng, 480)                             # the shorter file ends the zip
    recs1[7] = ("", "ACGT", "FFFF")                               # empty name
    recs1[9] = ("name_without_space", "", "")                     # empty read
    recs2[11] = ("x y", "ACG", "FFFFFF")                          # quality longer than the sequence
    p1 = _write_pair(tmp_path, recs1, recs2, gz=gz)
    opener = gzip.open if gz else open
    for bl in (0, 6, 42, 100):
        native = fastq.load_pairs_native(_args(p1, bc_read, bl, sampling), opener)
        assert native is not None
        general = fastq.load_pairs(_args(p1, bc_read, bl, sampling, python_fastq=True), opener)
        assert _columns(native) == _columns(general)
        assert np.array_equal(native.len, general.len)
        got = [bytes(native.buf[o:o + n]) for o, n in zip(native.off.tolist(), native.len.tolist())]
        want = [bytes(general.buf[o:o + n]) for o, n in zip(general.off.tolist(), general.len.tolist())]
        assert got == want
        assert fastq.count_containing(native.bc, "N") == fastq.count_containing(general.bc, "N")
        part = native.shard(100, 150)
        assert list(part.vdj) == list(general.vdj)[100:150] and len(part) == 50


@pytest.mark.parametrize("text", [
    "",                                                           # empty
    "@a\nACGT\n+\nFFFF",                                          # last line not terminated (readfq drops its last char)
    "@a\r\nACGT\r\n+\r\nFFFF\r\n",                                # CRLF
    "@a\nAC\nGT\n+\nFFFF\n",                                      # multi-line sequence
    "@a\nACGT\n+\nFF\nFF\n",                                      # multi-line quality
    ">a\nACGT\n>b\nACGT\n",                                       # FASTA records
    "@a\nACGT\n+\nFFFF\n@b\nACGT\n+\n",                           # partial last record
    "@a\n+CGT\n+\nFFFF\n",                                        # sequence line that readfq takes for a marker
    "@a\nACGT\n+\nFFFF\n\n@b\nACGT\n+\nFFFF\n",                   # blank line between records
    "@a\nAC\xc3\xa9T\n+\nFFFFF\n",                                # non-ASCII
])
def test_non_strict_layouts_are_handed_over(text):
    assert _lib.fastq_index(text.encode("latin-1")) is None


def test_tiny_files_of_the_reference(tmp_path):
    """The reference's own test files (bundled copy used by the golden tests) are strict and index identically."""
    here = os.path.dirname(os.path.abspath(__file__))
    cands = [os.path.join(here, "golden", "TINY_1.fq"), os.path.join(here, "golden", "TINY_1.fq.gz")]
    path = next((c for c in cands if os.path.exists(c)), None)
    if path is None:
        pytest.skip("TINY_1.fq not bundled")
    opener = gzip.open if path.endswith(".gz") else open
    native = fastq.load_pairs_native(_args(path), opener)
    general = fastq.load_pairs(_args(path, python_fastq=True), opener)
    assert native is not None and _columns(native) == _columns(general)


@pytest.mark.parametrize("pack_rc", [True, False])
@pytest.mark.parametrize("with_tail", [True, False])
def test_format_rows_matches_the_reference_row_assembly(pack_rc, with_tail):
    """dcb_format_rows against the literal row assembly of the main loop (decombine.py:1015-1039): both strands,
    IUPAC / lower-case symbols through the reverse complement, quality lines longer than their read, slices that run
    past the end."""
    from decombinator_b200.decombine import revcomp
    rng = np.random.default_rng(1)
    n = 3000
    text = bytearray()
    cols = {k: ([], []) for k in ("ids", "vdj", "q", "bc", "bcq", "tail")}

    def add(col, s):
        cols[col][0].append(len(text)); cols[col][1].append(len(s))
        text.extend(s.encode()); text.extend(b"\n")

    for i in range(n):
        L = int(rng.integers(20, 120))
        add("ids", "R%d" % i)
        add("vdj", "".join(rng.choice(list("ACGTNacgtRYKMbdhv"), L)))
        add("q", "".join(rng.choice(list("FI#,:"), L + int(rng.integers(0, 3)))))
        add("bc", "".join(rng.choice(list("ACGTN"), 12)))
        add("bcq", "F" * 12)
        add("tail", "".join(rng.choice(list("ACGT"), int(rng.integers(0, 31)))))
    data = bytes(text)
    C = {k: fastq.TextColumn(data, np.array(v[0], np.uint64), np.array(v[1], np.uint32)) for k, v in cols.items()}
    res = np.zeros(n, dtype=_lib.RESULT_DTYPE)
    for i in range(n):
        L = int(C["vdj"].len[i])
        a = int(rng.integers(0, L)); b = int(rng.integers(a, L + 5))
        ia = int(rng.integers(a, max(a + 1, b)))
        res[i] = (rng.random() < 0.7, rng.integers(0, 2), rng.integers(0, 200), rng.integers(0, 99), 0, rng.integers(0, 30000),
                  rng.integers(0, 12), ia, int(rng.integers(ia, b + 1)) if b >= ia else ia, a, b)[:len(res.dtype.names)] \
            if False else res[i]
        res[i]["status"] = rng.random() < 0.7; res[i]["frame"] = rng.integers(0, 2)
        res[i]["v"] = rng.integers(0, 200); res[i]["j"] = rng.integers(0, 99)
        res[i]["vdel"] = rng.integers(0, 30000); res[i]["jdel"] = rng.integers(0, 12)
        res[i]["v_seq_start"] = a; res[i]["j_seq_end"] = b
        res[i]["ins_start"] = ia; res[i]["ins_end"] = int(rng.integers(ia, b + 1)) if b >= ia else ia
    blob, nrows = _lib.format_rows(res, pack_rc, (C["ids"], C["vdj"], C["q"], C["bc"], C["bcq"], C["tail"] if with_tail else None), ", ")
    want = []
    for i in np.nonzero(res["status"])[0]:
        r = res[i]
        is_rev = pack_rc != bool(r["frame"])
        a, b = int(r["v_seq_start"]), int(r["j_seq_end"])
        vdj, q = C["vdj"][i], C["q"][i]
        o = revcomp(vdj) if is_rev else vdj
        row = [str(int(r["v"])), str(int(r["j"])), str(int(r["vdel"])), str(int(r["jdel"])),
               o[int(r["ins_start"]):int(r["ins_end"])], C["ids"][i], o[a:b], (q[::-1][a:b] if is_rev else q[a:b]),
               C["bc"][i], C["bcq"][i]]
        if with_tail:
            row.append(C["tail"][i])
        want.append(", ".join(row) + "\n")
    assert blob.decode() == "".join(want)
    assert nrows == len(want)


def _bgzf(data, block=60000):
    """A BGZF file (the format bgzip and Illumina's converters write): gzip members of at most 64 KB whose extra field
    'BC' holds the member's size, closed by an empty member."""
    import struct
    import zlib
    out = []
    for lo in list(range(0, len(data), block)) + [None]:
        chunk = b"" if lo is None else data[lo:lo + block]
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        body = c.compress(chunk) + c.flush()
        bsize = 12 + 6 + len(body) + 8
        out.append(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1) + body +
                   struct.pack("<II", zlib.crc32(chunk), len(chunk)))
    return b"".join(out)


def test_gunzip_file_equals_the_gzip_module(tmp_path):
    """fastq.gunzip_file: one member, several members, BGZF blocks (inflated in parallel), zero padding, an empty file; a
    damaged file is left to the gzip module (the reader the reference uses), whose error then surfaces."""
    import gzip
    import random
    from decombinator_b200 import fastq
    rng = random.Random(3)
    text = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, bytes(rng.choice(b"ACGT") for _ in range(rng.randrange(40, 200))), b"I" * 50)
                    for i in range(20000))
    cases = {"one.gz": gzip.compress(text), "many.gz": b"".join(gzip.compress(text[lo:lo + 300000]) for lo in range(0, len(text), 300000)),
             "bgzf.gz": _bgzf(text), "padded.gz": gzip.compress(text) + b"\0" * 512, "empty_member.gz": gzip.compress(b"")}
    for name, blob in cases.items():
        path = tmp_path / name
        path.write_bytes(blob)
        want = gzip.open(path, "rb").read()
        got = fastq.gunzip_file(str(path))
        assert bytes(got) == want, name
        assert bytes(fastq._file_bytes(str(path), gzip.open)) == want, name
    assert fastq._bgzf_blocks(cases["bgzf.gz"]) is not None and fastq._bgzf_blocks(cases["one.gz"]) is None
    assert bytes(fastq.gunzip_file(str(tmp_path / "bgzf.gz"), n_threads=3)) == text
    bad = bytearray(cases["bgzf.gz"])
    bad[5000] ^= 0x55
    (tmp_path / "bad.gz").write_bytes(bytes(bad))
    with pytest.raises(Exception):
        fastq._file_bytes(str(tmp_path / "bad.gz"), gzip.open)
    (tmp_path / "cut.gz").write_bytes(cases["one.gz"][:-100])
    with pytest.raises(EOFError):
        fastq._file_bytes(str(tmp_path / "cut.gz"), gzip.open)
