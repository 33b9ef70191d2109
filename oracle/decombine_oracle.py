"""CPU oracle for the decombine path: Python driver around oracle/dcr_oracle.c.

TEST INFRASTRUCTURE ONLY -- the checker for the CUDA path and the reported CPU baseline.
Nothing under decombinator_b200/ imports this module; only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs do.

Parity pinning: tests/test_oracle_golden.py checks this oracle against the reference's own
golden `.n12` files (reference tests/test_pipeline.py:63-84) and against fixtures recorded
from the unmodified reference source (oracle/make_golden.py).

Citations are file:line in /root/reference/src/decombinator/.
"""
import ctypes
import gzip
import json
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")
_BUNDLE = os.path.join(os.path.dirname(_HERE), "decombinator_b200", "data", "tagsets.json.gz")

COUNTER_NAMES = [
    "verr1", "verr2", "jerr1", "jerr2",
    "dcrfilter_intertagN", "dcrfilter_toolong_intertag", "dcrfilter_imposs_deletion", "dcrfilter_tag_overlap",
    "multiple_v_matches", "v_del_failed_tag_at_end", "v_del_failed", "foundv1notv2", "foundv2notv1",
    "no_vtags_found", "multiple_j_matches", "j_del_failed", "foundj1notj2", "foundj2notj1",
    "no_j_assigned", "VJ_assignment_failed",
]

RESULT_DTYPE = np.dtype([
    ("ok", "<i4"), ("v", "<i4"), ("j", "<i4"), ("vdel", "<i4"), ("jdel", "<i4"),
    ("ins_start", "<i4"), ("ins_end", "<i4"), ("v_seq_start", "<i4"), ("j_seq_end", "<i4"), ("frame", "<i4"),
])

ORIENTATION = {"reverse": 0, "forward": 1, "both": 2}


def build(force=False):
    """gcc the C restatement into oracle/liboracle.so (building the checker is not using it)."""
    src = os.path.join(_HERE, "dcr_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "dcr_oracle.h"))):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-pthread", "-o", _LIB, src])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB)
        L.orc_create.restype = ctypes.c_void_p
        L.orc_create.argtypes = [ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_int32),
                                 ctypes.POINTER(ctypes.c_char_p), ctypes.c_int] * 2 + [ctypes.c_int, ctypes.c_int]
        L.orc_destroy.argtypes = [ctypes.c_void_p]
        L.orc_dcr.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                              ctypes.c_void_p, ctypes.c_void_p]
        L.orc_decombine.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_uint64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                    ctypes.c_void_p, ctypes.c_int]
        L.orc_revcomp.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p]
        L.orc_findall.restype = ctypes.c_int
        L.orc_findall.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p,
                                  ctypes.c_void_p, ctypes.c_int]
        _lib = L
    return _lib


# ---------------------------------------------------------------------------------------------
# Tag / FASTA loading: the file-reading half of import_tcr_info (decombine.py:593-719)
# ---------------------------------------------------------------------------------------------
_CHAIN_ALIASES = {
    "a": ["A", "ALPHA", "TRA", "TCRA"], "b": ["B", "BETA", "TRB", "TCRB"],
    "g": ["G", "GAMMA", "TRG", "TCRG"], "d": ["D", "DELTA", "TRD", "TCRD"],
}


def resolve_chain(chain):
    """decombine.py:608-619"""
    for c, names in _CHAIN_ALIASES.items():
        if chain.upper() in names:
            return c
    raise SystemExit("TCR chain not recognised")


def resolve_tagset(tags, species, chain):
    """decombine.py:640-661: mouse and gamma/delta silently fall back to the original set;
    half split is (10, 10) for extended and (10, 6) for original."""
    if tags == "extended" and species == "mouse":
        tags = "original"
    if tags == "extended" and chain in ("g", "d"):
        tags = "original"
    if tags == "extended":
        return tags, 10, 10
    if tags == "original":
        return tags, 10, 6
    raise SystemExit("Tag set unrecognised")


_bundle_cache = None


def _read_tag_text(species, tagset, chain, gene, filetype, tagdir):
    """read_tcr_file (decombine.py:187-225) minus the download: cwd, then tagdir, then the
    bundled copy of the same data files."""
    global _bundle_cache
    name = "%s_%s_TR%s%s" % (species, tagset, chain.upper(), gene.upper())
    fn = name + "." + filetype
    for cand in (fn, os.path.join(tagdir or "", fn)):
        if os.path.isfile(cand):
            with open(cand, "rt") as fh:
                return fh.read()
    if _bundle_cache is None:
        with gzip.open(_BUNDLE, "rt") as fh:
            _bundle_cache = json.load(fh)
    return _bundle_cache[name][filetype]


def parse_fasta(text):
    """SeqIO.parse(..., "fasta") -> upper-cased record sequences (decombine.py:690-696)."""
    seqs, cur = [], None
    for line in text.splitlines():
        if line.startswith(">"):
            if cur is not None:
                seqs.append("".join(cur))
            cur = []
        elif cur is not None:
            cur.append(line.strip())
    if cur is not None:
        seqs.append("".join(cur))
    return [s.upper() for s in seqs]


def parse_tags(text):
    """get_v_tags / get_j_tags (decombine.py:820-866): column 0 tag, column 1 jump."""
    seqs, jumps = [], []
    for line in text.splitlines():
        parts = line.split()
        if not parts:  # the reference would raise IndexError on a blank line; shipped files have none
            continue
        seqs.append(parts[0])
        jumps.append(int(parts[1]))
    return seqs, jumps


class TagSet:
    def __init__(self, species="human", tags="extended", chain="b", tagdir=None):
        self.chain = resolve_chain(chain)
        self.species = species
        self.tags, self.v_split, self.j_split = resolve_tagset(tags, species, self.chain)
        g = {}
        for gene in ("v", "j"):
            g[gene + "_regions"] = parse_fasta(_read_tag_text(species, self.tags, self.chain, gene, "fasta", tagdir))
            g[gene + "_seqs"], g[gene + "_jump"] = parse_tags(
                _read_tag_text(species, self.tags, self.chain, gene, "tags", tagdir))
        self.__dict__.update(g)


def _carr(strings):
    arr = (ctypes.c_char_p * len(strings))()
    arr[:] = [s.encode() for s in strings]
    return arr


class Oracle:
    """dcr() and the batch orientation loop, backed by the C restatement."""

    def __init__(self, tagset: TagSet, allowNs=False, lenthreshold=130):
        self.ts = tagset
        self.allowNs = int(bool(allowNs))
        self.lenthreshold = int(lenthreshold)
        L = lib()
        nv, nj = len(tagset.v_seqs), len(tagset.j_seqs)
        self._keep = (_carr(tagset.v_seqs), (ctypes.c_int32 * nv)(*tagset.v_jump), _carr(tagset.v_regions[:nv]),
                      _carr(tagset.j_seqs), (ctypes.c_int32 * nj)(*tagset.j_jump), _carr(tagset.j_regions[:nj]))
        k = self._keep
        self._ctx = L.orc_create(k[0], k[1], k[2], nv, k[3], k[4], k[5], nj, tagset.v_split, tagset.j_split)
        self.counts = np.zeros(len(COUNTER_NAMES), dtype=np.uint64)

    def __del__(self):
        try:
            lib().orc_destroy(self._ctx)
        except Exception:
            pass

    def reset_counts(self):
        self.counts[:] = 0

    def counts_dict(self):
        return {k: int(v) for k, v in zip(COUNTER_NAMES, self.counts)}

    def dcr(self, read: str):
        """dcr(read, inputargs) (decombine.py:534-585): list of 7 or None; bumps self.counts."""
        res = np.zeros(1, dtype=RESULT_DTYPE)
        b = read.encode("latin-1")
        lib().orc_dcr(self._ctx, b, len(b), self.allowNs, self.lenthreshold, res.ctypes.data, self.counts.ctypes.data)
        r = res[0]
        if not r["ok"]:
            return None
        return [int(r["v"]), int(r["j"]), int(r["vdel"]), int(r["jdel"]),
                read[int(r["ins_start"]):int(r["ins_end"])], int(r["v_seq_start"]), int(r["j_seq_end"])]

    def findall(self, which, read: str, cap=256):
        kw = np.zeros(cap, dtype=np.int32)
        st = np.zeros(cap, dtype=np.int32)
        b = read.encode("latin-1")
        n = lib().orc_findall(self._ctx, which, b, len(b), kw.ctypes.data, st.ctypes.data, cap)
        return [(int(kw[i]), int(st[i])) for i in range(min(n, cap))]

    def decombine_arrays(self, ascii_buf: np.ndarray, off: np.ndarray, length: np.ndarray, orientation="reverse",
                         nthreads=1):
        """Orientation loop (decombine.py:999-1010) over reads stored in one uint8 buffer."""
        ascii_buf = np.ascontiguousarray(ascii_buf, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        length = np.ascontiguousarray(length, dtype=np.uint32)
        res = np.zeros(len(off), dtype=RESULT_DTYPE)
        lib().orc_decombine(self._ctx, ascii_buf.ctypes.data, off.ctypes.data, length.ctypes.data, len(off),
                            ORIENTATION[orientation], self.allowNs, self.lenthreshold, res.ctypes.data,
                            self.counts.ctypes.data, int(nthreads))
        return res

    def decombine_reads(self, reads, orientation="reverse", nthreads=1):
        bufs = [r.encode("latin-1") for r in reads]
        length = np.array([len(b) for b in bufs], dtype=np.uint32)
        off = np.zeros(len(bufs), dtype=np.uint64)
        if len(bufs) > 1:
            off[1:] = np.cumsum(length[:-1], dtype=np.uint64)
        buf = np.frombuffer(b"".join(bufs) + b"\0", dtype=np.uint8)
        return self.decombine_arrays(buf, off, length, orientation, nthreads)


def revcomp(read: str) -> str:
    """revcomp() (decombine.py:182-184)."""
    b = read.encode("latin-1")
    out = ctypes.create_string_buffer(len(b) + 1)
    lib().orc_revcomp(b, len(b), out)
    return out.raw[:len(b)].decode("latin-1")


# ---------------------------------------------------------------------------------------------
# readfq + main loop restated (small inputs only: pure-Python loops)
# ---------------------------------------------------------------------------------------------
def readfq(fp):
    """Heng Li's readfq as shipped (decombine.py:228-265), including the unconditional l[:-1]."""
    last = None
    while True:
        if not last:
            for l in fp:
                if l[0] in ">@":
                    last = l[:-1]
                    break
        if not last:
            break
        name, seqs, last = last[1:].partition(" ")[0], [], None
        for l in fp:
            if l[0] in "@+>":
                last = l[:-1]
                break
            seqs.append(l[:-1])
        if not last or last[0] != "+":
            yield name, "".join(seqs), None
            if not last:
                break
        else:
            seq, leng, seqs = "".join(seqs), 0, []
            for l in fp:
                seqs.append(l[:-1])
                leng += len(l) - 1
                if leng >= len(seq):
                    last = None
                    yield name, seq, "".join(seqs)
                    break
            if last:
                yield name, seq, None
                break


def decombinator_rows(infile, chain, bc_read="R2", bclength=42, orientation="reverse", tags="extended",
                      species="human", allowNs=False, lenthreshold=130, tagdir=None):
    """The barcoded branch of the main loop (decombine.py:950-1039) -> (rows, counts dict)."""
    ts = TagSet(species, tags, chain, tagdir)
    orc = Oracle(ts, allowNs, lenthreshold)
    counts = {"read_count": 0, "vj_count": 0, "dcrfilter_barcodeN": 0}
    opener = gzip.open if infile.endswith(".gz") else open
    fq1 = readfq(opener(infile, "rt"))
    fq2 = readfq(opener(infile.replace("1.f", "2.f"), "rt")) if bc_read == "R2" else fq1
    rows = []
    for record1, record2 in zip(fq1, fq2):
        if bc_read == "R2":
            readid, vdj, vdjqual = record1
            bc, bcQ = record2[1][:bclength], record2[2][:bclength]
        else:
            readid = record1[0]
            vdj, vdjqual = record1[1][bclength:], record1[2][bclength:]
            bc, bcQ = record1[1][0:bclength], record1[2][0:bclength]
        if "N" in bc and not allowNs:
            counts["dcrfilter_barcodeN"] += 1
        counts["read_count"] += 1
        recom, frame = None, None
        if orientation in ("reverse", "both"):
            recom, frame = orc.dcr(revcomp(vdj)), "reverse"
        if orientation == "forward" or (orientation == "both" and not recom):
            recom, frame = orc.dcr(vdj), "forward"
        if recom:
            counts["vj_count"] += 1
            if frame == "reverse":
                tcrseq = revcomp(vdj)[recom[5]:recom[6]]
                tcrQ = vdjqual[::-1][recom[5]:recom[6]]
            else:
                tcrseq = vdj[recom[5]:recom[6]]
                tcrQ = vdjqual[recom[5]:recom[6]]
            rows.append([str(recom[0]), str(recom[1]), str(recom[2]), str(recom[3]), recom[4], readid, tcrseq, tcrQ,
                         bc, bcQ])
    counts.update(orc.counts_dict())
    return rows, counts


def n12_text(rows):
    """write_out_intermediate's line format (io.py:494-496)."""
    return "".join(", ".join(map(str, line)) + "\n" for line in rows)
