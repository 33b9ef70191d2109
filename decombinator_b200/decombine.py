"""Drop-in for the reference's ``decombinator.decombine`` module, with the hot loop on the GPU.

Same entry points and contracts as /root/reference/src/decombinator/decombine.py:

* ``decombinator(inputargs) -> list[list[str]]``  (reference decombine.py:881): same argument dict,
  same rows (``[v, j, vdel, jdel, insert, readid, tcrseq, tcrQ, bc, bcQ]``), same summary CSV, same errors;
* ``import_tcr_info(inputargs)`` (decombine.py:593) and ``dcr(read, inputargs)`` (decombine.py:534): the
  per-read contract, here one-read batches through the same kernels;
* ``counts``: the module-level ``collections.Counter`` the reference exposes.

What changed is the body of the read loop: all V(D)J reads are 2-bit packed (``dcb_pack_reads``), the
batch is analysed by the CUDA kernels behind ``dcb_decombine_batch``, and only the positions come back;
the output strings are sliced on the host from the text it already holds.  There is no CPU
implementation of the matching here: without libdcb.so and a GPU these functions raise.
"""
import collections as coll
import os
from time import strftime, time

import numpy as np

from . import __version__, _lib, fastq, tags
from .fastq import opener_check, readfq  # noqa: F401  (re-exported like the reference module)

counts = coll.Counter()
chainnams = tags.CHAINNAMS
chain = None
_info = None      # TcrInfo of the last import_tcr_info()
_ctx_cache = {}   # (both_frames, allowNs, lenthreshold) -> _lib.Context

_COMP = bytes.maketrans(b"ACGTUMRWSYKVHDBNacgtumrwsykvhdbn", b"TGCAAKYWSRMBDHVNtgcaakywsrmbdhvn")


def revcomp(read):
    """Reverse complement with Bio.Seq's table (decombine.py:182-184); used for the output strings."""
    return read.encode("latin-1").translate(_COMP)[::-1].decode("latin-1")


def sort_permissions(fl):
    """decombine.py:869-873"""
    if oct(os.stat(fl).st_mode)[4:] != "666":
        os.chmod(fl, 0o666)


def import_tcr_info(inputargs):
    """Gather the TCR chain information and build the device tag tables (decombine.py:593-746)."""
    global counts, chain, _info, _ctx_cache
    counts = coll.Counter()
    _info = tags.TcrInfo(inputargs)
    chain = _info.chain
    if _info.chain_detected:
        counts["chain_detected"] = 1
    _ctx_cache = {}
    g = globals()
    for name in ("v_seqs", "j_seqs", "half1_v_seqs", "half2_v_seqs", "half1_j_seqs", "half2_j_seqs", "jump_to_end_v",
                 "jump_to_start_j", "v_regions", "j_regions", "v_half_split", "j_half_split"):
        g[name] = getattr(_info, name)
    return _info


_ctx_store = {}   # (tag set, device, both_frames, allowNs, lenthreshold) -> _lib.Context, kept across runs of one process


def _context(inputargs, both_frames):
    key = (bool(both_frames), bool(inputargs["allowNs"]), int(inputargs["lenthreshold"]))
    ctx = _ctx_cache.get(key)
    if ctx is None:
        device = int(os.environ.get("LOCAL_RANK", inputargs.get("device", 0) or 0))
        # the device tables depend on nothing but the tag set: a second run with the same set (the pipeline's next stage,
        # the next file of a batch job) reuses the context with its device buffers and page-locked staging
        tagkey = (tuple(_info.v_seqs), tuple(_info.j_seqs), tuple(_info.jump_to_end_v), tuple(_info.jump_to_start_j),
                  _info.v_half_split, _info.j_half_split, hash(tuple(_info.v_regions)), hash(tuple(_info.j_regions)), device) + key
        ctx = _ctx_store.get(tagkey)
        if ctx is None:
            if len(_ctx_store) >= 4:          # a handful of chains at most stay resident
                _ctx_store.pop(next(iter(_ctx_store))).close()
            vt, jt = _info.tables()
            ctx = _lib.Context(vt, jt, device=device, both_frames=key[0], allow_ns=key[1], lenthreshold=key[2])
            _ctx_store[tagkey] = ctx
        _ctx_cache[key] = ctx
    return ctx


def _orientation_plan(orientation):
    """-> (pack the reverse complement?, retry the other frame?) (decombine.py:999-1010)"""
    if orientation == "reverse":
        return True, False
    if orientation == "forward":
        return False, False
    if orientation == "both":
        return True, True
    raise UnboundLocalError("orientation must be forward, reverse or both")  # reference: unbound `recom`


def _add_device_counters(dev_counts):
    for name, val in zip(_lib.counter_names(), dev_counts):
        if val:
            counts[name] += int(val)


def dcr(read, inputargs):
    """Check one read (in the given frame) for a rearranged TCR (decombine.py:534-585).

    Returns ``[v, j, vdel, jdel, insert, v_seq_start, j_seq_end]`` or None and bumps ``counts`` like the
    reference.  ``import_tcr_info`` must have been called first."""
    ctx = _context(inputargs, False)
    packed = _lib.pack_strings([read], revcomp=False)
    res, dev_counts = ctx.decombine(packed)
    packed.free()
    _add_device_counters(dev_counts)
    r = res[0]
    if not r["status"]:
        return None
    return [int(r["v"]), int(r["j"]), int(r["vdel"]), int(r["jdel"]), read[int(r["ins_start"]):int(r["ins_end"])],
            int(r["v_seq_start"]), int(r["j_seq_end"])]


class RowsText:
    """The rows of a run as the text write_out_intermediate would produce from them (", " between fields, one row per
    line): what `decombinator decombine` hands to the writer instead of a list of lists when nothing else reads them."""

    __slots__ = ("text", "n_rows")

    def __init__(self, text: bytes, n_rows: int):
        self.text, self.n_rows = text, n_rows

    def __len__(self):
        return self.n_rows

    def rows(self):
        return [line.split(", ") for line in bytes(self.text).decode("ascii").split("\n")[:-1]]


class RowsColumns:
    """The rows of a run kept COLUMNAR: the result records of the kernels plus the FASTQ text columns they index.  This is
    what `pipeline` hands from decombine to collapse instead of a list of ten Python strings per read: collapse locates the
    barcodes straight from the text columns on the device and only builds strings for the rows that survive its filters.

    Anything that wants the reference's ``list[list[str]]`` still gets it: len(), iteration, indexing and ``rows()``
    materialise the rows (through dcb_format_rows), ``text`` is the .n12 text for the writer."""

    def __init__(self, res, hits, columns, pack_rc):
        self.res, self.hits, self.columns, self.pack_rc = res, hits, columns, pack_rc
        self._rows = self._text = None

    def __len__(self):
        return int(len(self.hits))

    @property
    def text(self):
        if self._text is None:
            self._text = _lib.format_rows(self.res, self.pack_rc, self.columns, ", ")[0]
        return self._text

    def barcode_columns(self):
        """(text, offsets, lengths) of the barcode regions and of their qualities, one entry per hit (dcb_barcodes' input)."""
        _, _, _, bc, bcq, _ = self.columns
        hits = self.hits
        return (np.frombuffer(bc.buf, dtype=np.uint8), np.asarray(bc.off)[hits], np.asarray(bc.len)[hits],
                np.frombuffer(bcq.buf, dtype=np.uint8), np.asarray(bcq.off)[hits], np.asarray(bcq.len)[hits])

    def subset_rows(self, keep):
        """list[list[str]] of the hits selected by the boolean mask `keep` (over the hits), in order."""
        res = self.res
        if not bool(np.all(keep)):
            res = res.copy()
            res["status"][self.hits[~np.asarray(keep, dtype=bool)]] = 0
        blob, _ = _lib.format_rows(res, self.pack_rc, self.columns, "\x1f")
        return [line.split("\x1f") for line in blob.tobytes().decode("ascii").split("\n")[:-1]]

    def collapse_lines(self, keep):
        """For the hits selected by `keep`, in order: (tcrseq, str(row[:5]), "|".join((str(row[:5]), tcrseq, tcrQ, id))) as
        three parallel lists of str -- what collapse files a row under, built by dcb_format_collapse_rows."""
        res = self.res
        if not bool(np.all(keep)):
            res = res.copy()
            res["status"][self.hits[~np.asarray(keep, dtype=bool)]] = 0
        blob, n = _lib.format_collapse_rows(res, self.pack_rc, self.columns)
        lines = blob.tobytes().decode("ascii").split("\n")
        return lines[0:3 * n:3], lines[1:3 * n:3], lines[2:3 * n:3]

    def rows(self):
        if self._rows is None:
            self._rows = self.subset_rows(np.ones(len(self.hits), dtype=bool))
        return self._rows

    def __iter__(self):
        return iter(self.rows())

    def __getitem__(self, i):
        return self.rows()[i]


def decombine_batch(batch: fastq.ReadBatch, inputargs):
    """GPU pass over every V(D)J read of the batch -> dcb_result array (one record per read)."""
    pack_rc, both = _orientation_plan(inputargs["orientation"])
    ctx = _context(inputargs, both)
    off = np.ascontiguousarray(batch.off, dtype=np.uint64)
    lens = np.asarray(batch.len)
    if len(lens) and int(lens.max()) > MAX_READ_LEN:
        raise ValueError("read %d has %d bases; this build analyses reads of up to %d bases (DCB_MAX_READ_LEN)"
                         % (int(lens.argmax()), int(lens.max()), MAX_READ_LEN))
    # A read slot is as wide as the longest read of a batch, and the fastest kernels take slots of up to 320 bases: a few
    # long reads in a file of short ones are analysed as a batch of their own instead of widening every slot.
    if len(lens) and int(lens.max()) > FAST_READ_LEN and float((lens <= FAST_READ_LEN).mean()) >= 0.5:
        res = np.zeros(len(lens), dtype=_lib.RESULT_DTYPE)
        for part in (np.nonzero(lens <= FAST_READ_LEN)[0], np.nonzero(lens > FAST_READ_LEN)[0]):
            sub = fastq.ReadBatch()
            sub.buf, sub.off, sub.len = batch.buf, off[part], np.ascontiguousarray(lens[part], dtype=np.uint32)
            res[part] = _decombine_columns(ctx, sub.buf, sub.off, sub.len, pack_rc, True)
        return res
    # the `decombine` command formats the rows and lets go of the records: it may read them where the device wrote them
    return _decombine_columns(ctx, batch.buf, off, batch.len, pack_rc, bool(inputargs.get("rows_as_text")))


MAX_READ_LEN = 4096      # DCB_MAX_READ_LEN of csrc/dcb_tables.h
FAST_READ_LEN = 320      # widest slot of the flat exact-tag kernel and the half-tag kernel (20 words)


def _decombine_columns(ctx, buf, off, length, pack_rc, transient=False):
    """(offset, length) columns into a text buffer -> dcb_result array; counters added to `counts`.

    The records arrive in a page-locked buffer of the context (copies into pageable memory would stall the chunk loop);
    transient: the caller is done with them before the context is used again, so the view itself is returned."""
    try:
        t1 = time()
        # the text goes to the GPU as it is, or 2-bit packed by the host threads (dcb_decombine_ascii shares the chunks; it
        # refuses offsets that run backwards inside a chunk)
        res, dev_counts = ctx.decombine_ascii(buf, off, length, pack_rc, pinned=True)
        if not transient:
            res = res.copy()
        if os.environ.get("DCB_TIMING"):
            print("\t[timing]   dcb_decombine_ascii %.3f s (chunks packed by host threads / device: %s)" % (time() - t1, ctx.last_pack_shares()))
    except _lib.DcbError:
        # reads out of text order, or more non-ACGT symbols than the device-side list holds: pack on the host threads
        packed = _lib.pack_arrays(buf, off, length, revcomp=pack_rc)
        res, dev_counts = ctx.decombine(packed)
        packed.free()
    _add_device_counters(dev_counts)
    return res


def _summary_path(inputargs, logpath, date, samplenam, number=None):
    name = logpath + date + "_"
    if inputargs["chain"]:
        name += chainnams[chain] + "_"
    name += samplenam + "_Decombinator_Summary" + ("" if number is None else str(number)) + ".csv"
    return name


def _open_summary(inputargs, summaryname, logpath, date, samplenam):
    """First free name among Summary.csv, Summary2.csv, ... (decombine.py:1082-1095)."""
    if not os.path.exists(summaryname):
        return summaryname, open(summaryname, "wt")
    for i in range(2, 10000):
        cand = _summary_path(inputargs, logpath, date, samplenam, i)
        if not os.path.exists(cand):
            return cand, open(cand, "wt")
    raise RuntimeError("no free summary file name")


def sample_name(inputargs):
    samplenam = str(inputargs["infile"].split(".")[0])
    if os.sep in samplenam:
        samplenam = samplenam.split(os.sep)[-1]
    return samplenam


def summary_location(inputargs, samplenam, date):
    """Logs/ directory (created) and the first-choice summary name (decombine.py:897-910)."""
    logpath = inputargs["outpath"] + f"Logs{os.sep}"
    if not os.path.exists(logpath):
        os.makedirs(logpath, exist_ok=True)
    return logpath, _summary_path(inputargs, logpath, date, samplenam)


def check_fastq(inputargs, opener, summaryname, logpath, date, samplenam):
    """The input check at the top of decombinator() (decombine.py:131-179, 912-918): fewer than four lines -> stub summary and
    ValueError.  Multi-GPU runs call it on rank 0 and broadcast the outcome (parallel.decombinator_shard)."""
    if summaryname is None:
        raise UnboundLocalError("cannot access local variable 'summaryname'")  # as the reference (-s without -dk)
    if not fastq.fastq_sanity(inputargs["infile"], opener):
        # stub summary for an empty input (decombine.py:131-161)
        inout_name = "_".join(f"{samplenam}".split("_")[:-1]) + f"_{chainnams[chain]}"
        summstr = "OutputFile," + inout_name + "\nNumberReadsInput," + "0"
        summaryname, fh = _open_summary(inputargs, summaryname, logpath, date, samplenam)
        print(summstr, file=fh)
        fh.close()
        sort_permissions(summaryname)
        raise ValueError(
            "There are fewer than four lines in this file, and thus it is not a valid FASTQ file. Please check input and try again."
        )


def write_summary(inputargs, summaryname, logpath, date, samplenam, timetaken):
    """The Decombinator summary CSV from the module's `counts` (decombine.py:1081-1200).  Multi-GPU runs call this on rank 0
    AFTER the per-rank counters have been summed (parallel.decombinator_shard), so the file reports the whole job."""
    summaryname, summaryfile = _open_summary(inputargs, summaryname, logpath, date, samplenam)
    inout_name = "_".join(f"{samplenam}".split("_")[:-1]) + f"_{chainnams[chain]}"
    summstr = ("Property,Value\nDirectory," + os.getcwd() + "\nInputFile," + inout_name + "\nOutputFile," + inout_name
               + "\nDateFinished," + date + "\nTimeFinished," + strftime("%H:%M:%S") + "\nTimeTaken(Seconds),"
               + str(round(timetaken, 2)) + "\n\nInputArguments:,\n")
    for s in ["species", "chain", "extension", "tags", "dontgzip", "allowNs", "orientation", "lenthreshold", "bc_read",
              "bclength"]:
        summstr = summstr + s + "," + str(inputargs[s]) + "\n"
    counts["pc_decombined"] = counts["vj_count"] / counts["read_count"]
    sections = [
        ("\nNumberReadsInput,", "read_count"), ("\nNumberReadsDecombined,", "vj_count"),
    ]
    for label, key in sections:
        summstr += label + str(counts[key])
    summstr += "\nPercentReadsDecombined," + str(round(counts["pc_decombined"], 3))
    summstr += "\n\nReadsAssignedUsingHalfTags:,"
    for label, key in (("V1error", "verr1"), ("V2error", "verr2"), ("J1error", "jerr1"), ("J2error", "jerr2")):
        summstr += "\n" + label + "," + str(counts[key])
    summstr += "\n\nReadsFilteredOut:,"
    for label, key in (("AmbiguousBaseCall(DCR)", "dcrfilter_intertagN"),
                       ("AmbiguousBaseCall(Barcode)", "dcrfilter_barcodeN"),
                       ("OverlongInterTagSeq", "dcrfilter_toolong_intertag"),
                       ("ImpossibleDeletions", "dcrfilter_imposs_deletion"),
                       ("OverlappingTagBoundaries", "dcrfilter_tag_overlap")):
        summstr += "\n" + label + "," + str(counts[key])
    summstr += "\n\nReadsFailedAssignment:,"
    for label, key in (("MultipleVtagMatches", "multiple_v_matches"), ("VTagAtEndRead", "v_del_failed_tag_at_end"),
                       ("VDeletionsUndetermined", "v_del_failed"), ("FoundV1HalfTagNotV2", "foundv1notv2"),
                       ("FoundV2HalfTagNotV1", "foundv2notv1"), ("NoVDetected", "no_vtags_found"),
                       ("MultipleJTagMatches", "multiple_j_matches"), ("JDeletionsUndermined", "j_del_failed"),
                       ("FoundJ1HalfTagNotJ2", "foundj1notj2"), ("FoundJ2HalfTagNotJ1", "foundj2notj1"),
                       ("NoJDetected", "no_j_assigned")):
        summstr += "\n" + label + "," + str(counts[key])
    print(summstr, file=summaryfile)
    summaryfile.close()
    sort_permissions(summaryname)


def decombinator(inputargs: dict) -> list:
    """Function wrapper for decombinator (decombine.py:881-1202)."""
    print("Running Decombinator version", __version__)
    opener = opener_check(inputargs)
    import_tcr_info(inputargs)

    samplenam = sample_name(inputargs)

    summaryname = logpath = None
    date = strftime("%Y_%m_%d")
    if inputargs["suppresssummary"] == False:  # noqa: E712
        logpath, summaryname = summary_location(inputargs, samplenam, date)

    if inputargs["dontcheck"] == False:  # noqa: E712
        check_fastq(inputargs, opener, summaryname, logpath, date, samplenam)

    counts["start_time"] = time()
    print("Decombining FASTQ data...")

    outdata = []
    _t = [time()]

    def _lap(what):   # DCB_TIMING=1: where the wall time of the stage goes
        if os.environ.get("DCB_TIMING"):
            _t.append(time())
            print("\t[timing] %-28s %.3f s" % (what, _t[-1] - _t[-2]))
    if inputargs["nobarcoding"] == False:  # noqa: E712
        batch = fastq.load_pairs(inputargs, opener)
        _lap("FASTQ text -> record index")
        if inputargs.get("shard"):   # multi-GPU run (parallel.py): this rank analyses one contiguous shard of the reads
            from .parallel import shard_bounds
            batch = batch.shard(*shard_bounds(len(batch), *inputargs["shard"]))
        n = len(batch)
        bc_n = None
        if inputargs["allowNs"] == False:  # noqa: E712
            # barcodes with an N are counted beside the GPU pass (another file's text, the native scan releases the GIL)
            import threading
            box = []
            bc_n = threading.Thread(target=lambda: box.append(fastq.count_containing(batch.bc, "N")))
            bc_n.start()
        counts["read_count"] += n
        if inputargs["dontcount"] == False:  # noqa: E712
            for k in range(100000, n + 1, 100000):
                print("\t read", k)
        try:
            res = decombine_batch(batch, inputargs) if n else np.zeros(0, dtype=_lib.RESULT_DTYPE)
        finally:
            if bc_n is not None:
                bc_n.join()
        if bc_n is not None:
            counts["dcrfilter_barcodeN"] += box[0] if box else fastq.count_containing(batch.bc, "N")
            if counts["dcrfilter_barcodeN"] == 0:
                del counts["dcrfilter_barcodeN"]
        _lap("text -> GPU -> records (+ barcode N count)")
        pack_rc, _ = _orientation_plan(inputargs["orientation"])
        text_only = bool(inputargs.get("rows_as_text")) and isinstance(batch.vdj, fastq.TextColumn) and not inputargs.get("rows_as_columns")
        # (the `decombine` command needs the number of hits, not their indices: the formatter walks the records itself)
        hits = range(int(np.count_nonzero(res["status"]))) if text_only else np.nonzero(res["status"])[0]
        counts["vj_count"] += int(len(hits))
        sampling = inputargs.get("sampling_analysis")
        if isinstance(batch.vdj, fastq.TextColumn) and len(hits):
            # native ingest: the rows are formatted by dcb_format_rows (all hits, multi-threaded) into one buffer whose
            # field separator cannot occur in the data, and split here at C speed
            sep = "\x1f"
            cols = (batch.ids, batch.vdj, batch.vdjqual, batch.bc, batch.bcq, batch.v_tail if sampling else None)
            if inputargs.get("rows_as_columns") and not any(c.buf.find(sep.encode()) != -1 for c in {id(c.buf): c for c in cols if c is not None}.values()):
                # `pipeline`: the next stage takes the columns themselves
                outdata = RowsColumns(res, hits, cols, pack_rc)
                hits = ()
            elif inputargs.get("rows_as_text"):
                # the `decombine` command only writes the rows: with the .n12 separator the buffer IS the file's text
                blob, nrows = _lib.format_rows(res, pack_rc, cols, ", ")
                assert nrows == len(hits)
                outdata = RowsText(blob, nrows)
                hits = ()
            elif not any(c.buf.find(sep.encode()) != -1 for c in {id(c.buf): c for c in cols if c is not None}.values()):
                blob, nrows = _lib.format_rows(res, pack_rc, cols, sep)
                outdata = [line.split(sep) for line in blob.tobytes().decode("ascii").split("\n")[:-1]]
                assert len(outdata) == nrows == len(hits)
                hits = ()
        for i in hits:
            r = res[i]
            # frame "reverse" <=> the analysed string was revcomp(vdj) (decombine.py:1015-1020)
            is_rev = pack_rc != bool(r["frame"])
            a, b = int(r["v_seq_start"]), int(r["j_seq_end"])
            vdj, vdjqual = batch.vdj[i], batch.vdjqual[i]
            oriented = revcomp(vdj) if is_rev else vdj
            tcrseq = oriented[a:b]
            tcrQ = vdjqual[::-1][a:b] if is_rev else vdjqual[a:b]
            row = [str(int(r["v"])), str(int(r["j"])), str(int(r["vdel"])), str(int(r["jdel"])),
                   oriented[int(r["ins_start"]):int(r["ins_end"])], batch.ids[i], tcrseq, tcrQ, batch.bc[i], batch.bcq[i]]
            if sampling:
                row.append(batch.v_tail[i])
            outdata.append(row)
    else:
        # the reference's read loop is nested under `nobarcoding == False` (decombine.py:950): nothing is read
        if inputargs["extension"] == "n12":
            print("Non-barcoding option selected, but default output file extension (n12) detected. "
                  "Automatically changing to 'nbc'.")

    _lap("row assembly")
    counts["end_time"] = time()
    timetaken = counts["end_time"] - counts["start_time"]

    print("Analysed", "{:,}".format(counts["read_count"]), "reads, finding", "{:,}".format(counts["vj_count"]),
          chainnams[chain], "VJ rearrangements")
    print("Reading from", inputargs["infile"] + ", writing to variable")
    print("Took", str(round(timetaken, 2)), "seconds")

    if inputargs["suppresssummary"] == False:  # noqa: E712
        write_summary(inputargs, summaryname, logpath, date, samplenam, timetaken)

    return outdata

