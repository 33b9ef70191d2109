"""The drop-in stage function decombinator(inputargs) and the per-read dcr() of decombinator_b200 against the
reference's own golden .n12 files and against whole-file runs recorded from the unmodified reference."""
import os
import shutil

import pytest

from decombinator_b200 import decombine, io

gpu = pytest.mark.gpu


def _args(infile, chain, outdir, **kw):
    a = io.create_args_dict(infile=infile, chain=chain, bc_read="R2", dontgzip=True, dontcount=True,
                            outpath=str(outdir) + os.sep, tagfastadir="Decombinator-Tags-FASTAs", command="decombine")
    a.update(kw)
    return a


@gpu
@pytest.mark.parametrize("mode", ["native", "rows_as_text", "python_fastq"])
@pytest.mark.parametrize("chain,name", [("a", "alpha"), ("b", "beta")])
def test_tiny_golden_n12_bytes(golden_dir, tmp_path, chain, name, mode):
    """reference tests/test_pipeline.py:63-84 / test_subparsers.py:27-47: the .n12 must be byte-identical -- through
    the native ingest + row formatter (list of rows), through the text hand-over of the `decombine` command, and
    through the general (Python) parser and row assembly."""
    for f in ("TINY_1.fq", "TINY_2.fq"):
        shutil.copy(os.path.join(golden_dir, f), tmp_path / f)
    args = _args(str(tmp_path / "TINY_1.fq"), chain, tmp_path)
    if mode != "native":
        args[mode] = True
    rows = decombine.decombinator(args)
    assert isinstance(rows, decombine.RowsText) == (mode == "rows_as_text")
    io.write_out_intermediate(rows, args, ".n12")
    out = tmp_path / ("dcr_TINY_1_%s.n12" % name)
    assert out.read_bytes() == open(os.path.join(golden_dir, "dcr_TINY_1_%s.n12" % name), "rb").read()
    # summary file exists, and the one line the reference's own log test pins (test_pipeline.py:138-159)
    logs = os.listdir(tmp_path / "Logs")
    assert len(logs) == 1 and logs[0].endswith("_%s_TINY_1_Decombinator_Summary.csv" % name)
    lines = (tmp_path / "Logs" / logs[0]).read_text().split("\n")
    assert lines[8] == "InputArguments:,"
    assert "NumberReadsInput,106" in lines


@gpu
def test_recorded_reference_runs(decombinator_runs, tmp_path):
    """Rows AND the counts Counter of whole-file runs (R1/R2 barcodes, -sa, both/forward, mouse, allowNs)."""
    for ri, run in enumerate(decombinator_runs["runs"]):
        a = dict(run["args"])
        f1 = tmp_path / os.path.basename(a["infile"])
        f1.write_text(run["fastq1"])
        (tmp_path / os.path.basename(a["infile"]).replace("1.f", "2.f")).write_text(run["fastq2"])
        a["infile"] = str(f1)
        a["tagfastadir"] = "Decombinator-Tags-FASTAs"
        for mode in ("native", "python_fastq", "rows_as_text"):   # the three host paths around the same kernels
            b = dict(a)
            if mode != "native":
                b[mode] = True
            rows = decombine.decombinator(b)
            if mode == "rows_as_text" and not isinstance(rows, list):
                rows = rows.rows()
            assert rows == run["rows"], (ri, mode)
            got = {k: int(v) for k, v in decombine.counts.items() if k not in ("start_time", "end_time")}
            assert got == run["counts"], (ri, mode, got, run["counts"])
        # the summary CSV, byte for byte (the values of the four lines that depend on where / when it ran are masked in
        # the recording, decombine.py:1081-1200), and its file name behind the date prefix
        logdir = tmp_path / "Logs"
        before = set(os.listdir(logdir)) if logdir.exists() else set()
        decombine.decombinator(dict(a, suppresssummary=False, dontcheck=False, outpath=str(tmp_path) + os.sep))
        new = sorted(set(os.listdir(logdir)) - before)
        assert len(new) == 1 and new[0].split("_", 3)[3] == run["summary_name"], (ri, new)
        text = (logdir / new[0]).read_text()
        masked = "\n".join(ln.split(",")[0] + ",<run>" if ln.split(",")[0] in ("Directory", "DateFinished", "TimeFinished", "TimeTaken(Seconds)")
                           else ln for ln in text.split("\n"))
        assert masked == run["summary"], (ri, masked, run["summary"])


@gpu
def test_dcr_per_read_contract(dcr_cases):
    g = dcr_cases["groups"][1]
    args = {"infile": "x", "chain": g["chain"], "tags": g["tags"], "species": g["species"], "tagfastadir": None,
            "allowNs": g["allowNs"], "lenthreshold": g["lenthreshold"]}
    decombine.import_tcr_info(args)
    for read, exp in list(zip(g["reads"], g["results"]))[:150]:
        got = decombine.dcr(decombine.revcomp(read), args)
        assert got == (exp[:7] if exp else None)


@gpu
def test_empty_fastq_raises_and_writes_stub_summary(tmp_path):
    """reference tests/test_decombine.py:10-94"""
    f = tmp_path / "empty_merge.fq"
    f.write_text("")
    args = _args(str(f), "a", tmp_path)
    with pytest.raises(ValueError):
        decombine.decombinator(args)
    logs = os.listdir(tmp_path / "Logs")
    assert (tmp_path / "Logs" / logs[0]).read_text() == "OutputFile,empty_alpha\nNumberReadsInput,0\n"
    with pytest.raises(ValueError):
        decombine.decombinator(args)
    assert any(n.endswith("Summary2.csv") for n in os.listdir(tmp_path / "Logs"))


def test_readfq_matches_reference_parser_quirks(tmp_path):
    from decombinator_b200.fastq import readfq
    import io as _io
    text = "@r1 desc\nACGT\nAC\n+\nFFFF\nFF\n@r2\nGG\n+\nFF\n>fa\nACGT\n@r3\nAA\n+\nF"
    got = list(readfq(_io.StringIO(text)))
    import decombine_oracle as O
    assert got == list(O.readfq(_io.StringIO(text)))
    assert got[0] == ("r1", "ACGTAC", "FFFFFF") and got[2] == ("fa", "ACGT", None)


def _sharded_worker(rank, world, port, tmp, out_path):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "oracle"), os.path.join(root, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    # two ranks share the one GPU of the test box: LOCAL_RANK stays 0, the rendezvous is gloo
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK="0", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import json
    import torch.distributed as dist
    from decombinator_b200 import collapse as C, decombine as D, parallel
    parallel.init_from_env("gloo")
    # the summary is switched ON: rank 0 must write it from the counters summed over both ranks (and run the input check)
    args = _args(os.path.join(tmp, "TINY_1.fq"), "b", os.path.join(tmp, "two"), oligo="M13", command="pipeline")
    mine, first = parallel.decombinator_shard(args)     # every rank keeps its own shard for the collapse all-to-all
    args["suppresssummary"] = True                      # (the collapse summary is not under test here)
    rows = parallel.gather_rows(mine)
    freq = parallel.collapsinator_sharded(args, data=mine, first_index=first)
    if rank == 0:
        json.dump({"rows": rows, "vj": int(D.counts["vj_count"]), "freq": freq}, open(out_path, "w"))
    dist.barrier()
    dist.destroy_process_group()


@gpu
def test_two_rank_pipeline_equals_golden(golden_dir, tmp_path):
    """decombine sharded over two ranks (contiguous shards, no collective) + collapse with the barcode-hash all-to-all
    reproduces the single-process goldens: .n12 rows in input order and the .freq rows."""
    import json
    import socket
    import torch.multiprocessing as mp
    for f in ("TINY_1.fq", "TINY_2.fq"):
        shutil.copy(os.path.join(golden_dir, f), tmp_path / f)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out_path = str(tmp_path / "out.json")
    (tmp_path / "two").mkdir()
    mp.spawn(_sharded_worker, args=(2, port, str(tmp_path), out_path), nprocs=2, join=True)
    got = json.load(open(out_path))
    # the two-rank summary CSV equals the single-process one (all counters are whole-job totals), time lines aside
    (tmp_path / "one").mkdir()
    decombine.decombinator(_args(str(tmp_path / "TINY_1.fq"), "b", tmp_path / "one"))

    def summary(d):
        logs = os.listdir(tmp_path / d / "Logs")
        assert len(logs) == 1, logs
        return [ln for ln in (tmp_path / d / "Logs" / logs[0]).read_text().split("\n")
                if ln.split(",")[0] not in ("TimeFinished", "TimeTaken(Seconds)")]
    assert summary("two") == summary("one") and "NumberReadsInput,106" in summary("two")
    want_n12 = open(os.path.join(golden_dir, "dcr_TINY_1_beta.n12")).read()
    assert "".join(", ".join(map(str, r)) + "\n" for r in got["rows"]) == want_n12
    assert got["vj"] == 48
    want_freq = open(os.path.join(golden_dir, "dcr_TINY_1_beta.freq")).read()
    assert "".join(", ".join(map(str, r)) + "\n" for r in got["freq"]) == want_freq


@gpu
@pytest.mark.parametrize("chain,name", [("a", "alpha"), ("b", "beta")])
def test_cli_commands_write_the_golden_files(golden_dir, tmp_path, chain, name, monkeypatch):
    """`decombinator decombine ...` and `decombinator pipeline ...` through the command-line entry point (reference
    tests/test_cli.py:76-97, test_subparsers.py:104-125): the .n12 (plain and gzipped) and the .freq must match the
    reference's golden files byte for byte.  The decombine command takes the native ingest and hands the rows to the
    writer as text."""
    import gzip
    import sys
    from decombinator_b200 import pipeline
    for f in ("TINY_1.fq", "TINY_2.fq"):
        shutil.copy(os.path.join(golden_dir, f), tmp_path / f)
    want_n12 = open(os.path.join(golden_dir, "dcr_TINY_1_%s.n12" % name), "rb").read()
    want_freq = open(os.path.join(golden_dir, "dcr_TINY_1_%s.freq" % name), "rb").read()
    monkeypatch.chdir(tmp_path)
    out1 = tmp_path / "o1"; out2 = tmp_path / "o2"; out3 = tmp_path / "o3"
    for d in (out1, out2, out3):
        d.mkdir()
    base = ["-in", str(tmp_path / "TINY_1.fq"), "-c", chain, "-br", "R2", "-bl", "42", "-dc"]
    monkeypatch.setattr(sys, "argv", ["decombinator", "decombine"] + base + ["-dz", "-op", str(out1) + os.sep])
    pipeline.main()
    assert (out1 / ("dcr_TINY_1_%s.n12" % name)).read_bytes() == want_n12
    monkeypatch.setattr(sys, "argv", ["decombinator", "decombine"] + base + ["-op", str(out2) + os.sep])
    pipeline.main()
    assert gzip.open(out2 / ("dcr_TINY_1_%s.n12.gz" % name), "rb").read() == want_n12
    monkeypatch.setattr(sys, "argv", ["decombinator", "pipeline"] + base + ["-dz", "-ol", "M13", "-op", str(out3) + os.sep])
    pipeline.main()
    assert (out3 / ("dcr_TINY_1_%s.n12" % name)).read_bytes() == want_n12
    assert (out3 / ("dcr_TINY_1_%s.freq" % name)).read_bytes() == want_freq
    # ... and the AIRR .tsv of the translate stage (reference tests/test_pipeline.py:39-60)
    assert (out3 / ("dcr_TINY_1_%s.tsv" % name)).read_bytes() == open(os.path.join(golden_dir, "dcr_TINY_1_%s.tsv" % name), "rb").read()
    assert len(os.listdir(out3 / "Logs")) == 3       # one summary per stage
    # `decombinator translate` on the .freq just written (reference tests/test_subparsers.py:78-100)
    out4 = tmp_path / "o4"
    out4.mkdir()
    monkeypatch.setattr(sys, "argv", ["decombinator", "translate", "-in", str(out3 / ("dcr_TINY_1_%s.freq" % name)), "-c", chain, "-dz",
                                      "-op", str(out4) + os.sep])
    pipeline.main()
    assert (out4 / ("dcr_TINY_1_%s.tsv" % name)).read_bytes() == open(os.path.join(golden_dir, "dcr_TINY_1_%s.tsv" % name), "rb").read()


@gpu
def test_columnar_hand_over_equals_row_hand_over(tmp_path):
    """`pipeline` hands decombine.RowsColumns to collapse (barcodes located on the device straight from the FASTQ text
    columns, strings only for surviving rows); the result, the counters and the .n12 text must equal the hand-over of
    list[list[str]] rows -- on reads with repeated UMIs, substitutions in the barcode region (fuzzy spacers: host path),
    low-quality barcodes and N."""
    import collections as coll
    import numpy as np
    from decombinator_b200 import _lib, collapse, tags
    info = tags.load("human", "extended", "b")
    n, L = 30000, 250
    syn = _lib.Synth([(info.v_regions, info.j_regions)], 77, L, 62, 0.005, 0.0005, 0.02, umi_pool=1500, sub_rate2=0.01)
    r1, r2 = syn.reads(0, n, want_r2=True)
    rng = np.random.default_rng(3)
    q2 = np.full((n, 62), ord("I"), dtype=np.uint8)
    low = rng.random(n) < 0.05
    q2[low, 22:28] = ord("#")
    for path, arr, ln, q in ((tmp_path / "s_1.fq", r1, L, None), (tmp_path / "s_2.fq", r2, 62, q2)):
        a = arr.reshape(n, ln)
        with open(path, "wb") as fh:
            fh.write(b"".join(b"@SYN:%d 1:N:0\n%s\n+\n%s\n" % (i, a[i].tobytes(), (q[i].tobytes() if q is not None else b"I" * ln))
                              for i in range(n)))
    base = _args(str(tmp_path / "s_1.fq"), "b", tmp_path, suppresssummary=True, dontcheck=True, oligo="M13", command="pipeline")
    out, cnts, texts = {}, {}, {}
    for mode in ("rows", "columns"):
        a = dict(base)
        if mode == "columns":
            a["rows_as_columns"] = True
        data = decombine.decombinator(a)
        assert isinstance(data, decombine.RowsColumns) == (mode == "columns")
        texts[mode] = bytes(data.text) if mode == "columns" else "".join(", ".join(r) + "\n" for r in data).encode()
        out[mode] = collapse.collapsinator(dict(a), data=data)
        cnts[mode] = {k: v for k, v in collapse.counts.items() if "time" not in k}
    assert texts["rows"] == texts["columns"]
    assert out["rows"] == out["columns"] and len(out["rows"]) > 500
    assert cnts["rows"] == cnts["columns"]
    assert cnts["rows"]["getbarcode_pass_regexmatch"] > 0 and cnts["rows"]["readdata_fail_low_barcode_quality"] > 0


@gpu
def test_a_few_long_reads_do_not_widen_every_slot(tmp_path):
    """A file of 250-nt reads with a handful of 600..3000-nt ones: same rows as the oracle; the short reads still go through
    the flat exact-tag kernel (the long ones are analysed as a batch of their own)."""
    import numpy as np
    import decombine_oracle as O
    from decombinator_b200 import _lib, tags
    info = tags.load("human", "extended", "b")
    n, L = 3000, 250
    syn = _lib.Synth([(info.v_regions, info.j_regions)], 5, L, 62, 0.01, 0.001, 0.02)
    r1, r2 = syn.reads(0, n, want_r2=True)
    reads = [bytes(r1[i * L:(i + 1) * L]).decode() for i in range(n)]
    rng = np.random.default_rng(2)
    for i in (7, 500, 501, 2999):
        reads[i] = "".join("ACGT"[k] for k in rng.integers(0, 4, int(rng.integers(600, 3000)))) + reads[i]
    with open(tmp_path / "m_1.fq", "w") as f1, open(tmp_path / "m_2.fq", "w") as f2:
        for i, r in enumerate(reads):
            f1.write("@R%d\n%s\n+\n%s\n" % (i, r, "I" * len(r)))
            f2.write("@R%d\n%s\n+\n%s\n" % (i, bytes(r2[i * 62:(i + 1) * 62]).decode(), "I" * 62))
    args = _args(str(tmp_path / "m_1.fq"), "b", tmp_path, suppresssummary=True, dontcheck=True)
    rows = decombine.decombinator(args)
    orc = O.Oracle(O.TagSet("human", "extended", "b"))
    want = orc.decombine_reads(reads, "reverse")
    assert len(rows) == int(want["ok"].sum())
    assert [int(r[5][1:]) for r in rows] == np.nonzero(want["ok"])[0].tolist()
    assert decombine._ctx_cache and list(decombine._ctx_cache.values())[0].exact_kernel_name() in ("dcb_exact_kernel", "dcb_exact_kernel_spec", "dcb_exact_kernel_flat")
    with open(tmp_path / "m_1.fq", "a") as f1, open(tmp_path / "m_2.fq", "a") as f2:
        f1.write("@RX\n%s\n+\n%s\n" % ("A" * 5000, "I" * 5000))
        f2.write("@RX\n%s\n+\n%s\n" % ("A" * 62, "I" * 62))
    with pytest.raises(ValueError, match="4096"):
        decombine.decombinator(args)
