"""The translate stage (decombinator_b200/translate.py, host code) against the reference's golden AIRR .tsv files:
reference tests/test_pipeline.py:39-60 compares dcr_TINY_1_{alpha,beta}.tsv byte for byte after pipeline.run, and
tests/test_subparsers.py:78-100,154 does the same for `decombinator translate` on the golden .freq.  CPU-only: this
stage has no kernel; the full pipeline (GPU) is checked in test_decombinator_api.py."""
import os
import shutil

import pytest

from decombinator_b200 import io, translate


def _args(infile, chain, outdir, command, **kw):
    a = io.create_args_dict(infile=str(infile), chain=chain, bc_read="R2", dontgzip=True, outpath=str(outdir) + os.sep,
                            tagfastadir="Decombinator-Tags-FASTAs", command=command)
    a.update(kw)
    return a


@pytest.mark.parametrize("chain,name", [("a", "alpha"), ("b", "beta")])
def test_translate_command_reproduces_the_golden_tsv(golden_dir, tmp_path, chain, name):
    """`decombinator translate -in dcr_TINY_1_<chain>.freq`: lines are split at ',' (fields keep their leading space)."""
    freq = tmp_path / ("dcr_TINY_1_%s.freq" % name)
    shutil.copy(os.path.join(golden_dir, freq.name), freq)
    args = _args(freq, chain, tmp_path, "translate")
    df = translate.cdr3translator(args)
    io.write_out_translated(df, args)
    want = open(os.path.join(golden_dir, "dcr_TINY_1_%s.tsv" % name), "rb").read()
    assert (tmp_path / ("dcr_TINY_1_%s.tsv" % name)).read_bytes() == want
    logs = os.listdir(tmp_path / "Logs")
    assert len(logs) == 1 and logs[0].endswith("_dcr_dcr_TINY_1_%s_%s_CDR3_Translation_Summary.csv" % (name, name))
    text = (tmp_path / "Logs" / logs[0]).read_text()
    assert "NumberUniqueDCRsInput,%d" % len(df) in text and "P_V-F,0" in text


@pytest.mark.parametrize("chain,name", [("a", "alpha"), ("b", "beta")])
def test_translate_in_memory_rows_reproduce_the_golden_tsv(golden_dir, tmp_path, chain, name):
    """The pipeline hand-over: rows as collapsinator returns them ([v, j, vdel, jdel, insert, count, cluster size])."""
    rows = []
    for line in open(os.path.join(golden_dir, "dcr_TINY_1_%s.freq" % name)):
        f = line.rstrip("\n").split(", ")
        rows.append(f[:5] + [int(f[5]), int(f[6])])
    args = _args(os.path.join(golden_dir, "TINY_1.fq"), chain, tmp_path, None, suppresssummary=True)
    df = translate.cdr3translator(args, data=rows)
    io.write_out_translated(df, args)
    want = open(os.path.join(golden_dir, "dcr_TINY_1_%s.tsv" % name), "rb").read()
    assert (tmp_path / ("dcr_TINY_1_%s.tsv" % name)).read_bytes() == want
    # gzip output and the non-productive filter
    args2 = _args(os.path.join(golden_dir, "TINY_1.fq"), chain, tmp_path, None, suppresssummary=True, dontgzip=False,
                  nonproductivefilter=True)
    df2 = translate.cdr3translator(args2, data=rows)
    assert len(df2) == int((df["productive"] == "T").sum()) and set(df2["productive"]) <= {"T"}
    io.write_out_translated(df2, args2)
    assert (tmp_path / ("dcr_TINY_1_%s.tsv.gz" % name)).exists()


def test_get_cdr3_contract_and_gene_tables():
    args = {"species": "human", "tags": "extended", "chain": "b", "tagfastadir": None, "command": None}
    info = translate.import_gene_information(args)
    assert len(info) == 12 and info[2][0].startswith("TRBV") and info[6][0] < 0
    out = translate.get_cdr3(["15", "10", "4", "1", "CTACCCCCGCGGAGAC"], translate.out_headers, args)
    assert (out["v_call"], out["j_call"], out["junction_aa"], out["productive"]) == ("TRBV20-1", "TRBJ2-5", "CSATTPAETQETQYF", "T")
    assert out["decombinator_id"] == "15, 10, 4, 1, CTACCCCCGCGGAGAC" and out["cdr1_aa"] == "DFQATT"
    # one base more: out of frame, nothing else filled in
    bad = translate.get_cdr3(["15", "10", "4", "1", "CTACCCCCGCGGAGACA"], translate.out_headers, args)
    assert bad["productive"] == "F" and bad["vj_in_frame"] == "F" and bad["junction_aa"] == "" and bad["cdr1_aa"] == ""
    # mouse and gamma/delta: no extended set, no CDR1/2 files
    margs = {"species": "mouse", "tags": "extended", "chain": "a", "tagfastadir": None, "command": None}
    minfo = translate.import_gene_information(margs)
    assert margs["tags"] == "original" and set(minfo[10]) == {""}
    with pytest.raises(SystemExit):
        translate.cdr3translator({"chain": "x", "infile": "f.freq", "species": "human", "tags": "extended", "command": "translate"})


def test_translate_nt_follows_biopython_rules():
    t = translate.translate_nt
    assert t("ATGGCCTAA") == "MA*" and t("atggcctaaG") == "MA*"          # case-insensitive, partial codon dropped
    assert t("GCN") == "A" and t("TAR") == "*" and t("TAN") == "X" and t("NNN") == "X"
    assert t("RAY") == "B" and t("SAA") == "Z" and t("MTT") == "J" and t("AUG") == "M"
    with pytest.raises(ValueError):
        t("AC!")
