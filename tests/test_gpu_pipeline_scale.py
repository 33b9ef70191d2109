"""decombine + collapse at a size the row-by-row fixtures cannot hold (BASELINE configs[3] shape): the paired FASTQ is
regenerated from its seed (oracle/synth_pairs.py) and the SHA-256 of the .n12 text and of the .freq rows must equal what
the UNMODIFIED reference produced for the same files (tests/golden/pipeline_digest.json, recorded in the build container
by oracle/make_golden_pipeline_digest.py) -- through the row hand-over, the columnar hand-over of `pipeline`, and the
`collapse` command reading the .n12 file it was given."""
import hashlib
import json
import os

import pytest

import synth_pairs
from decombinator_b200 import collapse, decombine, io

pytestmark = pytest.mark.gpu


def _digest_cases(golden_dir):
    with open(os.path.join(golden_dir, "pipeline_digest.json")) as fh:
        return json.load(fh)["cases"]


def _sha(text):
    return hashlib.sha256(text.encode()).hexdigest()


@pytest.mark.parametrize("index", [0, 1])
def test_decombine_and_collapse_equal_the_reference_digests(golden_dir, tmp_path, index):
    case = _digest_cases(golden_dir)[index]
    species, tagset, chain, n, pool, L, sub1, sub2, seed, oligo = case["spec"]
    f1, f2 = str(tmp_path / "s_1.fq"), str(tmp_path / "s_2.fq")
    synth_pairs.write_pairs(f1, f2, species, tagset, chain, n, pool, L, sub1, sub2, seed)
    base = io.create_args_dict(infile=f1, chain=chain, bc_read="R2", dontgzip=True, dontcount=True, suppresssummary=True, dontcheck=True,
                               outpath=str(tmp_path) + os.sep, species=species, tags=tagset, oligo=oligo, command="pipeline")
    # rows as the reference hands them over
    rows = decombine.decombinator(dict(base))
    assert len(rows) == case["n12_rows"]
    n12 = "".join(", ".join(r[:10]) + "\n" for r in rows)
    assert _sha(n12) == case["n12_sha256"]
    freq = collapse.collapsinator(dict(base), data=[list(r) for r in rows])
    assert [list(map(str, r)) for r in freq[:3]] == case["freq_head"]
    assert len(freq) == case["freq_rows"]
    assert _sha("".join(", ".join(map(str, r)) + "\n" for r in freq)) == case["freq_sha256"]
    # the columnar hand-over of `pipeline`
    a = dict(base)
    a["rows_as_columns"] = True
    data = decombine.decombinator(a)
    assert isinstance(data, decombine.RowsColumns) and _sha(bytes(data.text).decode()) == case["n12_sha256"]
    freq2 = collapse.collapsinator(dict(a), data=data)
    assert freq2 == freq
    # the `collapse` command on the .n12 file (columns over the file's text)
    path = str(tmp_path / "rows.n12")
    with open(path, "w") as fh:
        fh.write(n12)
    c = vars(io.create_parser().parse_args(["collapse", "-in", path, "-ol", oligo, "-c", chain, "-op", str(tmp_path) + os.sep, "-dz", "-dc"]))
    freq3 = collapse.collapsinator(c)
    assert freq3 == freq
