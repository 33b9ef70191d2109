"""Checks of decombinator_b200.collapse shared by the CPU suite (distances supplied by the oracle, to test the HOST
logic) and the GPU suite (distances from the CUDA kernels through the C ABI)."""
import collections as coll
import os

from decombinator_b200 import collapse, io as dio


def check_case(case):
    """One recorded run of the reference's collapse stage: every intermediate and the .freq rows must match."""
    args = dict(case["args"])
    rows = [list(r) for r in case["rows"]]
    collapse.counts = coll.Counter()
    qp = [args["minbcQ"], args["bcQbelowmin"], args["avgQthreshold"]]
    frac = args["percentlevdist"] / 100
    groups = collapse.read_in_data([list(r) for r in rows], args, qp, frac, True, open)
    if "counts_read_in_data" in case:                 # the reference's counters after read_in_data (barcode location, filters)
        got = {k: v for k, v in sorted(collapse.counts.items()) if not k.startswith("time_") and isinstance(v, int)}
        assert got == case["counts_read_in_data"]
    assert list(groups.keys()) == case["group_keys"]
    assert [len(v) for v in groups.values()] == case["group_sizes"]
    _, blist, umi_proto = collapse.create_clustering_objs(groups)
    assert [u for u, _ in umi_proto] == case["umis"]
    matches = collapse.make_merge_groups(umi_proto, args["bcthreshold"], True)
    assert [[int(i), int(j)] for i, j in zip(matches.row, matches.col)] == case["pairs"]
    clusters = collapse.make_clusters(matches, blist, frac)
    assert list(clusters.keys()) == case["cluster_keys"]
    assert [len(v) for v in clusters.values()] == case["cluster_sizes"]
    out = collapse.collapsinator(dict(args), data=[list(r) for r in rows])
    assert out == case["freq"]


def check_tiny_freq(golden_dir, tmp_path, chain, name):
    """`decombinator collapse -in dcr_TINY_1_<chain>.n12 -ol M13 -c <c> -dz` reproduces the reference's golden .freq
    byte for byte (reference tests/test_subparsers.py:51-74, 128-149)."""
    n12 = os.path.join(golden_dir, "dcr_TINY_1_%s.n12" % name)
    args = dio.create_parser().parse_args(["collapse", "-in", n12, "-ol", "M13", "-c", chain, "-op", str(tmp_path) + os.sep, "-dz"])
    args = vars(args)
    data = collapse.collapsinator(args)
    dio.write_out_intermediate(data, args, ".freq")
    got = open(os.path.join(str(tmp_path), "dcr_TINY_1_%s.freq" % name)).read()
    want = open(os.path.join(golden_dir, "dcr_TINY_1_%s.freq" % name)).read()
    assert got == want


def check_reference_unit_answers():
    """Known answers of the reference's own tests/test_collapse.py (:10-52, :198-314)."""
    import pytest
    with pytest.raises(ValueError):
        collapse.cluster_UMIs(coll.defaultdict(list), {}, 0, 0, False)
    groups = {"AAAA|0|AAAA": ["AAAA"], "GGGG|0|GGGG": ["GGGG"], "AAAG|0|AAAG": ["AAAG"], "AAAA|1|GGGG": ["GGGG"]}
    clusters = collapse.cluster_UMIs(groups, {"writeclusters": False}, 2, 0.25, True)
    assert clusters == {"AAAA|0|AAAA": ["AAAA", "AAAG"], "GGGG|0|GGGG": ["GGGG"], "AAAA|1|GGGG": ["GGGG"]}
    assert list(clusters) == ["AAAA|0|AAAA", "GGGG|0|GGGG", "AAAA|1|GGGG"]

    pipe_args = {"command": "pipeline", "lenthreshold": 130, "minbcQ": 20, "bcQbelowmin": 1, "avgQthreshold": 30,
                 "oligo": "M13", "sampling_analysis": False}
    collapse.counts = coll.Counter()
    with pytest.raises(ValueError):
        collapse.read_in_data([], pipe_args, None, None, None, None)
    pre, post = "ATCCTGAAGACAGCAGCTTCTACATCTGCAGTGCTAGAG", "CAGCCCCAGCATTTTGGTGATGGGACTCGACTC"
    q = "IIIIIIIIIIIIIII-II-IIIIIIIIIIIIIIIIIIIIIIIIIIIII-IIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII"
    rid, bc, bq, tail = "LH00409:259:22JJCFLT4:8:1149:38410:17375", "GTCGTGACTGGGAAAACCCTGGCACCCGGTCGTGATCTGACT", "I" * 42, "GACACAACTCTCCCCAGAGAAGGTGGTGTGA"
    rows = [["15", "4", "1", "7", ins, rid, pre + ins + post, q, bc, bq, tail]
            for ins in ("CCCCCAGGGGGCTC", "CCCCCAGGGGGCTG", "AAAAAAAAAAAAAA")]
    qp = [20, 1, 30]
    assert collapse.read_in_data(rows, pipe_args, qp, lev_threshold_fraction=0.1, dont_count=False, opener=open) == {}
    got = collapse.read_in_data(rows, pipe_args, qp, lev_threshold_fraction=1, dont_count=False, opener=open)
    key = "CACCCGCTGACT|0|" + pre + "CCCCCAGGGGGCTC" + post
    assert list(got) == [key]
    assert got[key] == ["%s|%s|%s|%s" % (str(r[:5]), r[6], q, rid) for r in rows]
