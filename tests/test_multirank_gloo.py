"""World-size-2 runs of the multi-GPU host logic on CPU (gloo): shard bounds, the barcode-hash all-to-all, and the
sharded collapse stage, whose result must equal the single-process run recorded from the reference.  The distances are
supplied by the oracle here (no GPU in this container); the kernels themselves are covered by the -m gpu tests."""
import gzip
import json
import os
import socket
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case_index, out_path):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    import collapse_oracle as CO
    from decombinator_b200 import collapse, parallel
    parallel.init_from_env("gloo")
    collapse._dist = CO.OracleDist()     # test only: the checker stands in for the GPU context
    with gzip.open(os.path.join(ROOT, "tests", "golden", "collapse_cases.json.gz"), "rt") as fh:
        case = json.load(fh)["cases"][case_index]
    rows = [list(r) for r in case["rows"]]
    # all-to-all self check: rank r sends "r->d" * (d + 1) to every rank d
    got = parallel.all_to_all_bytes([("%d->%d;" % (rank, d)).encode() * (d + 1) for d in range(world)])
    assert got == [("%d->%d;" % (s, rank)).encode() * (rank + 1) for s in range(world)], got
    lo, hi = parallel.shard_bounds(len(rows), rank, world)
    out = parallel.collapsinator_sharded(dict(case["args"]), data=rows[lo:hi], first_index=lo)
    assert parallel._last_exchange["form"] == "records"          # 64-byte records, not pickled rows
    # the UMI-code all-gather: every rank ends up with every rank's codes, in rank order
    import numpy as np
    codes, sizes = parallel.all_gather_codes(np.arange(3 + 2 * rank, dtype=np.uint64) + np.uint64(100 * rank))
    assert sizes == [3 + 2 * r for r in range(world)]
    assert codes.tolist() == [100 * r + k for r in range(world) for k in range(3 + 2 * r)]
    # with -wc the rows travel pickled (quality strings and read ids are needed again), same result
    if case_index == 0:
        args_wc = dict(case["args"], writeclusters=True, chain="b")
        cwd = os.getcwd()
        os.chdir(os.path.dirname(out_path))
        try:
            out_wc = parallel.collapsinator_sharded(args_wc, data=rows[lo:hi], first_index=lo)
        finally:
            os.chdir(cwd)
        assert parallel._last_exchange["form"] == "pickled rows"
        if rank == 0:
            assert out_wc == out
    if rank == 0:
        with open(out_path, "w") as fh:
            json.dump({"freq": out, "counts": {k: v for k, v in collapse.counts.items() if isinstance(v, int)}}, fh)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_in_order():
    from decombinator_b200.parallel import shard_bounds
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def test_barcode_owner_is_stable():
    from decombinator_b200.parallel import barcode_owner
    assert barcode_owner("ACGTACGTACGT", 1) == 0
    assert all(0 <= barcode_owner("ACGTACGTACGT", w) < w for w in (2, 3, 4, 8))
    assert barcode_owner("ACGTACGTACGT", 8) == barcode_owner("ACGTACGTACGT", 8)
    owners = {barcode_owner("".join("ACGT"[(i >> (2 * k)) & 3] for k in range(6)), 4) for i in range(4096)}
    assert owners == {0, 1, 2, 3}


@pytest.mark.parametrize("case_index", [0, 2])
def test_sharded_collapse_equals_single_process(tmp_path, case_index):
    world = 2
    out_path = str(tmp_path / "out.json")
    mp.spawn(_worker, args=(world, _free_port(), case_index, out_path), nprocs=world, join=True)
    with gzip.open(os.path.join(ROOT, "tests", "golden", "collapse_cases.json.gz"), "rt") as fh:
        case = json.load(fh)["cases"][case_index]
    got = json.load(open(out_path))
    assert got["freq"] == case["freq"]
    assert got["counts"]["readdata_input_dcrs"] == len(case["rows"])
    assert got["counts"]["readdata_barcode_dcretc_keys"] == len(case["group_keys"])
