"""Shared helpers for the parity tests."""
import numpy as np

import decombine_oracle as O


def oriented_read(read, frame_forward):
    return read if frame_forward else O.revcomp(read)


def record_to_list(read, rec, orientation):
    """dcb_result / orc_result record -> the list dcr() returns (+ frame), or None."""
    ok = rec["status"] if "status" in rec.dtype.names else rec["ok"]
    if not ok:
        return None
    if orientation == "both":
        fwd = int(rec["frame"]) == 1
    else:
        fwd = orientation == "forward"
    s = oriented_read(read, fwd)
    return [int(rec["v"]), int(rec["j"]), int(rec["vdel"]), int(rec["jdel"]), s[int(rec["ins_start"]):int(rec["ins_end"])],
            int(rec["v_seq_start"]), int(rec["j_seq_end"]), 1 if fwd else 0]


FIELDS = ("v", "j", "vdel", "jdel", "ins_start", "ins_end", "v_seq_start", "j_seq_end")


def assert_records_equal(res, ores, orientation, what=""):
    """CUDA (dcb_result) vs oracle (orc_result) arrays: bit-exact on every field of every decombined read."""
    assert len(res) == len(ores)
    st = res["status"].astype(np.int32)
    assert np.array_equal(st, ores["ok"]), "%s: status differs at %s" % (what, np.nonzero(st != ores["ok"])[0][:10])
    m = ores["ok"] == 1
    for f in FIELDS:
        a, b = res[f][m].astype(np.int64), ores[f][m].astype(np.int64)
        assert np.array_equal(a, b), "%s: field %s differs at %s" % (what, f, np.nonzero(a != b)[0][:10])
    if orientation == "both":
        assert np.array_equal(res["frame"][m].astype(np.int32), ores["frame"][m]), what + ": frame differs"


def synth_batch(info, n, L, sub=0.0, nrate=0.0, junk=0.0, seed=1, first=0, sets=None):
    from decombinator_b200 import _lib
    syn = _lib.Synth(sets or [(info.v_regions, info.j_regions)], seed, L, 0, sub, nrate, junk)
    r1, _ = syn.reads(first, n)
    off = np.arange(n, dtype=np.uint64) * L
    ln = np.full(n, L, dtype=np.uint32)
    return r1, off, ln
