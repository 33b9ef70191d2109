#!/usr/bin/env python
"""Time the UNMODIFIED reference's decombinator() -- TEST / MEASUREMENT INFRASTRUCTURE.

Runs only in the build container (needs /root/reference; the GPU box does not have it, so bench.py cannot run this
and quotes the committed result instead).  The reference source is imported as it is through oracle/refenv.py; its five
absent wheels are the stand-ins of oracle/standins/, with the tag search behind `acora` done by a compiled
Aho-Corasick automaton (oracle/standins/_ac.c) so that it costs what a compiled automaton costs.

    python oracle/time_reference.py [--reads 100000]  ->  profiles/r02_reference_cpu_timing.json

Input: the first `reads` read pairs of the configs[1] stream (synthetic 250-nt reads, human beta, extended tags, seed
20260002) written as FASTQ text; timed region: wall time of decombinator(args) with dontcount / suppresssummary /
dontcheck, one process, one core; tag loading and interpreter start excluded.
"""
import argparse
import json
import os
import platform
import subprocess
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return platform.processor()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=100_000)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_reference_cpu_timing.json"))
    args = ap.parse_args()
    os.makedirs(os.path.join(HERE, "_ref"), exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", os.path.join(HERE, "standins", "_ac.c"), "-o",
                           os.path.join(HERE, "_ref", "libac.so")])
    import refenv
    from decombinator_b200 import _lib, tags
    info = tags.load("human", "extended", "b")
    n, L = args.reads, 250
    syn = _lib.Synth([(info.v_regions, info.j_regions)], 20260002, L, 62, 0.0, 0.0, 0.0)
    r1, r2 = syn.reads(0, n, want_r2=True)
    tmp = tempfile.mkdtemp(prefix="dcbref")
    f1 = os.path.join(tmp, "syn_1.fq")
    for path, arr, ln in ((f1, r1, L), (f1.replace("1.f", "2.f"), r2, 62)):
        a = arr.reshape(n, ln)
        with open(path, "wb") as fh:
            fh.write(b"".join(b"@SYN:%d 1:N:0\n%s\n+\n%s\n" % (i, a[i].tobytes(), b"I" * ln) for i in range(n)))
    out = {"what": "unmodified reference decombinator() (innate2adaptive/decombinator @ 71f78d1) + stand-ins for its five absent wheels",
           "workload": "first %d read pairs of the configs[1] stream (250-nt reads, human beta, extended tags, -br R2 -bl 42), FASTQ text" % n,
           "cores": 1, "cpu": cpu_model(), "where": "build container (the GPU box has no /root/reference)", "runs": {}}
    for label, env in (("compiled_aho_corasick", None), ("pure_python_acora_standin", "1")):
        if env:
            os.environ["DCB_ACORA_PURE_PYTHON"] = env
        else:
            os.environ.pop("DCB_ACORA_PURE_PYTHON", None)
        ref = refenv.load()                  # puts the stand-ins on sys.path
        import acora
        acora._LIB = None                    # re-decide compiled / pure Python for this run
        d = ref["decombine"]
        a = ref["io"].create_args_dict(infile=f1, chain="b", bc_read="R2", dontcount=True, suppresssummary=True, dontcheck=True,
                                       tagfastadir=refenv.REF_TAGDIR, outpath=tmp + os.sep)
        t0 = time.perf_counter()
        rows = d.decombinator(a)
        dt = time.perf_counter() - t0
        out["runs"][label] = {"seconds": round(dt, 3), "reads_per_s": round(n / dt, 1), "rows": len(rows)}
        print(label, out["runs"][label])
    out["value"] = out["runs"]["compiled_aho_corasick"]["reads_per_s"]
    out["unit"] = "reads/s on one core"
    with open(args.out, "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
