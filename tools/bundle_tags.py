#!/usr/bin/env python
"""Bundle a Decombinator-Tags-FASTAs directory into decombinator_b200/data/tagsets.json.gz.

The reference downloads `<species>_<tagset>_TR<chain><gene>.{tags,fasta,translate,cdrs}` from
GitHub when they are not found locally (decombine.py:187-225).  There is no network on the GPU
box, so the same DATA files (innate2adaptive/Decombinator-Tags-FASTAs @ 20efd39) are shipped
as one compressed bundle; `-tfdir` still takes precedence exactly as in the reference.

usage: python tools/bundle_tags.py /root/reference/tests/resources/Decombinator-Tags-FASTAs
"""
import gzip
import json
import os
import sys


def main(src):
    out = {}
    for fn in sorted(os.listdir(src)):
        stem, ext = os.path.splitext(fn)
        if ext not in (".tags", ".fasta", ".translate", ".cdrs"):
            continue
        with open(os.path.join(src, fn), "rt") as fh:
            out.setdefault(stem, {})[ext[1:]] = fh.read()
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "decombinator_b200", "data", "tagsets.json.gz")
    with gzip.GzipFile(dst, "wb", mtime=0) as fh:
        fh.write(json.dumps(out, sort_keys=True).encode())
    print("wrote", dst, len(out), "sets", os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main(sys.argv[1])
