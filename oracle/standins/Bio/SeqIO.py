"""Stand-in for Bio.SeqIO.parse(path, "fasta") (decombine.py:690, translate.py:183)."""
from .Seq import Seq


class SeqRecord:
    def __init__(self, id_, description, seq):
        self.id = id_
        self.name = id_
        self.description = description
        self.seq = Seq(seq)


def parse(handle, fmt):
    assert fmt == "fasta"
    close = False
    if isinstance(handle, (str, bytes)):
        handle = open(handle, "rt")
        close = True
    try:
        recs = []
        header, chunks = None, []
        for line in handle:
            line = line.rstrip("\r\n")
            if line.startswith(">"):
                if header is not None:
                    recs.append((header, "".join(chunks)))
                header, chunks = line[1:], []
            elif header is not None:
                chunks.append(line.strip())
        if header is not None:
            recs.append((header, "".join(chunks)))
    finally:
        if close:
            handle.close()
    for header, seq in recs:
        parts = header.split(None, 1)
        yield SeqRecord(parts[0] if parts else "", header, seq)
