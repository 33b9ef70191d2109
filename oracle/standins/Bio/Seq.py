"""Stand-in for Bio.Seq.Seq (decombine.py:184, translate.py:308)."""

_COMP = str.maketrans(
    "ACGTUMRWSYKVHDBNacgtumrwsykvhdbn",
    "TGCAAKYWSRMBDHVNtgcaakywsrmbdhvn",
)

_BASES = "TCAG"
_AAS = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"
_CODON = {
    a + b + c: _AAS[16 * i + 4 * j + k]
    for i, a in enumerate(_BASES)
    for j, b in enumerate(_BASES)
    for k, c in enumerate(_BASES)
}


class Seq:
    def __init__(self, data):
        self._d = str(data)

    def __str__(self):
        return self._d

    def __repr__(self):
        return "Seq(%r)" % self._d

    def __len__(self):
        return len(self._d)

    def __eq__(self, other):
        return str(self) == str(other)

    def __hash__(self):
        return hash(self._d)

    def __getitem__(self, i):
        r = self._d[i]
        return Seq(r) if isinstance(i, slice) else r

    def __add__(self, other):
        return Seq(self._d + str(other))

    def __radd__(self, other):
        return Seq(str(other) + self._d)

    def __contains__(self, x):
        return str(x) in self._d

    def upper(self):
        return Seq(self._d.upper())

    def lower(self):
        return Seq(self._d.lower())

    def find(self, *a):
        return self._d.find(*[str(x) if not isinstance(x, int) else x for x in a])

    def reverse_complement(self):
        return Seq(self._d.translate(_COMP)[::-1])

    def translate(self, table=1, to_stop=False):
        s = self._d.upper().replace("U", "T")
        out = []
        for i in range(0, len(s) - len(s) % 3, 3):
            aa = _CODON.get(s[i : i + 3], "X")
            if to_stop and aa == "*":
                break
            out.append(aa)
        return Seq("".join(out))
