"""Stand-in for acora==2.4 (only what decombine.py:109,722-746,275.. calls).

Semantics (acora docs): ``AcoraBuilder.add(*keywords)``, ``.build()`` returns a
matcher whose ``findall(s)`` lists every (possibly overlapping) occurrence of
every keyword as ``(keyword, start)``, in the order the automaton reports them,
i.e. ascending END position (a longer keyword before its own suffix at the
same end).  Test infrastructure only.
"""


class _Matcher:
    def __init__(self, keywords):
        self._kw = list(keywords)

    def findall(self, s):
        hits = []
        for kw in self._kw:
            start = s.find(kw)
            while start != -1:
                hits.append((start + len(kw), -len(kw), kw, start))
                start = s.find(kw, start + 1)
        hits.sort()
        return [(kw, start) for _, _, kw, start in hits]

    def finditer(self, s):
        return iter(self.findall(s))


class AcoraBuilder:
    def __init__(self, *keywords):
        self._kw = {}
        self.add(*keywords)

    def add(self, *keywords):
        for k in keywords:
            if k:
                self._kw.setdefault(k, None)

    def update(self, keywords):
        self.add(*keywords)

    def build(self, ignore_case=None, acora=None):
        return _Matcher(self._kw.keys())
