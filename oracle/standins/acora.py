"""Stand-in for acora==2.4 (only what decombine.py:109,722-746,275.. calls).

Semantics (acora docs): ``AcoraBuilder.add(*keywords)``, ``.build()`` returns a
matcher whose ``findall(s)`` lists every (possibly overlapping) occurrence of
every keyword as ``(keyword, start)``, in the order the automaton reports them,
i.e. ascending END position (a longer keyword before its own suffix at the
same end).  Test infrastructure only.

Two matchers with the same results: a compiled Aho-Corasick automaton (``_ac.c``, loaded from ``oracle/_ref/libac.so``
when that has been built -- ``oracle/time_reference.py`` builds it, so that timing the reference charges the tag search
what a compiled automaton costs, like the real Cython acora) and a pure-Python one (``str.find`` per keyword).
"""
import ctypes
import os

_LIB = None


def _load_compiled():
    global _LIB
    if _LIB is None:
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "_ref", "libac.so")
        _LIB = False
        if os.path.exists(path) and not os.environ.get("DCB_ACORA_PURE_PYTHON"):
            lib = ctypes.CDLL(path)
            lib.ac_build.restype = ctypes.c_void_p
            lib.ac_build.argtypes = [ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_int32), ctypes.c_int]
            lib.ac_findall.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int32), ctypes.c_int]
            lib.ac_free.argtypes = [ctypes.c_void_p]
            _LIB = lib
    return _LIB


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


class _CompiledMatcher:
    def __init__(self, keywords, lib):
        self._kw = list(keywords)
        self._lib = lib
        enc = [k.encode("latin-1") for k in self._kw]
        arr = (ctypes.c_char_p * len(enc))(*enc)
        lens = (ctypes.c_int32 * len(enc))(*[len(k) for k in enc])
        self._h = lib.ac_build(arr, lens, len(enc))
        self._buf = (ctypes.c_int32 * 128)()

    def findall(self, s):
        b = s.encode("latin-1") if isinstance(s, str) else bytes(s)
        n = self._lib.ac_findall(self._h, b, len(b), self._buf, 64)
        if n > 64:
            big = (ctypes.c_int32 * (2 * n))()
            self._lib.ac_findall(self._h, b, len(b), big, n)
            return [(self._kw[big[2 * i]], big[2 * i + 1]) for i in range(n)]
        buf = self._buf
        return [(self._kw[buf[2 * i]], buf[2 * i + 1]) for i in range(n)]

    def finditer(self, s):
        return iter(self.findall(s))

    def __del__(self):
        try:
            self._lib.ac_free(self._h)
        except Exception:
            pass


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
        lib = _load_compiled()
        if lib:
            return _CompiledMatcher(self._kw.keys(), lib)
        return _Matcher(self._kw.keys())
