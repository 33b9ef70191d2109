"""Import the UNMODIFIED reference (`/root/reference/src/decombinator`) in the build container.

TEST INFRASTRUCTURE ONLY.  Used by the scripts that record what the reference
itself returns (`oracle/make_golden.py`, `make_golden_collapse.py`,
`make_golden_collapse_oligos.py`, `make_golden_pipeline_digest.py`,
`time_reference.py`); nothing that runs on the GPU box imports it.  The missing third-party
wheels are replaced by the pure-Python stand-ins in `oracle/standins/`, and
`importlib.metadata.version("decombinator")` (decombine.py:884) is patched to
return a string because the reference is not pip-installed.
"""
import importlib
import os
import sys
from importlib import metadata as _md

REFERENCE_ROOT = os.environ.get("DCB_REFERENCE_ROOT", "/root/reference")
REF_SRC = os.path.join(REFERENCE_ROOT, "src")
REF_TAGDIR = os.path.join(REFERENCE_ROOT, "tests", "resources", "Decombinator-Tags-FASTAs")
_STANDINS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "standins")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_SRC, "decombinator"))


def load():
    """Return the reference package modules (decombine, collapse, io, pipeline, translate)."""
    if not available():
        raise RuntimeError("reference source not present at %s" % REF_SRC)
    if _STANDINS not in sys.path:
        sys.path.insert(0, _STANDINS)
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    _orig = _md.version

    def _version(name):
        if name == "decombinator":
            return "reference-src"
        return _orig(name)

    _md.version = _version
    import decombinator  # noqa: F401  (the reference package)

    mods = {}
    for m in ("decombine", "collapse", "io", "pipeline", "translate"):
        mods[m] = importlib.import_module("decombinator." + m)
    return mods
