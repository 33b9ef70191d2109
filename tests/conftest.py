import gzip
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _load(name):
    with gzip.open(os.path.join(GOLDEN, name), "rt") as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def dcr_cases():
    """Per-read dcr() results recorded from the unmodified reference (oracle/make_golden.py)."""
    return _load("dcr_cases.json.gz")


@pytest.fixture(scope="session")
def decombinator_runs():
    """Whole-file decombinator() runs recorded from the unmodified reference."""
    return _load("decombinator_runs.json.gz")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
