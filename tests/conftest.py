import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "dna-kmeres-parallel_b200"),
          os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "reference_golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    """The CPU restatement (test-side checker; never used by the product)."""
    import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def kmerlib():
    """libkmerb200.so through ctypes; built on demand."""
    import kmerb200
    if not os.path.exists(kmerb200.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    kmerb200.lib()
    return kmerb200


@pytest.fixture(scope="session")
def ctx(kmerlib):
    """A context on cuda:0 — fails loudly (no skip, no fallback) without a B200."""
    c = kmerlib.Context(0)
    yield c
    c.close()
