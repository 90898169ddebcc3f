import json
import os
import struct
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def from_bits(s):
    if s == "inf":
        return float("inf")
    return struct.unpack("<d", struct.pack("<Q", int(s, 16)))[0]


def golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def bits_of(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def assert_bit_equal(a, b, what=""):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    bad = np.flatnonzero(a.view(np.uint64).ravel() != b.view(np.uint64).ravel())
    assert bad.size == 0, f"{what}: {bad.size} of {a.size} values differ, first at {bad[0]}: " \
                          f"{a.ravel()[bad[0]]!r} vs {b.ravel()[bad[0]]!r}"


@pytest.fixture(scope="session")
def ctx():
    """The CUDA context of the product library.  GPU tests must run the CUDA path: no skip-on-missing here."""
    import eqsolver_b200 as E
    return E.Context(0)
