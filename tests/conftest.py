"""Test configuration.

`-m "not gpu"` (CPU box): oracle vs golden vectors / brute force / the reference's compiled
libsais, host logic, and that libpss_b200.so loads and exports every symbol of include/pss.h.
`-m gpu` (B200): parity of the CUDA path against the oracle, through the C ABI and the
Python API.  /root/reference is never read at test time.
"""
import ctypes as C
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
LIB_PATH = os.path.join(ROOT, "pysubstringsearch_b200", "libpss_b200.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_available():
    try:
        lib = C.CDLL(LIB_PATH)
        return lib.pss_device_count() > 0
    except OSError:
        return False


HAVE_GPU = _cuda_available()


def pytest_collection_modifyitems(config, items):
    if HAVE_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this environment")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def kats():
    with open(os.path.join(GOLDEN, "reference_kats.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def vectors():
    with open(os.path.join(GOLDEN, "sa_vectors.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    return O


@pytest.fixture(scope="session")
def pss():
    from pysubstringsearch_b200 import capi
    return capi
