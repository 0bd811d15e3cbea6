import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def fixtures():
    """the reference's own test fixtures (tests/golden/make_fixtures.py)"""
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_fixtures.npz"))


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def K():
    """the product, initialised on cuda:0 -- fails loudly without the library or a GPU"""
    import kmerlr_b200 as K
    K.init(0)
    return K


def cat(fixtures, a, b):
    """fg || bg concatenation of two fixture sets -> (buffer, offsets, labels)"""
    sa, oa = fixtures[a + "_seq"], fixtures[a + "_off"]
    sb, ob = fixtures[b + "_seq"], fixtures[b + "_off"]
    buf = np.concatenate([sa, sb])
    off = np.concatenate([oa, ob[1:] + oa[-1]])
    labels = np.concatenate([np.ones(len(oa) - 1, dtype=np.uint8), np.zeros(len(ob) - 1, dtype=np.uint8)])
    return buf, off, labels
