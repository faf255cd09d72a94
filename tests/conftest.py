import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(tag):
    g = np.load(os.path.join(GOLDEN, tag + ".npz"))
    X = sp.csr_matrix((g["data"], g["indices"], g["indptr"]), shape=tuple(g["shape"]))
    return g, X


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.fixture(scope="session")
def golden_c1_zipf():
    return load_golden("c1_zipf")


@pytest.fixture(scope="session")
def golden_c1_planted():
    return load_golden("c1_planted")


@pytest.fixture(scope="session")
def golden_small():
    return load_golden("small_cases")
