import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.join(ROOT, "tests") not in sys.path:
    sys.path.insert(1, os.path.join(ROOT, "tests"))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) device; run with `-m gpu` on the GPU box")


def load_golden(name):
    import scipy.sparse as sp
    d = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    for m in ("X2F", "UFull"):
        d[m] = sp.csr_matrix((d[m + "_val"], d[m + "_col"], d[m + "_ptr"]), shape=tuple(int(v) for v in d[m + "_shape"]))
    return d


@pytest.fixture(scope="session")
def ctx():
    """CUDA context through the C-ABI. GPU tests must fail (not skip) when the library or the device is missing:
    a silent fallback would void every parity claim."""
    from metriko_b200 import capi
    return capi.Context(0)


@pytest.fixture(scope="session")
def ico4():
    from hostglue import Session
    return Session().make_sphere(20, 4, bump=0.08)
