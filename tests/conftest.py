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


def load_ref(name):
    """tests/golden/ref_<name>.npz: what the REFERENCE'S OWN CODE computed for the mesh of the golden `name` (oracle/_ref: metriko's headers compiled
    in place against Eigen stand-ins; generator tests/golden/make_golden_ref.py; tests/test_oracle_ref.py ties the oracle goldens to it)."""
    return dict(np.load(os.path.join(GOLDEN, "ref_" + name + ".npz")))


def assert_field_matches_reference_code(name, z, tol):
    """A field solved from the golden's u0 against the 50th iterate of the reference's own FaceRosyField(mesh, 4, Smoothest) built from (V, F, u0)
    (frf.h:17-44 through hmesh.h:313-569 and sls.h:29-54): per-face direction after removing the global phase."""
    from oracle.field import field_angle_error
    err = float(field_angle_error(z, load_ref(name)["frf_compressed"], 4).max())
    assert err < tol, f"{name}: {err} rad off the reference code's field"


def assert_rounding_matches_reference_code(name, fixedIdcs, fixedVals, fullx, tol=2e-6):
    """A rounding result against the reference's own iterative_rounding (ir.h:33-113) on the golden's inputs (`fullx`) and against its whole
    example_igm.cpp pipeline from (V, F, u0, seam) (`ex_fullx`): the integers exactly, fullx within `tol` (the CUDA path is held to 1e-6 against
    the oracle golden, the oracle golden is within 5e-8 of both)."""
    r = load_ref(name)
    idx = np.asarray(fixedIdcs, np.int64)
    val = np.asarray(fixedVals, np.int64)
    for key in ("fullx", "ex_fullx"):
        fx = r[key]
        assert np.array_equal(np.round(fx[idx]).astype(np.int64), val), f"{name}: integers differ from the reference code's ({key})"
        d = float(np.abs(np.asarray(fullx) - fx).max())
        assert d < tol, f"{name}: fullx {d} off the reference code's ({key})"


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
