"""Golden vectors of the localInjectivity (barrier) path of the oracle, derived from the committed fixtures:
    python tests/golden/make_golden_barrier.py      ->  tests/golden/<name>_barrier.npz
The reference ships no fixtures for this path; these pin the oracle's restatement (oracle/barrier.py, oracle/igm.py) so that the CUDA barrier
path has a fixed target that travels to the GPU box. tests/test_oracle_ref.py::test_barrier_golden_agrees_with_reference ties them to the
reference's own injective_barrier.h / LM code compiled in place (oracle/_ref): fullx within 4e-13 / 3e-12, the same number of factorisations."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import oracle  # noqa: E402
from conftest import load_golden  # noqa: E402
from oracle.barrier import InjectiveBarrier  # noqa: E402

for name in ("torus_24x16", "sphere_ico3"):
    g = load_golden(name)
    N = 4
    st = {}
    out = oracle.iterative_rounding_x2f(g["fixedIdcs"], g["fixedVals"], g["singularIdcs"], g["integerIdcs"], float(g["length"]), g["UFull"], g["X2F"],
                                        N, 2, True, bool(g["round_seams"]), g["F2"], stats=st, localInjectivity=True)
    B = InjectiveBarrier(N, len(g["F2"]) // (2 * N), g["F2"])
    r = B.ratio((g["X2F"] * float(g["length"])) @ out["x_small"])[0]
    np.savez_compressed(os.path.join(HERE, name + "_barrier.npz"), fullx=out["fullx"], x_small=out["x_small"], x_ni=out["x_ni"], F2_ni=out["F2_ni"],
                        out_fixedIdcs=np.asarray(out["fixedIdcs"]), out_fixedVals=np.asarray(out["fixedVals"]), n_solves=out["n_solves"],
                        success=out["success"], min_ratio=r.min(), n_below_s=int((r < 0.5).sum()), lm_iterations=st["lm_iterations"],
                        factorizations=st["factorizations"])
    print(name, "success", out["success"], "solves", out["n_solves"], "min ratio %.6f" % r.min(), "below s", int((r < 0.5).sum()), st)
