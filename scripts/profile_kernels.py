"""Light driver for ncu captures (a run under ncu is never a bench value).

  python scripts/profile_kernels.py prep /tmp/wl.npz [levels]   build the 1M-vertex workload once (NOT under ncu) and save it
  ncu ... python scripts/profile_kernels.py run /tmp/wl.npz     few launches of every hot kernel on that workload:
        field SpMV (plain + fused dot), 2 inverse-iteration steps (PCG + V-cycle + recycle kernels), IGM SpMV,
        IGM Poisson solve + rounding loop
"""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(1, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from metriko_b200 import api, capi  # noqa: E402
from hostglue import Session  # noqa: E402

FIELD_KEYS = ("face2half", "next", "twin", "face", "heArgOnFace", "faceArea")
IGM_KEYS = ("fixedIdcs", "fixedVals", "singularIdcs", "integerIdcs", "F2")


def prep(path, levels):
    ctx = capi.Context(0)
    s = Session().make_sphere(8, levels, bump=0.08)
    fi = s.field_inputs()
    nF = s.dims["nF"]
    rng = np.random.default_rng(0)
    u0 = rng.uniform(-1, 1, nF) + 1j * rng.uniform(-1, 1, nF)
    F = api.FaceRosyField(ctx, *[fi[k] for k in FIELD_KEYS], 4)
    z = F.solveSmallestEig(u0, 50)
    s.field_post(z, 4, "curl")
    s.igm_setup(0.02)
    inp = s.igm_inputs()
    X, U = inp["X2F"].tocsr(), inp["UFull"].tocsr()
    d = {k: fi[k] for k in FIELD_KEYS}
    d.update({k: inp[k] for k in IGM_KEYS})
    d.update(u0=u0, length=np.float64(inp["length"]), Xp=X.indptr, Xc=X.indices, Xv=X.data, Xs=np.asarray(X.shape),
             Up=U.indptr, Uc=U.indices, Uv=U.data, Us=np.asarray(U.shape))
    np.savez(path, **d)
    print("saved", path, s.dims)


def run(path, what):
    d = np.load(path)
    ctx = capi.Context(0)
    if "field" in what:
        F = api.FaceRosyField(ctx, *[d[k] for k in FIELD_KEYS], 4)
        print("field spmv us", F.spmv_bench(1, False) * 1e3, F.spmv_bench(1, True) * 1e3)
        F.solveSmallestEig(d["u0"], 3)
        print("field 3 outer:", F.stats)
    if "igm" in what:
        X = sp.csr_matrix((d["Xv"], d["Xc"], d["Xp"]), shape=tuple(d["Xs"]))
        U = sp.csr_matrix((d["Uv"], d["Uc"], d["Up"]), shape=tuple(d["Us"]))
        prob = api.IgmProblem(ctx, X, U, d["F2"], float(d["length"]))
        print("igm spmv us", prob.spmv_bench(1, False) * 1e3)
        t = time.time()
        if "round" in what:
            out = prob.round(d["fixedIdcs"], d["fixedVals"], d["singularIdcs"], d["integerIdcs"], True, False)
            print("round", time.time() - t, out["stats"], out["n_solves"])
        else:
            prob.poisson(d["fixedIdcs"], d["fixedVals"])
            print("poisson", time.time() - t, prob.stats)


if __name__ == "__main__":
    if sys.argv[1] == "prep":
        prep(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 9)
    else:
        run(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "field,igm")
