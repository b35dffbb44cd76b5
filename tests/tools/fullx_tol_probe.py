"""GPU probe: how far is the device rounding loop's fullx from the oracle's at a BASELINE-size golden as a function of the PCG tolerance?
The oracle's LM re-solves the penalty system until |d| < tol (iterative refinement), so its fullx carries ~eps*kappa of forward error; PCG
stopped at relres 1e-10 carries up to kappa*1e-10 in the lowest modes. Prints one JSON line per tolerance.

    python tests/tools/fullx_tol_probe.py octa9 1e-10 1e-11 1e-12
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(1, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from golden.make_golden_large import make_session  # noqa: E402
from metriko_b200 import api, capi  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "octa9"
tols = [float(t) for t in sys.argv[2:]] or [1e-10, 1e-11, 1e-12]
g = dict(np.load(os.path.join(ROOT, "tests", "golden", f"large_{name}.npz")))
s = make_session(json.loads(str(g["recipe"])))
fi = s.field_inputs()
nF = s.dims["nF"]
ctx = capi.Context(0)
F = api.FaceRosyField(ctx, fi["face2half"], fi["next"], fi["twin"], fi["face"], fi["heArgOnFace"], fi["faceArea"], 4)
z = F.solveSmallestEig(oracle.random_u0(nF, int(g["seed"])), 50)
sel = g["z50_idx"] if len(g["z50_idx"]) else np.arange(nF)
dd = np.exp(1j * (np.angle(z[sel]) - g["z50_arg"]))
ph = dd.mean() / abs(dd.mean())
ferr = np.abs(np.angle(dd * np.conj(ph))) / 4
print(json.dumps(dict(name=name, stage="field", removed_global_phase_rad=float(np.angle(ph)), err_max=float(ferr.max()), err_median=float(np.median(ferr)),
                      relres_max=float(F.stats["relres_max"]), inner=int(F.stats["inner_iterations"]), lam=float(F.stats["lambda_"]), golden_lam=float(g["lam"]))), flush=True)
F.close()
s.field_post(z, 4, "curl")
s.igm_setup(float(g["gridscale"]))
inp = s.igm_inputs()
ref = dict(zip(g["out_fixedIdcs"].tolist(), g["out_fixedVals"].tolist()))
for tol in tols:
    prob = api.IgmProblem(ctx, inp["X2F"], inp["UFull"], inp["F2"], inp["length"], tol=tol)
    if "x0_val" in g:
        xs0, xf0 = prob.poisson(inp["fixedIdcs"], inp["fixedVals"])
        d0 = np.abs(xf0[g["fullx_idx"]] - g["x0_val"])
        print(json.dumps(dict(name=name, tol=tol, stage="poisson", x0_max_diff=float(d0.max()), x0_median_diff=float(np.median(d0)), stats=prob.stats)), flush=True)
    try:
        out = prob.round(inp["fixedIdcs"], inp["fixedVals"], inp["singularIdcs"], inp["integerIdcs"], True, bool(g["round_seams"]))
    except Exception as e:  # e.g. MK_ERR_NOT_CONVERGED below the fp64 floor of the residual
        print(json.dumps(dict(name=name, tol=tol, error=repr(e), stats=prob.stats)), flush=True)
        continue
    finally:
        prob.close()
    got = dict(zip([int(i) for i in out["fixedIdcs"]], [int(v) for v in out["fixedVals"]]))
    d = np.abs(out["fullx"][g["fullx_idx"]] - g["fullx_val"])
    print(json.dumps(dict(name=name, tol=tol, success=bool(out["success"]), same_integers=(got == ref), n_solves=out["n_solves"],
                          golden_n_solves=int(g["n_solves"]), fullx_max_diff=float(d.max()), fullx_median_diff=float(np.median(d)),
                          fullx_absmax=float(np.abs(g["fullx_val"]).max()), inner=int(out["stats"]["inner_iterations"]),
                          relres_max=float(out["stats"]["relres_max"]), status=int(out["stats"]["status"]),
                          solve_ms=float(out["stats"]["solve_ms"]))), flush=True)
