"""Size sweep on one B200 (north_star: "report for synthetic meshes of 10k to 8M vertices"): field + IGM step time, inner iteration
counts and the SpMV roofline fraction per size. One JSON line per mesh; `--torus` adds BASELINE configs[1] (320x320 torus, ~100k
vertices). Workload generation (host glue) is untimed; a step is what bench.py times (assemble + hierarchy + 50 inverse iterations,
A0 + hierarchy + Poisson + rounding loop), measured once after one warm-up step.

  python scripts/bench_sweep.py --levels 6 7 8 9 10 --torus
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(1, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from metriko_b200 import api, capi  # noqa: E402
from hostglue import Session  # noqa: E402

FIELD_KEYS = ("face2half", "next", "twin", "face", "heArgOnFace", "faceArea")


def peak():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"])
    except Exception:
        return 6650.0


def one(ctx, s, name, igm=True):
    import torch
    fi = s.field_inputs()
    nF, nV = s.dims["nF"], s.dims["nV"]
    rng = np.random.default_rng(0)
    u0 = rng.uniform(-1, 1, nF) + 1j * rng.uniform(-1, 1, nF)
    F = api.FaceRosyField(ctx, *[fi[k] for k in FIELD_KEYS], 4)
    z = F.solveSmallestEig(u0, 50)
    out = {"mesh": name, "verts": nV, "faces": nF}
    inp = None
    if igm:
        s.field_post(z, 4, "curl")
        s.igm_setup(0.02)
        inp = s.igm_inputs()
        X2F, UFull = inp["X2F"].tocsr(), inp["UFull"].tocsr()
        X2F.sort_indices()
        UFull.sort_indices()

    def step():
        F.assemble()
        F.solveSmallestEig(u0, 50)
        st = dict(F.stats)
        r = {"field_setup_ms": st["setup_ms"], "field_solve_ms": st["solve_ms"], "field_inner": st["inner_iterations"],
             "field_relres_max": st["relres_max"], "field_status": st["status"]}
        if igm:
            prob = api.IgmProblem(ctx, X2F, UFull, inp["F2"], inp["length"])
            o = prob.round(inp["fixedIdcs"], inp["fixedVals"], inp["singularIdcs"], inp["integerIdcs"], True, False)
            prob.close()
            assert o["success"]
            r.update(igm_setup_ms=o["stats"]["setup_ms"], igm_solve_ms=o["stats"]["solve_ms"], igm_inner=o["stats"]["inner_iterations"],
                     igm_resolves=o["n_solves"], igm_unknowns=int(UFull.shape[1]))
        return r

    step()
    torch.cuda.synchronize()
    t = time.perf_counter()
    r = step()
    torch.cuda.synchronize()
    out["step_ms"] = (time.perf_counter() - t) * 1e3
    out.update(r)
    us = F.spmv_bench(200, False) * 1e3
    out["spmv_us"] = us
    out["spmv_gbs_algorithmic"] = 116 * nF / (us * 1e-6) / 1e9
    out["spmv_frac_of_peak"] = out["spmv_gbs_algorithmic"] / peak()
    out["spmv_working_set_mb"] = 92 * nF / 1e6
    out["verts_per_s"] = nV / (out["step_ms"] * 1e-3)
    F.close()
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--levels", type=int, nargs="*", default=[6, 7, 8, 9])
    ap.add_argument("--torus", action="store_true")
    ap.add_argument("--field-only-above", type=int, default=10, help="levels >= this skip the IGM stage (host glue time)")
    args = ap.parse_args()
    ctx = capi.Context(0)
    if args.torus:
        one(ctx, Session().make_torus(320, 320, bump=0.1), "torus 320x320 (BASELINE configs[1])")
    for lv in args.levels:
        one(ctx, Session().make_sphere(8, lv, bump=0.08), f"octahedron x{lv}", igm=lv < args.field_only_above)


if __name__ == "__main__":
    main()
