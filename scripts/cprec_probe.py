"""Development tool: the constraint preconditioner (MK_IGM_CPREC=1, csrc/igm.cuh) against the projected one on ONE workload:
same rounded integers, inner iterations and ms of the IGM stage. Prints one JSON line per configuration."""
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
CONFIGS = [{}, {"MK_IGM_CPREC": 1}, {"MK_IGM_CPREC": 1, "MK_AMG_TYPED": 1}]


def run(ctx, levels, reps):
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
    X2F, UFull = inp["X2F"].tocsr(), inp["UFull"].tocsr()
    X2F.sort_indices()
    UFull.sort_indices()
    ref = None
    for cfg in CONFIGS:
        for k in ("MK_IGM_CPREC", "MK_AMG_TYPED"):
            os.environ.pop(k, None)
        for k, v in cfg.items():
            os.environ[k] = str(v)
        for _ in range(reps):
            prob = api.IgmProblem(ctx, X2F, UFull, inp["F2"], inp["length"])
            t = time.time()
            o = prob.round(inp["fixedIdcs"], inp["fixedVals"], inp["singularIdcs"], inp["integerIdcs"], True, False)
            wall = time.time() - t
            prob.close()
        ints = dict(zip(map(int, o["fixedIdcs"]), map(int, o["fixedVals"])))
        if ref is None:
            ref = (ints, o["fullx"].copy())
        print(json.dumps({"levels": levels, "cfg": cfg, "ok": bool(o["success"]), "same_integers": ints == ref[0],
                          "fullx_maxdiff": float(np.abs(o["fullx"] - ref[1]).max()), "igm_solve_ms": round(o["stats"]["solve_ms"], 1),
                          "igm_inner": o["stats"]["inner_iterations"], "relres_max": o["stats"]["relres_max"], "wall_ms": round(wall * 1e3, 1)}),
              flush=True)


if __name__ == "__main__":
    ctx = capi.Context(0)
    for lv in [int(a) for a in sys.argv[1:]] or [5, 9]:
        run(ctx, lv, 2 if lv >= 8 else 1)
