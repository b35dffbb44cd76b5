"""Development tool: solve-phase time / iteration counts of the 1M-vertex step under different multigrid parameters (env MK_AMG_*),
on ONE workload built once. Prints one line per configuration."""
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
levels = int(sys.argv[1]) if len(sys.argv) > 1 else 9
configs = [{}] + [json.loads(a) for a in sys.argv[2:]]
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
X2F, UFull = inp["X2F"].tocsr(), inp["UFull"].tocsr()
X2F.sort_indices()
UFull.sort_indices()
for cfg in configs:
    for k in list(os.environ):
        if k.startswith("MK_AMG_") or k in ("MK_FP64_TAIL", "MK_PRECOND_FP32"):
            del os.environ[k]
    for k, v in cfg.items():
        os.environ[k] = str(v)
    res = {}
    for rep in range(2):
        F.assemble()
        F.solveSmallestEig(u0, 50)
        st = dict(F.stats)
        prob = api.IgmProblem(ctx, X2F, UFull, inp["F2"], inp["length"])
        o = prob.round(inp["fixedIdcs"], inp["fixedVals"], inp["singularIdcs"], inp["integerIdcs"], True, False)
        prob.close()
        res = dict(field_setup=round(st["setup_ms"]), field_solve=round(st["solve_ms"]), field_inner=st["inner_iterations"],
                   igm_setup=round(o["stats"]["setup_ms"]), igm_solve=round(o["stats"]["solve_ms"]), igm_inner=o["stats"]["inner_iterations"],
                   ok=bool(o["success"]))
    res["total"] = res["field_setup"] + res["field_solve"] + res["igm_setup"] + res["igm_solve"]
    print(json.dumps({"cfg": cfg, **res}), flush=True)
