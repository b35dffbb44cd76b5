"""Where does a PCG iteration's time go? Live (graph-replay) timing of the preconditioner, level by level, at the bench workload.
    python scripts/cycle_probe.py [levels=9]      -> one JSON line (field + IGM): hierarchy, us per application / per level, us per iteration
Environment knobs (MK_TAIL_ROWS, MK_AMG_*) are read when the hierarchies are built, so variants are separate processes."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(1, os.path.join(ROOT, "tests"))
from metriko_b200 import api, capi  # noqa: E402
from hostglue import Session  # noqa: E402

levels = int(sys.argv[1]) if len(sys.argv) > 1 else 9
ctx = capi.Context(0)
s = Session().make_sphere(8, levels, bump=0.08)
fi = s.field_inputs()
nF = s.dims["nF"]
rng = np.random.default_rng(0)
u0 = rng.uniform(-1, 1, nF) + 1j * rng.uniform(-1, 1, nF)
F = api.FaceRosyField(ctx, fi["face2half"], fi["next"], fi["twin"], fi["face"], fi["heArgOnFace"], fi["faceArea"], 4)
z = F.solveSmallestEig(u0, 50)
out = {"env": {k: v for k, v in os.environ.items() if k.startswith("MK_")}, "nF": nF}
st = F.stats
out["field"] = {"hierarchy": F.hierarchy(), "cycle_us (apply, from L1.., [xr+handover, cycle+dot, update_p])": F.profile_cycle(200), "spmv_dot_us": F.spmv_bench(200, True) * 1e3,
                "solve_ms": st["solve_ms"], "inner": st["inner_iterations"], "us_per_iteration": st["solve_ms"] * 1e3 / st["inner_iterations"]}
s.field_post(z, 4, "curl")
s.igm_setup(0.02)
inp = s.igm_inputs()
prob = api.IgmProblem(ctx, inp["X2F"], inp["UFull"], inp["F2"], inp["length"])
r = prob.round(inp["fixedIdcs"], inp["fixedVals"], inp["singularIdcs"], inp["integerIdcs"], True, False)
st = r["stats"]
out["igm"] = {"hierarchy": prob.hierarchy(), "cycle_us (apply, from L1..)": prob.profile_cycle(200), "spmv_dot_us": prob.spmv_bench(200, True) * 1e3,
              "solve_ms": st["solve_ms"], "inner": st["inner_iterations"], "us_per_iteration": st["solve_ms"] * 1e3 / st["inner_iterations"]}
print(json.dumps(out))
