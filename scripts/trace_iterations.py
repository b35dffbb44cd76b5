"""Real timeline of the PCG iterations (CUPTI through torch.profiler: every kernel of the process, also inside CUDA-graph replays, with its
start time and duration -- not serialised, not cache-flushed like an ncu launch list).
    python scripts/trace_iterations.py field|igm [levels=9] [outer=2]   ->  gpurun_out/trace_<what>.csv (name, start us, duration us)"""
import csv
import json
import os
import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(1, os.path.join(ROOT, "tests"))
from metriko_b200 import api, capi  # noqa: E402
from hostglue import Session  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "field"
levels = int(sys.argv[2]) if len(sys.argv) > 2 else 9
outer = int(sys.argv[3]) if len(sys.argv) > 3 else 2
torch.cuda.init()
ctx = capi.Context(0)
s = Session().make_sphere(8, levels, bump=0.08)
fi = s.field_inputs()
nF = s.dims["nF"]
rng = np.random.default_rng(0)
u0 = rng.uniform(-1, 1, nF) + 1j * rng.uniform(-1, 1, nF)
F = api.FaceRosyField(ctx, fi["face2half"], fi["next"], fi["twin"], fi["face"], fi["heArgOnFace"], fi["faceArea"], 4)
if what == "field":
    F.solveSmallestEig(u0, 1)   # warm: hierarchy, graph
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        F.solveSmallestEig(u0, outer)
        torch.cuda.synchronize()
else:
    z = F.solveSmallestEig(u0, 50)
    s.field_post(z, 4, "curl")
    s.igm_setup(0.02)
    inp = s.igm_inputs()
    prob = api.IgmProblem(ctx, inp["X2F"], inp["UFull"], inp["F2"], inp["length"])
    prob.poisson(inp["fixedIdcs"], inp["fixedVals"])
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        if what == "igm_round":  # the whole rounding loop (Poisson + every re-solve)
            prob.round(inp["fixedIdcs"], inp["fixedVals"], inp["singularIdcs"], inp["integerIdcs"], True, False)
        else:
            prob.poisson(inp["fixedIdcs"], inp["fixedVals"])
        torch.cuda.synchronize()
path = os.path.join(ROOT, "gpurun_out", f"trace_{what}.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"] if ev else 0
with open(os.path.join(ROOT, "gpurun_out", f"trace_{what}.csv"), "w", newline="") as fh:
    w = csv.writer(fh)
    w.writerow(["name", "start_us", "dur_us"])
    for e in ev:
        w.writerow([e["name"][:90], round(e["ts"] - t0, 3), round(e["dur"], 3)])
os.remove(path)
print(what, "kernels traced:", len(ev))
