import os, sys
import numpy as np
ROOT='/root/repo'
sys.path.insert(0, ROOT); sys.path.insert(1, os.path.join(ROOT,'tests'))
from metriko_b200 import api, capi
from hostglue import Session
ctx = capi.Context(0)
s = Session().make_sphere(8, 9, bump=0.08)
fi = s.field_inputs(); nF = s.dims["nF"]
rng = np.random.default_rng(0)
u0 = rng.uniform(-1, 1, nF) + 1j * rng.uniform(-1, 1, nF)
F = api.FaceRosyField(ctx, fi["face2half"], fi["next"], fi["twin"], fi["face"], fi["heArgOnFace"], fi["faceArea"], 4)
mode = sys.argv[1]
if mode == "two":
    F.solveSmallestEig(u0, 1)
print("=== timed solve", flush=True)
sys.stderr.write("=== timed solve\n"); sys.stderr.flush()
z = F.solveSmallestEig(u0, 50)
print(mode, F.stats, flush=True)
