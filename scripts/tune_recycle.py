"""Development tool: field solve at 1M vertices for different recycle-space sizes (per-outer inner iteration counts with MK_VERBOSE=1)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(1, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from metriko_b200 import api, capi
from hostglue import Session
FIELD_KEYS = ("face2half", "next", "twin", "face", "heArgOnFace", "faceArea")
ctx = capi.Context(0)
s = Session().make_sphere(8, int(sys.argv[1]) if len(sys.argv) > 1 else 9, bump=0.08)
fi = s.field_inputs(); nF = s.dims["nF"]
rng = np.random.default_rng(0)
u0 = rng.uniform(-1, 1, nF) + 1j * rng.uniform(-1, 1, nF)
F = api.FaceRosyField(ctx, *[fi[k] for k in FIELD_KEYS], 4)
F.solveSmallestEig(u0, 3)
for rec in [int(a) for a in sys.argv[2:]] or [32]:
    F.solveSmallestEig(u0, 50, recycle=rec)
    st = F.stats
    print(f"recycle {rec}: solve {st['solve_ms']:.0f} ms, inner {st['inner_iterations']}, max {st['inner_max']}, relres {st['relres_max']:.2e}", flush=True)
