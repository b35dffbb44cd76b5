"""Development tool (CPU only): iteration counts of the mirrored rounding loop (tests/amg_mirror.py rounding_loop) on an octahedron sphere of
the given subdivision level, projected preconditioner P V P against the constraint preconditioner, untyped / typed aggregation.
    python tests/tools/mirror_rounding.py 7        # 65k vertices, ~2 min
The field comes from the oracle's direct solver (levels <= 7)."""
import os, sys, time
import numpy as np, scipy.sparse as sp
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
from amg_mirror import VCycle, describe, rounding_loop
from metriko_b200 import capi
from hostglue import Session

lv = int(sys.argv[1]) if len(sys.argv) > 1 else 6
s = Session().make_sphere(8, lv, bump=0.08)
fi = s.field_inputs(); nF = s.dims["nF"]
L = oracle.connection_laplacian(fi["twin"], fi["face"], fi["heArgOnFace"], nF, 4).tocsr()
M = oracle.galerkin_mass_matrix(fi["faceArea"])
z = oracle.solve_smallest_eig(L, M, oracle.random_u0(nF, 0), 50)
s.field_post(z, 4, "curl"); s.igm_setup(0.02); inp = s.igm_inputs()
U = sp.csr_matrix(inp["UFull"]); X = (sp.csr_matrix(inp["X2F"]) * float(inp["length"])).tocsr()
A = (X.T @ X + 1e-4 * sp.identity(X.shape[1])).tocsr(); A.sort_indices()
f = X.T @ inp["F2"]
print("unknowns", A.shape[0], "singular dofs", len(inp["singularIdcs"]), flush=True)
seq = None
for tname, types in (("untyped", None), ("typed", capi.igm_dof_types(U))):
    H = capi.amg_host_hierarchy(A, types=types)
    V = VCycle(H, deg=1)
    for mode in ("proj", "cprec"):
        t0 = time.time()
        its, ints, x = rounding_loop(A, f, U, inp["fixedIdcs"], inp["fixedVals"], inp["singularIdcs"], V, mode, X=X, replay=seq)
        seq = seq or ints
        print(f"{tname:8s} {mode:6s} total {sum(its):5d}  per solve {its}  ({time.time() - t0:.0f}s)  levels {describe(H)}", flush=True)
