"""Development tool (CPU only): PCG iteration counts of one fresh solve with the IGM operator A = length^2 X2F^T X2F + 1e-4 I (tol 1e-10)
under variants of the multigrid cycle / hierarchy parameters, typed aggregation (the device default).
    python tests/tools/igm_cycle_study.py 7 [theta ...]
"""
import os, sys, time
import numpy as np, scipy.sparse as sp
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "tools"))
import oracle
from amg_mirror import describe, pcg
from metriko_b200 import capi
from hostglue import Session

src = open(os.path.join(ROOT, "tests/tools/field_cycle_study.py")).read().split("lv = int(sys.argv[1])")[0]
exec(src)  # class Cycle

lv = int(sys.argv[1]) if len(sys.argv) > 1 else 6
s = Session().make_sphere(8, lv, bump=0.08)
fi = s.field_inputs(); nF = s.dims["nF"]
L = oracle.connection_laplacian(fi["twin"], fi["face"], fi["heArgOnFace"], nF, 4).tocsr()
M = oracle.galerkin_mass_matrix(fi["faceArea"])
if lv <= 7:
    z = oracle.solve_smallest_eig(L, M, oracle.random_u0(nF, 0), 50)
else:  # a smooth field is all the operator needs: a few hundred Jacobi sweeps on a random start
    z = oracle.random_u0(nF, 0)
    for _ in range(400):
        z = z - (L @ z) / 6.0
        z /= np.abs(z).max()
s.field_post(z, 4, "curl"); s.igm_setup(0.02); inp = s.igm_inputs()
U = sp.csr_matrix(inp["UFull"]); X = (sp.csr_matrix(inp["X2F"]) * float(inp["length"])).tocsr()
A = (X.T @ X + 1e-4 * sp.identity(X.shape[1])).tocsr(); A.sort_indices()
f = X.T @ inp["F2"]
types = capi.igm_dof_types(U)
for theta in (sys.argv[2:] or ["0.1"]):
    os.environ["MK_AMG_THETA"] = theta
    H = capi.amg_host_hierarchy(A, types=types)
    print("unknowns", A.shape[0], "theta", theta, "levels", describe(H), flush=True)
    for name, kw in [("V deg 1/2 (device default)", dict()), ("V deg 2/2", dict(deg0=2)), ("W from level 1", dict(gamma_from=1)),
                     ("two-grid (exact level 1)", dict(exact_from=1)), ("three-grid (exact level 2)", dict(exact_from=2))]:
        t0 = time.time()
        x, it = pcg(A, f, np.zeros(A.shape[0]), Cycle(H, **kw), tol=1e-10)
        print(f"  {name:32s} {it:4d} iterations  ({time.time() - t0:.0f}s)", flush=True)
