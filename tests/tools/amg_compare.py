"""Development tool: PCG iteration counts of the mirrored V-cycle for different aggregation modes (CPU only)."""
import os, sys, time
import numpy as np, scipy.sparse as sp
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
from amg_mirror import VCycle, pcg, describe
from metriko_b200 import capi
from hostglue import Session

lv = int(sys.argv[1]) if len(sys.argv) > 1 else 6
s = Session().make_sphere(8, lv, bump=0.08)
fi = s.field_inputs(); nF = s.dims["nF"]
L = oracle.connection_laplacian(fi["twin"], fi["face"], fi["heArgOnFace"], nF, 4).tocsr()
M = oracle.galerkin_mass_matrix(fi["faceArea"])
u0 = oracle.random_u0(nF, 0)
b = M @ u0
res = {}
for mode in sys.argv[2:] or ["0", "1"]:
    os.environ["MK_AMG_AGG"] = mode
    t = time.time(); H = capi.amg_host_hierarchy(L); t = time.time() - t
    x, it = pcg(L, b, np.zeros_like(b), VCycle(H))
    cx = sum(h[0].nnz for h in H) / H[0][0].nnz
    print(f"field agg={mode}: its {it}  levels {describe(H)}  opcx {cx:.2f} setup {t:.2f}s", flush=True)
if os.environ.get("SKIP_IGM"): sys.exit()
from scipy.sparse.linalg import splu
z = oracle.solve_smallest_eig(L, M, u0, 50)
s.field_post(z, 4, "curl"); s.igm_setup(0.02); inp = s.igm_inputs()
X = (inp["X2F"] * float(inp["length"])).tocsr()
A = (X.T @ X + 1e-4 * sp.identity(X.shape[1])).tocsr(); rhs = X.T @ inp["F2"]
for mode in sys.argv[2:] or ["0", "1"]:
    os.environ["MK_AMG_AGG"] = mode
    t = time.time(); H = capi.amg_host_hierarchy(A); t = time.time() - t
    x, it = pcg(A, rhs, np.zeros_like(rhs), VCycle(H))
    cx = sum(h[0].nnz for h in H) / H[0][0].nnz
    print(f"igm   agg={mode}: its {it}  levels {describe(H)}  opcx {cx:.2f} setup {t:.2f}s", flush=True)
