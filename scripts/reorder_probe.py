"""Does the numbering of the reduced IGM dofs limit the A0 SpMV? Times y = A0 x (fp64 SELL kernel, mk_spd_spmv_bench) for the numbering the
host stages produce (subdivision order: a vertex's neighbours are far away in memory) and for a reverse Cuthill-McKee renumbering.
    python scripts/reorder_probe.py [levels=9]"""
import json
import os
import sys

import numpy as np
import scipy.sparse as sp
from scipy.sparse.csgraph import reverse_cuthill_mckee

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
F.close()
s.field_post(z, 4, "curl")
s.igm_setup(0.02)
inp = s.igm_inputs()
prob = api.IgmProblem(ctx, inp["X2F"], inp["UFull"], inp["F2"], inp["length"], precond=0)
A = prob.matrix().tocsr()
prob.close()
out = {"rows": A.shape[0], "nnz": int(A.nnz)}
bytes_alg = A.nnz * 12 + (A.shape[0] + 1) * 4 + A.shape[0] * 16


def bench(M, tag):
    w = api.CholeskyWrapper(ctx, precond=0)
    assert w.factorize(M.tocsc())
    ms = w.spmv_bench(200, False)
    out[tag] = {"us": ms * 1e3, "gbs_algorithmic": bytes_alg / (ms * 1e-3) / 1e9}
    w.close()


bench(A, "host numbering")
# keep the dense dofs (rows with > 64 entries) at the end: they are served by their own CTAs anyway
lens = np.diff(A.indptr)
sparse_rows = np.flatnonzero(lens <= 64)
sub = A[sparse_rows][:, sparse_rows]
perm_sub = reverse_cuthill_mckee(sp.csr_matrix(sub), symmetric_mode=True)
perm = np.concatenate([sparse_rows[perm_sub], np.flatnonzero(lens > 64)])
bench(A[perm][:, perm], "reverse Cuthill-McKee")
rnd = rng.permutation(A.shape[0])
bench(A[rnd][:, rnd], "random numbering")
print(json.dumps(out))
