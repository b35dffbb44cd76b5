"""Development tool (CPU only): PCG iteration counts of one fresh field solve L x = M u0 (tol 1e-10) under variants of the multigrid cycle,
on the host mirror's hierarchy of an octahedron sphere: V(1,1) with Chebyshev degree 1 on level 0 and 2 below (the device default),
W-cycle from level 1 / level 2, more smoothing, and the two-grid bound (exact level-1 solve).
    python tests/tools/field_cycle_study.py 7
"""
import os, sys, time
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
from amg_mirror import VCycle, describe, pcg
from metriko_b200 import capi
from hostglue import Session


class Cycle(VCycle):
    def __init__(self, levels, deg0=1, degc=2, gamma_from=None, gamma=2, exact_from=None, ratio=8.0):
        super().__init__(levels, deg=deg0, ratio=ratio)
        self.deg0, self.degc, self.gamma_from, self.gamma, self.exact_from = deg0, degc, gamma_from, gamma, exact_from
        self.lu = spla.splu(sp.csc_matrix(self.A[exact_from])) if exact_from is not None else None

    def cycle(self, l, b):
        if self.exact_from is not None and l == self.exact_from:
            return self.lu.solve(b)
        if l == len(self.A) - 1:
            if self.cinv is None:
                self.cinv_lu = getattr(self, "cinv_lu", None) or spla.splu(sp.csc_matrix(self.A[-1]))
                return self.cinv_lu.solve(b)
            return self.cinv @ b
        self.deg = self.deg0 if l == 0 else self.degc
        x, r = self.smooth(l, b, None, True)
        rc = self.R[l] @ r
        xc = self.cycle(l + 1, rc)
        if self.gamma_from is not None and l + 1 >= self.gamma_from and l + 1 < len(self.A) - 1:
            for _ in range(self.gamma - 1):  # W-cycle: a second visit of the child on the child's residual
                xc = xc + self.cycle(l + 1, rc - self.A[l + 1] @ xc)
        x = x + self.P[l] @ xc
        self.deg = self.deg0 if l == 0 else self.degc
        x, _ = self.smooth(l, b, x, False)
        return x


lv = int(sys.argv[1]) if len(sys.argv) > 1 else 6
s = Session().make_sphere(8, lv, bump=0.08)
fi = s.field_inputs(); nF = s.dims["nF"]
L = oracle.connection_laplacian(fi["twin"], fi["face"], fi["heArgOnFace"], nF, 4).tocsr(); L.sort_indices()
M = oracle.galerkin_mass_matrix(fi["faceArea"])
b = M @ oracle.random_u0(nF, 0)
H = capi.amg_host_hierarchy(L)
print("faces", nF, "levels", describe(H), flush=True)
variants = [("V deg 1/2 (device default)", dict()), ("V deg 1/1", dict(degc=1)), ("V deg 1/3", dict(degc=3)), ("V deg 2/2", dict(deg0=2)), ("W from level 1", dict(gamma_from=1)),
            ("W from level 2", dict(gamma_from=2)), ("W from 1, deg 1/1", dict(gamma_from=1, degc=1)), ("two-grid (exact level 1)", dict(exact_from=1)),
            ("three-grid (exact level 2)", dict(exact_from=2))]
for name, kw in variants:
    V = Cycle(H, **kw)
    t0 = time.time()
    x, it = pcg(L, b, np.zeros(nF, complex), V, tol=1e-10)
    print(f"{name:32s} {it:4d} iterations  ({time.time() - t0:.0f}s)", flush=True)
