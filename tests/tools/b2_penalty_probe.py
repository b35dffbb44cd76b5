"""GPU probe: seam B2 (CholeskyWrapper = mk_spd_factorize / mk_spd_solve) on the matrix the reference's LMSolver really hands to it
(lm.h:184-189): dampJ^T dampJ of a SeamlessIntegration step, J = [gInteg; gClose; gConst] with wConst = 10e4 (ir_loop.h:102-124), i.e.
1e10 penalty rows on top of X^T X + 1e-4 I (condition number ~1e14), and the right-hand side -J^T E of that step.
    python tests/tools/b2_penalty_probe.py [golden=torus_24x16] [n_fixed=6]"""
import json
import os
import sys

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(1, os.path.join(ROOT, "tests"))
from conftest import load_golden  # noqa: E402
from metriko_b200 import api, capi  # noqa: E402


def si_system(g, n_fixed):
    """J and E of the SI objective (ir_loop.h:77-86) after `n_fixed` dofs of the golden's rounding sequence were fixed, at x = the Poisson solution."""
    U = g["UFull"].tocsr()
    X = (g["X2F"] * float(g["length"])).tocsr()
    cols = X.shape[1]
    fi = np.asarray(g["out_fixedIdcs"][:n_fixed])
    fv = np.asarray(g["out_fixedVals"][:n_fixed], float)
    wInteg, wConst, wClose = 1.0, 10e4, 0.01
    J = sp.vstack([X * wInteg, sp.identity(cols) * wClose, U[fi, :] * wConst]).tocsr()
    x = spla.lsqr(U, g["x0_full"], atol=1e-14, btol=1e-14)[0]  # reduced Poisson solution
    E = np.concatenate([(X @ x - g["F2"]) * wInteg, np.zeros(cols), ((U @ x)[fi] - fv) * wConst])
    return J, E


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "torus_24x16"
    n_fixed = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    g = load_golden(name)
    J, E = si_system(g, n_fixed)
    A = (J.T @ J).tocsc()
    rhs = -(J.T @ E)
    d_ref = spla.splu(A).solve(rhs)
    ctx = capi.Context(0)
    for precond in (1, 0):
        w = api.CholeskyWrapper(ctx, precond=precond)
        ok = w.factorize(A)
        try:
            d = w.solve(rhs)
            err = None
        except api.MkError as e:
            d, err = None, repr(e)
        out = dict(golden=name, n=A.shape[0], n_fixed=n_fixed, precond=precond, factorize=ok, error=err, stats=w.stats,
                   diag_max=float(A.diagonal().max()), diag_min=float(A.diagonal().min()))
        if d is not None:
            out.update(relres=float(np.linalg.norm(A @ d - rhs) / np.linalg.norm(rhs)), max_diff_vs_lu=float(np.abs(d - d_ref).max()),
                       d_absmax=float(np.abs(d_ref).max()), lu_relres=float(np.linalg.norm(A @ d_ref - rhs) / np.linalg.norm(rhs)))
        print(json.dumps(out), flush=True)
        w.close()


if __name__ == "__main__":
    main()
