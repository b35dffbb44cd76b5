"""CPU mirror of the penalty-slack emulation of the rounding loop (csrc/capi_igm.cu `IgmRun::slack_*`).

The reference pins every rounded dof by a penalty row of weight rho = wConst^2 = 1e10 (ir_loop.h:102-124), so its iterate satisfies
C x = d only up to the slack  C x_rho - d = mu / rho + O(rho^-2),  mu = the Lagrange multipliers of the exactly constrained problem.
A candidate dof whose UFull row u lies in the row space of the rows already fixed (twins across a seam, dofs tied by a node constraint)
therefore does not sit ON an integer in the reference but  u . x_1 / rho  away from it,  u . x_1 = beta^T (T^T T)^-1 nu  with
beta = Q u, nu = Q (h - H x), T = C Q^T  (Q = orthonormal basis of the fixed rows).  `minDif > 10e-7` (ir.h:100) then decides whether the
reference re-solves -- and because every re-solve is a proximal step (wClose |x - xPrev|^2) the final fullx depends on that decision
(7e-4 at 1M vertices for ONE extra re-solve). The device eliminates constraints exactly, so it evaluates the decision on the emulated slack.

This script runs the oracle (penalty, SuperLU) and an exact-elimination loop with the emulated slack side by side on a small golden and
prints, for every fixed dof, both distances: they must agree to O(1e-9) wherever the dof is dependent.

    python tests/tools/slack_mirror.py [tests/golden/sphere_ico3.npz]
"""
import os
import sys

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

RHO = 1e10


def csr(g, name):
    return sp.csr_matrix((g[name + "_val"], g[name + "_col"], g[name + "_ptr"]), shape=tuple(g[name + "_shape"]))


def exact_loop(g, round_seams):
    """Exact elimination + emulated slack; returns the trace [(dof, value incl. slack, distance, resolved)] and fullx."""
    UFull, X2F0, F2 = csr(g, "UFull"), csr(g, "X2F"), g["F2"]
    length = float(g["length"])
    fixedIdcs, fixedVals = g["fixedIdcs"], g["fixedVals"]
    XSmall, B = oracle.igm.naive_poisson(X2F0, UFull, F2, fixedIdcs, fixedVals, length)
    F2_ = (1e8 * length * (B @ XSmall) + F2) / (1e8 + 1.0)
    X = (B * length).tocsr()
    H = (X.T @ X + 1e-4 * sp.identity(X.shape[1])).tocsc()
    base = X.T @ F2_
    left = [int(v) for v in (g["integerIdcs"] if round_seams else g["singularIdcs"])]
    rounded_singulars = False
    fixed_i, fixed_v = [], []
    Q = np.zeros((0, UFull.shape[1]))  # orthonormal rows (dense: small case)
    x = XSmall.copy()
    y = np.zeros(0)  # (T^T T)^-1 nu of the last solve
    x1 = None        # first-order term of EVERY dof: H x_1 + Q^T m = 0, Q x_1 = y  (IgmRun::compute_x1)
    trace = []
    n_solves = 0
    while left:
        vals = {}
        for li in left:
            u = UFull[li].toarray().ravel()
            v = u @ x
            if Q.shape[0]:
                beta = Q @ u
                if np.linalg.norm(u - Q.T @ beta) <= 1e-9 * np.linalg.norm(u) and len(y):
                    v += beta @ y / RHO  # the reference's penalty slack, first order
            vals[li] = v
        idx = min(left, key=lambda li: abs(vals[li] - np.round(vals[li])))
        d = abs(vals[idx] - np.round(vals[idx]))
        val = int(np.round(vals[idx]))
        left.remove(idx)
        if not left and not round_seams and not rounded_singulars:
            left = [int(v) for v in g["integerIdcs"]]
            rounded_singulars = True
        fixed_i.append(idx)
        fixed_v.append(val)
        u = UFull[idx].toarray().ravel()
        r = u - (Q.T @ (Q @ u) if Q.shape[0] else 0)
        r = r - (Q.T @ (Q @ r) if Q.shape[0] else 0)
        if np.linalg.norm(r) > 1e-9 * np.linalg.norm(u):
            Q = np.vstack([Q, r / np.linalg.norm(r)])
        resolve = d > 10e-7
        full = float(UFull[idx].toarray().ravel() @ x + (UFull[idx].toarray().ravel() @ x1 / RHO if x1 is not None else 0.0))
        trace.append((idx, float(vals[idx]), float(d), resolve, full))
        if not resolve:
            continue
        n_solves += 1
        xprev = x.copy()
        h = base + 1e-4 * xprev
        C = UFull[np.asarray(fixed_i), :]
        dvec = np.asarray(fixed_v, float)
        # exact constrained minimiser: x = x_p + Z z with Z = null space through Q (dense, small): project
        k = Q.shape[0]
        Qs = sp.csr_matrix(Q)
        # KKT with the orthonormal rows: [[H, Q^T], [Q, 0]] [x; m] = [h; Q-targets]; targets from a particular solution of C x = d
        xp = np.linalg.lstsq(C.toarray(), dvec, rcond=None)[0]
        big = sp.bmat([[H, Qs.T], [Qs, None]], format="csc")
        sol = spla.splu(big).solve(np.concatenate([h, Q @ xp]))
        x = sol[: H.shape[0]]
        res = h - H @ x  # = C^T mu, in the row space
        nu = Q @ res
        T = C.toarray() @ Q.T
        y = np.linalg.solve(T.T @ T, nu)
        sol1 = spla.splu(big).solve(np.concatenate([np.zeros(H.shape[0]), y]))
        x1 = sol1[: H.shape[0]]
    return trace, UFull @ x, n_solves


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "sphere_ico3.npz")
    g = dict(np.load(path))
    rs = bool(int(g["round_seams"]))
    otrace = []
    ref = oracle.iterative_rounding_x2f(g["fixedIdcs"], g["fixedVals"], g["singularIdcs"], g["integerIdcs"], float(g["length"]), csr(g, "UFull"),
                                        csr(g, "X2F"), int(g["N"]), int(g["n"]), True, rs, g["F2"], ni_mode="closed", reuse_factor=True, trace=otrace)
    etrace, fullx, n_solves = exact_loop(g, rs)
    print(f"oracle re-solves {ref['n_solves']}, exact + slack re-solves {n_solves}; fullx max diff {np.abs(fullx - ref['fullx']).max():.3e}")
    worst = 0.0
    worst_all_exact = worst_all_x1 = 0.0
    for (i, v, d, _), (j, w, e, res, full) in zip(otrace, etrace):
        tag = "re-solve" if res else "skipped "
        print(f"  dof {i:6d}/{j:6d} {tag} oracle value {v:+.10f} dist {d:.3e} | emulated {w:+.10f} dist {e:.3e} | with x_1: value diff {abs(v - full):.2e}")
        if i == j and not res:
            worst = max(worst, abs(d - e))
        if i == j:
            worst_all_exact = max(worst_all_exact, abs(v - w))
            worst_all_x1 = max(worst_all_x1, abs(v - full))
    print(f"worst |oracle distance - emulated distance| over skipped dofs: {worst:.3e}")
    print(f"worst |oracle value - device value| over ALL chosen dofs: {worst_all_exact:.3e} without x_1, {worst_all_x1:.3e} with the first-order term x_1 / rho")


if __name__ == "__main__":
    main()
