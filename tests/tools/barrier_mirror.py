"""numpy mirror of the device barrier LMs (csrc/barrier.cuh, BarrierRun in csrc/capi_igm.cu): the NI stage with the field unknowns eliminated
per face (Schur complement) and the rounding-loop LM with the fixed dofs eliminated exactly (KKT) instead of penalised, both against the
oracle's golden vectors (tests/golden/*_barrier.npz). Development tool: it is how the algebra was validated before the kernels were written.
    python tests/tools/barrier_mirror.py [sphere_ico3 torus_24x16]"""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from conftest import load_golden  # noqa: E402
from oracle.barrier import InjectiveBarrier  # noqa: E402


def kkt_solve(H, C, rhs):
    """argmin 1/2 d^T H d - rhs^T d  s.t.  C d = 0 (the exact-elimination counterpart of the reference's 1e10 penalty rows)"""
    nc = C.shape[0]
    K = sp.bmat([[H, C.T], [C, None]], format="csc") if nc else H.tocsc()
    sol = spla.splu(K).solve(np.concatenate([rhs, np.zeros(nc)]))
    return sol[: H.shape[0]]


def independent_rows(C):
    if C.shape[0] == 0:
        return C
    Q, R, piv = __import__("scipy.linalg").linalg.qr(C.toarray().T, pivoting=True, mode="economic")
    rank = int((np.abs(np.diag(R)) > 1e-9 * abs(R[0, 0])).sum())
    return sp.csr_matrix(C.toarray()[np.sort(piv[:rank])])


def lm(x0, lam0, maxit, ftol, energy, direction, add, norm, post=None):
    prevx, lam, it = x0, lam0, 0
    while True:
        d = direction(prevx, lam)
        if norm(d) < ftol:
            return prevx, it
        prevE2, newE2 = energy(prevx), energy(add(prevx, d))
        x = add(prevx, d) if prevE2 > newE2 else prevx
        if prevE2 > newE2 and abs(prevE2 - newE2) < ftol:
            return x, it
        lam = lam / 10 if (prevE2 > newE2 and np.isfinite(newE2)) else lam * 10
        if post is not None and post(x):
            return x, it
        it += 1
        prevx = x
        if not it <= maxit:
            return x, it


def run(name):
    g = load_golden(name)
    gb = dict(np.load(os.path.join(ROOT, "tests/golden", name + "_barrier.npz")))
    N, F2, l = 4, g["F2"], float(g["length"])
    nF = len(F2) // (2 * N)
    UFull, X = g["UFull"].tocsr(), (g["X2F"] * l).tocsr()
    cols = X.shape[1]
    A0 = (X.T @ X).tocsr()
    dA0 = A0.diagonal()
    bar = InjectiveBarrier(N, nF, F2)
    xs, _ = oracle.naive_poisson(g["X2F"], UFull, F2, g["fixedIdcs"], g["fixedVals"], l)
    Cfix = UFull[g["fixedIdcs"], :]
    csq = np.asarray(Cfix.multiply(Cfix).sum(axis=0)).ravel()
    Cind = independent_rows(Cfix)

    def G_of(f):
        fB, barDer, S = bar.update(f, True)
        return fB, (sp.diags(barDer) @ bar.image_field(S, len(f))).tocsr()

    # ---- NI: unknown (x, f), f eliminated per face
    def energy_ni(xf):
        x, f = xf
        return 1e8 * np.sum((f - X @ x) ** 2) + np.sum((f - F2) ** 2) + 1e-8 * np.sum(bar.update(f, False)[0] ** 2)

    def dir_ni(xf, lam):
        x, f = xf
        fB, G = G_of(f)
        r = f - X @ x
        gx, gf = -1e8 * (X.T @ r), 1e8 * r + (f - F2) + G.T @ fB
        M = (G.T @ G).toarray() if nF * 8 <= 4000 else None
        GtG = (G.T @ G).tocsr()
        blocks = np.zeros((nF, 8, 8))
        coo = GtG.tocoo()
        blocks[coo.row // 8, coo.row % 8, coo.col % 8] = coo.data
        A = 1e8 * blocks
        idx = np.arange(8)
        A[:, idx, idx] *= (1.0 + lam)
        k0, r0 = (1e8 + 1) * (1 + lam), 1 + lam * (1e8 + 1)
        K = A + k0 * np.eye(8)
        Kinv = np.linalg.inv(K)
        W = 1e8 * np.einsum("fij,fjk->fik", A + r0 * np.eye(8), Kinv)
        bd = lambda B: sp.block_diag([b for b in B], format="csr") if nF <= 3000 else sp.bsr_matrix((B, np.arange(nF), np.arange(nF + 1)), shape=(8 * nF, 8 * nF)).tocsr()
        Wm, Kim = bd(W), bd(Kinv)
        S = (X.T @ Wm @ X + lam * sp.diags(1e8 * dA0 + 1e8 * csq)).tocsc()
        rhs = -gx - 1e8 * (X.T @ (Kim @ gf))
        dx = kkt_solve(S, Cind, rhs)
        df = Kim @ (-gf + 1e8 * (X @ dx))
        return dx, df

    add2 = lambda a, d: (a[0] + d[0], a[1] + d[1])
    nrm2 = lambda d: np.sqrt(d[0] @ d[0] + d[1] @ d[1])
    t = time.time()
    (xni, fni), it = lm((xs.copy(), F2.copy()), 0.01, 100, 10e-10, energy_ni, dir_ni, add2, nrm2)
    print(f"{name}: NI LM iterations {it}, |x_ni - golden| {np.abs(xni - gb['x_ni']).max():.2e}, |F2_ni - golden| {np.abs(fni - gb['F2_ni']).max():.2e}  ({time.time()-t:.1f}s)")

    # ---- rounding loop with exact elimination
    F2_, xCurr, xInit = fni, xni.copy(), xni.copy()
    integerIdcs, singularIdcs = [int(v) for v in g["integerIdcs"]], [int(v) for v in g["singularIdcs"]]
    rs = bool(g["round_seams"])
    left, roundedS = list(integerIdcs if rs else singularIdcs), False
    fixed_i, fixed_v, nsolves = [], [], 0
    while left:
        xPrev = xCurr.copy()
        Xf = UFull @ xCurr
        dist = [abs(Xf[li] - np.round(Xf[li])) for li in left]
        imin = int(np.argmin(dist))
        dmin = dist[imin]
        idx = left.pop(imin)
        fixed_i.append(idx)
        fixed_v.append(int(np.round(Xf[idx])))
        if not left and not rs and not roundedS:
            left, roundedS = list(integerIdcs), True
        if not dmin > 10e-7:
            continue
        Cf = UFull[np.asarray(fixed_i), :]
        fv = np.asarray(fixed_v, float)
        Ci = independent_rows(Cf)
        csq2 = np.asarray(Cf.multiply(Cf).sum(axis=0)).ravel()
        x0 = xInit + np.linalg.lstsq(Cf.toarray(), fv - Cf @ xInit, rcond=None)[0]     # onto the constraint set (least-norm move)

        def energy_si(x):
            fld = X @ x
            return np.sum((fld - F2_) ** 2) + 1e-4 * np.sum((x - xPrev) ** 2) + 1e-8 * np.sum(bar.update(fld, False)[0] ** 2)

        def dir_si(x, lam):
            fld = X @ x
            fB, G = G_of(fld)
            grad = X.T @ ((fld - F2_) + 1e-8 * (G.T @ fB)) + 1e-4 * (x - xPrev)
            Jb = (G @ X).tocsr()
            D = dA0 + 1e-4 + 1e-8 * np.asarray(Jb.multiply(Jb).sum(axis=0)).ravel() + 1e10 * csq2
            H = (A0 + 1e-4 * sp.identity(cols) + 1e-8 * (Jb.T @ Jb) + lam * sp.diags(D)).tocsc()
            return kkt_solve(H, Ci, -grad)

        xl, it = lm(x0, 0.01, 100, 1e-7, energy_si, dir_si, lambda a, d: a + d, np.linalg.norm)
        nsolves += 1
        xCurr, xInit = xl.copy(), xl.copy()
    fullx = UFull @ xCurr
    same = dict(zip(fixed_i, fixed_v)) == dict(zip(gb["out_fixedIdcs"].tolist(), gb["out_fixedVals"].tolist()))
    print(f"{name}: rounding loop {nsolves} LM solves (golden {int(gb['n_solves'])}), integers identical {same}, |fullx - golden| {np.abs(fullx - gb['fullx']).max():.2e}  ({time.time()-t:.1f}s)")


if __name__ == "__main__":
    for nm in sys.argv[1:] or ["sphere_ico3", "torus_24x16"]:
        run(nm)
