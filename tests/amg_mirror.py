"""CPU prototype of the device V-cycle + PCG (mirrors amg.cuh / solver.cuh) on hierarchies built by the host set-up code,
to study iteration counts without a GPU. Development tool only."""
import os, sys, time
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metriko_b200 import capi


class VCycle:
    def __init__(self, levels, A0=None, deg=2, ratio=8.0):
        self.lv = levels
        self.deg, self.ratio = deg, ratio
        self.A = [l[0] for l in levels]
        if A0 is not None:
            self.A[0] = A0
        self.P = [l[1] for l in levels]
        self.R = [p.conj().T.tocsr() if p is not None else None for p in self.P]
        self.dinv = [l[3] if len(l) > 3 else 1.0 / a.diagonal() for l, a in zip(levels, self.A)]
        self.lmax = [l[2] for l in levels]
        Ac = self.A[-1].toarray()
        self.cinv = np.linalg.inv(Ac) if Ac.shape[0] <= 1500 else None

    def smooth(self, l, b, x, zero):
        A, dinv = self.A[l], self.dinv[l]
        lmax = self.lmax[l]; lmin = lmax / self.ratio
        theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
        sigma = theta / delta; rho = 1.0 / sigma
        r = b.copy() if zero else b - A @ x
        d = dinv * r / theta
        if zero: x = np.zeros_like(b)
        for k in range(self.deg):
            x = x + d
            r = r - A @ d
            rho1 = 1.0 / (2 * sigma - rho)
            d = rho1 * rho * d + 2 * rho1 / delta * dinv * r
            rho = rho1
        return x, r

    def cycle(self, l, b):
        if l == len(self.A) - 1:
            if self.cinv is not None: return self.cinv @ b
            save = self.deg; self.deg = 8
            x, _ = self.smooth(l, b, None, True); self.deg = save
            return x
        x, r = self.smooth(l, b, None, True)
        xc = self.cycle(l + 1, self.R[l] @ r)
        x = x + self.P[l] @ xc
        x, _ = self.smooth(l, b, x, False)
        return x

    def __call__(self, r):
        return self.cycle(0, r)


def pcg(A, b, x0, M, tol=1e-10, maxit=5000, proj=None):
    x = x0.copy()
    if proj: b = proj(b)
    r = b - (proj(A @ x) if proj else A @ x)
    bn = np.linalg.norm(b)
    z = M(r) if M else r
    if proj: z = proj(z)
    p = z.copy(); rho = np.vdot(r, z).real
    hist = []
    for it in range(maxit):
        if np.linalg.norm(r) <= tol * bn: return x, it
        q = A @ p
        if proj: q = proj(q)
        alpha = rho / np.vdot(p, q).real
        x += alpha * p; r -= alpha * q
        z = M(r) if M else r
        if proj: z = proj(z)
        rho1 = np.vdot(r, z).real
        p = z + (rho1 / rho) * p; rho = rho1
    return x, maxit


def describe(levels):
    return " | ".join(f"{l[0].shape[0]}:{l[0].nnz/ l[0].shape[0]:.1f}" for l in levels)


def projected_pcg(A, b, x0, apply_prec, proj, tol=1e-10, maxit=2000):
    """PCG in the feasible subspace (x0 feasible, directions projected): mirror of Pcg<T>::iteration_body with a projector."""
    x = x0.copy(); b = proj(b)
    r = b - proj(A @ x); bn = np.linalg.norm(b)
    z = apply_prec(r); p = z.copy(); rho = r @ z
    for it in range(maxit):
        if np.linalg.norm(r) <= tol * bn: return x, it
        q = proj(A @ p); alpha = rho / (p @ q)
        x += alpha * p; r -= alpha * q
        z = apply_prec(r); rho1 = r @ z
        p = z + (rho1 / rho) * p; rho = rho1
    return x, maxit


def rounding_loop(A, f, UFull, pins, pin_vals, left, V, mode, X=None, replay=None):
    """Mirror of the device rounding loop (capi_igm.cu mk_igm_round, simplified: every candidate is rounded and re-solved) with the
    projected preconditioner P V P (mode "proj") or the constraint preconditioner V - W (Q W)^-1 Q V, W = V Q^T (mode "cprec",
    csrc/igm.cuh ConstraintPrecond). X (= length X2F): after the Poisson solve the field is replaced by the NI closed form
    f* ~ X x_poisson, as the real loop does. replay: a [(dof, integer)] sequence to impose instead of selecting (greedy rounding is path dependent and
    symmetric meshes have selection ties, so solver variants are compared on ONE sequence). Returns (iterations per solve, [(dof, integer)], x)."""
    n = A.shape[0]
    rows = [int(i) for i in pins]; vals = [float(v) for v in pin_vals]
    left = [int(v) for v in left]
    x = np.zeros(n); its = []; ints = []; first = True
    replay = list(replay) if replay is not None else None
    while True:
        C = UFull[np.asarray(rows), :].toarray()
        Qt, R = np.linalg.qr(C.T)                  # C^T = Qt R, orthonormal constraint rows Q = Qt^T
        Q = Qt.T
        t = np.linalg.solve(R.T, np.asarray(vals))
        proj = lambda v, Q=Q: v - Q.T @ (Q @ v)
        x0 = x - Q.T @ (Q @ x - t)
        b = f + (0.0 if first else 1e-4) * x
        if mode == "proj":
            prec = lambda r: proj(V(r))
        else:
            W = np.stack([V(Q[j]) for j in range(Q.shape[0])], axis=1)
            Sinv = np.linalg.inv(Q @ W)
            def prec(r, W=W, Sinv=Sinv, Q=Q):
                z = V(r)
                return proj(z - W @ (Sinv @ (Q @ z)))
        x, it = projected_pcg(A, b, x0, prec, proj)
        its.append(it)
        if first and X is not None: f = X.T @ (X @ x)
        first = False
        if replay is not None:
            if len(ints) == len(replay): break
            idx, val = replay[len(ints)]
        else:
            if not left: break
            Xf = UFull @ x
            d = np.abs(Xf[left] - np.round(Xf[left])); i = int(np.argmin(d))
            idx = left.pop(i); val = int(np.round(Xf[idx]))
        rows.append(int(idx)); vals.append(float(val)); ints.append((int(idx), int(val)))
    return its, ints, x
