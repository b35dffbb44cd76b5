"""Oracle: Integer-Grid-Map Poisson solve + iterative rounding. Restates
include/metriko/core/igm/iterative_rounding/{iter_rounding.h:33-113, iter_rounding_init.h:60-72,173-232,
iter_rounding_loop.h:29-124, iter_rounding_common.h:60-94} and solver/levenberg_marquardt.h:13-231
with scipy's SuperLU standing in for Eigen::SparseLU / SimplicialLLT (test infrastructure, see
oracle/__init__.py). `UFull` (ir_init.h:98-154, dense FullPivLU kernel of the node constraints) is taken
as an input: it is host-side setup shared verbatim by the oracle and the CUDA path."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


class _Factor:
    """CholeskyWrapper (ir.h:14-20): factorize(A) + solve(rhs). SuperLU on the SPD normal matrix."""

    def __init__(self):
        self.lu = None
        self.count = 0

    def factorize(self, A) -> bool:
        self.lu = spla.splu(sp.csc_matrix(A))
        self.count += 1
        return True

    def solve(self, rhs):
        return self.lu.solve(rhs)


def lm_solve(objective, jacobian, x0, lam0=0.0, max_iterations=100, func_tol=10e-10, foo_tol=10e-10, post_iteration=None,
             stats=None, reuse_factor=False):
    """LMSolver::solve (lm.h:140-231) with DiagonalDamping (lm.h:13-88), verbose = false.

    objective(x) -> E vector; jacobian(x) -> sparse J. Quirks kept: the first-order-optimality test never
    breaks when not verbose (lm.h:171-177); `|direction| < funcTol` returns True with x = prevx
    (lm.h:193-197); a small energy change breaks and returns False (lm.h:208-211); lambda * 10 stays 0
    when lam0 = 0 so a non-decreasing step re-solves the same system until maxIterations (lm.h:58-62).

    reuse_factor = True skips a factorisation whose input is BIT-IDENTICAL to the previous one of this call: with lambda = 0 and a
    constant Jacobian (the same matrix object, linear problem) dampJ^T dampJ does not change, SuperLU is deterministic, so the factor --
    and therefore every iterate -- is the same to the last bit (tests/test_oracle.py asserts it). It only removes redundant work; it is
    what makes the oracle affordable at the BASELINE sizes (tests/golden/make_golden_large.py). `factorizations` still counts the
    factorisations the reference would perform; `factorizations_done` the ones actually run."""
    LS = _Factor()
    lam = lam0
    prevx = np.array(x0, np.float64)
    x = prevx.copy()

    def damp(J):
        # DiagonalDamping::init/update: dampJ = [J; diag(sqrt(lam * colsum(J.^2)))]
        d = lam * np.asarray(J.multiply(J).sum(axis=0)).ravel()
        return sp.vstack([J, sp.diags(np.sqrt(d))]).tocsc()

    E = objective(prevx)
    J = jacobian(prevx)
    dampJ = damp(J)
    it = 0
    ret = False
    factored_J = None  # the Jacobian object the current factor was computed from (only tracked while lambda == 0)
    skipped = 0
    while True:
        rhs = -(J.T @ E)
        if np.max(np.abs(rhs)) < foo_tol:
            x = prevx.copy()  # no break unless verbose
        if reuse_factor and lam == 0.0 and factored_J is J and LS.lu is not None:
            skipped += 1
        elif not LS.factorize((dampJ.T @ dampJ).tocsc()):
            ret = False
            break
        else:
            factored_J = J if lam == 0.0 else None
        direction = LS.solve(rhs)
        if np.linalg.norm(direction) < func_tol:
            x = prevx.copy()
            ret = True
            break
        prevE2 = float(np.sum(objective(prevx) ** 2))
        newE2 = float(np.sum(objective(prevx + direction) ** 2))
        if prevE2 > newE2:
            x = prevx + direction
            if abs(prevE2 - newE2) < func_tol:
                ret = False
                break
        else:
            x = prevx.copy()
        E = objective(x)
        J = jacobian(x)
        # DT->update (lm.h:43-83): re-evaluates both energies, lambda /= 10 on progress else *= 10
        if prevE2 > newE2 and np.isfinite(newE2):
            lam /= 10.0
        else:
            lam *= 10.0
        dampJ = damp(J)
        if post_iteration is not None and post_iteration(x):
            ret = True
            break
        it += 1
        prevx = x.copy()
        if not (it <= max_iterations):
            ret = False
            break
    if stats is not None:
        stats["factorizations"] = stats.get("factorizations", 0) + LS.count + skipped
        stats["factorizations_done"] = stats.get("factorizations_done", 0) + LS.count
        stats["lm_iterations"] = stats.get("lm_iterations", 0) + it
    return x, ret


def naive_poisson(X2F, UFull, F2, fixedIdcs, fixedVals, length):
    """NaiveIntegration ctor, ir_init.h:173-212 with X2F = (G2 UFull).pruned() given: E = X2F^T X2F * length;
    f = X2F^T F2; KKT [[E, C^T], [C, 0]] with C = UFull.rows(fixedIdcs); one sparse LU solve."""
    UFull = sp.csr_matrix(UFull)
    X2F = sp.csr_matrix(X2F)
    E = (X2F.T @ X2F) * length
    f = X2F.T @ F2
    C = UFull[np.asarray(fixedIdcs), :]
    nc = C.shape[0]
    big = sp.bmat([[E, C.T], [C, sp.csr_matrix((nc, nc))]], format="csc")
    rhs = np.concatenate([f, np.asarray(fixedVals, np.float64)])
    sol = spla.splu(big).solve(rhs)
    return sol[: UFull.shape[1]], X2F


def iterative_rounding(fixedIdcs, fixedVals, singularIdcs, integerIdcs, length, UFull, G2p, N, n, nF, seamless,
                       roundSeams, F2, ni_mode="lm", stats=None):
    """ir.h:33-113 for localInjectivity = false, taking G2p = G2 * fullx2Nfunc as the reference does (integ.ipp:72)."""
    X2F = sp.csr_matrix(sp.csr_matrix(G2p) @ sp.csr_matrix(UFull))  # ir_init.h:173
    X2F.eliminate_zeros()
    return iterative_rounding_x2f(fixedIdcs, fixedVals, singularIdcs, integerIdcs, length, UFull, X2F, N, n, seamless, roundSeams, F2,
                                  ni_mode=ni_mode, stats=stats)


def iterative_rounding_x2f(fixedIdcs, fixedVals, singularIdcs, integerIdcs, length, UFull, X2F, N, n, seamless, roundSeams, F2,
                           ni_mode="lm", stats=None, localInjectivity=False, iter_cb=None, reuse_factor=False, progress=None, trace=None):
    """The body of ir.h:33-113. G2 enters the algorithm only through X2F = (G2 UFull).pruned() (ir_init.h:173,
    ir_loop.h:123; gInteg * UExt = [-length X2F | I], ir_init.h:40-55), so X2F is the operator input here.
    Returns dict(success, fullx, x_small, fixedIdcs, fixedVals, n_solves, min_margin, ...).

    ni_mode = "lm"     : the NI stage runs the reference's LM over [x_small; f] (ir.h:82-83, ir_init.h:60-72);
    ni_mode = "closed" : uses the closed form the linear LM converges to:
                         x = Poisson solution, f* = (1e8 * length * B x + F2) / (1e8 + 1). Agreement of the two
                         modes is ~1e-8 relative (the reference's own round-off floor) and is tested on small inputs.

    localInjectivity = True (example_qgp / example_qex flags): every LM carries the per-face injectivity barrier
    (oracle/barrier.py; iter_rounding_common.h:70-85, weight wBarrier = 1e-4, ir_init.h:87) and starts from lambda = 0.01
    (ir.h:57,96,102); the NI stage always runs the LM (there is no closed form) and stops as soon as iter_cb(fullx) is true
    (ir_init.h:28; the caller's locally_injectivity_check, integ.ipp:46-63); with !seamless the NI result is returned (ir.h:86-88).
    """
    if stats is None:
        stats = {}
    UFull = sp.csr_matrix(UFull)
    F2 = np.asarray(F2, np.float64)
    fixedIdcs = np.asarray(fixedIdcs)
    fixedVals = np.asarray(fixedVals, np.float64)
    cols = UFull.shape[1]
    nfull = UFull.shape[0]
    l1 = F2.shape[0]

    XSmall, X2F_nolen = naive_poisson(X2F, UFull, F2, fixedIdcs, fixedVals, length)
    x0_full = UFull @ XSmall
    out = dict(success=True, fixedIdcs=[], fixedVals=[], n_solves=0, min_margin=np.inf, x0_full=x0_full, x_poisson=XSmall)
    if not seamless and not localInjectivity:  # ir.h:78-81
        out.update(fullx=x0_full, x_small=XSmall)
        return out
    barrier = None
    wBarrier = 1e-4  # ir_init.h:87
    lam0 = 0.01 if localInjectivity else 0.0  # ir.h:57
    if localInjectivity:
        from .barrier import InjectiveBarrier
        barrier = InjectiveBarrier(N, l1 // (2 * N), F2)  # ir.h:60
        ni_mode = "lm"

    # ---- NI stage (Integration weights ir_init.h:87: wInteg = 10e3, wConst = 10e3, wClose = 1)
    wInteg, wConst, wClose = 10e3, 10e3, 1.0
    B = X2F_nolen  # G2 * UFull
    if ni_mode == "lm":
        UExt = sp.block_diag([UFull, sp.identity(l1)], format="csr")  # ir_init.h:214-223
        gInteg = sp.hstack([-length * B, sp.identity(l1)]).tocsr() * wInteg  # ([-length G2 | I] UExt), ir_init.h:30-58
        gClose = sp.hstack([sp.csr_matrix((l1, cols)), sp.identity(l1)]).tocsr() * wClose
        sel = sp.csr_matrix((np.ones(len(fixedIdcs)), (np.arange(len(fixedIdcs)), fixedIdcs)), shape=(len(fixedIdcs), nfull + l1))
        gConst = sel @ UExt * wConst
        Jni = sp.vstack([gInteg, gClose, gConst]).tocsr()

        def obj_ni(xf_):  # ir_init.h:60-72 + ir_common.h:60-94
            xf = UExt @ xf_
            x, f = xf[:nfull], xf[nfull:]
            fInteg = f - length * (B @ xf_[:cols])  # = f - length G2 x
            fClose = f - F2
            fConst = x[fixedIdcs] - fixedVals
            parts = [fInteg * wInteg, fClose * wClose, fConst * wConst]
            if barrier is not None:
                parts.append(barrier.update(f, False)[0] * wBarrier)
            return np.concatenate(parts)

        def jac_ni(xf_):
            if barrier is None:
                return Jni
            f = (UExt @ xf_)[nfull:]
            # loop = false: gBarrier * imageField * gClose * wInteg (sic, ir_common.h:79)
            return sp.vstack([Jni, barrier.jacobian_block(f, gClose, wInteg)]).tocsr()

        post = None
        if iter_cb is not None:
            post = lambda xf_: bool(iter_cb((UExt @ xf_)[:nfull]))  # ir_init.h:28
        xf, _ = lm_solve(obj_ni, jac_ni, np.concatenate([XSmall, F2]), lam0, 100, 10e-10, 10e-10, post_iteration=post, stats=stats,
                         reuse_factor=reuse_factor)
        X_ = xf[:cols]
        F2_ = xf[cols:]
        if not seamless:  # ir.h:86-88
            out.update(fullx=UFull @ X_, x_small=X_)
            return out
    else:
        X_ = XSmall
        F2_ = (1e8 * length * (B @ XSmall) + F2) / (1e8 + 1.0)
    out["x_ni"] = X_
    out["F2_ni"] = F2_

    # ---- SI stage (ir_loop.h:102-124: wInteg = 1, wConst = 10e4, wClose = 0.01; X2F = G2 UFull length)
    wInteg, wConst, wClose = 1.0, 10e4, 0.01
    X2F = (B * length).tocsr()
    X2F.eliminate_zeros()
    integerIdcs = [int(v) for v in integerIdcs]
    singularIdcs = [int(v) for v in singularIdcs]
    left = list(integerIdcs if roundSeams else singularIdcs)
    roundedSingulars = False
    fixed_i: list[int] = []
    fixed_v: list[int] = []
    xCurr = X_.copy()
    xInit = X_.copy()
    x_last = None
    success = True
    rounded = False
    gInteg = X2F * wInteg
    gClose = sp.identity(cols, format="csr") * wClose
    near_ties = []
    while left:
        # initFixedIndices (ir_loop.h:29-74)
        xPrev = xCurr.copy()
        X = UFull @ xCurr
        minDif, minIdx = 3276700.0, -1
        second = np.inf
        for i, li in enumerate(left):
            v = X[li]
            d = abs(v - np.round(v))
            if d < minDif:
                second = minDif
                minIdx, minDif = i, d
            elif d < second:
                second = d
        idx = left[minIdx]
        val = int(np.round(X[idx]))
        frac_part = abs(X[idx] - np.floor(X[idx]) - 0.5)
        fixed_i.append(idx)
        fixed_v.append(val)
        del left[minIdx]
        if trace is not None:  # every fixed dof, re-solved or not: (dof, value before rounding, distance, runner-up distance)
            trace.append((idx, float(X[idx]), float(minDif), float(second)))
        if not left and not roundSeams and not roundedSingulars:
            left = list(integerIdcs)
            roundedSingulars = True
        if not (minDif > 10e-7):
            continue  # ir.h:100
        out["min_margin"] = min(out["min_margin"], frac_part)
        near_ties.append((idx, float(minDif), float(second)))
        rounded = True
        Cfix = UFull[np.asarray(fixed_i), :]
        gConst = Cfix * wConst
        Jsi = sp.vstack([gInteg, gClose, gConst]).tocsr()
        fv = np.asarray(fixed_v, np.float64)
        fi = np.asarray(fixed_i)

        def obj_si(xs, xPrev=xPrev, fi=fi, fv=fv):  # ir_loop.h:77-86
            Xf = UFull @ xs
            field = X2F @ xs
            parts = [(field - F2_) * wInteg, (xs - xPrev) * wClose, (Xf[fi] - fv) * wConst]
            if barrier is not None:
                parts.append(barrier.update(field, False)[0] * wBarrier)
            return np.concatenate(parts)

        def jac_si(xs, Jsi=Jsi):
            if barrier is None:
                return Jsi
            # loop = true: gBarrier * imageField * X2F * wBarrier (ir_common.h:78)
            return sp.vstack([Jsi, barrier.jacobian_block(X2F @ xs, X2F, wBarrier)]).tocsr()

        x_last, _ = lm_solve(obj_si, jac_si, xInit, lam0, 100, 1e-7, 1e-7, stats=stats, reuse_factor=reuse_factor)  # lambda reset per rounding step, ir.h:102
        if progress is not None:
            progress(f"re-solve {out['n_solves'] + 1}: dof {idx} -> {val} (minDif {minDif:.3e}), {len(left)} left")
        out["n_solves"] += 1
        # post_checking (ir_loop.h:89-100)
        fullx = UFull @ x_last
        mx = max(abs(fullx[i] - np.round(fullx[i])) for i in fixed_i)
        xCurr = x_last.copy()
        xInit = x_last.copy()
        if not (mx <= 10e-7):
            success = False
            break
    out["success"] = success
    out["fixedIdcs"] = fixed_i
    out["fixedVals"] = fixed_v
    out["selection"] = near_ties
    if rounded:
        out["fullx"] = UFull @ x_last
        out["x_small"] = x_last
    else:
        out["fullx"] = x0_full  # ir.h:111
        out["x_small"] = XSmall
    return out
