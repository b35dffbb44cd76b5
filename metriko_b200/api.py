"""Host-side mirror of metriko's solver seams above the C-ABI (names and argument meaning follow the reference so the
parity tests read like calls into metriko itself):

  solveSmallestEig(L, M, nIter, u0)       solver/sparse_linear_solver.h:29      (B1)
  CholeskyWrapper.factorize / .solve      igm/iterative_rounding/iter_rounding.h:14-20 (B2, LMSolver's LinearSolver)
  iterative_rounding(...)                 igm/iterative_rounding/iter_rounding.h:33-113 (B4)
  FaceRosyField                           vectorfield/face_rosy_field.h:17-84   (assembly + solve + decode on device)

Everything computes on the GPU through libmetriko_b200.so; nothing here falls back to the CPU."""
from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.sparse as sp

from . import capi
from .capi import Context, MkError, Stats, as_c128, as_f64, as_i32, default_opts, ptr


def _hierarchy(ctx, fn, h):
    """[(rows, entries of A_l, entries of P_l)] per multigrid level, from mk_*_hierarchy."""
    cap = 16
    n = C.c_int()
    rows, na, npp = (np.zeros(cap, np.int64) for _ in range(3))
    ctx.check(fn(h, cap, C.byref(n), ptr(rows), ptr(na), ptr(npp)))
    return [(int(rows[l]), int(na[l]), int(npp[l])) for l in range(min(n.value, cap))]


def _profile_cycle(ctx, fn, h, reps):
    cap = 16
    n = C.c_int()
    us = np.zeros(cap, np.float64)
    ctx.check(fn(h, reps, cap, C.byref(n), ptr(us)))
    return us[: min(n.value, cap)].tolist()


class FaceRosyField:
    """FaceRosyField(mesh, N, FieldType::Smoothest) (frf.h:17-44): connectionLaplacian + galerkinMassMatrix are
    assembled on the device from the half-edge arrays, then solveSmallestEig runs there."""

    def __init__(self, ctx: Context, face2half, next_, twin, face, heArgOnFace, faceArea, N: int = 4):
        self.ctx = ctx
        self.N = N
        self.nF = len(faceArea)
        self.nH = len(twin)
        self.h = C.c_void_p()
        a = [as_i32(face2half), as_i32(next_), as_i32(twin), as_i32(face), as_f64(heArgOnFace), as_f64(faceArea)]
        ctx.check(ctx.L.mk_field_create(ctx.h, self.nF, self.nH, N, *[ptr(v) for v in a], C.byref(self.h)))
        ctx.check(ctx.L.mk_field_assemble(self.h))
        self.stats = None

    def assemble(self):
        """Re-run connectionLaplacian + galerkinMassMatrix assembly on the device (drops the cached hierarchy)."""
        self.ctx.check(self.ctx.L.mk_field_assemble(self.h))

    def operator(self):
        """(L, M) as scipy CSR / diagonal, copied back from the device (tests)."""
        rowptr = np.empty(self.nF + 1, np.int32)
        col = np.empty(4 * self.nF, np.int32)
        val = np.empty(4 * self.nF, np.complex128)
        mass = np.empty(self.nF, np.float64)
        self.ctx.check(self.ctx.L.mk_field_get_operator(self.h, ptr(rowptr), ptr(col), ptr(val), ptr(mass)))
        return sp.csr_matrix((val, col, rowptr), shape=(self.nF, self.nF)), mass

    def solveSmallestEig(self, u0, nIter: int = 50, out=None, **opts):
        """`out`: preallocated (e.g. pinned) complex128 array of length nF for the result."""
        u0 = as_c128(u0)
        x = out if out is not None else np.empty(self.nF, np.complex128)
        st = Stats()
        o = default_opts(**opts)
        rc = self.ctx.L.mk_field_solve(self.h, ptr(u0), nIter, C.byref(o), ptr(x), C.byref(st))
        self.stats = st.as_dict()
        self.ctx.check(rc, allow=(-3,))   # MK_ERR_NOT_CONVERGED: result written, stats["status"] / ["relres_max"] say how far it got
        return x

    def solveCurvatureAligned(self, heLen, dihedralArg, return_rhs: bool = False, **opts):
        """FieldType::CurvatureAligned (frf.h:25-33): principalCurvatureDir -> rhs -> solveSquare(L, rhs), all on the device."""
        ln, da = as_f64(heLen), as_f64(dihedralArg)
        assert len(ln) == self.nH and len(da) == self.nH
        x = np.empty(self.nF, np.complex128)
        rhs = np.empty(self.nF, np.complex128) if return_rhs else None
        st = Stats()
        o = default_opts(**opts)
        rc = self.ctx.L.mk_field_solve_curvature_aligned(self.h, ptr(ln), ptr(da), C.byref(o), ptr(rhs), ptr(x), C.byref(st))
        self.stats = st.as_dict()
        self.ctx.check(rc)
        return (x, rhs) if return_rhs else x

    def computeMatching(self, mesh: dict, matching: str = "curl"):
        """computeMatching(Curl | Principal) + computeSingular (frf.h:100-202) on the device, from the last solution. `mesh`: the Hmesh
        flat arrays pos, tail, head, edge2face, faceBasisX, faceBasisY, edge, prev, vert2half, angleDefect. Returns (matching, effort, singular)."""
        pos, bx, by = as_f64(mesh["pos"]).reshape(-1), as_f64(mesh["faceBasisX"]).reshape(-1), as_f64(mesh["faceBasisY"]).reshape(-1)
        tail, head, edge, prev = (as_i32(mesh[k]) for k in ("tail", "head", "edge", "prev"))
        e2f, v2h, dfc = as_i32(mesh["edge2face"]).reshape(-1), as_i32(mesh["vert2half"]), as_f64(mesh["angleDefect"])
        nV, nE = len(v2h), len(e2f) // 2
        m, eff, sing = np.empty(nE, np.int32), np.empty(nE, np.float64), np.empty(nV, np.int32)
        self.ctx.check(self.ctx.L.mk_field_matching(self.h, nV, nE, 1 if matching == "curl" else 0, ptr(pos), ptr(tail), ptr(head), ptr(e2f), ptr(bx),
                                                    ptr(by), ptr(edge), ptr(prev), ptr(v2h), ptr(dfc), ptr(m), ptr(eff), ptr(sing)))
        return m, eff, sing

    def solveSquare(self, rhs, **opts):
        """solveSquare(L, rhs) (sls.h:16-27) on the assembled operator."""
        b = as_c128(rhs)
        x = np.empty(self.nF, np.complex128)
        st = Stats()
        o = default_opts(**opts)
        rc = self.ctx.L.mk_field_solve_square(self.h, ptr(b), C.byref(o), ptr(x), C.byref(st))
        self.stats = st.as_dict()
        self.ctx.check(rc)
        return x

    def solve_dev(self, u0_dev_ptr: int, nIter: int = 50, x_dev_ptr: int | None = None, **opts):
        st = Stats()
        o = default_opts(**opts)
        rc = self.ctx.L.mk_field_solve_dev(self.h, C.c_void_p(u0_dev_ptr), nIter, C.byref(o),
                                           C.c_void_p(x_dev_ptr) if x_dev_ptr else None, C.byref(st))
        self.stats = st.as_dict()
        self.ctx.check(rc, allow=(-3,))
        return self.stats

    def decode(self):
        out = np.empty((self.nF, self.N), np.complex128)
        self.ctx.check(self.ctx.L.mk_field_decode(self.h, ptr(out)))
        return out

    def spmv(self, x):
        x = as_c128(x)
        y = np.empty_like(x)
        self.ctx.check(self.ctx.L.mk_field_spmv(self.h, ptr(x), ptr(y)))
        return y

    def spmv_bench(self, reps: int = 50, fused_dot: bool = False) -> float:
        ms = C.c_double()
        self.ctx.check(self.ctx.L.mk_field_spmv_bench(self.h, reps, int(fused_dot), C.byref(ms)))
        return ms.value

    def hierarchy(self):
        return _hierarchy(self.ctx, self.ctx.L.mk_field_hierarchy, self.h)

    def profile_cycle(self, reps: int = 100):
        """[us of one preconditioner application, us of the cycle from level 1, from level 2, ...] measured live (graph replays)."""
        return _profile_cycle(self.ctx, self.ctx.L.mk_field_profile_cycle, self.h, reps)

    def close(self):
        if self.h:
            self.ctx.L.mk_field_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def solveSmallestEig(ctx: Context, L, M, nIter: int = 50, u0=None, return_stats: bool = False, **opts):
    """Drop-in for `VecXc solveSmallestEig(SprsC& L, SprsC& M, int nIter = 50)` (sls.h:29): L, M scipy sparse
    (converted to compressed CSC exactly as Eigen holds them). `u0` replaces VecXc::Random (must be given)."""
    if u0 is None:
        raise ValueError("u0 must be given explicitly (the reference draws it from an unseeded std::rand)")
    Lc = sp.csc_matrix(L, dtype=np.complex128)
    Mc = sp.csc_matrix(M, dtype=np.complex128)
    Lc.sort_indices()
    Mc.sort_indices()
    n = Lc.shape[0]
    u0 = as_c128(u0)
    x = np.empty(n, np.complex128)
    st = Stats()
    o = default_opts(**opts)
    a = [as_i32(Lc.indptr), as_i32(Lc.indices), as_c128(Lc.data), as_i32(Mc.indptr), as_i32(Mc.indices), as_c128(Mc.data)]
    rc = ctx.L.mk_solve_smallest_eig_csc(ctx.h, n, *[ptr(v) for v in a], ptr(u0), nIter, C.byref(o), ptr(x), C.byref(st))
    ctx.check(rc, allow=(-3,))   # MK_ERR_NOT_CONVERGED (fp64 floor of large meshes): x is the best iterate, stats["relres_max"] says how far it got
    return (x, st.as_dict()) if return_stats else x


class CholeskyWrapper:
    """The LinearSolver duck type of LMSolver (lm.h:90): `factorize(A) -> bool`, `solve(rhs, x) -> bool`
    (ir.h:14-20). The 'factor' is a device-resident operator + multigrid hierarchy; solve is PCG to opts.tol."""

    def __init__(self, ctx: Context, **opts):
        self.ctx = ctx
        self.opts = opts
        self.h = C.c_void_p()
        self.n = 0
        self.stats = None

    def factorize(self, A) -> bool:
        self.close()
        Ac = sp.csc_matrix(A, dtype=np.float64)
        Ac.sort_indices()
        self.n = Ac.shape[0]
        o = default_opts(**self.opts)
        rc = self.ctx.L.mk_spd_factorize(self.ctx.h, self.n, ptr(as_i32(Ac.indptr)), ptr(as_i32(Ac.indices)), ptr(as_f64(Ac.data)),
                                         C.byref(o), C.byref(self.h))
        return rc == capi.MK_OK

    def solve(self, rhs, x_guess=None):
        rhs = as_f64(rhs)
        x = np.empty(self.n, np.float64)
        g = as_f64(x_guess) if x_guess is not None else None
        st = Stats()
        rc = self.ctx.L.mk_spd_solve(self.h, ptr(rhs), ptr(g), ptr(x), C.byref(st))
        self.stats = st.as_dict()
        self.ctx.check(rc)
        return x

    def spmv_bench(self, reps: int = 50, fused_dot: bool = False) -> float:
        ms = C.c_double()
        self.ctx.check(self.ctx.L.mk_spd_spmv_bench(self.h, reps, int(fused_dot), C.byref(ms)))
        return ms.value

    def close(self):
        if self.h:
            self.ctx.L.mk_spd_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class IgmProblem:
    """Device-resident IGM problem (mk_igm): A0 = length^2 X2F^T X2F assembled on the device, UFull, F2."""

    def __init__(self, ctx: Context, X2F, UFull, F2, length: float, raw=None, dev=None, **opts):
        """`raw` = (x2f_ptr i64, x2f_col i32, x2f_val f64, uf_ptr i64, uf_col i32, uf_val f64, F2 f64): the C-ABI arrays
        themselves (e.g. pinned buffers), used as they are; X2F / UFull then only give the shapes.
        `dev` = (x2f_ptr, x2f_col, x2f_val, F2) raw DEVICE pointers (ints) of the same arrays already resident in HBM
        (mk_igm_create_dev); `raw` must then be given too (its UFull arrays are read on the host)."""
        self.ctx = ctx
        if dev is not None:
            assert raw is not None
            a = list(raw)
            X, U = X2F, UFull
            self.rows, self.cols = X.shape
            self.ufull_rows = U.shape[0]
            self.h = C.c_void_p()
            o = default_opts(**opts)
            ctx.check(ctx.L.mk_igm_create_dev(ctx.h, self.rows, self.cols, C.c_void_p(dev[0]), C.c_void_p(dev[1]), C.c_void_p(dev[2]), self.ufull_rows,
                                              ptr(a[3]), ptr(a[4]), ptr(a[5]), C.c_void_p(dev[3]), float(length), C.byref(o), C.byref(self.h)))
            self.stats = None
            return
        if raw is None:
            X = sp.csr_matrix(X2F, dtype=np.float64)
            U = sp.csr_matrix(UFull, dtype=np.float64)
            X.sort_indices()
            U.sort_indices()
            a = [np.ascontiguousarray(X.indptr, np.int64), as_i32(X.indices), as_f64(X.data),
                 np.ascontiguousarray(U.indptr, np.int64), as_i32(U.indices), as_f64(U.data), as_f64(F2)]
        else:
            X, U = X2F, UFull
            a = list(raw)
            assert a[0].dtype == np.int64 and a[1].dtype == np.int32 and a[2].dtype == np.float64
            assert a[3].dtype == np.int64 and a[4].dtype == np.int32 and a[5].dtype == np.float64 and a[6].dtype == np.float64
        self.rows, self.cols = X.shape
        self.ufull_rows = U.shape[0]
        assert U.shape[1] == self.cols
        self.h = C.c_void_p()
        o = default_opts(**opts)
        ctx.check(ctx.L.mk_igm_create(ctx.h, self.rows, self.cols, ptr(a[0]), ptr(a[1]), ptr(a[2]), self.ufull_rows, ptr(a[3]), ptr(a[4]),
                                      ptr(a[5]), ptr(a[6]), float(length), C.byref(o), C.byref(self.h)))
        self.stats = None

    @classmethod
    def from_mesh(cls, ctx: Context, N: int, cutF, cutPos, fullx2Nfunc, UFull, length: float, F2=None, field: "FaceRosyField | None" = None,
                  comb_turns=None, raw=None, **opts):
        """RosyParameterization::integ() from the arrays it reads itself (mk_igm_create_from_mesh): cut-mesh corner table and positions,
        fullx2Nfunc, UFull and either F2 = rawF2Vec (host) or (field, comb_turns): the eigen-solve result still resident in `field` plus the
        per-face combing turns. G2 / F2 / X2F are assembled on the device. `raw` = (f2n_ptr i64, f2n_col i32, f2n_val f64, uf_ptr i64, uf_col
        i32, uf_val f64, cutF i32, cutPos f64[, F2 f64 or comb_turns i32]) passes prepared (e.g. pinned) buffers as they are."""
        self = cls.__new__(cls)
        self.ctx = ctx
        if raw is None:
            Fn = sp.csr_matrix(fullx2Nfunc, dtype=np.float64)
            U = sp.csr_matrix(UFull, dtype=np.float64)
            Fn.sort_indices()
            U.sort_indices()
            a = [np.ascontiguousarray(Fn.indptr, np.int64), as_i32(Fn.indices), as_f64(Fn.data),
                 np.ascontiguousarray(U.indptr, np.int64), as_i32(U.indices), as_f64(U.data), as_i32(np.asarray(cutF).reshape(-1)),
                 as_f64(np.asarray(cutPos).reshape(-1))]
            f2 = as_f64(F2) if F2 is not None else None
            turns = as_i32(comb_turns) if comb_turns is not None else None
        else:
            Fn, U = fullx2Nfunc, UFull
            a = list(raw[:8])
            f2 = raw[8] if F2 is not None else None
            turns = raw[8] if F2 is None else None
        nF, cut_nV = len(a[6]) // 3, len(a[7]) // 3
        assert Fn.shape[0] == N * cut_nV and U.shape[0] == Fn.shape[1]
        self.rows, self.cols = 2 * N * nF, U.shape[1]
        self.ufull_rows = U.shape[0]
        self.N, self.nF, self.cut_nV = N, nF, cut_nV
        self.h = C.c_void_p()
        o = default_opts(**opts)
        ctx.check(ctx.L.mk_igm_create_from_mesh(ctx.h, N, nF, cut_nV, ptr(a[6]), ptr(a[7]), Fn.shape[1], ptr(a[0]), ptr(a[1]), ptr(a[2]), U.shape[1],
                                                ptr(a[3]), ptr(a[4]), ptr(a[5]), ptr(f2), field.h if (f2 is None and field is not None) else None,
                                                ptr(turns), float(length), C.byref(o), C.byref(self.h)))
        self.stats = None
        return self

    def x2f(self):
        """(X2F, F2) assembled on the device (mk_igm_create_from_mesh handles), copied back for tests."""
        nnz = C.c_int64()
        self.ctx.check(self.ctx.L.mk_igm_get_x2f(self.h, C.byref(nnz), None, None, None, None))
        p64 = np.empty(self.rows + 1, np.int64)
        col = np.empty(nnz.value, np.int32)
        val = np.empty(nnz.value, np.float64)
        f2 = np.empty(self.rows, np.float64)
        self.ctx.check(self.ctx.L.mk_igm_get_x2f(self.h, C.byref(nnz), ptr(p64), ptr(col), ptr(val), ptr(f2)))
        return sp.csr_matrix((val, col, p64), shape=(self.rows, self.cols)), f2

    def scatter(self, fullx=None, want_cfn: bool = True, out=None):
        """Result scatter (integ.ipp:88-98) on the device: (nfn[cut_nV x N], cfn[nF x 3N]); fullx None = the last rounding result, still resident.
        out = (nfn, cfn) preallocated (e.g. pinned) float64 arrays of those shapes."""
        nfn = out[0] if out is not None else np.empty((self.cut_nV, self.N), np.float64)
        cfn = (out[1] if out is not None else np.empty((self.nF, 3 * self.N), np.float64)) if want_cfn else None
        fx = as_f64(fullx) if fullx is not None else None
        self.ctx.check(self.ctx.L.mk_igm_scatter(self.h, ptr(fx), ptr(nfn), ptr(cfn)))
        return nfn, cfn

    def matrix(self):
        nnz = C.c_int64()
        self.ctx.check(self.ctx.L.mk_igm_get_matrix(self.h, C.byref(nnz), None, None, None))
        rowptr = np.empty(self.cols + 1, np.int32)
        col = np.empty(nnz.value, np.int32)
        val = np.empty(nnz.value, np.float64)
        self.ctx.check(self.ctx.L.mk_igm_get_matrix(self.h, C.byref(nnz), ptr(rowptr), ptr(col), ptr(val)))
        return sp.csr_matrix((val, col, rowptr), shape=(self.cols, self.cols))

    def poisson(self, fixedIdcs, fixedVals):
        fi, fv = as_i32(fixedIdcs), as_f64(fixedVals)
        xs = np.empty(self.cols, np.float64)
        xf = np.empty(self.ufull_rows, np.float64)
        st = Stats()
        rc = self.ctx.L.mk_igm_poisson(self.h, len(fi), ptr(fi), ptr(fv), ptr(xs), ptr(xf), C.byref(st))
        self.stats = st.as_dict()
        self.ctx.check(rc)
        return xs, xf

    def round(self, fixedIdcs, fixedVals, singularIdcs, integerIdcs, seamless=True, roundSeams=False, localInjectivity=False, use_iter_cb=False):
        """iterative_rounding (ir.h:33-113). localInjectivity: the barrier LMs (mk_igm_round_injective); use_iter_cb: stop the NI LM as soon as the
        map is locally injective (from_mesh handles)."""
        fi, fv = as_i32(fixedIdcs), as_f64(fixedVals)
        si, ii = as_i32(singularIdcs), as_i32(integerIdcs)
        m = len(si) + len(ii)
        fullx = np.empty(self.ufull_rows, np.float64)
        out_i = np.zeros(max(m, 1), np.int32)
        out_v = np.zeros(max(m, 1), np.int32)
        resolve = np.zeros(max(m, 1), np.int32)
        ties = np.zeros(max(m, 1), np.int32)
        nfix = C.c_int()
        st = Stats()
        if localInjectivity:
            rc = self.ctx.L.mk_igm_round_injective(self.h, len(fi), ptr(fi), ptr(fv), len(si), ptr(si), len(ii), ptr(ii), int(seamless), int(roundSeams),
                                                   int(use_iter_cb), ptr(fullx), ptr(out_i), ptr(out_v), C.byref(nfix), ptr(resolve), ptr(ties), C.byref(st))
        else:
            rc = self.ctx.L.mk_igm_round(self.h, len(fi), ptr(fi), ptr(fv), len(si), ptr(si), len(ii), ptr(ii), int(seamless), int(roundSeams),
                                         ptr(fullx), ptr(out_i), ptr(out_v), C.byref(nfix), ptr(resolve), ptr(ties), C.byref(st))
        self.stats = st.as_dict()
        self.ctx.check(rc, allow=(-5,))
        k = nfix.value
        return dict(success=(rc == capi.MK_OK), fullx=fullx, fixedIdcs=out_i[:k].tolist(), fixedVals=out_v[:k].tolist(),
                    resolve_flags=resolve[:k].tolist(), tie_flags=ties[:k].tolist(), n_solves=int(resolve[:k].sum()), stats=self.stats)

    def spmv_bench(self, reps: int = 50, fused_dot: bool = False) -> float:
        ms = C.c_double()
        self.ctx.check(self.ctx.L.mk_igm_spmv_bench(self.h, reps, int(fused_dot), C.byref(ms)))
        return ms.value

    def hierarchy(self):
        return _hierarchy(self.ctx, self.ctx.L.mk_igm_hierarchy, self.h)

    def profile_cycle(self, reps: int = 100):
        return _profile_cycle(self.ctx, self.ctx.L.mk_igm_profile_cycle, self.h, reps)

    def close(self):
        if self.h:
            self.ctx.L.mk_igm_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def iterative_rounding(ctx: Context, fixedIdcs, fixedVals, singularIdcs, integerIdcs, length, UFull, X2F, N, n, nF, seamless,
                       roundSeams, intrinsicField, **opts):
    """Drop-in for `iterative_rounding(...)` (ir.h:33-51) with localInjectivity = false. Differences from the reference
    signature: UFull and X2F = (G2 UFull).pruned() are passed in (the host builds them, ir_init.h:98-173) instead of
    Cfull and G2; the result dict carries fullx plus the rounding trace instead of an out-parameter."""
    prob = IgmProblem(ctx, X2F, UFull, intrinsicField, length, **opts)
    try:
        return prob.round(fixedIdcs, fixedVals, singularIdcs, integerIdcs, seamless, roundSeams)
    finally:
        prob.close()
