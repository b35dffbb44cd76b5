"""ctypes binding of oracle/_ref/libmetriko_ref.so: the reference's OWN headers (hmesh.h, face_rosy_field.h, parameterization.h,
iter_rounding*.h, levenberg_marquardt.h, injective_barrier.h, sparse_linear_solver.h -- included in place from /root/reference) compiled against the Eigen / libigl stand-ins of
oracle/ref_shim/ by oracle/Makefile. TEST INFRASTRUCTURE like the rest of oracle/: only tests/ (and the fixture generator
tests/golden/make_golden_ref.py) import this module. /root/reference does not exist on the GPU box and the library is built from it, so
everything here is used in this container only; what travels are the fixtures tests/golden/ref_*.npz it produced.

The direct solvers behind Eigen::SparseLU / SimplicialLLT are hooks (eigen_standin.hpp) filled here with scipy's SuperLU -- the solver
the restated oracle uses too, so a difference between the two is a difference in the restated LOGIC (loop semantics, weights, index
conventions, tolerances), which is what this pin is for."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libmetriko_ref.so")
REFERENCE = os.environ.get("METRIKO_REFERENCE", "/root/reference")

_FACTOR = C.CFUNCTYPE(C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double))
_SOLVE = C.CFUNCTYPE(C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double))
_RELEASE = C.CFUNCTYPE(None, C.c_int)
_ITER_CB = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_double), C.c_long)


def build() -> bool:
    """make -C oracle (needs the reference's sources); True when the library exists afterwards."""
    if os.path.isdir(os.path.join(REFERENCE, "include", "metriko")):
        subprocess.run(["make", "-C", _HERE, f"REF={REFERENCE}"], check=True, capture_output=True)
        if os.path.exists(os.path.join(_HERE, "..", "metriko_b200", "libmetriko_b200.so")):
            subprocess.run(["make", "-C", _HERE, f"REF={REFERENCE}", "overlay"], check=False, capture_output=True)  # optional: needs a GPU to run
    return os.path.exists(LIB_PATH)


def available() -> bool:
    return os.path.exists(LIB_PATH)


class Ref:
    def __init__(self):
        self.L = C.CDLL(LIB_PATH)
        self._factors: dict[int, tuple] = {}
        self._next = 0
        self.factor_log: list[tuple[int, int]] = []  # (kind, n) of every factorisation requested

        def factor(kind, is_complex, n, outer, inner, values):
            try:
                op = np.ctypeslib.as_array(outer, (n + 1,)).copy()
                nnz = int(op[n])
                ii = np.ctypeslib.as_array(inner, (nnz,)).copy()
                if is_complex:
                    vv = np.ctypeslib.as_array(values, (2 * nnz,)).copy().view(np.complex128)
                else:
                    vv = np.ctypeslib.as_array(values, (nnz,)).copy()
                A = sp.csc_matrix((vv, ii, op), shape=(n, n))
                lu = spla.splu(A)
                h = self._next
                self._next += 1
                self._factors[h] = (lu, n, bool(is_complex))
                self.factor_log.append((int(kind), int(n)))
                return h
            except Exception:  # singular matrix etc.: Eigen's info() != Success
                return -1

        def solve(h, rhs, x):
            lu, n, cplx = self._factors[h]
            if cplx:
                b = np.ctypeslib.as_array(rhs, (2 * n,)).view(np.complex128)
                np.ctypeslib.as_array(x, (2 * n,))[:] = lu.solve(b.copy()).view(np.float64)
            else:
                b = np.ctypeslib.as_array(rhs, (n,))
                np.ctypeslib.as_array(x, (n,))[:] = lu.solve(b.copy())
            return 0

        def release(h):
            self._factors.pop(h, None)

        self._cbs = (_FACTOR(factor), _SOLVE(solve), _RELEASE(release))  # keep the thunks alive
        self.L.ref_set_solver_hooks(*self._cbs)
        self.L.ref_factor_count.restype = C.c_long

    # ---- helpers
    @staticmethod
    def _coo(M):
        M = sp.coo_matrix(M)
        return (int(M.shape[0]), int(M.shape[1]), C.c_long(M.nnz), _ip(M.row.astype(np.int32)), _ip(M.col.astype(np.int32)),
                _dp(M.data.astype(np.float64)))

    def iterative_rounding(self, fixedIdcs, fixedVals, singularIdcs, integerIdcs, length, Cfull, G2p, N, n, nF, seamless, roundSeams,
                           localInjectivity, F2, iter_cb=None, verbose=False):
        """iter_rounding.h:33-113. Returns dict(success, fullx, factorizations)."""
        fi = np.ascontiguousarray(fixedIdcs, np.int32)
        fv = np.ascontiguousarray(fixedVals, np.float64)
        si = np.ascontiguousarray(singularIdcs, np.int32)
        ii = np.ascontiguousarray(integerIdcs, np.int32)
        F2 = np.ascontiguousarray(F2, np.float64)
        assert F2.size == 2 * N * nF
        cap = int(Cfull.shape[1]) + 8
        fullx = np.zeros(cap)
        flen = C.c_long(0)
        cb = _ITER_CB(lambda p, ln: int(bool(iter_cb(np.ctypeslib.as_array(p, (ln,)).copy())))) if iter_cb is not None else _ITER_CB()
        self.L.ref_factor_count(1)
        self.factor_log.clear()
        ok = self.L.ref_iterative_rounding(
            len(fi), _ip(fi), _dp(fv), len(si), _ip(si), len(ii), _ip(ii), C.c_double(length), *self._coo(Cfull), *self._coo(G2p),
            int(N), int(n), int(nF), int(bool(seamless)), int(bool(roundSeams)), int(bool(localInjectivity)), int(bool(verbose)), _dp(F2), cb,
            C.c_long(cap), _dp(fullx), C.byref(flen))
        return dict(success=bool(ok), fullx=fullx[: flen.value].copy(), factorizations=int(self.L.ref_factor_count(0)),
                    factor_log=list(self.factor_log))

    def naive_integration(self, fixedIdcs, fixedVals, length, Cfull, G2p, N, n, nF, F2):
        """NaiveIntegration's constructor (ir_init.h:74-232): UFull, X2F = (G2 UFull).pruned(), x0, XSmall."""
        fi = np.ascontiguousarray(fixedIdcs, np.int32)
        fv = np.ascontiguousarray(fixedVals, np.float64)
        F2 = np.ascontiguousarray(F2, np.float64)
        dims = (C.c_long * 5)()
        args = (len(fi), _ip(fi), _dp(fv), C.c_double(length), *self._coo(Cfull), *self._coo(G2p), int(N), int(n), int(nF), _dp(F2), dims)
        null_i, null_d = C.POINTER(C.c_int)(), C.POINTER(C.c_double)()
        self.L.ref_naive_integration(*args, C.c_long(0), null_i, null_i, null_d, C.c_long(0), null_i, null_i, null_d, null_d, null_d)
        ur, uc, un, xr, xn = (int(v) for v in dims)
        Ur, Uc, Uv = np.zeros(un, np.int32), np.zeros(un, np.int32), np.zeros(un)
        Xr, Xc, Xv = np.zeros(xn, np.int32), np.zeros(xn, np.int32), np.zeros(xn)
        x0, xs = np.zeros(ur), np.zeros(uc)
        ok = self.L.ref_naive_integration(*args, C.c_long(un), _ip(Ur), _ip(Uc), _dp(Uv), C.c_long(xn), _ip(Xr), _ip(Xc), _dp(Xv), _dp(x0), _dp(xs))
        return dict(ok=bool(ok), UFull=sp.csr_matrix((Uv, (Ur, Uc)), shape=(ur, uc)), X2F=sp.csr_matrix((Xv, (Xr, Xc)), shape=(xr, uc)), x0=x0,
                    XSmall=xs)

    def barrier_update(self, N, nF, F2, field, updateJ=True):
        """InjectiveBarrier::update (injective_barrier.h:54-109): fBarrier, diag(gBarrier), gen_image_filed as 4 (col, value) pairs per row."""
        F2 = np.ascontiguousarray(F2, np.float64)
        field = np.ascontiguousarray(field, np.float64)
        fB, gD = np.zeros(N * nF), np.zeros(N * nF)
        cols, vals = np.zeros(4 * N * nF, np.int32), np.zeros(4 * N * nF)
        self.L.ref_barrier_update(int(N), int(nF), _dp(F2), _dp(field), int(bool(updateJ)), _dp(fB), _dp(gD), _ip(cols), _dp(vals))
        return fB, gD, cols.reshape(-1, 4), vals.reshape(-1, 4)

    def solve_smallest_eig(self, L, M, u0, n_iter=50):
        """solveSmallestEig(SprsC& L, SprsC& M, nIter) (sls.h:29-54, SparseLU branch) with VecXc::Random := u0."""
        L = sp.coo_matrix(L).astype(np.complex128)
        M = sp.coo_matrix(M).astype(np.complex128)
        n = L.shape[0]
        u0 = np.ascontiguousarray(u0, np.complex128)
        x = np.zeros(n, np.complex128)

        def coo_c(A):
            return (C.c_long(A.nnz), _ip(A.row.astype(np.int32)), _ip(A.col.astype(np.int32)), _dp(np.ascontiguousarray(A.data).view(np.float64)))

        self.L.ref_solve_smallest_eig(int(n), *coo_c(L), *coo_c(M), _dp(u0.view(np.float64)), int(n_iter), _dp(x.view(np.float64)))
        return x


    _INT = ("twin", "next", "prev", "face", "tail", "head", "edge", "crnr", "face2half", "vert2half", "edge2face", "edge2vert", "face2edge", "matching",
            "singular", "cmatching", "dims", "L_rows", "L_cols", "M_rows", "cutF", "fixedIdcs", "integerIdcs", "singularIdcs", "he2matching", "he2transidx",
            "vtrans2cut_shape", "vtrans2cut_rows", "vtrans2cut_cols", "uncompress_shape", "uncompress_rows", "uncompress_cols", "constraint_shape",
            "constraint_rows", "constraint_cols")
    _CPLX = ("L_vals", "M_vals", "compressed", "field", "connection", "cfield", "principalCurvatureDir")

    def face_rosy_field(self, V, F, N, u0, matching="curl", seam=None, names=None, parameterize=None, field_type="smoothest"):
        """Hmesh(V, F) (hmesh.h:313-569) + FaceRosyField(mesh, N, Smoothest) (frf.h:17-44) + computeMatching (frf.h:100-202) [+ the combed field
        (frf.h:207-249) when a seam is given] run by the reference's own code; returns the requested arrays (default: all).
        parameterize = dict(gridscale, seamless, localInjectivity, roundSeams): the rest of example_igm.cpp:17-33 as well (compute_cut_mesh,
        RosyParameterization, setup(), integ()): adds integ_ok, fullx, nfn, cfn, the index lists and setup()'s matrices."""
        V = np.ascontiguousarray(V, np.float64).reshape(-1, 3)
        F = np.ascontiguousarray(F, np.int32).reshape(-1, 3)
        u0 = np.ascontiguousarray(u0, np.complex128)
        assert u0.size == F.shape[0]
        self.L.ref_field_create2.restype = C.c_void_p
        self.L.ref_field_get.restype = C.c_long
        sm = None if seam is None else np.ascontiguousarray(seam, np.int32)
        h = C.c_void_p(self.L.ref_field_create2(int(V.shape[0]), _dp(V.reshape(-1)), int(F.shape[0]), _ip(F.reshape(-1)), int(N),
                                                _dp(u0.view(np.float64)), 1 if matching == "curl" else 0,
                                                _ip(sm) if sm is not None else C.POINTER(C.c_int)(), 2 if field_type == "curvature" else 1))
        all_names = ("dims", "twin", "next", "prev", "face", "tail", "head", "edge", "crnr", "face2half", "vert2half", "heArgOnFace", "faceArea",
                     "angleDefect", "dihedralArg", "heLen", "faceBasisX", "faceBasisY", "faceNormal", "edge2face", "edge2vert", "face2edge", "L_rows",
                     "L_cols", "L_vals", "M_rows", "M_vals", "compressed", "field", "connection", "matching", "singular", "ext") + \
                    (("cfield", "cmatching") if seam is not None else ())
        out = {}
        if parameterize is not None:
            assert seam is not None
            out["integ_ok"] = int(self.L.ref_field_parameterize(h, C.c_double(parameterize.get("gridscale", 0.02)), int(parameterize.get("seamless", True)),
                                                               int(parameterize.get("localInjectivity", False)), int(parameterize.get("roundSeams", False))))
            all_names += ("cext", "cutF", "cutPos", "fullx", "nfn", "cfn", "fixedIdcs", "fixedVals", "integerIdcs", "singularIdcs", "he2matching",
                          "he2transidx") + tuple(f"{m}_{w}" for m in ("vtrans2cut", "uncompress", "constraint") for w in ("shape", "rows", "cols", "vals"))
        try:
            for k in (names or all_names):
                n = self.L.ref_field_get(h, k.encode(), C.c_long(0), C.POINTER(C.c_double)())
                assert n >= 0, k
                buf = np.zeros(n)
                self.L.ref_field_get(h, k.encode(), C.c_long(n), _dp(buf))
                if k in self._INT:
                    buf = buf.astype(np.int64)
                elif k in self._CPLX:
                    buf = buf.view(np.complex128)
                out[k] = buf
        finally:
            self.L.ref_field_destroy(h)
        return out


_keep: list = []


def _ip(a):
    a = np.ascontiguousarray(a, np.int32)
    _keep.append(a)
    if len(_keep) > 256:
        del _keep[:128]
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _dp(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    _keep.append(a)
    if len(_keep) > 256:
        del _keep[:128]
    return a.ctypes.data_as(C.POINTER(C.c_double))
