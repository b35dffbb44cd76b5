"""ctypes binding of libmetriko_b200.so (include/metriko_b200.h). Thin: every function maps 1:1 to a C-ABI entry
point; numpy arrays are passed as plain pointers. There is no CPU fallback -- loading fails loudly when the CUDA
library has not been built, and every compute call fails when no sm_100a device is present."""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmetriko_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "metriko_b200.h")

MK_OK = 0
STATUS = {0: "MK_OK", -1: "MK_ERR_CUDA", -2: "MK_ERR_ARG", -3: "MK_ERR_NOT_CONVERGED", -4: "MK_ERR_NOT_SPD",
          -5: "MK_ERR_ROUNDING", -9: "MK_ERR_INTERNAL"}


class MkError(RuntimeError):
    def __init__(self, status: int, msg: str = ""):
        self.status = status
        super().__init__(f"{STATUS.get(status, status)}: {msg}")


class SolveOpts(C.Structure):
    _fields_ = [("tol", C.c_double), ("max_iter", C.c_int), ("precond", C.c_int), ("use_graph", C.c_int), ("recycle", C.c_int)]


class Stats(C.Structure):
    _fields_ = [("outer_iterations", C.c_int), ("inner_iterations", C.c_int64), ("inner_max", C.c_int),
                ("relres_max", C.c_double), ("lambda_", C.c_double), ("kernel_launches", C.c_int64),
                ("setup_ms", C.c_double), ("solve_ms", C.c_double), ("status", C.c_int)]

    def as_dict(self):
        return dict(outer_iterations=self.outer_iterations, inner_iterations=self.inner_iterations, inner_max=self.inner_max,
                    relres_max=self.relres_max, lambda_=self.lambda_, kernel_launches=self.kernel_launches,
                    setup_ms=self.setup_ms, solve_ms=self.solve_ms, status=self.status)


def declared_symbols() -> list[str]:
    """Every entry point include/metriko_b200.h declares (used by the CPU test that the library exports them all)."""
    src = open(HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mk_[a-z0-9_]+)\s*\(", src)))


_LIB = None


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback for the CUDA path)")
        L = C.CDLL(LIB_PATH)
        vp, ip, dp, i64p = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_int64)
        L.mk_default_opts.argtypes = [C.POINTER(SolveOpts)]
        L.mk_default_opts.restype = None
        L.mk_device_count.restype = C.c_int
        L.mk_ctx_create.argtypes = [C.c_int, vp, C.POINTER(vp)]
        L.mk_ctx_destroy.argtypes = [vp]
        L.mk_last_error.argtypes = [vp]
        L.mk_last_error.restype = C.c_char_p
        L.mk_ctx_sync.argtypes = [vp]
        L.mk_field_create.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, C.POINTER(vp)]
        L.mk_field_destroy.argtypes = [vp]
        L.mk_field_assemble.argtypes = [vp]
        L.mk_field_get_operator.argtypes = [vp, vp, vp, vp, vp]
        L.mk_field_solve.argtypes = [vp, vp, C.c_int, C.POINTER(SolveOpts), vp, C.POINTER(Stats)]
        L.mk_field_solve_dev.argtypes = [vp, vp, C.c_int, C.POINTER(SolveOpts), vp, C.POINTER(Stats)]
        L.mk_field_decode.argtypes = [vp, vp]
        L.mk_field_matching.argtypes = [vp, C.c_int, C.c_int, C.c_int] + [vp] * 13
        L.mk_field_solve_curvature_aligned.argtypes = [vp, vp, vp, C.POINTER(SolveOpts), vp, vp, C.POINTER(Stats)]
        L.mk_field_solve_square.argtypes = [vp, vp, C.POINTER(SolveOpts), vp, C.POINTER(Stats)]
        L.mk_field_spmv.argtypes = [vp, vp, vp]
        L.mk_field_spmv_bench.argtypes = [vp, C.c_int, C.c_int, dp]
        L.mk_field_hierarchy.argtypes = [vp, C.c_int, ip, vp, vp, vp]
        L.mk_field_profile_cycle.argtypes = [vp, C.c_int, C.c_int, ip, vp]
        L.mk_solve_smallest_eig_csc.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp, vp, C.c_int, C.POINTER(SolveOpts), vp, C.POINTER(Stats)]
        L.mk_spd_factorize.argtypes = [vp, C.c_int, vp, vp, vp, C.POINTER(SolveOpts), C.POINTER(vp)]
        L.mk_spd_solve.argtypes = [vp, vp, vp, vp, C.POINTER(Stats)]
        L.mk_spd_destroy.argtypes = [vp]
        L.mk_spd_spmv_bench.argtypes = [vp, C.c_int, C.c_int, dp]
        L.mk_igm_create.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, C.c_int, vp, vp, vp, vp, C.c_double, C.POINTER(SolveOpts), C.POINTER(vp)]
        L.mk_igm_create_dev.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, C.c_int, vp, vp, vp, vp, C.c_double, C.POINTER(SolveOpts), C.POINTER(vp)]
        L.mk_igm_destroy.argtypes = [vp]
        L.mk_igm_create_from_mesh.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, vp, vp, vp, C.c_int, vp, vp, vp, vp, vp, vp, C.c_double,
                                              C.POINTER(SolveOpts), C.POINTER(vp)]
        L.mk_igm_get_x2f.argtypes = [vp, i64p, vp, vp, vp, vp]
        L.mk_igm_scatter.argtypes = [vp, vp, vp, vp]
        L.mk_igm_poisson.argtypes = [vp, C.c_int, vp, vp, vp, vp, C.POINTER(Stats)]
        L.mk_igm_round.argtypes = [vp, C.c_int, vp, vp, C.c_int, vp, C.c_int, vp, C.c_int, C.c_int, vp, vp, vp, ip, vp, vp, C.POINTER(Stats)]
        L.mk_igm_round_injective.argtypes = [vp, C.c_int, vp, vp, C.c_int, vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, ip, vp, vp, C.POINTER(Stats)]
        L.mk_igm_get_matrix.argtypes = [vp, i64p, vp, vp, vp]
        L.mk_igm_spmv_bench.argtypes = [vp, C.c_int, C.c_int, dp]
        L.mk_igm_hierarchy.argtypes = [vp, C.c_int, ip, vp, vp, vp]
        L.mk_igm_profile_cycle.argtypes = [vp, C.c_int, C.c_int, ip, vp]
        L.mk_dist_unique_id.argtypes = [C.c_char_p, C.c_int]
        L.mk_dist_create.argtypes = [vp, C.c_int, C.c_int, C.c_char_p, C.c_int, C.POINTER(vp)]
        L.mk_dist_destroy.argtypes = [vp]
        L.mk_dfield_create.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, C.POINTER(vp)]
        L.mk_dfield_destroy.argtypes = [vp]
        L.mk_dfield_set_global.argtypes = [vp, C.c_int, vp, vp, vp]
        L.mk_dfield_spmv.argtypes = [vp, vp, vp]
        L.mk_dfield_spmv_bench.argtypes = [vp, C.c_int, C.c_int, dp]
        L.mk_dfield_pcg.argtypes = [vp, vp, vp, C.POINTER(SolveOpts), vp, C.POINTER(Stats)]
        L.mk_dfield_pcg_dev.argtypes = [vp, vp, vp, C.POINTER(SolveOpts), C.POINTER(Stats)]
        L.mk_dfield_solve.argtypes = [vp, vp, C.c_int, C.POINTER(SolveOpts), vp, C.POINTER(Stats)]
        L.mk_amg_host_build.argtypes = [C.c_int, vp, vp, vp, C.c_int, vp, C.POINTER(vp)]
        L.mk_amg_dev_build.argtypes = [vp, C.c_int, vp, vp, vp, C.c_int, C.POINTER(vp)]
        L.mk_dev_sparse_op.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, C.c_int, C.c_int, vp, vp, vp, C.POINTER(vp)]
        L.mk_amg_host_build_typed.argtypes = [C.c_int, vp, vp, vp, C.c_int, vp, C.POINTER(vp)]
        L.mk_amg_dev_build_typed.argtypes = [vp, C.c_int, vp, vp, vp, C.c_int, vp, C.POINTER(vp)]
        L.mk_amg_host_num_levels.argtypes = [vp]
        L.mk_amg_host_level_info.argtypes = [vp, C.c_int, C.c_int, ip, ip, i64p, dp]
        L.mk_amg_host_level_get.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp]
        L.mk_amg_host_destroy.argtypes = [vp]
        _LIB = L
    return _LIB


def ptr(a: np.ndarray | None):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def default_opts(**kw) -> SolveOpts:
    o = SolveOpts()
    lib().mk_default_opts(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def as_i32(a):
    return np.ascontiguousarray(a, np.int32)


def as_f64(a):
    return np.ascontiguousarray(a, np.float64)


def as_c128(a):
    return np.ascontiguousarray(a, np.complex128)


class Context:
    """mk_ctx: one device + stream. Pass `stream` (a raw cudaStream_t int) to run on e.g. torch's current stream."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self.L = lib()
        self.h = C.c_void_p()
        rc = self.L.mk_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(self.h))
        if rc != MK_OK:
            raise MkError(rc, f"mk_ctx_create(device={device}) -- a B200 (sm_100a) device is required; there is no CPU fallback")
        self.device = device

    def check(self, rc: int, allow=()):
        if rc != MK_OK and rc not in allow:
            raise MkError(rc, self.L.mk_last_error(self.h).decode())
        return rc

    def sync(self):
        self.check(self.L.mk_ctx_sync(self.h))

    def close(self):
        if self.h:
            self.L.mk_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _read_hierarchy(L, h, dt):
    import scipy.sparse as sp
    out = []
    try:
        nl = L.mk_amg_host_num_levels(h)
        for l in range(nl):
            mats = []
            lm = C.c_double()
            for which in (0, 1):
                r, cc, z = C.c_int(), C.c_int(), C.c_int64()
                if L.mk_amg_host_level_info(h, l, which, C.byref(r), C.byref(cc), C.byref(z), C.byref(lm)) != MK_OK:
                    mats.append(None)
                    continue
                p_ = np.empty(r.value + 1, np.int32)
                c_ = np.empty(z.value, np.int32)
                v_ = np.empty(z.value, dt)
                L.mk_amg_host_level_get(h, l, which, ptr(p_), ptr(c_), ptr(v_))
                mats.append(sp.csr_matrix((v_, c_, p_), shape=(r.value, cc.value)))
            dinv = np.empty(mats[0].shape[0], dt)
            if L.mk_amg_host_level_get(h, l, 2, None, None, ptr(dinv)) != MK_OK:
                dinv = None
            out.append((mats[0], mats[1], lm.value, dinv))
    finally:
        L.mk_amg_host_destroy(h)
    return out


def _canonical_csr(A):
    import scipy.sparse as sp
    A = sp.csr_matrix(A)
    A.sum_duplicates()
    A.sort_indices()
    cx = np.iscomplexobj(A.data)
    return A, cx, (np.complex128 if cx else np.float64)


def amg_host_hierarchy(A, cand=None, types=None):
    """Builds the solvers' smoothed-aggregation hierarchy with the host mirror (no device) and returns
    [(A_l, P_l or None, lmax_l, dinv_l)] (scipy CSR matrices, smoother diagonal) -- for CPU tests of the set-up.
    types: optional dof-type label per row (typed aggregation, what mk_igm_create uses for the IGM operator)."""
    A, cx, dt = _canonical_csr(A)
    L = lib()
    h = C.c_void_p()
    val = np.ascontiguousarray(A.data, dt)
    c = as_f64(cand) if cand is not None else None
    if types is not None:
        assert cand is None
        ty = as_i32(types)
        rc = L.mk_amg_host_build_typed(A.shape[0], ptr(as_i32(A.indptr)), ptr(as_i32(A.indices)), ptr(val), int(cx), ptr(ty), C.byref(h))
    else:
        rc = L.mk_amg_host_build(A.shape[0], ptr(as_i32(A.indptr)), ptr(as_i32(A.indices)), ptr(val), int(cx), ptr(c), C.byref(h))
    if rc != MK_OK:
        raise MkError(rc, "mk_amg_host_build")
    return _read_hierarchy(L, h, dt)


def amg_dev_hierarchy(ctx: "Context", A, types=None):
    """The same hierarchy built by the device set-up (what the solvers run), copied back level by level."""
    A, cx, dt = _canonical_csr(A)
    h = C.c_void_p()
    val = np.ascontiguousarray(A.data, dt)
    ty = as_i32(types) if types is not None else None
    ctx.check(ctx.L.mk_amg_dev_build_typed(ctx.h, A.shape[0], ptr(as_i32(A.indptr)), ptr(as_i32(A.indices)), ptr(val), int(cx), ptr(ty), C.byref(h)))
    return _read_hierarchy(ctx.L, h, dt)


def dev_sparse_op(ctx: "Context", op: str, A, B=None):
    """Device set-up algebra for tests: op = 'matmul' (A @ B), 'T' (A^T) or 'H' (A^H); returns scipy CSR."""
    A, cxa, _ = _canonical_csr(A)
    cx = cxa
    if B is not None:
        B, cxb, _ = _canonical_csr(B)
        cx = cxa or cxb
    dt = np.complex128 if cx else np.float64
    code = {"matmul": 0, "T": 1, "H": 2}[op]
    h = C.c_void_p()
    av = np.ascontiguousarray(A.data, dt)
    args = [ctx.h, code, int(cx), A.shape[0], A.shape[1], ptr(as_i32(A.indptr)), ptr(as_i32(A.indices)), ptr(av)]
    keep = []
    if B is not None:
        bv = np.ascontiguousarray(B.data, dt)
        bp, bc = as_i32(B.indptr), as_i32(B.indices)
        keep = [bv, bp, bc]
        args += [B.shape[0], B.shape[1], ptr(bp), ptr(bc), ptr(bv)]
    else:
        args += [0, 0, None, None, None]
    ctx.check(ctx.L.mk_dev_sparse_op(*args, C.byref(h)))
    return _read_hierarchy(ctx.L, h, dt)[0][0]


def igm_dof_types(UFull, n: int = 2) -> np.ndarray:
    """Dof-type labels of the reduced IGM unknowns as mk_igm_create derives them: the full dof a column of UFull contributes to most
    strongly has index n * vertex + k; the label is k (u- or v-type). Mirror of igm_types_from_ufull (csrc/capi_igm.cu) for the tests."""
    import scipy.sparse as sp
    U = sp.csr_matrix(UFull).tocoo()
    best = np.zeros(U.shape[1])
    ty = np.zeros(U.shape[1], np.int32)
    order = np.lexsort((U.col, U.row))           # row-major traversal, like the CSR loop of the C++ side
    for r, c, v in zip(U.row[order], U.col[order], np.abs(U.data[order])):
        if v > best[c]:
            best[c] = v
            ty[c] = r % n
    return ty
