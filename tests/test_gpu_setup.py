"""GPU tests of the device set-up algebra (csrc/devcsr.cuh, csrc/amg_dev.cuh) through the C-ABI: the symbolic/numeric SpGEMM and
the transposition against scipy, the device-built multigrid hierarchy against its host mirror (the same algorithm on the CPU),
and bit-reproducibility of the device build (no float atomics anywhere in the set-up)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

import oracle
from conftest import load_golden
from metriko_b200 import api, capi
from hostglue import Session

pytestmark = pytest.mark.gpu


def _rand(rng, m, n, density, cx, long_rows=(), long_len=0, empty_rows=()):
    A = sp.random(m, n, density=density, random_state=rng, format="lil", dtype=np.float64)
    for r in long_rows:
        cols = rng.choice(n, size=min(long_len, n), replace=False)
        A[r, cols] = rng.standard_normal(len(cols))
    for r in empty_rows:
        A[r, :] = 0
    A = A.tocsr()
    A.eliminate_zeros()
    if cx:
        A = A.astype(np.complex128)
        A.data = A.data + 1j * rng.standard_normal(A.nnz)
    return A


def _same(C, R, tol=1e-12):
    C, R = sp.csr_matrix(C), sp.csr_matrix(R)
    assert C.shape == R.shape
    C.sort_indices()
    assert np.all(np.diff(C.indptr) >= 0)
    for i in range(0, C.shape[0], max(1, C.shape[0] // 50)):   # sampled rows: strictly increasing columns
        c = C.indices[C.indptr[i]:C.indptr[i + 1]]
        assert np.all(np.diff(c) > 0)
    scale = max(1.0, abs(R).max() if R.nnz else 1.0)
    assert abs(C - R).max() <= tol * scale if (C - R).nnz else True


@pytest.mark.parametrize("cx", [False, True])
def test_spgemm_all_bins_against_scipy(ctx, cx):
    rng = np.random.default_rng(7)
    # A: mostly short rows; rows whose product bound lands in the 128 / 1024 / 8192 bins; rows longer than 512 and products
    # beyond 8192 candidates (dense-accumulator path); empty rows
    A = _rand(rng, 3000, 2500, 0.002, cx, long_rows=(5, 900, 1800), long_len=700, empty_rows=(0, 17, 2999))
    A = sp.lil_matrix(A)
    for r, ln in ((40, 40), (41, 120), (42, 400)):
        cols = rng.choice(2500, size=ln, replace=False)
        A[r, cols] = 1.0 + rng.random(ln)
    A = A.tocsr()
    B = _rand(rng, 2500, 4000, 0.004, cx, long_rows=(3, 1000), long_len=1500, empty_rows=(1, 2))
    C = capi.dev_sparse_op(ctx, "matmul", A, B)
    R = (A @ B).tocsr()
    _same(C, R)
    # the symbolic pattern is the structural product (explicit zeros from cancellation are kept)
    S = (abs(sp.csr_matrix(A)).sign() @ abs(sp.csr_matrix(B)).sign()).tocsr()
    assert C.nnz == S.nnz
    C2 = capi.dev_sparse_op(ctx, "matmul", A, B)
    assert np.array_equal(C.indices, C2.indices) and np.array_equal(C.data.view(np.uint8), C2.data.view(np.uint8))  # bit-reproducible


@pytest.mark.parametrize("cx", [False, True])
def test_transpose_against_scipy(ctx, cx):
    rng = np.random.default_rng(11)
    A = _rand(rng, 4000, 1500, 0.003, cx, long_rows=(7,), long_len=1400, empty_rows=(0,))
    A = sp.lil_matrix(A)
    A[:, 3] = (rng.random((4000, 1)) > 0.25) * 1.0      # a column with ~3000 entries: a row of the transpose in the CTA bin
    big = sp.random(12000, 1500, density=0.0005, random_state=rng, format="lil")
    big[:, 9] = 1.0                                       # 12000 entries: the global-scratch sort
    A = sp.vstack([A.tocsr().astype(np.complex128 if cx else np.float64), big.tocsr()]).tocsr()
    T = capi.dev_sparse_op(ctx, "T", A)
    _same(T, A.T.tocsr(), 0.0)
    H = capi.dev_sparse_op(ctx, "H", A)
    _same(H, A.conj().T.tocsr(), 0.0)


def _igm_matrix(g):
    X = (g["X2F"] * float(g["length"])).tocsr()
    return (X.T @ X + 1e-4 * sp.identity(X.shape[1])).tocsr()


def _compare_hierarchies(Hd, Hh):
    assert [h[0].shape for h in Hd] == [h[0].shape for h in Hh]
    for (Ad, Pd, lmd, dd), (Ah, Ph, lmh, dh) in zip(Hd, Hh):
        sc = abs(Ah).max()
        assert abs(Ad - Ah).max() <= 1e-11 * sc if (Ad - Ah).nnz else True
        assert abs(lmd - lmh) <= 1e-9 * lmh
        assert np.abs(dd - dh).max() <= 1e-11 * np.abs(dh).max()
        if Ph is not None:
            assert Pd is not None and Pd.shape == Ph.shape
            assert abs(Pd - Ph).max() <= 1e-11 if (Pd - Ph).nnz else True


def test_device_hierarchy_equals_host_mirror(ctx):
    g = load_golden("sphere_ico3")
    A = _igm_matrix(g)
    L = oracle.connection_laplacian(g["twin"], g["face"], g["heArgOnFace"], len(g["faceArea"]), 4).tocsr()
    for M in (A, L):
        _compare_hierarchies(capi.amg_dev_hierarchy(ctx, M), capi.amg_host_hierarchy(M))


def test_device_hierarchy_with_dense_dofs(ctx):
    """A 16k-vertex sphere: the IGM operator has rows with hundreds of entries (seam translations / constraint-kernel columns);
    the device build must carry them as singletons exactly like the host mirror, and build the same face-operator hierarchy."""
    s = Session().make_sphere(8, 6, bump=0.08)
    fi = s.field_inputs()
    nF = s.dims["nF"]
    F = api.FaceRosyField(ctx, fi["face2half"], fi["next"], fi["twin"], fi["face"], fi["heArgOnFace"], fi["faceArea"], 4)
    z = F.solveSmallestEig(oracle.random_u0(nF, 0), 50)
    L = oracle.connection_laplacian(fi["twin"], fi["face"], fi["heArgOnFace"], nF, 4).tocsr()
    _compare_hierarchies(capi.amg_dev_hierarchy(ctx, L), capi.amg_host_hierarchy(L))
    s.field_post(z, 4, "curl")
    s.igm_setup(0.02)
    inp = s.igm_inputs()
    A = _igm_matrix(dict(X2F=inp["X2F"], length=inp["length"]))
    assert (np.diff(A.indptr) > 64).sum() > 0
    Hd, Hh = capi.amg_dev_hierarchy(ctx, A), capi.amg_host_hierarchy(A)
    _compare_hierarchies(Hd, Hh)
    Hd2 = capi.amg_dev_hierarchy(ctx, A)
    for a, b in zip(Hd, Hd2):   # bit-reproducible set-up
        assert np.array_equal(a[0].data.view(np.uint8), b[0].data.view(np.uint8))
    # typed aggregation (u- / v-type dofs never share an aggregate; what mk_igm_create builds): device == host mirror, and
    # it is a different hierarchy from the untyped one
    ty = capi.igm_dof_types(inp["UFull"])
    Ht_d, Ht_h = capi.amg_dev_hierarchy(ctx, A, types=ty), capi.amg_host_hierarchy(A, types=ty)
    _compare_hierarchies(Ht_d, Ht_h)
    assert [h[0].shape for h in Ht_h] != [h[0].shape for h in Hh]


def test_host_and_device_setup_give_the_same_rounding(ctx):
    g = load_golden("sphere_octa4_roundseams")
    outs = []
    for mode in ("device", "host"):
        os.environ["MK_SETUP"] = mode
        try:
            prob = api.IgmProblem(ctx, g["X2F"], g["UFull"], g["F2"], float(g["length"]))
            X = (g["X2F"] * float(g["length"])).tocsr()
            assert abs(prob.matrix() - X.T @ X).max() < 1e-11
            outs.append(prob.round(g["fixedIdcs"], g["fixedVals"], g["singularIdcs"], g["integerIdcs"], True, bool(g["round_seams"])))
            prob.close()
        finally:
            os.environ.pop("MK_SETUP", None)
    a, b = outs
    assert a["success"] and b["success"]
    # same integers; the ORDER may differ among dofs that are already integral to round-off (selection ties, flagged by the API)
    assert dict(zip(map(int, a["fixedIdcs"]), map(int, a["fixedVals"]))) == dict(zip(map(int, b["fixedIdcs"]), map(int, b["fixedVals"])))
    assert np.abs(a["fullx"] - b["fullx"]).max() < 1e-7


def test_device_resident_inputs_and_precisions_give_the_same_rounding(ctx):
    """mk_igm_create_dev (X2F / F2 already in HBM: the `value` path of bench.py) against the host-buffer entry point, and the fp32
    V-cycle (default) against the fp64 one (MK_PRECOND_FP32=0): identical integers, fullx within the post_checking tolerance."""
    import torch
    g = load_golden("sphere_octa4_roundseams")
    X = sp.csr_matrix(g["X2F"], dtype=np.float64)
    X.sort_indices()
    U = sp.csr_matrix(g["UFull"], dtype=np.float64)
    U.sort_indices()
    raw = (X.indptr.astype(np.int64), X.indices.astype(np.int32), X.data.astype(np.float64),
           U.indptr.astype(np.int64), U.indices.astype(np.int32), U.data.astype(np.float64), np.ascontiguousarray(g["F2"], np.float64))
    dev_t = [torch.from_numpy(a).cuda() for a in (raw[0], raw[1], raw[2], raw[6])]
    outs = {}
    for name, kw, env in (("host", {}, None), ("dev", {"dev": tuple(int(t.data_ptr()) for t in dev_t)}, None), ("fp64", {}, "0")):
        if env is not None:
            os.environ["MK_PRECOND_FP32"] = env
        try:
            prob = api.IgmProblem(ctx, X, U, g["F2"], float(g["length"]), raw=raw, **kw)
            outs[name] = prob.round(g["fixedIdcs"], g["fixedVals"], g["singularIdcs"], g["integerIdcs"], True, bool(g["round_seams"]))
            assert prob.stats["relres_max"] <= 1.001e-10
            prob.close()
        finally:
            os.environ.pop("MK_PRECOND_FP32", None)
    ref = dict(zip(g["out_fixedIdcs"].tolist(), g["out_fixedVals"].tolist()))
    for name, o in outs.items():
        assert o["success"], name
        assert dict(zip(map(int, o["fixedIdcs"]), map(int, o["fixedVals"]))) == ref, name
        assert np.abs(o["fullx"] - g["fullx"]).max() < 1e-6, name


def test_field_fp32_and_fp64_cycles_agree(ctx, ico4):
    fi = ico4.field_inputs()
    nF = ico4.dims["nF"]
    u0 = oracle.random_u0(nF, 0)
    z = {}
    for mode in ("1", "0"):
        os.environ["MK_PRECOND_FP32"] = mode
        try:
            F = api.FaceRosyField(ctx, fi["face2half"], fi["next"], fi["twin"], fi["face"], fi["heArgOnFace"], fi["faceArea"], 4)
            z[mode] = F.solveSmallestEig(u0, 50)
            assert F.stats["relres_max"] <= 1.001e-10 and F.stats["status"] == 0
            F.close()
        finally:
            os.environ.pop("MK_PRECOND_FP32", None)
    from oracle.field import field_angle_error
    assert field_angle_error(z["1"], z["0"], 4).max() < 1e-7
