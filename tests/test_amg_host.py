"""CPU tests of the multigrid set-up (host side of the solvers; no device needed): Galerkin products, dense-dof handling
and the quality of the hierarchy, measured by a numpy mirror of the device V-cycle + PCG (tests/amg_mirror.py)."""
import numpy as np
import scipy.sparse as sp

import oracle
from amg_mirror import VCycle, pcg, rounding_loop
from conftest import load_golden
from metriko_b200 import capi


def _igm_matrix(g):
    X = (g["X2F"] * float(g["length"])).tocsr()
    return (X.T @ X + 1e-4 * sp.identity(X.shape[1])).tocsr(), X.T @ g["F2"]


def test_galerkin_products_real_and_complex():
    g = load_golden("sphere_ico3")
    A, _ = _igm_matrix(g)
    L = oracle.connection_laplacian(g["twin"], g["face"], g["heArgOnFace"], len(g["faceArea"]), 4).tocsr()
    for M in (A, L):
        H = capi.amg_host_hierarchy(M)
        assert len(H) >= 2 and abs(H[0][0] - M).max() < 1e-14
        for (Al, P, _, _), (Ac, _, _, _) in zip(H[:-1], H[1:]):
            assert P.shape == (Al.shape[0], Ac.shape[0])
            G = (P.conj().T @ Al @ P).tocsr()
            assert abs(G - Ac).max() <= 1e-12 * abs(Ac).max()
            assert abs(Ac - Ac.conj().T).max() <= 1e-12 * abs(Ac).max()      # Hermitian stays Hermitian


def test_dense_dofs_are_carried_as_singletons():
    """The IGM operator has a few rows with hundreds to thousands of entries (seam translations / constraint-kernel
    columns). They must reach the coarsest level as their own unknowns instead of swallowing whole aggregates."""
    from hostglue import Session
    s = Session().make_sphere(8, 5, bump=0.08)
    fi = s.field_inputs()
    nF = s.dims["nF"]
    L = oracle.connection_laplacian(fi["twin"], fi["face"], fi["heArgOnFace"], nF, 4)
    z = oracle.solve_smallest_eig(L, oracle.galerkin_mass_matrix(fi["faceArea"]), oracle.random_u0(nF, 0), 50)
    s.field_post(z, 4, "curl")
    s.igm_setup(0.02)
    inp = s.igm_inputs()
    A, b = _igm_matrix(dict(X2F=inp["X2F"], length=inp["length"], F2=inp["F2"]))
    rowlen = np.diff(A.indptr)
    dense = np.where(rowlen > 64)[0]
    assert len(dense) > 0
    H = capi.amg_host_hierarchy(A)
    P0 = H[0][1]
    # every dense dof owns a coarse unknown of its own: its dominant coarse columns are pairwise distinct
    P0 = P0.tocsr()
    dom = np.array([P0.indices[P0.indptr[i]:P0.indptr[i + 1]][np.argmax(abs(P0.data[P0.indptr[i]:P0.indptr[i + 1]]))] for i in range(A.shape[0])])
    owners = dom[dense]
    assert len(set(owners.tolist())) == len(dense)
    assert H[1][0].shape[0] < A.shape[0] / 4
    x, it = pcg(A, b, np.zeros_like(b), VCycle(H))
    assert it <= 45 and np.linalg.norm(A @ x - b) <= 1.01e-10 * np.linalg.norm(b)


def test_hierarchy_quality_by_mirrored_pcg():
    """PCG with the mirrored V-cycle must reach 1e-10 in a few dozen iterations on the fixtures (plain CG needs 3-6x more even at these tiny sizes, 75x more at 131k unknowns)."""
    for name, bound in (("sphere_ico3", 40), ("torus_24x16", 60), ("sphere_octa4_roundseams", 50)):
        g = load_golden(name)
        A, b = _igm_matrix(g)
        H = capi.amg_host_hierarchy(A)
        x, it = pcg(A, b, np.zeros_like(b), VCycle(H))
        assert it <= bound, (name, it)
        assert np.linalg.norm(A @ x - b) <= 1e-10 * np.linalg.norm(b) * 1.01
        _, it_plain = pcg(A, b, np.zeros_like(b), None, maxit=5000)
        assert it_plain > 3 * it
    g = load_golden("sphere_ico3")
    nF = len(g["faceArea"])
    L = oracle.connection_laplacian(g["twin"], g["face"], g["heArgOnFace"], nF, 4).tocsr()
    M = oracle.galerkin_mass_matrix(g["faceArea"])
    b = M @ g["u0"]
    x, it = pcg(L, b, np.zeros_like(b), VCycle(capi.amg_host_hierarchy(L)))
    assert it <= 30 and np.linalg.norm(L @ x - b) <= 1.01e-10 * np.linalg.norm(b)


def test_typed_aggregation_keeps_u_and_v_dofs_apart():
    """The IGM hierarchy is built with a dof-type label per reduced unknown (u / v, derived from UFull): no aggregate mixes the two
    on any level (every prolongator column is supported on ONE type of the level below), the Galerkin relation still holds, and the
    mirrored PCG needs no more iterations than with untyped aggregates (it needs ~20 % fewer from 16k vertices up)."""
    g = load_golden("sphere_octa4_roundseams")
    A, b = _igm_matrix(g)
    A.sort_indices()
    ty = capi.igm_dof_types(g["UFull"])
    assert ty.shape == (A.shape[0],) and set(np.unique(ty).tolist()) == {0, 1}
    H = capi.amg_host_hierarchy(A, types=ty)
    Hu = capi.amg_host_hierarchy(A)
    assert len(H) >= 2
    rowlen = np.diff(A.indptr)
    t = ty.copy()
    for lvl, ((Al, P, _, _), (Ac, _, _, _)) in enumerate(zip(H[:-1], H[1:])):
        G = (P.T @ Al @ P).tocsr()
        assert abs(G - Ac).max() <= 1e-12 * abs(Ac).max()
        # dominant entry of a prolongator row = the aggregate of that dof (the smoothing only adds neighbours)
        Pc = P.tocsr()
        agg = np.array([Pc.indices[Pc.indptr[i]:Pc.indptr[i + 1]][np.argmax(abs(Pc.data[Pc.indptr[i]:Pc.indptr[i + 1]]))] for i in range(P.shape[0])])
        tc = np.full(P.shape[1], -1)
        for i, a in enumerate(agg):
            assert tc[a] in (-1, t[i]), f"level {lvl}: aggregate {a} mixes dof types"
            tc[a] = t[i]
        assert (tc >= 0).all()
        t = tc
    _, it_typed = pcg(A, b, np.zeros_like(b), VCycle(H))
    _, it_plain = pcg(A, b, np.zeros_like(b), VCycle(Hu))
    assert it_typed <= it_plain + 1, (it_typed, it_plain)


def test_constraint_preconditioner_in_the_mirrored_rounding_loop():
    """The re-solves of the rounding loop run PCG in the feasible subspace. Preconditioning it with P V P (V-cycle, then projector)
    costs iterations per constraint; the constraint preconditioner V - W (Q W)^-1 Q V with W = V Q^T (csrc/igm.cuh) does not:
    same solutions on the same rounding sequence, clearly fewer iterations, and typed aggregation only pays with it
    (B200, 1M vertices: 1347 -> 841 -> 698 inner iterations, profiles/cprec_probe_r01.txt)."""
    import scipy.sparse as sp
    for name in ("sphere_ico3", "torus_24x16"):
        g = load_golden(name)
        X = (g["X2F"] * float(g["length"])).tocsr()
        A, f = _igm_matrix(g)
        A.sort_indices()
        U = sp.csr_matrix(g["UFull"])
        left = [int(i) for i in g["singularIdcs"]]
        res, seq = {}, None
        for tname, types in (("untyped", None), ("typed", capi.igm_dof_types(U))):
            V = VCycle(capi.amg_host_hierarchy(A, types=types), deg=1)
            for mode in ("proj", "cprec"):
                its, ints, x = rounding_loop(A, f, U, g["fixedIdcs"], g["fixedVals"], left, V, mode, X=X, replay=seq)
                seq = seq or ints
                assert ints == seq
                res[tname, mode] = (sum(its), x, its)
        totals = {k: v[0] for k, v in res.items()}
        base = res["untyped", "proj"]
        for key, (tot, x, its) in res.items():
            assert np.abs(x - base[1]).max() < 1e-6, (name, key)
        assert totals["untyped", "cprec"] < 0.85 * totals["untyped", "proj"], (name, totals)
        assert totals["typed", "cprec"] < 0.75 * totals["typed", "proj"], (name, totals)
        assert totals["typed", "cprec"] <= totals["untyped", "cprec"], (name, totals)
        # no growth with the number of constraints any more: the last re-solves are the cheapest
        its = res["typed", "cprec"][2]
        assert max(its[-4:]) <= its[1], (name, its)
