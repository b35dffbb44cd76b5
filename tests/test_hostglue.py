"""CPU tests of the host-glue stage (the reference's unchanged bookkeeping restated in C++): index conventions of
SURVEY.md Appendix A and the reference's own runtime invariants."""
import numpy as np
import pytest

from hostglue import Session


@pytest.mark.parametrize("maker,euler", [(lambda: Session().make_sphere(20, 2), 2), (lambda: Session().make_sphere(8, 3), 2),
                                         (lambda: Session().make_sphere(4, 3), 2), (lambda: Session().make_torus(24, 16), 0)])
def test_halfedge_topology(maker, euler):
    s = maker()
    d = s.dims
    assert d["euler"] == euler and d["nE"] * 2 == d["nH"] == 3 * d["nF"]
    twin, nxt, prev, face, tail, head, edge = (s.get(k) for k in ("twin", "next", "prev", "face", "tail", "head", "edge"))
    h = np.arange(d["nH"])
    assert np.array_equal(twin[twin], h) and np.array_equal(nxt[nxt[nxt]], h) and np.array_equal(prev[nxt], h)
    assert np.array_equal(tail[twin], head) and np.array_equal(face[nxt], face)
    # dcel numbering: half 2e is canonical with tail = vmin of the edge (hmesh.h:256-262, A.1)
    e2v = s.get("edge2vert").reshape(-1, 2)
    assert np.array_equal(edge, h // 2) and np.array_equal(tail[0::2], e2v[:, 0]) and np.array_equal(tail[1::2], e2v[:, 1])
    assert np.all(e2v[:, 0] < e2v[:, 1]) and np.all(np.diff(e2v[:, 0] * d["nV"] + e2v[:, 1].astype(np.int64)) > 0)
    # face2half[f] runs F(f,0) -> F(f,1) (hmesh.h:379)
    idx = s.get("idx").reshape(-1, 3)
    f2h = s.get("face2half")
    assert np.array_equal(tail[f2h], idx[:, 0]) and np.array_equal(head[f2h], idx[:, 1])
    # crnr[h] is the corner opposite to h (hmesh.h:430-441)
    crnr = s.get("crnr")
    cv = idx.reshape(-1)[crnr]
    assert np.all(cv != tail) and np.all(cv != head) and np.array_equal(crnr // 3, face)


def test_geometry_caches():
    s = Session().make_sphere(20, 3, bump=0.08)
    d = s.dims
    pos = s.get("pos").reshape(-1, 3)
    idx = s.get("idx").reshape(-1, 3)
    area = s.get("faceArea")
    bx, by, nn = (s.get(k).reshape(-1, 3) for k in ("faceBasisX", "faceBasisY", "faceNormal"))
    e1, e2 = pos[idx[:, 1]] - pos[idx[:, 0]], pos[idx[:, 2]] - pos[idx[:, 0]]
    assert np.allclose(area, 0.5 * np.linalg.norm(np.cross(e1, e2), axis=1), rtol=1e-14)
    assert np.allclose(np.cross(nn, bx), by, atol=1e-13) and np.allclose(np.einsum("ij,ij->i", bx, by), 0, atol=1e-13)
    # heArgOnFace equals atan2(vec.Y, vec.X) mod 2pi (A.2)
    tail, head, face, farg = s.get("tail"), s.get("head"), s.get("face"), s.get("heArgOnFace")
    v = pos[head] - pos[tail]
    ang = np.arctan2(np.einsum("ij,ij->i", v, by[face]), np.einsum("ij,ij->i", v, bx[face])) % (2 * np.pi)
    diff = np.abs((ang - farg + np.pi) % (2 * np.pi) - np.pi)
    assert diff.max() < 1e-12
    # Gauss-Bonnet: sum of angle defects = 2 pi chi
    assert abs(s.get("angleDefect").sum() - 2 * np.pi * d["euler"]) < 1e-9


def test_rejects_open_mesh():
    V = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], float)
    F = np.array([[0, 1, 2]], np.int32)
    with pytest.raises(RuntimeError):
        Session().set_mesh(V, F)


def _pipeline(s, seed=0):
    import oracle
    fi = s.field_inputs()
    nF = s.dims["nF"]
    L = oracle.connection_laplacian(fi["twin"], fi["face"], fi["heArgOnFace"], nF, 4)
    M = oracle.galerkin_mass_matrix(fi["faceArea"])
    z = oracle.solve_smallest_eig(L, M, oracle.random_u0(nF, seed), 50)
    s.field_post(z, 4, "curl")
    return z


@pytest.mark.parametrize("maker,chi", [(lambda: Session().make_sphere(20, 3, bump=0.08), 2), (lambda: Session().make_torus(24, 16), 0)])
def test_singularities_and_seam(maker, chi):
    s = maker()
    _pipeline(s)
    sing = s.get("singular")
    assert sing.sum() == 4 * chi  # sum of indices = N * chi (frf.h:100-107)
    assert s.scalar("comb_visited") == s.dims["nF"]  # the cut mesh is connected
    # after combing every non-seam edge has matching 0 (A.6)
    cm, seam = s.get("cmatching"), s.get("seam")
    assert np.count_nonzero(cm[seam == 0]) == 0
    d = s.igm_setup(0.02)
    # cutting along the seam makes a disk: V - E + F = 1 with seam edges doubled
    nseam = int(seam.sum())
    assert d["cut_nV"] - (s.dims["nE"] + nseam) + s.dims["nF"] == 1
    inp = s.igm_inputs()
    # UFull spans the kernel of the node constraints (ir_init.h:98-154)
    C = s.csr("Cfull")
    assert abs(C @ inp["UFull"]).max() < 1e-12
    assert inp["UFull"].shape[1] == inp["UFull"].shape[0] - C.shape[0]
    # sign symmetry: rows k and k + n of X2F are negatives (param.h:84-90)
    X = inp["X2F"].toarray() if inp["X2F"].shape[0] < 20000 else None
    if X is not None:
        Xr = X.reshape(-1, 4, 2, X.shape[1])
        assert np.abs(Xr[:, 0] + Xr[:, 2]).max() < 1e-12 and np.abs(Xr[:, 1] + Xr[:, 3]).max() < 1e-12
