"""Generates the committed golden fixtures under tests/golden/.

The reference (komietty/metriko) ships no golden vectors and cannot be built or imported in this image
(C++ with Eigen/libigl absent), so these fixtures are NOT outputs of the reference: they are outputs of the CPU
oracle (oracle/, a line-by-line restatement with SuperLU direct solves) on inputs produced by the host-glue
stage. They pin (1) the oracle against regressions and (2) the CUDA path on the GPU box, where neither
/root/reference nor a long oracle run is available. tests/golden/make_golden_ref.py / tests/test_oracle_ref.py tie them to the
reference's own rounding-loop / LM / inverse-iteration code compiled against Eigen stand-ins (ref_<name>.npz).

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(1, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from hostglue import Session  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def csr_parts(name, A):
    A = A.tocsr()
    A.sort_indices()
    return {f"{name}_ptr": A.indptr.astype(np.int64), f"{name}_col": A.indices.astype(np.int32), f"{name}_val": A.data.astype(np.float64),
            f"{name}_shape": np.asarray(A.shape, np.int64)}


def make(name, session, seed, round_seams=False):
    s = session
    fi = s.field_inputs()
    nF = s.dims["nF"]
    u0 = oracle.random_u0(nF, seed)
    L = oracle.connection_laplacian(fi["twin"], fi["face"], fi["heArgOnFace"], nF, 4)
    M = oracle.galerkin_mass_matrix(fi["faceArea"])
    z = oracle.solve_smallest_eig(L, M, u0, 50)
    lam = (np.vdot(z, L @ z) / np.vdot(z, M @ z)).real
    s.field_post(z, 4, "curl")
    dims = s.igm_setup(0.02)
    inp = s.igm_inputs()
    out = oracle.iterative_rounding(inp["fixedIdcs"], inp["fixedVals"], inp["singularIdcs"], inp["integerIdcs"], inp["length"],
                                    inp["UFull"], inp["G2p"], inp["N"], inp["n"], inp["nF"], True, round_seams, inp["F2"], ni_mode="lm")
    fin = s.igm_finish(out["fullx"])
    assert out["success"] and fin["flipped_faces"] == 0, (name, out["success"], fin["flipped_faces"])
    d = dict(
        dims=np.asarray([s.dims[k] for k in ("nV", "nF", "nE", "nH", "euler")], np.int64),
        twin=fi["twin"], face=fi["face"], face2half=fi["face2half"], next=fi["next"], heArgOnFace=fi["heArgOnFace"], faceArea=fi["faceArea"],
        u0=u0, z50=z, lam=np.float64(lam), singular=s.get("singular"),
        F2=inp["F2"], fixedIdcs=inp["fixedIdcs"], fixedVals=inp["fixedVals"], integerIdcs=inp["integerIdcs"], singularIdcs=inp["singularIdcs"],
        length=np.float64(inp["length"]), N=np.int64(inp["N"]), n=np.int64(inp["n"]), round_seams=np.int64(round_seams),
        x0_full=out["x0_full"], fullx=out["fullx"], out_fixedIdcs=np.asarray(out["fixedIdcs"], np.int64),
        out_fixedVals=np.asarray(out["fixedVals"], np.int64), n_solves=np.int64(out["n_solves"]), min_margin=np.float64(out["min_margin"]),
        igm_dims=np.asarray(list(dims.values()), np.int64),
    )
    d.update(csr_parts("X2F", inp["X2F"]))
    d.update(csr_parts("UFull", inp["UFull"]))
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **d)
    print(name, "nF", nF, "lambda", lam, "singular", int(np.count_nonzero(d["singular"])), "solves", out["n_solves"], "margin", out["min_margin"],
          "bytes", os.path.getsize(path))


if __name__ == "__main__":
    make("sphere_ico3", Session().make_sphere(20, 3, bump=0.08), seed=0)
    make("torus_24x16", Session().make_torus(24, 16, bump=0.1), seed=1)
    make("sphere_octa4_roundseams", Session().make_sphere(8, 4, bump=0.1, phase=0.7), seed=2, round_seams=True)
