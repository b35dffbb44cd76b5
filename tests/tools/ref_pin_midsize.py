"""One-off: the oracle/_ref pin (the reference's own example_igm.cpp pipeline compiled against the Eigen stand-ins) at a size between the unit
fixtures (<= 5k faces) and the BASELINE configurations: an octahedron subdivided `levels` times (6: 32,768 faces). Prints the agreement of the
reference's code with host glue + oracle; too slow for the test suite (the reference's LM re-factorises without reuse). Round 2: levels 5 and 6
were stopped after 27 / 47 CPU-minutes (profiles/README.md); the committed pin ends at the 5,120 faces of icosphere.obj.

    python tests/tools/ref_pin_midsize.py [levels]  
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(1, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from oracle import ref as oref  # noqa: E402
from hostglue import Session  # noqa: E402

levels = int(sys.argv[1]) if len(sys.argv) > 1 else 6
assert oref.build()
R = oref.Ref()
s = Session().make_sphere(8, levels, bump=0.08)
fi = s.field_inputs()
nF = s.dims["nF"]
u0 = oracle.random_u0(nF, 0)
t = time.time()
L = oracle.connection_laplacian(fi["twin"], fi["face"], fi["heArgOnFace"], nF, 4)
z = oracle.solve_smallest_eig(L, oracle.galerkin_mass_matrix(fi["faceArea"]), u0, 50)
s.field_post(z, 4, "curl")
s.igm_setup(0.02)
inp, m = s.igm_inputs(), s.igm_mesh_inputs()
st = {}
o = oracle.iterative_rounding_x2f(inp["fixedIdcs"], inp["fixedVals"], inp["singularIdcs"], inp["integerIdcs"], inp["length"], inp["UFull"], inp["X2F"],
                                  inp["N"], inp["n"], True, False, inp["F2"], ni_mode="lm", stats=st, reuse_factor=True)
fin = s.igm_finish(o["fullx"])
t_oracle = time.time() - t
t = time.time()
r = R.face_rosy_field(s.get("pos"), s.get("idx"), 4, u0, "curl", seam=s.get("seam"),
                      parameterize=dict(gridscale=0.02, seamless=True, localInjectivity=False, roundSeams=False))
t_ref = time.time() - t
print(f"octahedron x{levels}: nV {s.dims['nV']} nF {nF}; host glue + oracle {t_oracle:.1f} s (LM factorisations {st['factorizations']}, "
      f"{st['factorizations_done']} run), the reference's code {t_ref:.1f} s (factorisations {R.L.ref_factor_count(0)})")
for k in ("twin", "next", "face", "face2half", "matching", "singular", "cmatching"):
    assert np.array_equal(r[k], s.get(k).reshape(-1)), k
assert np.array_equal(r["heArgOnFace"], fi["heArgOnFace"]) and np.array_equal(r["faceArea"], fi["faceArea"])
assert np.array_equal(r["cutF"], m["cutF"].reshape(-1)) and all(np.array_equal(r[k], inp[k]) for k in ("fixedIdcs", "integerIdcs", "singularIdcs"))
print("half-edge arrays, per-half angles, areas, matchings, singularities, cut mesh, index lists: identical")
print(f"|z50 - oracle| {np.abs(r['compressed'] - z).max():.2e}   integ() {r['integ_ok']}   integers identical "
      f"{bool(np.array_equal(np.round(r['fullx'][np.asarray(o['fixedIdcs'])]).astype(np.int64), np.asarray(o['fixedVals'])))}   "
      f"|fullx - oracle| {np.abs(r['fullx'] - o['fullx']).max():.2e}   |cfn - host glue| {np.abs(r['cfn'] - fin['cfn'].reshape(-1)).max():.2e}")
