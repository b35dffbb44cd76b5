"""The reference's example_igm.cpp pipeline with its two solver seams on the B200 (oracle/_ref/libmetriko_ref_overlay.so: metriko's own
face_rosy_field.h / parameterization.h compiled in place with solveSmallestEig (frf.h:22) and iterative_rounding (integ.ipp:65) redirected to
include/metriko_b200/metriko_cuda_shims.hpp, tests/cpp/ref_overlay_test.cpp) on the reference's icosphere.obj at scale 0.03, compared with the
oracle golden tests/golden/large_icosphere.npz and with the reference's own CPU run tests/golden/ref_large_icosphere.npz.

    python scripts/run_ref_overlay.py          (on a GPU box; the library is prebuilt in this container by `make -C oracle overlay`)

Round 2, one B200 (profiles/overlay_r02_reference_example_igm_on_b200.log, 3.7 s for the whole script): integ() true, the integers identical to
the oracle golden, fullx within 7.2e-9 of the golden and of the reference's own CPU run. The CPU side (compile, link, the C-ABI symbols the
reference's code now reaches, loud failure without a device) is tests/test_shim.py."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(1, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from golden.make_golden_large import RECIPES, make_session  # noqa: E402


def main():
    so = os.path.join(ROOT, "oracle", "_ref", "libmetriko_ref_overlay.so")
    lib = C.CDLL(so)
    recipe, seed, rs, scale, _ = RECIPES["icosphere"]
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "large_icosphere.npz")))
    r = dict(np.load(os.path.join(ROOT, "tests", "golden", "ref_large_icosphere.npz")))
    s = make_session(recipe)
    fi = s.field_inputs()
    nF = s.dims["nF"]
    u0 = oracle.random_u0(nF, seed)
    # the seam comes from the host glue on the ORACLE's field (compute_seam is libigl); the device field agrees with it to ~1e-9 rad
    L = oracle.connection_laplacian(fi["twin"], fi["face"], fi["heArgOnFace"], nF, 4)
    z = oracle.solve_smallest_eig(L, oracle.galerkin_mass_matrix(fi["faceArea"]), u0, 50)
    s.field_post(z, 4, "curl")
    V, F = np.ascontiguousarray(s.get("pos"), np.float64), np.ascontiguousarray(s.get("idx"), np.int32)
    seam = np.ascontiguousarray(s.get("seam"), np.int32)
    u0f = np.ascontiguousarray(u0).view(np.float64)
    cap = 4 * int(s.dims["nV"]) + 4096
    fullx, flen, cfn = np.zeros(cap), C.c_long(0), np.zeros(nF * 12)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    rc = lib.ref_overlay_example_igm(int(s.dims["nV"]), V.ctypes.data_as(dp), int(nF), F.ctypes.data_as(ip), 4, u0f.ctypes.data_as(dp),
                                     seam.ctypes.data_as(ip), C.c_double(scale), 1, 0, int(rs), C.c_long(cap), fullx.ctypes.data_as(dp), C.byref(flen),
                                     cfn.ctypes.data_as(dp))
    fx = fullx[: flen.value]
    print("integ() through the overlay:", rc, " len(fullx)", flen.value)
    if rc != 1:
        print("no result (rc -1: no usable sm_100a device, or the overlay reported a failure)")
        return 1
    print("integers identical to the oracle golden:", bool(np.array_equal(np.round(fx[g["out_fixedIdcs"]]).astype(np.int64), g["out_fixedVals"])))
    print("|fullx - oracle golden sample|", float(np.abs(fx[g["fullx_idx"]] - g["fullx_val"]).max()))
    print("|fullx - the reference's own CPU run|", float(np.abs(fx - r["ex_igm_fullx"]).max()))
    return 0 if rc == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
