"""ctypes wrapper of libmetriko_hostglue.so: the reference's unchanged host stages (mesh topology,
matching / seam / combing, IGM index bookkeeping) restated so that hot-path inputs can be generated in
this image (no Eigen / libigl). CPU only; TEST-INPUT INFRASTRUCTURE: it lives under tests/, is not part of the product package
(metriko_b200/) and is linked by nothing there; tests, bench.py and smoke() use it to generate the arrays a metriko host would pass."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import scipy.sparse as sp

_LIB = None
_DT = {0: np.float64, 1: np.int32, 2: np.complex128, 3: np.int64}
_DT_INV = {np.dtype(np.float64): 0, np.dtype(np.int32): 1, np.dtype(np.complex128): 2, np.dtype(np.int64): 3}


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmetriko_hostglue.so")
        if not os.path.exists(path):
            from metriko_b200 import build
            build.build_hostglue()
        L = C.CDLL(path)
        L.hg_session_new.restype = C.c_void_p
        L.hg_session_free.argtypes = [C.c_void_p]
        L.hg_last_error.restype = C.c_char_p
        L.hg_last_error.argtypes = [C.c_void_p]
        L.hg_bag_put.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_int64]
        L.hg_bag_get.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_int)]
        L.hg_bag_has.argtypes = [C.c_void_p, C.c_char_p]
        L.hg_make_mesh.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]
        L.hg_set_mesh.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.hg_field_post.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.hg_igm_setup.argtypes = [C.c_void_p, C.c_double]
        L.hg_igm_finish.argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


class Session:
    """One mesh and everything derived from it on the host."""

    def __init__(self):
        self._L = _lib()
        self._h = C.c_void_p(self._L.hg_session_new())

    def __del__(self):
        try:
            if self._h:
                self._L.hg_session_free(self._h)
                self._h = None
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError("hostglue: " + self._L.hg_last_error(self._h).decode())

    # -- bag access -------------------------------------------------------------------------------
    def put(self, name: str, arr) -> None:
        a = np.ascontiguousarray(arr)
        if a.dtype not in _DT_INV:
            raise TypeError(f"unsupported dtype {a.dtype}")
        self._chk(self._L.hg_bag_put(self._h, name.encode(), _DT_INV[a.dtype], a.ctypes.data_as(C.c_void_p), a.size))

    def has(self, name: str) -> bool:
        return bool(self._L.hg_bag_has(self._h, name.encode()))

    def get(self, name: str) -> np.ndarray:
        p = C.c_void_p()
        n = C.c_int64()
        dt = C.c_int()
        self._chk(self._L.hg_bag_get(self._h, name.encode(), C.byref(p), C.byref(n), C.byref(dt)))
        dtype = np.dtype(_DT[dt.value])
        if n.value == 0:
            return np.zeros(0, dtype)
        buf = (C.c_char * (n.value * dtype.itemsize)).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype).copy()

    def scalar(self, name: str):
        return self.get(name)[0]

    def csr(self, name: str) -> sp.csr_matrix:
        shape = tuple(int(v) for v in self.get(name + "_shape"))
        return sp.csr_matrix((self.get(name + "_val"), self.get(name + "_col"), self.get(name + "_ptr")), shape=shape)

    # -- stages -----------------------------------------------------------------------------------
    def make_torus(self, nu: int, nv: int, bump: float = 0.1, phase: float = 0.0) -> "Session":
        self._chk(self._L.hg_make_mesh(self._h, 0, nu, nv, bump, phase))
        return self

    def make_sphere(self, kind: int, levels: int, bump: float = 0.08, phase: float = 0.0) -> "Session":
        """kind = 4 / 8 / 20 (tetra / octa / icosahedron) subdivided `levels` times."""
        self._chk(self._L.hg_make_mesh(self._h, kind, levels, 0, bump, phase))
        return self

    def set_mesh(self, V, F) -> "Session":
        V = np.ascontiguousarray(V, np.float64)
        F = np.ascontiguousarray(F, np.int32)
        self._chk(self._L.hg_set_mesh(self._h, V.ctypes.data_as(C.c_void_p), V.shape[0], F.ctypes.data_as(C.c_void_p), F.shape[0]))
        return self

    @property
    def dims(self):
        nV, nF, nE, nH, euler = (int(v) for v in self.get("dims"))
        return dict(nV=nV, nF=nF, nE=nE, nH=nH, euler=euler)

    def field_inputs(self):
        """The flat arrays connectionLaplacian / galerkinMassMatrix read (frf.h:46-84)."""
        return dict(twin=self.get("twin"), face=self.get("face"), face2half=self.get("face2half"), next=self.get("next"),
                    heArgOnFace=self.get("heArgOnFace"), faceArea=self.get("faceArea"))

    def curvature_inputs(self):
        """Per-half arrays principalCurvatureDir reads (frf.h:86-98): half lengths and dihedral angles (hmesh.h:215, 508-519)."""
        return dict(heLen=self.get("heLen"), dihedralArg=self.get("dihedralArg"))

    def field_post(self, compressed, N: int = 4, matching: str = "curl") -> None:
        self.put("compressed", np.asarray(compressed, np.complex128))
        self._chk(self._L.hg_field_post(self._h, N, 1 if matching == "curl" else 0))

    def igm_setup(self, gridscale: float = 0.02) -> dict:
        self._chk(self._L.hg_igm_setup(self._h, gridscale))
        d = [int(v) for v in self.get("igm_dims")]
        keys = ["N", "n", "nF", "nR", "nT", "nS", "cut_nV", "ufull_cols", "n_constraint_blocks", "cfull_rows", "kernel_dim"]
        return dict(zip(keys, d))

    def igm_inputs(self) -> dict:
        """Everything `iterative_rounding` (ir.h:33-51) receives, plus UFull / X2F (ir_init.h:98-173)."""
        d = [int(v) for v in self.get("igm_dims")]
        return dict(
            N=d[0], n=d[1], nF=d[2], length=float(self.scalar("length")),
            G2p=self.csr("G2p"), Cfull=self.csr("Cfull"), UFull=self.csr("UFull"), X2F=self.csr("X2F"),
            F2=self.get("F2"), fixedIdcs=self.get("fixedIdcs"), fixedVals=self.get("fixedVals"),
            integerIdcs=self.get("integerIdcs"), singularIdcs=self.get("singularIdcs"),
        )

    def igm_mesh_inputs(self) -> dict:
        """What RosyParameterization::integ() itself reads (integ.ipp:3-43), for mk_igm_create_from_mesh: the cut mesh (corner table in F's
        corner order + positions, param.h:126-152), fullx2Nfunc (setup.ipp:52-170, integ.ipp:4) and the per-face combing turns (frf.h:207-249)."""
        cut2raw = self.get("cut2raw")
        pos = self.get("pos").reshape(-1, 3)
        return dict(cutF=self.get("cutF").reshape(-1, 3), cutPos=np.ascontiguousarray(pos[cut2raw]), fullx2Nfunc=self.csr("fullx2Nfunc"),
                    comb_turns=self.get("comb_turns"))

    def igm_finish(self, fullx) -> dict:
        self.put("fullx", np.asarray(fullx, np.float64))
        self._chk(self._L.hg_igm_finish(self._h))
        return dict(flipped_faces=int(self.scalar("flipped_faces")), seam_rot_dev=float(self.scalar("seam_rot_dev")),
                    seam_trans_frac=float(self.scalar("seam_trans_frac")), cfn=self.get("cfn"), nfn=self.get("nfn"))
