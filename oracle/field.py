"""Oracle: N-RoSy field eigen-solve. Restates include/metriko/core/vectorfield/face_rosy_field.h:17-84
and include/metriko/core/solver/sparse_linear_solver.h:29-54 (test infrastructure, see oracle/__init__.py)."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def random_u0(nF: int, seed: int = 0) -> np.ndarray:
    """Stand-in for `VecXc::Random` (sls.h:32,44): real and imaginary parts uniform in (-1, 1). The
    reference draws from an unseeded std::rand, so u0 is an explicit, shared input here."""
    rng = np.random.default_rng(seed)
    return rng.uniform(-1.0, 1.0, nF) + 1j * rng.uniform(-1.0, 1.0, nF)


def connection_laplacian(twin, face, heArgOnFace, nF: int, N: int) -> sp.csc_matrix:
    """frf.h:46-76. transport[h] = -e^{i(farg(twin h) - farg(h))} (l.52-60, both halves of an edge);
    per face f and half h of f: (f,f) += 1, (f, face(twin h)) += -transport[twin h]^N (l.62-71);
    returns S + 1e-9 I (l.75)."""
    twin = np.asarray(twin)
    face = np.asarray(face)
    farg = np.asarray(heArgOnFace, np.float64)
    nH = twin.shape[0]
    v = np.exp(1j * farg)
    transport = np.empty(nH, np.complex128)
    canon = np.arange(0, nH, 2)  # edge2half[e] = 2e (hmesh.h:378)
    r = -v[twin[canon]] / v[canon]
    transport[canon] = r
    transport[twin[canon]] = 1.0 / r
    h = np.arange(nH)
    rows = np.concatenate([face[h], face[h]])
    cols = np.concatenate([face[h], face[twin[h]]])
    vals = np.concatenate([np.ones(nH, np.complex128), -(transport[twin[h]] ** N)])
    S = sp.coo_matrix((vals, (rows, cols)), shape=(nF, nF)).tocsc()
    return (S + 1e-9 * sp.identity(nF, dtype=np.complex128, format="csc")).tocsc()


def galerkin_mass_matrix(faceArea) -> sp.csc_matrix:
    """frf.h:78-84: diag(faceArea) as a complex sparse matrix."""
    a = np.asarray(faceArea, np.float64).astype(np.complex128)
    return sp.diags(a, format="csc")


def solve_smallest_eig(L, M, u0, n_iter: int = 50):
    """sls.h:42-52 (the no-SuiteSparse branch): factor L once with a sparse LU, then n_iter times
    x = L^{-1}(M u); s = sqrt(|x^H M x|); x /= s; u = x. No convergence test. u0 replaces VecXc::Random."""
    lu = spla.splu(sp.csc_matrix(L))
    u = np.array(u0, np.complex128)
    x = u.copy()
    for _ in range(n_iter):
        x = lu.solve(M @ u)
        s = np.sqrt(abs(np.vdot(x, M @ x)))
        x = x / s
        u = x
    return x


def principal_curvature_dir(face2half, next_, heArgOnFace, heLen, dihedralArg) -> np.ndarray:
    """frf.h:86-98: D_f = 1/4 sum over the halves h of f of -c c / l * darg(h), c = polar(1, farg(h)) * l, l = len(h)."""
    h = np.asarray(face2half, np.int64).copy()
    nxt = np.asarray(next_)
    farg, ln, da = (np.asarray(a, np.float64) for a in (heArgOnFace, heLen, dihedralArg))
    d = np.zeros(len(h), np.complex128)
    for _ in range(3):
        c = np.exp(1j * farg[h]) * ln[h]
        d += -c * c / ln[h] * da[h]
        h = nxt[h]
    return 0.25 * d


def curvature_aligned(L, M, D, N: int):
    """frf.h:25-33 with the reference's lambda = 0: D squared for N = 4, rhs = M D / sqrt|D^H M D|, x = solveSquare(L - 0 M, rhs)
    (sls.h:16-27: a sparse LU solve). Returns (x, rhs)."""
    assert N in (2, 4)
    D = np.asarray(D, np.complex128)
    if N == 4:
        D = D * D
    rhs = (M @ D) / np.sqrt(abs(np.vdot(D, M @ D)))
    x = spla.splu(sp.csc_matrix(L)).solve(rhs)
    return x, rhs


def decode_field(z, N: int) -> np.ndarray:
    """frf.h:37-43: field(f,i) = (e^{i arg z_f})^{1/N} * (e^{2 pi i / N})^i."""
    z = np.asarray(z, np.complex128)
    base = np.exp(1j * np.angle(z) / N)
    rot = np.exp(2j * np.pi * np.arange(N) / N)
    return base[:, None] * rot[None, :]


def field_angle_error(z_a, z_b, N: int) -> np.ndarray:
    """Per-face direction difference in radians after removing the global phase (north_star tolerance
    1e-5 rad). The eigenvector is defined up to a global unit complex factor; the decoded direction is
    arg(z)/N, so the per-face error is |arg(z_a conj(z_b) e^{-i phi})| / N with phi the mean phase."""
    za = np.asarray(z_a, np.complex128)
    zb = np.asarray(z_b, np.complex128)
    g = np.vdot(zb, za)
    g = g / abs(g)
    return np.abs(np.angle(za * np.conj(zb * g))) / N
