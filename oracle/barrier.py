"""Oracle: the per-face injectivity barrier of the localInjectivity LM path. Restates
include/metriko/core/igm/iterative_rounding/injective_barrier.h:19-109 and the localInjectivity branch of
Integration::compute_jacobian (iter_rounding_common.h:70-85) in numpy. Test infrastructure only (oracle/__init__.py): the checker of
mk_igm_round_injective (csrc/barrier.cuh). Pinned against the reference's own injective_barrier.h compiled in place (oracle/_ref,
tests/test_oracle_ref.py::test_live_barrier_update and the barrier fixtures tests/golden/ref_*.npz)."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


class InjectiveBarrier:
    """injective_barrier.h:10-109. Field layout: 2N doubles per face, [x0, y0, ..., x_{N-1}, y_{N-1}] (integ.ipp:4-43)."""

    S_PARAM = 0.5  # `s` of update(), injective_barrier.h:62

    def __init__(self, N: int, nF: int, F2):
        self.N, self.nF = N, nF
        f = np.asarray(F2, np.float64).reshape(nF, N, 2)
        # V(i, j) = |v_j| |v_{j+1}| of the INPUT field (injective_barrier.h:32-39): the volume the ratio is normalised by
        self.V = np.linalg.norm(f, axis=2) * np.linalg.norm(np.roll(f, -1, axis=1), axis=2)
        i = np.arange(nF)[:, None]
        j = np.arange(N)[None, :]
        base = 2 * N * i
        # J(i N + j, :) = columns of v_j.x, v_j.y, v_{j+1}.x, v_{j+1}.y (injective_barrier.h:23-30)
        self.cols = np.stack([base + 2 * j, base + 2 * j + 1, base + 2 * ((j + 1) % N), base + 2 * ((j + 1) % N) + 1], axis=2).reshape(nF * N, 4)

    def ratio(self, field):
        f = np.asarray(field, np.float64).reshape(self.nF, self.N, 2)
        v0, v1 = f, np.roll(f, -1, axis=1)
        return (v0[..., 0] * v1[..., 1] - v0[..., 1] * v1[..., 0]) / self.V, v0, v1

    def update(self, field, updateJ: bool):
        """Returns (fBarrier, barDer or None, S or None): fBarrier = 1 / bar(r) - 1 with the cubic spline
        bar(r) = (r/s)^3 - 3 (r/s)^2 + 3 r/s, +inf for r <= 0, 0 for r >= s (injective_barrier.h:64-77);
        barDer = -bar'(r) / bar(r)^2, 0 where |fBarrier| < 10e-9, +inf where fBarrier is (injective_barrier.h:79-100);
        S = d r / d (v0, v1) per (face, j) row (injective_barrier.h:84-88)."""
        s = self.S_PARAM
        r, v0, v1 = self.ratio(field)
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            bar = (r / s) ** 3 - 3.0 * (r / s) ** 2 + 3.0 * r / s
            bar2 = 1.0 / bar - 1.0
            bar2 = np.where(r <= 0, np.inf, bar2)
            bar2 = np.where(r >= s, 0.0, bar2)
            fB = bar2.reshape(-1)
            if not updateJ:
                return fB, None, None
            d = 3.0 * (r * r / s ** 3) - 6.0 * (r / s ** 2) + 3.0 / s
            d = np.where(r <= 0, np.inf, d)
            d = np.where(r >= s, 0.0, d)
            barDer = (-d / (bar * bar)).reshape(-1)
        barDer = np.where(np.abs(fB) < 10e-9, 0.0, barDer)
        barDer = np.where(fB == np.inf, np.inf, barDer)
        V = self.V
        S = np.stack([v1[..., 1] / V, -v1[..., 0] / V, -v0[..., 1] / V, v0[..., 0] / V], axis=2).reshape(-1, 4)
        return fB, barDer, S

    def image_field(self, S, ncols: int):
        """gen_image_filed (injective_barrier.h:42-51): (N nF) x ncols sparse matrix with S at the four field columns of each row."""
        rows = np.repeat(np.arange(self.nF * self.N), 4)
        return sp.csr_matrix((S.reshape(-1), (rows, self.cols.reshape(-1))), shape=(self.nF * self.N, ncols))

    def jacobian_block(self, field, right, weight: float):
        """diag(barDer) * imageField(field) * right * weight -- the gBarrier row block of compute_jacobian
        (iter_rounding_common.h:76-79): right = X2F, weight = wBarrier in the rounding loop; right = gClose, weight = wInteg (sic) in NI."""
        _, barDer, S = self.update(field, True)
        G = sp.diags(barDer) @ self.image_field(S, len(field))
        return (G @ right * weight).tocsr()
