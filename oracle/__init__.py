"""CPU oracle of metriko's sparse-linear-algebra hot path (numpy + scipy direct solvers).

TEST INFRASTRUCTURE ONLY. Nothing in the product path (`metriko_b200/`) imports this package; it is used
by `tests/`, by `__graft_entry__.smoke()` and by `bench.py`'s `cpu_baseline` / `--impl reference` legs,
always as the checker or the reported CPU baseline, never as the thing shipped.

PARITY UNPINNED: the reference (komietty/metriko @ f2b618d) ships no tests, golden vectors or fixtures,
and cannot be compiled in this image (Eigen, libigl v2.5.0, SuiteSparse, uuid/uuid.h absent; typedef.h:10
includes a libc++-only header), so this restatement cannot be checked against outputs of the reference
itself. It follows the reference line by line (every function cites the file:line it restates) and is
pinned only by the reference's own runtime invariants. The sparse direct solver is
scipy's SuperLU (`splu`), the closest available relative of `Eigen::SparseLU` (sparse_linear_solver.h:42).
"""
from .field import connection_laplacian, galerkin_mass_matrix, solve_smallest_eig, decode_field, random_u0  # noqa: F401
from .field import principal_curvature_dir, curvature_aligned  # noqa: F401
from . import igm  # noqa: F401
from .igm import iterative_rounding, iterative_rounding_x2f, naive_poisson  # noqa: F401
