"""CPU oracle of metriko's sparse-linear-algebra hot path (numpy + scipy direct solvers).

TEST INFRASTRUCTURE ONLY. Nothing in the product path (`metriko_b200/`) imports this package; it is used
by `tests/`, by `__graft_entry__.smoke()` and by `bench.py`'s `cpu_baseline` / `--impl reference` legs,
always as the checker or the reported CPU baseline, never as the thing shipped.

PARITY: the reference (komietty/metriko @ f2b618d) ships no tests, golden vectors or fixtures, and its build cannot run in
this image (Eigen, libigl v2.5.0, SuiteSparse, uuid/uuid.h absent; typedef.h:10 includes a libc++-only header). Its headers do
compile against stand-ins, though: oracle/Makefile builds oracle/_ref/libmetriko_ref.so from the reference's own hmesh.h,
face_rosy_field.h, parameterization.h (+ setup.ipp, integ.ipp), iter_rounding*.h, levenberg_marquardt.h, injective_barrier.h and
sparse_linear_solver.h (included in place; oracle/ref_harness.cpp, oracle/ref_shim/, binding oracle/ref.py), so both example
programs of the reference run here end to end, and tests/test_oracle_ref.py pins this restatement (and the host glue that
generates every test input) against what THAT code computes (fixtures tests/golden/ref_*.npz, incl. the reference's
icosphere.obj with the example_igm and example_qgp flags). Not covered by that pin: the direct solvers (hooks filled with scipy's
SuperLU -- Eigen's factorisations are absent), SparseQR's row choice in reduce_to_linearly_independent, and libigl's compute_seam
(the seam is an input). DESIGN.md section 5 has the numbers.
The sparse direct solver is scipy's SuperLU (`splu`), the closest available relative of `Eigen::SparseLU` (sparse_linear_solver.h:42).
"""
from .field import connection_laplacian, galerkin_mass_matrix, solve_smallest_eig, decode_field, random_u0  # noqa: F401
from .field import principal_curvature_dir, curvature_aligned  # noqa: F401
from . import igm  # noqa: F401
from .igm import iterative_rounding, iterative_rounding_x2f, naive_poisson  # noqa: F401
