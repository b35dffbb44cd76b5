/* metriko_b200.h -- C ABI of libmetriko_b200.so: the B200 (sm_100a) implementation of the sparse-linear-algebra
 * hot path of komietty/metriko (N-RoSy field eigen-solve, IGM Poisson solve, iterative-rounding re-solves).
 *
 * Conventions: plain pointers and sizes, no C++ / torch types; every entry point returns an int
 * status (0 = ok, negative = mk_status error; no exceptions cross the ABI -- the C++ shim in
 * include/metriko_b200/metriko_cuda_shims.hpp rethrows std::runtime_error / returns false as each reference seam
 * does); the caller owns host buffers, the library owns device buffers inside handles; indices are int32, values
 * fp64, complex numbers interleaved (re, im) exactly like std::complex<double> / Eigen::SparseMatrix<complex>;
 * one host thread per context, contexts on different devices are independent. There is NO CPU fallback: every
 * compute entry point fails with MK_ERR_CUDA when no sm_100a device is usable.
 *
 * Each entry point cites the reference interface it replaces (paths relative to include/metriko/core/).
 */
#ifndef METRIKO_B200_H
#define METRIKO_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mk_ctx mk_ctx;       /* one per device + stream */
typedef struct mk_field mk_field;   /* field problem: stencil operator + mass + work vectors */
typedef struct mk_spd mk_spd;       /* real SPD solver handle (LM inner solver) */
typedef struct mk_igm mk_igm;       /* IGM problem: A0 = X2F^T X2F, UFull, rounding state */

enum mk_status {
    MK_OK = 0,
    MK_ERR_CUDA = -1,        /* CUDA runtime error or no device */
    MK_ERR_ARG = -2,         /* bad argument */
    MK_ERR_NOT_CONVERGED = -3, /* inner solve missed the requested residual (result still written) */
    MK_ERR_NOT_SPD = -4,     /* p^T A p <= 0 met (CHOLMOD "matrix is not positive definite", ss/suitesparse_pdefinite.h:38) */
    MK_ERR_ROUNDING = -5,    /* post_checking failed ("Rounding has failed!", igm/parameterization_integ.ipp:85) */
    MK_ERR_INTERNAL = -9
};

typedef struct mk_solve_opts {
    double tol;        /* relative residual ||b - A x|| / ||b|| of every inner solve; default 1e-10 */
    int max_iter;      /* per inner solve; default 100000 */
    int precond;       /* 0 = none, 1 = smoothed-aggregation multigrid (default) */
    int use_graph;     /* replay the PCG iteration as a CUDA graph (default 1) */
    int recycle;       /* eigen-solve: how many earlier solution increments are kept as a Galerkin start space for the next
                          inverse-iteration solve (default 64, capped by the number of outer iterations; 0 = plain warm start); n * 16 B each on
                          the device. Measured at 1M vertices, 50 outer steps: 4018 inner iterations without, 1559 with 32 slots (the space
                          freezes when full and the counts climb again), 1187 with 50 */
} mk_solve_opts;

typedef struct mk_stats {
    int outer_iterations;     /* inverse-iteration steps or rounding re-solves */
    int64_t inner_iterations; /* total PCG iterations */
    int inner_max;            /* max PCG iterations of one solve */
    double relres_max;        /* worst true relative residual over all inner solves */
    double lambda;            /* field: Rayleigh quotient at exit */
    int64_t kernel_launches;  /* kernels launched by the call */
    double setup_ms, solve_ms;/* host wall clock split (setup = upload + assembly + hierarchy) */
    int status;               /* mk_status of the worst inner solve */
} mk_stats;

void mk_default_opts(mk_solve_opts* o);

/* ---- context ---- */
/* `stream` is a cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream) or NULL to create a private stream. */
int mk_ctx_create(int device, void* stream, mk_ctx** out);
int mk_ctx_destroy(mk_ctx* ctx);
const char* mk_last_error(mk_ctx* ctx);
int mk_ctx_sync(mk_ctx* ctx);
int mk_device_count(void);

/* ---- N-RoSy field: FaceRosyField ctor, vectorfield/face_rosy_field.h:17-44 ---- */
/* Upload the half-edge arrays connectionLaplacian()/galerkinMassMatrix() read (frf.h:46-84; hmesh/hmesh.h:98-146).
 * nH = number of halfs, nF = number of faces; closed meshes (README.md:27). */
int mk_field_create(mk_ctx* ctx, int nF, int nH, int N, const int* face2half, const int* next, const int* twin, const int* face,
                    const double* heArgOnFace, const double* faceArea, mk_field** out);
int mk_field_destroy(mk_field* f);
/* connectionLaplacian() + galerkinMassMatrix() (frf.h:46-84) as one atomic-free per-face kernel. */
int mk_field_assemble(mk_field* f);
/* Copy the assembled operator back as CSR (row f: diagonal then the three neighbours) for inspection/tests:
 * rowptr[nF+1], col[4 nF], val[4 nF] complex interleaved, mass[nF]. */
int mk_field_get_operator(mk_field* f, int* rowptr, int* col, double* val, double* mass);
/* solveSmallestEig(L, M, nIter) (solver/sparse_linear_solver.h:29-54): n_iter steps of x = L^{-1}(M u);
 * x /= sqrt|x^H M x|; u = x, starting from the explicit u0 (the reference's VecXc::Random is unseeded).
 * u0, x: host, complex interleaved, length nF. */
int mk_field_solve(mk_field* f, const double* u0, int n_iter, const mk_solve_opts* opts, double* x, mk_stats* stats);
/* Same, with u0 / x already resident on the device (HBM-resident timing path; pointers are device pointers). */
int mk_field_solve_dev(mk_field* f, const double* u0_dev, int n_iter, const mk_solve_opts* opts, double* x_dev, mk_stats* stats);
/* FieldType::CurvatureAligned (frf.h:25-33): rhs = M D / sqrt|D^H M D| with D = principalCurvatureDir() (frf.h:86-98, squared for
 * N = 4) from the per-half lengths and dihedral angles (hmesh/hmesh.h:215, 508-519), then solveSquare(L - lambda M, rhs) with the
 * reference's lambda = 0 (sls.h:16-27) as one multigrid-preconditioned CG solve. heLen, dihedralArg: host, nH; x (and rhs_out, may be
 * NULL): host, complex interleaved, nF. The solution stays on the device for mk_field_decode. */
int mk_field_solve_curvature_aligned(mk_field* f, const double* heLen, const double* dihedralArg, const mk_solve_opts* opts, double* rhs_out, double* x,
                                     mk_stats* stats);
/* solveSquare(lhs, rhs) (sls.h:16-27) for lhs = the assembled L (lambda = 0, frf.h:27): one Hermitian solve, host vectors. */
int mk_field_solve_square(mk_field* f, const double* rhs, const mk_solve_opts* opts, double* x, mk_stats* stats);
/* computeMatching + computeSingular (frf.h:100-202) on the last solution, which stays on the device: per edge the connection
 * (frf.h:109-117), the matching (matching_type 1 = Curl frf.h:169-202, 0 = Principal frf.h:138-167) and the effort; per vertex the
 * singularity index (frf.h:100-107). Inputs are the Hmesh flat arrays (hmesh/hmesh.h:98-146): pos[3 nV], tail / head / edge / prev [nH],
 * edge2face[2 nE], faceBasisX / faceBasisY [3 nF], vert2half[nV], angleDefect[nV]; closed meshes (half 2e is the canonical half of edge e,
 * hmesh.h:378). Outputs (host): matching[nE], effort[nE], singular[nV]. */
int mk_field_matching(mk_field* f, int nV, int nE, int matching_type, const double* pos, const int* tail, const int* head, const int* edge2face,
                      const double* faceBasisX, const double* faceBasisY, const int* edge, const int* prev, const int* vert2half,
                      const double* angleDefect, int* matching, double* effort, int* singular);
/* Field decode (frf.h:37-43): field[nF x N] complex interleaved from the last solution. Host output. */
int mk_field_decode(mk_field* f, double* field);
/* y = L x on host vectors (tests) and a timed device loop for the roofline (reps launches, returns ms per launch). */
int mk_field_spmv(mk_field* f, const double* x, double* y);
int mk_field_spmv_bench(mk_field* f, int reps, int fused_dot, double* ms_per_launch);
/* The multigrid hierarchy built for the last solve (it stands where the reference factorises, sls.h:42-43): number of levels and, for
 * the first `cap` of them, rows, stored entries of A_l and of the prolongator P_l. Used by bench.py for the per-iteration byte count. */
int mk_field_hierarchy(mk_field* f, int cap, int* n_levels, int64_t* rows, int64_t* nnz_a, int64_t* nnz_p);
/* Live timing of the preconditioner of the last solve, replayed `reps` times as a CUDA graph between two events (warm, back to back, as
 * inside a solve): us[0] = one whole application (hand-over + V-cycle + hand-back), us[l] = the cycle from level l downwards, 1 <= l < levels;
 * then, when the fused PCG body is in use, three more entries: the x / r update fused with the hand-over, the cycle with the fused <r,z>
 * reduction, the direction update. */
int mk_field_profile_cycle(mk_field* f, int reps, int cap, int* n_out, double* us);

/* One-shot drop-in for solveSmallestEig(SprsC& L, SprsC& M, int nIter) on Eigen CSC arrays (sls.h:29): L Hermitian
 * n x n compressed CSC (outerIndexPtr/innerIndexPtr/valuePtr), M diagonal given as CSC too (only its diagonal is
 * read, frf.h:78-84). H2D, hierarchy set-up, solve and D2H all inside the call. */
int mk_solve_smallest_eig_csc(mk_ctx* ctx, int n, const int* L_outer, const int* L_inner, const double* L_values,
                              const int* M_outer, const int* M_inner, const double* M_values, const double* u0, int n_iter,
                              const mk_solve_opts* opts, double* x, mk_stats* stats);

/* ---- real SPD solves: the LinearSolver duck type of LMSolver (solver/levenberg_marquardt.h:90,184-189), implemented
 * in the reference by CholeskyWrapper{factorize, solve} = Eigen::SimplicialLLT (igm/iterative_rounding/iter_rounding.h:14-20) ---- */
/* factorize(A): A symmetric n x n CSC (= CSR); uploaded and a multigrid hierarchy is built instead of a factor. */
int mk_spd_factorize(mk_ctx* ctx, int n, const int* outer, const int* inner, const double* values, const mk_solve_opts* opts, mk_spd** out);
/* solve(rhs, x): x_guess may be NULL (zero start). */
int mk_spd_solve(mk_spd* s, const double* rhs, const double* x_guess, double* x, mk_stats* stats);
int mk_spd_destroy(mk_spd* s);
int mk_spd_spmv_bench(mk_spd* s, int reps, int fused_dot, double* ms_per_launch);

/* ---- IGM: iterative_rounding(...) (igm/iterative_rounding/iter_rounding.h:33-113) for localInjectivity = false ---- */
/* X2F = (G2 * UFull).pruned() WITHOUT the length factor (iter_rounding_init.h:173), CSR rows = 2 N nF; UFull CSR
 * (n nR x cols, iter_rounding_init.h:154); F2 = rawF2Vec (parameterization_integ.ipp:30-43). */
int mk_igm_create(mk_ctx* ctx, int rows, int cols, const int64_t* x2f_ptr, const int* x2f_col, const double* x2f_val,
                  int ufull_rows, const int64_t* uf_ptr, const int* uf_col, const double* uf_val, const double* F2,
                  double length, const mk_solve_opts* opts, mk_igm** out);
/* Same, with X2F (ptr / col / val) and F2 already resident on the device (device pointers; the HBM-resident timing path). UFull stays a
 * host matrix: the rounding loop reads candidate rows from it on the host. X2F rows must have strictly increasing columns. */
int mk_igm_create_dev(mk_ctx* ctx, int rows, int cols, const int64_t* x2f_ptr_dev, const int* x2f_col_dev, const double* x2f_val_dev,
                      int ufull_rows, const int64_t* uf_ptr, const int* uf_col, const double* uf_val, const double* F2_dev,
                      double length, const mk_solve_opts* opts, mk_igm** out);
/* RosyParameterization::integ() (igm/parameterization_integ.ipp:3-98) from the arrays integ() itself reads, instead of a host-built X2F:
 * the gradient stencil G2 (integ.ipp:9-28), rawF2Vec (integ.ipp:30-43) and X2F = (G2 * fullx2Nfunc * UFull).pruned() (integ.ipp:72,
 * iter_rounding_init.h:173) are assembled on the device, one thread per row, then the same set-up as mk_igm_create runs.
 *   cutF[3 nF], cutPos[3 cut_nV]  the cut mesh (param.h:126-152): corner table in F's corner order and positions; the face frames are
 *                                 recomputed from them exactly as Hmesh does (hmesh/hmesh.h:461-489)
 *   f2n_*                         fullx2Nfunc = vtrans2cut * uncompress (setup.ipp:52-170, integ.ipp:4): CSR, N cut_nV rows, f2n_cols = n nR
 *   uf_*                          UFull (iter_rounding_init.h:98-154): CSR, f2n_cols rows, ufull_cols columns
 *   F2                            rawF2Vec [2 N nF] on the host, or NULL with (field, comb_turns): the eigen-solve result still resident in
 *                                 `field` and the per-face combing turn of compute_combbed_field (frf.h:207-249), from which F2 is formed on
 *                                 the device (the combed extrinsic field is unit, frf.h:258-261)
 * H2D at 1M vertices: 0.17 GB (0.31 GB with a host F2) against 1.0 GB for mk_igm_create. */
int mk_igm_create_from_mesh(mk_ctx* ctx, int N, int nF, int cut_nV, const int* cutF, const double* cutPos, int f2n_cols, const int64_t* f2n_ptr,
                            const int* f2n_col, const double* f2n_val, int ufull_cols, const int64_t* uf_ptr, const int* uf_col,
                            const double* uf_val, const double* F2, mk_field* field, const int* comb_turns, double length,
                            const mk_solve_opts* opts, mk_igm** out);
/* The device-assembled X2F (CSR, 2 N nF rows; NULL arrays to query nnz) and F2 (may be NULL), for tests. */
int mk_igm_get_x2f(mk_igm* g, int64_t* nnz, int64_t* ptr, int* col, double* val, double* F2);
/* Result scatter (integ.ipp:88-98): NF = fullx2Nfunc * fullx, nfn[cut_nV x N], cfn[nF x 3N] = nfn gathered by the corner table. fullx NULL:
 * the result of the last mk_igm_round / mk_igm_poisson, still on the device. Either output may be NULL. mk_igm_create_from_mesh handles only. */
int mk_igm_scatter(mk_igm* g, const double* fullx, double* nfn, double* cfn);
int mk_igm_destroy(mk_igm* g);
/* NaiveIntegration Poisson solve (iter_rounding_init.h:173-212): min ||length X2F x - F2|| s.t. UFull[fixedIdcs] x = fixedVals.
 * x_small: cols, x0_full: ufull_rows (either may be NULL). */
int mk_igm_poisson(mk_igm* g, int n_fixed, const int* fixedIdcs, const double* fixedVals, double* x_small, double* x0_full, mk_stats* stats);
/* The NI closed form + SeamlessIntegration rounding loop (iter_rounding.h:78-112, iter_rounding_loop.h:29-124).
 * Outputs: fullx[ufull_rows]; fixed_idcs/fixed_vals[n_sing + n_int] in the order dofs were fixed, n_fixed_out their
 * count; resolve_flags[i] = 1 if fixing entry i triggered a re-solve; tie_flags[i] bit0 = rounding near-tie
 * (|frac - 0.5| < 1e-6), bit1 = selection near-tie (runner-up distance within 1e-8 and above the 10e-7 skip threshold), bit2 = re-solve
 * near-tie ("minDif > 10e-7", iter_rounding.h:100, met within a factor 2). Values that close to a decision are evaluated on the REFERENCE's
 * iterate: its 1e10 penalty rows hold the fixed dofs only up to mu / rho, and that slack (up to 1.3e-6 at 1M vertices) is what decides whether
 * a dof the fixed rows already determine triggers a re-solve; the library emulates it to first order in 1 / rho.
 * Returns MK_ERR_ROUNDING when post_checking (iter_rounding_loop.h:89-100) fails. */
int mk_igm_round(mk_igm* g, int n_fixed, const int* fixedIdcs, const double* fixedVals, int n_sing, const int* singularIdcs,
                 int n_int, const int* integerIdcs, int seamless, int roundSeams, double* fullx, int* fixed_idcs, int* fixed_vals,
                 int* n_fixed_out, int* resolve_flags, int* tie_flags, mk_stats* stats);
/* The same driver with localInjectivity = true (iter_rounding.h:57-104; the flags of example_qgp.cpp:26-29 and example_qex.cpp:59-63): every
 * LM carries the per-face injectivity barrier (injective_barrier.h:19-109, iter_rounding_common.h:70-85) and starts from lambda = 0.01 with
 * DiagonalDamping (levenberg_marquardt.h:13-88). The NI stage is the LM over [x; f] (iter_rounding_init.h:60-72) with f eliminated per face;
 * with seamless = 0 its result is returned (iter_rounding.h:86-88). use_iter_cb = 1: the NI LM stops as soon as the map is locally injective
 * (locally_injectivity_check, parameterization_integ.ipp:46-63, iter_rounding_init.h:28) -- needs a mk_igm_create_from_mesh handle. Fixed dofs
 * are eliminated exactly instead of penalised, as in mk_igm_round. */
int mk_igm_round_injective(mk_igm* g, int n_fixed, const int* fixedIdcs, const double* fixedVals, int n_sing, const int* singularIdcs,
                           int n_int, const int* integerIdcs, int seamless, int roundSeams, int use_iter_cb, double* fullx, int* fixed_idcs,
                           int* fixed_vals, int* n_fixed_out, int* resolve_flags, int* tie_flags, mk_stats* stats);
/* Copy A0 = length^2 X2F^T X2F (assembled on the device) back as CSR for tests: pass NULLs to query nnz. */
int mk_igm_get_matrix(mk_igm* g, int64_t* nnz, int* rowptr, int* col, double* val);
int mk_igm_spmv_bench(mk_igm* g, int reps, int fused_dot, double* ms_per_launch);
int mk_igm_hierarchy(mk_igm* g, int cap, int* n_levels, int64_t* rows, int64_t* nnz_a, int64_t* nnz_p);
int mk_igm_profile_cycle(mk_igm* g, int reps, int cap, int* n_out, double* us);

/* ---- partitioned face operator: one process per GPU, NCCL over NVLink (BASELINE configs[3]). The reference has no
 * distributed code; these entry points serve the same recurrence (sls.h:44-51) on a row-partitioned L. The host-side plan
 * (contiguous row blocks, interior-first local order, halo slots grouped by peer) is metriko_b200/partition.py. ---- */
typedef struct mk_dist mk_dist;       /* communicator of the partitioned mode (rank, world, NCCL) */
typedef struct mk_dfield mk_dfield;   /* this rank's rows of the face operator + halo plan + distributed PCG state */
/* id: MK_DIST_ID_BYTES bytes from rank 0's mk_dist_unique_id, distributed by the host application (MPI, torch.distributed, a file). */
#define MK_DIST_ID_BYTES 128
int mk_dist_unique_id(char* id, int id_bytes);
int mk_dist_create(mk_ctx* ctx, int rank, int world, const char* id, int id_bytes, mk_dist** out);
int mk_dist_destroy(mk_dist* d);
/* nbr_local[3][n_local]: column indices into x_ext = [owned rows | halo slots]; val[3][n_local] complex; rows [0, n_interior)
 * reference no halo slot; send_counts / recv_counts[world]; send_idx: local rows to pack, grouped by peer. */
int mk_dfield_create(mk_dist* d, int n_local, int n_interior, int n_halo, const int* nbr_local, const double* val, const double* mass,
                     const int* send_counts, const int* recv_counts, const int* send_idx, mk_dfield** out);
int mk_dfield_destroy(mk_dfield* f);
/* Global multigrid hierarchy for the partitioned solves: every rank passes the SAME global stencil (nbr / val [3][nF], global face
 * ids; the layout mk_field_get_operator reads back) and local_to_global[n_local], the global row of each local row. The hierarchy is
 * built redundantly on every rank by the deterministic device set-up; level 0 runs distributed (halo exchange), the restricted
 * residual is summed with one NCCL all-reduce and the coarse levels are replicated. Collective. */
int mk_dfield_set_global(mk_dfield* f, int nF, const int* nbr, const double* val, const int* local_to_global);
/* y = (L x) on the owned rows: pack -> NCCL send/recv (own stream) overlapped with the interior rows -> boundary rows. Collective. */
int mk_dfield_spmv(mk_dfield* f, const double* x_local, double* y_local);
int mk_dfield_spmv_bench(mk_dfield* f, int reps, int overlap, double* ms_per_launch);
/* one distributed PCG solve L x = b, and the 50-step inverse iteration. opts.precond: 0 none; 1 the global hierarchy of
 * mk_dfield_set_global (falls back to 2 when it was not called); 2 a multigrid V-cycle on each rank's diagonal block (block-Jacobi
 * across ranks). Collective: every rank calls with its rows. */
int mk_dfield_pcg(mk_dfield* f, const double* b_local, const double* x0_local, const mk_solve_opts* opts, double* x_local, mk_stats* stats);
int mk_dfield_solve(mk_dfield* f, const double* u0_local, int n_iter, const mk_solve_opts* opts, double* x_local, mk_stats* stats);
/* mk_dfield_pcg with the owned rows of b and x already in device memory (x_dev: start vector in, solution out): no host <-> device copies. */
int mk_dfield_pcg_dev(mk_dfield* f, const double* b_dev, double* x_dev, const mk_solve_opts* opts, mk_stats* stats);

/* ---- multigrid set-up, exposed for tests. mk_amg_host_build is the host mirror (no device needed) of the device set-up the solvers
 * run: it builds the same smoothed-aggregation hierarchy for a Hermitian (is_complex = 1, interleaved values) or symmetric CSR
 * matrix. cand = near-kernel candidate (real operators) or NULL. which: 0 = A_l, 1 = P_l (level l -> l+1),
 * 2 (level_get only) = the smoother's inverse diagonal of level l, written to val. ---- */
typedef struct mk_amg_host mk_amg_host;
int mk_amg_host_build(int n, const int* ptr, const int* col, const double* val, int is_complex, const double* cand, mk_amg_host** out);
/* The same hierarchy built by the device set-up (what the solvers run); its levels are copied back into the same container so the
 * accessors below serve both. Rows must have strictly increasing columns. */
int mk_amg_dev_build(mk_ctx* ctx, int n, const int* ptr, const int* col, const double* val, int is_complex, mk_amg_host** out);
/* The device set-up algebra, exposed for tests: op 0: C = A B (symbolic/numeric SpGEMM), op 1: C = A^T, op 2: C = A^H. CSR with
 * strictly increasing columns per row; the result is level 0 / which 0 of the returned container (B ignored for op 1, 2). */
int mk_dev_sparse_op(mk_ctx* ctx, int op, int is_complex, int a_rows, int a_cols, const int* a_ptr, const int* a_col, const double* a_val, int b_rows,
                     int b_cols, const int* b_ptr, const int* b_col, const double* b_val, mk_amg_host** out);
/* The same two builds with a dof-type label per row (types[n], may be NULL): dofs of different type never share an aggregate, on any
 * level. mk_igm_create derives the labels of the IGM operator (u-type / v-type reduced dofs) from UFull and uses them internally. */
int mk_amg_host_build_typed(int n, const int* ptr, const int* col, const double* val, int is_complex, const int* types, mk_amg_host** out);
int mk_amg_dev_build_typed(mk_ctx* ctx, int n, const int* ptr, const int* col, const double* val, int is_complex, const int* types, mk_amg_host** out);
int mk_amg_host_num_levels(mk_amg_host* h);
int mk_amg_host_level_info(mk_amg_host* h, int level, int which, int* rows, int* cols, int64_t* nnz, double* lmax);
int mk_amg_host_level_get(mk_amg_host* h, int level, int which, int* ptr, int* col, double* val);
int mk_amg_host_destroy(mk_amg_host* h);

#ifdef __cplusplus
}
#endif
#endif /* METRIKO_B200_H */
