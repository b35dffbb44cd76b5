// ref_harness.cpp -- TEST INFRASTRUCTURE (oracle/): a C entry for each piece of the REFERENCE'S OWN hot-path code that can be compiled in this
// image. The reference headers are included where they lie (-I /root/reference/include; nothing is copied); Eigen and libigl, which the image
// does not have, are the stand-ins of oracle/ref_shim/ (read eigen_standin.hpp for what that does and does not pin). oracle/Makefile builds
// this file into oracle/_ref/libmetriko_ref.so; only tests/ and tests/golden/make_golden_ref.py load it.
//
//   ref_iterative_rounding   igm/iterative_rounding/iter_rounding.h:33-113 (NaiveIntegration ir_init.h:74-232, SeamlessIntegration
//                            ir_loop.h:9-125, Integration ir_common.h:12-94, LMSolver / DiagonalDamping lm.h:13-231, InjectiveBarrier
//                            injective_barrier.h:9-109, sparse_block mops.h:32-57)
//   ref_naive_integration    the NaiveIntegration constructor alone: UFull (ir_init.h:98-154), x0 (ir_init.h:173-212), X2F
//   ref_barrier_update       InjectiveBarrier::update + gen_image_filed (injective_barrier.h:37-109)
//   ref_solve_smallest_eig   solver/sparse_linear_solver.h:29-54 (the no-SuiteSparse branch) on given L, M with an explicit u0
#include <cstring>

#include "metriko/core/igm/iterative_rounding/iter_rounding.h"
#include "metriko/core/solver/sparse_linear_solver.h"

using namespace metriko;

namespace {
SprsD from_coo(int rows, int cols, long nnz, const int* r, const int* c, const double* v) {
    std::vector<TripD> t;
    t.reserve(nnz);
    for (long k = 0; k < nnz; k++) t.emplace_back(r[k], c[k], v[k]);
    SprsD m(rows, cols);
    m.setFromTriplets(t.begin(), t.end());
    return m;
}
SprsC from_coo_c(int rows, int cols, long nnz, const int* r, const int* c, const double* v) {
    std::vector<TripC> t;
    t.reserve(nnz);
    for (long k = 0; k < nnz; k++) t.emplace_back(r[k], c[k], complex(v[2 * k], v[2 * k + 1]));
    SprsC m(rows, cols);
    m.setFromTriplets(t.begin(), t.end());
    return m;
}
VecXi ivec(int n, const int* p) { VecXi v(n); for (int i = 0; i < n; i++) v(i) = p[i]; return v; }
VecXd dvec(long n, const double* p) { VecXd v(n); for (long i = 0; i < n; i++) v(i) = p[i]; return v; }
// copy a sparse matrix out as COO (column-major order); returns nnz; arrays may be null to query
long to_coo(const SprsD& m, long cap, int* r, int* c, double* v) {
    long k = 0;
    for (long j = 0; j < m.outerSize(); j++)
        for (SprsD::InnerIterator it(m, j); it; ++it) {
            if (r && k < cap) { r[k] = it.row(); c[k] = it.col(); v[k] = it.value(); }
            k++;
        }
    return k;
}
}  // namespace

extern "C" {

void ref_set_solver_hooks(int (*factor)(int, int, int, const int*, const int*, const double*), int (*solve)(int, const double*, double*),
                          void (*release)(int)) {
    refshim::hooks.factor = factor;
    refshim::hooks.solve = solve;
    refshim::hooks.release = release;
}

long ref_factor_count(int reset) {
    const long c = refshim::factor_count;
    if (reset) refshim::factor_count = 0;
    return c;
}

// iterative_rounding(...) exactly as RosyParameterization::integ calls it (igm/parameterization_integ.ipp:65-83). Cfull and G2 (= G2 *
// fullx2Nfunc, integ.ipp:72) arrive as COO. iter_cb may be null (= never stop early). Returns the function's bool (1 / 0); fullx_len receives
// fullx.size(), fullx (capacity fullx_cap) its coefficients.
int ref_iterative_rounding(int n_fixed, const int* fixedIdcs, const double* fixedVals, int n_sing, const int* singularIdcs, int n_int,
                           const int* integerIdcs, double length, int c_rows, int c_cols, long c_nnz, const int* c_r, const int* c_c,
                           const double* c_v, int g_rows, int g_cols, long g_nnz, const int* g_r, const int* g_c, const double* g_v, int N, int n,
                           int nF, int seamless, int roundSeams, int localInjectivity, int verbose, const double* F2,
                           int (*iter_cb)(const double* x, long len), long fullx_cap, double* fullx, long* fullx_len) {
    const VecXi fi = ivec(n_fixed, fixedIdcs), si = ivec(n_sing, singularIdcs), ii = ivec(n_int, integerIdcs);
    const VecXd fv = dvec(n_fixed, fixedVals), f2 = dvec((long)2 * N * nF, F2);
    const SprsD Cfull = from_coo(c_rows, c_cols, c_nnz, c_r, c_c, c_v), G2 = from_coo(g_rows, g_cols, g_nnz, g_r, g_c, g_v);
    std::function<bool(const VecXd&)> cb = [iter_cb](const VecXd& x) { return iter_cb ? iter_cb(x.data(), x.size()) != 0 : false; };
    VecXd out;
    const bool ok = iterative_rounding(fi, fv, si, ii, length, Cfull, G2, N, n, nF, seamless != 0, roundSeams != 0, localInjectivity != 0,
                                       verbose != 0, f2, cb, out);
    *fullx_len = out.size();
    for (long i = 0; i < out.size() && i < fullx_cap; i++) fullx[i] = out(i);
    return ok ? 1 : 0;
}

// NaiveIntegration's constructor (ir_init.h:74-232) alone. Outputs: UFull and X2F as COO (pass cap 0 / nulls to query the counts), x0
// (UFull.rows()), XSmall (UFull.cols()). dims = {UFull.rows, UFull.cols, nnz(UFull), X2F.rows, nnz(X2F)}.
int ref_naive_integration(int n_fixed, const int* fixedIdcs, const double* fixedVals, double length, int c_rows, int c_cols, long c_nnz,
                          const int* c_r, const int* c_c, const double* c_v, int g_rows, int g_cols, long g_nnz, const int* g_r, const int* g_c,
                          const double* g_v, int N, int n, int nF, const double* F2, long* dims, long u_cap, int* u_r, int* u_c, double* u_v,
                          long x_cap, int* x_r, int* x_c, double* x_v, double* x0, double* xsmall) {
    const VecXi fi = ivec(n_fixed, fixedIdcs);
    const VecXd fv = dvec(n_fixed, fixedVals), f2 = dvec((long)2 * N * nF, F2);
    const SprsD Cfull = from_coo(c_rows, c_cols, c_nnz, c_r, c_c, c_v), G2 = from_coo(g_rows, g_cols, g_nnz, g_r, g_c, g_v);
    InjectiveBarrier barrier(N, nF, f2);
    NaiveIntegration ni(G2, Cfull, f2, fi, fv, length, false, nF, N, n, [](const VecXd&) { return false; }, &barrier);
    dims[0] = ni.UFull.rows();
    dims[1] = ni.UFull.cols();
    dims[2] = to_coo(ni.UFull, u_cap, u_r, u_c, u_v);
    dims[3] = ni.X2F.rows();
    dims[4] = to_coo(ni.X2F, x_cap, x_r, x_c, x_v);
    if (x0) for (long i = 0; i < ni.x0.size(); i++) x0[i] = ni.x0(i);
    if (xsmall) for (long i = 0; i < ni.UFull.cols(); i++) xsmall[i] = ni.XF_Small(i);
    return ni.x0.size() == ni.UFull.rows() ? 1 : 0;  // 0: "Solver failed..." (ir_init.h:203-206) left the members unset
}

// InjectiveBarrier(N, nF, F2).update(field, updateJ, fBarrier, gBarrier) (injective_barrier.h:54-109): fBarrier[N nF], the diagonal of gBarrier
// [N nF] and the 4 entries per row of gen_image_filed(field) (values S, columns J; injective_barrier.h:27-33, 44-52, 85-89).
int ref_barrier_update(int N, int nF, const double* F2, const double* field, int updateJ, double* fBarrier, double* gDiag, int* imgCols,
                       double* imgVals) {
    const VecXd f2 = dvec((long)2 * N * nF, F2), fld = dvec((long)2 * N * nF, field);
    InjectiveBarrier b(N, nF, f2);
    VecXd fB;
    SprsD gB;
    b.update(fld, updateJ != 0, fB, gB);
    for (long i = 0; i < fB.size(); i++) fBarrier[i] = fB(i);
    if (updateJ) {
        for (long i = 0; i < (long)N * nF; i++) gDiag[i] = 0.0;
        for (long j = 0; j < gB.outerSize(); j++)
            for (SprsD::InnerIterator it(gB, j); it; ++it)
                if (it.row() == it.col()) gDiag[it.row()] = it.value();
        // gen_image_filed sums duplicate (row, col) pairs; none occur for N >= 3, and the dense copy below follows the triplet order
        const SprsD img = b.gen_image_filed(fld);
        std::vector<int> fill((size_t)N * nF, 0);
        const SprsD imgT = img.transpose();  // column j of the transpose = row j of img, columns ascending
        for (long j = 0; j < imgT.outerSize(); j++)
            for (SprsD::InnerIterator it(imgT, j); it; ++it) {
                const int k = fill[j]++;
                if (k < 4) { imgCols[4 * j + k] = it.row(); imgVals[4 * j + k] = it.value(); }
            }
    }
    return 1;
}

// solveSmallestEig(SprsC& L, SprsC& M, nIter) (sls.h:29-54, Eigen::SparseLU branch). VecXc::Random is fed from u0 (2 n doubles).
int ref_solve_smallest_eig(int n, long l_nnz, const int* l_r, const int* l_c, const double* l_v, long m_nnz, const int* m_r, const int* m_c,
                           const double* m_v, const double* u0, int nIter, double* x) {
    SprsC L = from_coo_c(n, n, l_nnz, l_r, l_c, l_v), M = from_coo_c(n, n, m_nnz, m_r, m_c, m_v);
    refshim::random_source = u0;
    refshim::random_pos = 0;
    const VecXc r = solveSmallestEig(L, M, nIter);
    refshim::random_source = nullptr;
    for (long i = 0; i < r.size(); i++) { x[2 * i] = r(i).real(); x[2 * i + 1] = r(i).imag(); }
    return 1;
}

}  // extern "C"
