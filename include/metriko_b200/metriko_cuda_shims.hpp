// metriko_cuda_shims.hpp -- header overlay that plugs libmetriko_b200.so into metriko's own solver seams.
//
// metriko selects its sparse back end at compile time (`#ifdef GC_HAVE_SUITESPARSE`, solver/sparse_linear_solver.h:10-13,17,30).
// This overlay adds one more switch, METRIKO_HAVE_CUDA, and supplies drop-ins with the reference's own signatures:
//
//   B1  metriko::cuda::solveSmallestEig(SprsC& L, SprsC& M, int nIter, const VecXc& u0)   <- sls.h:29
//   B2  metriko::cuda::CholeskyWrapper { bool factorize(const SprsD&); bool solve(const VecXd&, VecXd&) const; }   <- ir.h:14-20
//   B4  metriko::cuda::iterative_rounding(...)  same argument list as ir.h:33-51 (localInjectivity must be false)
//
// Only Eigen's public accessors are used (rows(), outerIndexPtr(), innerIndexPtr(), valuePtr(), makeCompressed(), data(), resize(); for B4 also
// InnerIterator, setFromTriplets, operator*, pruned(), the RowMajor converting copy and FullPivLU::kernel()), so the overlay compiles against any
// Eigen 3.3+/libigl the metriko build already fetches. All three seams are compiled and run in this repo against tests/cpp/mini_eigen.hpp
// (tests/test_shim.py; the image has no Eigen). Errors follow each seam's convention:
// B1 throws std::runtime_error like the CHOLMOD path (suitesparse/suitesparse_pdefinite.h:35-40), B2/B4 return false.
#pragma once
#include <complex>
#include <cstdio>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "metriko_b200.h"

#ifndef METRIKO_B200_SHIM_TYPES_PROVIDED
#include "metriko/core/common/typedef.h"  // VecXc, VecXd, VecXi, SprsC, SprsD, MatXd
#endif

namespace metriko::cuda {

// One context (device + stream) per host thread; metriko itself is single-threaded.
class Context {
public:
    explicit Context(int device = 0) {
        const int rc = mk_ctx_create(device, nullptr, &h_);
        if (rc != MK_OK) throw std::runtime_error("metriko_b200: no usable sm_100a device (mk_ctx_create = " + std::to_string(rc) + ")");
    }
    ~Context() { mk_ctx_destroy(h_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    mk_ctx* get() const { return h_; }
    std::string last_error() const { return mk_last_error(h_); }
    static Context& thread_default() {
        thread_local Context c(0);
        return c;
    }

private:
    mk_ctx* h_ = nullptr;
};

// B1: VecXc solveSmallestEig(SprsC& L, SprsC& M, int nIter = 50)   (sls.h:29-54)
// u0 replaces VecXc::Random (unseeded std::rand in the reference): pass VecXc::Random(L.rows()) to keep the old behaviour.
inline VecXc solveSmallestEig(SprsC& L, SprsC& M, int nIter, const VecXc& u0, mk_stats* stats = nullptr, Context& ctx = Context::thread_default()) {
    L.makeCompressed();  // as toCholmod does, suitesparse/suitesparse_utilities.h:81
    M.makeCompressed();
    const int n = (int)L.rows();
    VecXc x(n);
    static_assert(sizeof(std::complex<double>) == 2 * sizeof(double), "interleaved complex layout");
    const int rc = mk_solve_smallest_eig_csc(ctx.get(), n, L.outerIndexPtr(), L.innerIndexPtr(), reinterpret_cast<const double*>(L.valuePtr()),
                                             M.outerIndexPtr(), M.innerIndexPtr(), reinterpret_cast<const double*>(M.valuePtr()),
                                             reinterpret_cast<const double*>(u0.data()), nIter, nullptr, reinterpret_cast<double*>(x.data()), stats);
    // MK_ERR_NOT_CONVERGED: the inner solves stopped at the fp64 floor above the requested residual (meshes beyond ~4M vertices); x is the
    // best iterate and the reference, which runs no convergence test at all (sls.h:44-51), would not have noticed -- warn, do not throw
    if (rc == MK_ERR_NOT_CONVERGED) std::fprintf(stderr, "metriko_b200 solveSmallestEig: inner residual above tolerance (%s)\n", ctx.last_error().c_str());
    else if (rc != MK_OK) throw std::runtime_error("metriko_b200 solveSmallestEig failed (" + std::to_string(rc) + "): " + ctx.last_error());
    return x;
}

// B2: the LinearSolver duck type of LMSolver<LinearSolver, SolverTraits, DampingTraits> (levenberg_marquardt.h:90,184-189)
class CholeskyWrapper {
public:
    explicit CholeskyWrapper(Context& ctx = Context::thread_default()) : ctx_(&ctx) {}
    ~CholeskyWrapper() { release(); }
    CholeskyWrapper(const CholeskyWrapper&) = delete;
    CholeskyWrapper& operator=(const CholeskyWrapper&) = delete;

    // A is an Eigen temporary (dampJ^T * dampJ, lm.h:184): it is uploaded here, nothing of it is referenced afterwards.
    bool factorize(const SprsD& A) {
        release();
        SprsD Ac = A;
        Ac.makeCompressed();
        n_ = (int)Ac.rows();
        return mk_spd_factorize(ctx_->get(), n_, Ac.outerIndexPtr(), Ac.innerIndexPtr(), Ac.valuePtr(), nullptr, &h_) == MK_OK;
    }
    bool solve(const VecXd& rhs, VecXd& x) const {
        if (!h_ || rhs.size() != n_) return false;
        x.resize(n_);
        return mk_spd_solve(h_, rhs.data(), nullptr, x.data(), nullptr) == MK_OK;
    }

private:
    void release() {
        if (h_) mk_spd_destroy(h_);
        h_ = nullptr;
    }
    Context* ctx_;
    mk_spd* h_ = nullptr;
    int n_ = 0;
};

#ifndef METRIKO_B200_SHIM_NO_B4
namespace detail {
// CSR arrays (int64 row pointers) of a column-major Eigen matrix, through its row-major copy
struct CsrCopy {
    std::vector<int64_t> ptr;
    std::vector<int> col;
    std::vector<double> val;
    int rows = 0, cols = 0;
    explicit CsrCopy(const SprsD& A) {
        Eigen::SparseMatrix<double, Eigen::RowMajor> R = A;
        R.makeCompressed();
        rows = (int)R.rows();
        cols = (int)R.cols();
        ptr.assign(R.outerIndexPtr(), R.outerIndexPtr() + rows + 1);
        col.assign(R.innerIndexPtr(), R.innerIndexPtr() + R.nonZeros());
        val.assign(R.valuePtr(), R.valuePtr() + R.nonZeros());
    }
};

// UFull = basis of ker(Cfull): identity on the columns no node constraint touches, a dense FullPivLU kernel on the few that
// are touched (what NaiveIntegration's constructor builds, iter_rounding_init.h:98-154). Host-side, tiny.
inline SprsD kernel_basis(const SprsD& Cfull) {
    const int nc = (int)Cfull.cols();
    std::vector<char> used(nc, 0);
    for (int k = 0; k < Cfull.outerSize(); ++k)
        for (SprsD::InnerIterator it(Cfull, k); it; ++it) used[it.col()] = 1;
    std::vector<int> part, free_cols, mask(nc, -1);
    for (int j = 0; j < nc; j++) (used[j] ? part : free_cols).push_back(j);
    for (size_t i = 0; i < part.size(); i++) mask[part[i]] = (int)i;
    MatXd K(0, 0);
    if (Cfull.rows() != 0 && !part.empty()) {
        MatXd Cs = MatXd::Zero(Cfull.rows(), (int)part.size());
        for (int k = 0; k < Cfull.outerSize(); ++k)
            for (SprsD::InnerIterator it(Cfull, k); it; ++it) Cs(it.row(), mask[it.col()]) = it.value();
        Eigen::FullPivLU<MatXd> lu(Cs);
        K = lu.kernel();
    }
    std::vector<Eigen::Triplet<double>> T;
    for (size_t i = 0; i < free_cols.size(); i++) T.emplace_back(free_cols[i], (int)i, 1.0);
    for (int i = 0; i < K.rows(); i++)
        for (int j = 0; j < K.cols(); j++) T.emplace_back(part[i], (int)free_cols.size() + j, K(i, j));
    SprsD U(nc, (int)free_cols.size() + (int)K.cols());
    U.setFromTriplets(T.begin(), T.end());
    return U;
}
}  // namespace detail

// B4: bool iterative_rounding(...)   (iter_rounding.h:33-113). Same arguments except iter_cb: with localInjectivity = true the barrier LMs
// (injective_barrier.h) run on the device too (mk_igm_round_injective), but the reference's host callback (the NI LM's early stop
// "locally injective now", iter_rounding_init.h:28) cannot be evaluated from here -- the integ()-level entry mk_igm_create_from_mesh +
// mk_igm_round_injective(use_iter_cb = 1) evaluates that check on the device (seam B6).
inline bool iterative_rounding(const VecXi& fixedIdcs, const VecXd& fixedVals, const VecXi& singularIdcs, const VecXi& integerIdcs,
                               const double length, const SprsD& Cfull, const SprsD& G2, const int N, const int n, const int nF,
                               const bool seamless, const bool roundSeams, const bool localInjectivity, const bool verbose,
                               const VecXd& intrinsicField, VecXd& fullx, Context& ctx = Context::thread_default()) {
    (void)N; (void)n; (void)nF; (void)verbose;
    const SprsD UFull = detail::kernel_basis(Cfull);
    SprsD X2F = (G2 * UFull).pruned();  // iter_rounding_init.h:173 (without `length`)
    detail::CsrCopy X(X2F), U(UFull);
    mk_igm* g = nullptr;
    if (mk_igm_create(ctx.get(), X.rows, X.cols, X.ptr.data(), X.col.data(), X.val.data(), U.rows, U.ptr.data(), U.col.data(), U.val.data(),
                      intrinsicField.data(), length, nullptr, &g) != MK_OK)
        return false;
    fullx.resize(U.rows);
    const int m = (int)(singularIdcs.size() + integerIdcs.size());
    std::vector<int> oi(m + 1), ov(m + 1), rf(m + 1), tf(m + 1);
    int nfix = 0;
    const int rc = localInjectivity
        ? mk_igm_round_injective(g, (int)fixedIdcs.size(), fixedIdcs.data(), fixedVals.data(), (int)singularIdcs.size(), singularIdcs.data(),
                                 (int)integerIdcs.size(), integerIdcs.data(), seamless ? 1 : 0, roundSeams ? 1 : 0, 0, fullx.data(), oi.data(),
                                 ov.data(), &nfix, rf.data(), tf.data(), nullptr)
        : mk_igm_round(g, (int)fixedIdcs.size(), fixedIdcs.data(), fixedVals.data(), (int)singularIdcs.size(), singularIdcs.data(),
                       (int)integerIdcs.size(), integerIdcs.data(), seamless ? 1 : 0, roundSeams ? 1 : 0, fullx.data(), oi.data(),
                       ov.data(), &nfix, rf.data(), tf.data(), nullptr);
    mk_igm_destroy(g);
    return rc == MK_OK;  // MK_ERR_ROUNDING <-> "Rounding has failed!" (parameterization_integ.ipp:85)
}

// The reference's EXACT argument list (iter_rounding.h:33-51), so that the call in parameterization_integ.ipp:65-83 compiles unchanged against the
// overlay (a one-word edit there, or `#define iterative_rounding cuda::iterative_rounding` around the include: tests/cpp/ref_overlay_test.cpp
// compiles the reference's own headers that way). iter_cb is accepted and not called (see above).
inline bool iterative_rounding(const VecXi& fixedIdcs, const VecXd& fixedVals, const VecXi& singularIdcs, const VecXi& integerIdcs,
                               const double length, const SprsD& Cfull, const SprsD& G2, const int N, const int n, const int nF,
                               const bool seamless, const bool roundSeams, const bool localInjectivity, const bool verbose,
                               const VecXd& intrinsicField, const std::function<bool(const VecXd&)>& /*iter_cb*/, VecXd& fullx) {
    return iterative_rounding(fixedIdcs, fixedVals, singularIdcs, integerIdcs, length, Cfull, G2, N, n, nF, seamless, roundSeams, localInjectivity,
                              verbose, intrinsicField, fullx, Context::thread_default());
}
#endif  // METRIKO_B200_SHIM_NO_B4

}  // namespace metriko::cuda
