// mini_eigen.hpp -- a stand-in for the slice of Eigen's public API that include/metriko_b200/metriko_cuda_shims.hpp touches, so that the
// whole overlay (seams B1, B2 AND B4: detail::CsrCopy, detail::kernel_basis, cuda::iterative_rounding) goes through a compiler and runs in
// this image, which has no Eigen. Same names, same semantics as Eigen 3.4 for what is used:
//   Eigen::SparseMatrix<T, Options> (compressed, ColMajor default / RowMajor), InnerIterator, setFromTriplets (duplicates summed),
//   operator* (sparse x sparse), pruned(), storage-order converting copy, outerIndexPtr / innerIndexPtr / valuePtr / nonZeros / makeCompressed
//   Eigen::Triplet<T>, Eigen::Matrix<T, Dynamic, 1> (data / size / resize / operator()), dense Eigen::MatrixXd (Zero, operator(), rows, cols)
//   Eigen::FullPivLU<MatrixXd>::kernel()  (complete pivoting, first maximum in column-major order, rank by |pivot| > eps * diagSize * |maxpivot|,
//                                          K = P_c [ -U1^{-1} U2 ; I ], one zero column when the kernel is trivial)
// Test infrastructure only: nothing under metriko_b200/ or include/ includes it.
#pragma once
#include <algorithm>
#include <cmath>
#include <complex>
#include <limits>
#include <numeric>
#include <vector>

namespace Eigen {

enum { ColMajor = 0, RowMajor = 1, Dynamic = -1 };

template <class T>
struct Triplet {
    int r = 0, c = 0;
    T v{};
    Triplet() = default;
    Triplet(int r_, int c_, const T& v_) : r(r_), c(c_), v(v_) {}
    int row() const { return r; }
    int col() const { return c; }
    const T& value() const { return v; }
};

template <class T, int Options = ColMajor>
class SparseMatrix {
public:
    SparseMatrix() : outer_(1, 0) {}
    SparseMatrix(long rows, long cols) : rows_((int)rows), cols_((int)cols), outer_((Options == RowMajor ? rows : cols) + 1, 0) {}
    // storage-order converting copy (what `SparseMatrix<double, RowMajor> R = A;` does in Eigen): a transposition of the index structure
    template <int O2>
    SparseMatrix(const SparseMatrix<T, O2>& o) { assign(o); }
    template <int O2>
    SparseMatrix& operator=(const SparseMatrix<T, O2>& o) { assign(o); return *this; }

    long rows() const { return rows_; }
    long cols() const { return cols_; }
    long outerSize() const { return (long)outer_.size() - 1; }
    long innerSize() const { return Options == RowMajor ? cols_ : rows_; }
    long nonZeros() const { return (long)inner_.size(); }
    void makeCompressed() {}
    const int* outerIndexPtr() const { return outer_.data(); }
    const int* innerIndexPtr() const { return inner_.data(); }
    const T* valuePtr() const { return val_.data(); }
    int* outerIndexPtr() { return outer_.data(); }
    int* innerIndexPtr() { return inner_.data(); }
    T* valuePtr() { return val_.data(); }

    class InnerIterator {
    public:
        InnerIterator(const SparseMatrix& m, long outer) : m_(&m), outer_((int)outer), k_(m.outer_[outer]), end_(m.outer_[outer + 1]) {}
        explicit operator bool() const { return k_ < end_; }
        InnerIterator& operator++() { ++k_; return *this; }
        const T& value() const { return m_->val_[k_]; }
        int index() const { return m_->inner_[k_]; }
        int row() const { return Options == RowMajor ? outer_ : m_->inner_[k_]; }
        int col() const { return Options == RowMajor ? m_->inner_[k_] : outer_; }
    private:
        const SparseMatrix* m_;
        int outer_, k_, end_;
    };

    template <class It>
    void setFromTriplets(It b, It e) {
        const int no = (int)outer_.size() - 1;
        std::vector<std::vector<std::pair<int, T>>> buckets(no);
        for (It t = b; t != e; ++t) {
            const int o = Options == RowMajor ? t->row() : t->col(), i = Options == RowMajor ? t->col() : t->row();
            buckets[o].push_back({i, t->value()});
        }
        inner_.clear();
        val_.clear();
        for (int o = 0; o < no; o++) {
            auto& bk = buckets[o];
            std::stable_sort(bk.begin(), bk.end(), [](const auto& x, const auto& y) { return x.first < y.first; });
            for (size_t k = 0; k < bk.size();) {
                const int i = bk[k].first;
                T s{};
                while (k < bk.size() && bk[k].first == i) s += bk[k++].second;
                inner_.push_back(i);
                val_.push_back(s);
            }
            outer_[o + 1] = (int)inner_.size();
        }
    }

    // entries with |v| <= |ref| * eps dropped (Eigen: pruned(ref = 1, eps = NumTraits::dummy_precision() = 1e-12))
    SparseMatrix pruned(const T& ref = T(1), double eps = 1e-12) const {
        SparseMatrix R(rows_, cols_);
        for (long o = 0; o < outerSize(); o++) {
            for (int k = outer_[o]; k < outer_[o + 1]; k++)
                if (std::abs(val_[k]) > std::abs(ref) * eps) { R.inner_.push_back(inner_[k]); R.val_.push_back(val_[k]); }
            R.outer_[o + 1] = (int)R.inner_.size();
        }
        return R;
    }

    // column-major product C = A * B, column by column with a dense accumulator (inner indices come out sorted)
    friend SparseMatrix operator*(const SparseMatrix& A, const SparseMatrix& B) {
        static_assert(Options == ColMajor, "mini_eigen: product implemented for the default storage order only");
        SparseMatrix Cm(A.rows_, B.cols_);
        std::vector<T> acc(A.rows_, T{});
        std::vector<char> hit(A.rows_, 0);
        std::vector<int> touched;
        for (int j = 0; j < B.cols_; j++) {
            touched.clear();
            for (int kb = B.outer_[j]; kb < B.outer_[j + 1]; kb++) {
                const int k = B.inner_[kb];
                const T b = B.val_[kb];
                for (int ka = A.outer_[k]; ka < A.outer_[k + 1]; ka++) {
                    const int i = A.inner_[ka];
                    if (!hit[i]) { hit[i] = 1; touched.push_back(i); }
                    acc[i] += A.val_[ka] * b;
                }
            }
            std::sort(touched.begin(), touched.end());
            for (int i : touched) { Cm.inner_.push_back(i); Cm.val_.push_back(acc[i]); acc[i] = T{}; hit[i] = 0; }
            Cm.outer_[j + 1] = (int)Cm.inner_.size();
        }
        return Cm;
    }

    template <class, int> friend class SparseMatrix;

private:
    template <int O2>
    void assign(const SparseMatrix<T, O2>& o) {
        rows_ = o.rows_;
        cols_ = o.cols_;
        if (O2 == Options) { outer_ = o.outer_; inner_ = o.inner_; val_ = o.val_; return; }
        const int no = Options == RowMajor ? rows_ : cols_;
        outer_.assign(no + 1, 0);
        for (int i : o.inner_) outer_[i + 1]++;
        for (int i = 0; i < no; i++) outer_[i + 1] += outer_[i];
        inner_.resize(o.inner_.size());
        val_.resize(o.val_.size());
        std::vector<int> pos(outer_.begin(), outer_.end() - 1);
        for (int oo = 0; oo + 1 < (int)o.outer_.size(); oo++)
            for (int k = o.outer_[oo]; k < o.outer_[oo + 1]; k++) {
                const int p = pos[o.inner_[k]]++;
                inner_[p] = oo;
                val_[p] = o.val_[k];
            }
    }
    int rows_ = 0, cols_ = 0;
    std::vector<int> outer_, inner_;
    std::vector<T> val_;
};

template <class T>
class Vector {  // Eigen::Matrix<T, Dynamic, 1>
public:
    Vector() = default;
    explicit Vector(long n) : v_(n) {}
    T* data() { return v_.data(); }
    const T* data() const { return v_.data(); }
    long size() const { return (long)v_.size(); }
    long rows() const { return (long)v_.size(); }
    void resize(long n) { v_.resize(n); }
    T& operator()(long i) { return v_[i]; }
    const T& operator()(long i) const { return v_[i]; }
    T& operator[](long i) { return v_[i]; }
    const T& operator[](long i) const { return v_[i]; }
private:
    std::vector<T> v_;
};

class MatrixXd {  // column-major dense
public:
    MatrixXd() = default;
    MatrixXd(long r, long c) : r_((int)r), c_((int)c), v_((size_t)r * c, 0.0) {}
    static MatrixXd Zero(long r, long c) { return MatrixXd(r, c); }
    long rows() const { return r_; }
    long cols() const { return c_; }
    double& operator()(long i, long j) { return v_[(size_t)j * r_ + i]; }
    const double& operator()(long i, long j) const { return v_[(size_t)j * r_ + i]; }
private:
    int r_ = 0, c_ = 0;
    std::vector<double> v_;
};

template <class M>
class FullPivLU;

template <>
class FullPivLU<MatrixXd> {
public:
    explicit FullPivLU(const MatrixXd& A) : lu_(A), rows_((int)A.rows()), cols_((int)A.cols()) {
        const int size = std::min(rows_, cols_);
        colperm_.resize(cols_);
        std::iota(colperm_.begin(), colperm_.end(), 0);
        nonzero_pivots_ = size;
        maxpivot_ = 0.0;
        for (int k = 0; k < size; k++) {
            // biggest coefficient of the bottom-right corner; column-major visit, the first maximum wins
            int pr = k, pc = k;
            double big = -1.0;
            for (int j = k; j < cols_; j++)
                for (int i = k; i < rows_; i++)
                    if (std::fabs(lu_(i, j)) > big) { big = std::fabs(lu_(i, j)); pr = i; pc = j; }
            if (big == 0.0) { nonzero_pivots_ = k; break; }
            if (big > maxpivot_) maxpivot_ = big;
            if (pr != k) for (int j = 0; j < cols_; j++) std::swap(lu_(k, j), lu_(pr, j));
            if (pc != k) { for (int i = 0; i < rows_; i++) std::swap(lu_(i, k), lu_(i, pc)); std::swap(colperm_[k], colperm_[pc]); }
            for (int i = k + 1; i < rows_; i++) lu_(i, k) /= lu_(k, k);
            for (int j = k + 1; j < cols_; j++)
                for (int i = k + 1; i < rows_; i++) lu_(i, j) -= lu_(i, k) * lu_(k, j);
        }
    }
    double threshold() const { return std::numeric_limits<double>::epsilon() * std::min(rows_, cols_); }
    int rank() const {
        int r = 0;
        for (int i = 0; i < nonzero_pivots_; i++) r += std::fabs(lu_(i, i)) > std::fabs(maxpivot_) * threshold();
        return r;
    }
    MatrixXd kernel() const {
        const int rk = rank(), dimker = cols_ - rk;
        if (dimker == 0) return MatrixXd::Zero(cols_, 1);
        // positions of the accepted pivots, moved to the front (they already are unless a small pivot sits between large ones)
        std::vector<int> piv;
        for (int i = 0; i < nonzero_pivots_; i++)
            if (std::fabs(lu_(i, i)) > std::fabs(maxpivot_) * threshold()) piv.push_back(i);
        std::vector<int> colorder = piv;  // columns of U in the order [pivot columns | the rest]
        {
            std::vector<char> isp(cols_, 0);
            for (int p : piv) isp[p] = 1;
            for (int j = 0; j < cols_; j++) if (!isp[j]) colorder.push_back(j);
        }
        // U1 (rk x rk, upper triangular over the pivot rows/columns), U2 (rk x dimker): X = U1^{-1} U2
        MatrixXd X(rk, dimker);
        for (int j = 0; j < dimker; j++)
            for (int i = rk - 1; i >= 0; i--) {
                double s = u(piv[i], colorder[rk + j]);
                for (int l = i + 1; l < rk; l++) s -= u(piv[i], piv[l]) * X(l, j);
                X(i, j) = s / lu_(piv[i], piv[i]);
            }
        MatrixXd K = MatrixXd::Zero(cols_, dimker);
        for (int i = 0; i < rk; i++)
            for (int j = 0; j < dimker; j++) K(colperm_[colorder[i]], j) = -X(i, j);
        for (int j = 0; j < dimker; j++) K(colperm_[colorder[rk + j]], j) = 1.0;
        return K;
    }
private:
    double u(int i, int j) const { return j >= i ? lu_(i, j) : 0.0; }  // the upper factor
    MatrixXd lu_;
    int rows_, cols_, nonzero_pivots_ = 0;
    double maxpivot_ = 0.0;
    std::vector<int> colperm_;
};

}  // namespace Eigen

// the aliases of metriko/core/common/typedef.h that the overlay uses
namespace metriko {
using VecXc = Eigen::Vector<std::complex<double>>;
using VecXd = Eigen::Vector<double>;
using VecXi = Eigen::Vector<int>;
using SprsC = Eigen::SparseMatrix<std::complex<double>>;
using SprsD = Eigen::SparseMatrix<double>;
using MatXd = Eigen::MatrixXd;
}  // namespace metriko
