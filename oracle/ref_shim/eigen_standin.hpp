// eigen_standin.hpp -- TEST INFRASTRUCTURE (oracle/): a stand-in for the slice of Eigen's public API that the reference's own headers
//   include/metriko/core/hmesh/hmesh.h, vectorfield/{base_field.h, face_rosy_field.h}, igm/{parameterization.h, parameterization_setup.ipp,
//   parameterization_integ.ipp}, igm/iterative_rounding/{iter_rounding.h, iter_rounding_init.h, iter_rounding_loop.h, iter_rounding_common.h,
//   injective_barrier.h}, solver/{levenberg_marquardt.h, matrix_ops.h, sparse_linear_solver.h}, common/typedef.h
// touch, so that THOSE FILES THEMSELVES (included in place from /root/reference, never copied) go through g++ and run in this image, which
// has no Eigen. oracle/Makefile builds oracle/_ref/libmetriko_ref.so from oracle/ref_harness.cpp with this directory
// in front of the include path; tests/test_oracle_ref.py pins the numpy restatement (oracle/igm.py, oracle/barrier.py, oracle/field.py) and
// the host glue (tests/hostglue) against what the reference's code computes.
//
// What is NOT Eigen here, and therefore what the pin does not cover:
//   * everything is evaluated eagerly (no expression templates): same values, same operation order per coefficient for the operations
//     used (sums run in index order as Eigen's scalar paths do; Eigen's vectorised reductions may associate differently: last-bit effects);
//   * SparseLU / SimplicialLLT / SimplicialLDLT are hooks the harness' caller fills (tests: scipy's SuperLU, the solver the restated
//     oracle uses as well) -- Eigen's own factorisations are absent from the image, any direct solver is a stand-in;
//   * FullPivLU follows Eigen 3.4's complete pivoting (first maximum in column-major order, rank threshold eps * diagSize * |maxpivot|,
//     kernel = P_c [-U1^{-1} U2; I]) so that UFull comes out with Eigen's column choice;
//   * SparseQR (reduce_to_linearly_independent, matrix_ops.h:12-30) picks independent columns in natural order (not COLAMD / Householder):
//     an equivalent set of constraint rows, the same one tests/hostglue picks;
//   * Matrix::Random() reads from refshim::random_source when the harness sets it (the reference's VecXc::Random is unseeded: u0 is an
//     explicit input everywhere in this repository).
#pragma once
// frf.h:150 calls an unqualified abs() on doubles. With only <cmath> / <cstdlib>, libstdc++ leaves just the C `int abs(int)` in the global
// namespace and the call silently truncates (seen: a principal matching of -1 instead of 3 on one edge); the reference's own toolchain (libc++,
// typedef.h:10) and real Eigen put the floating-point overloads there. The C++ wrappers of these two headers do the same.
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <cassert>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <initializer_list>
#include <iostream>
#include <limits>
#include <memory>
#include <numeric>
#include <queue>
#include <set>
#include <type_traits>
#include <vector>

namespace refshim {
// direct-solver hooks: kind 0 = general LU (Eigen::SparseLU), 1 = SPD (SimplicialLLT / LDLT); is_complex: values interleaved (re, im).
// The matrix arrives as compressed columns. factor returns a handle >= 0 or -1 on failure.
struct SolverHooks {
    int (*factor)(int kind, int is_complex, int n, const int* outer, const int* inner, const double* values) = nullptr;
    int (*solve)(int handle, const double* rhs, double* x) = nullptr;
    void (*release)(int handle) = nullptr;
};
inline SolverHooks hooks{};
inline const double* random_source = nullptr;  // consumed by Matrix::Random (doubles; complex takes two per coefficient)
inline long random_pos = 0;
inline long factor_count = 0;  // factorisations requested since the harness reset it
}  // namespace refshim

namespace Eigen {

enum { ColMajor = 0, RowMajor = 1 };
constexpr int Dynamic = -1;
constexpr int Infinity = -1;
enum ComputationInfo { Success = 0, NumericalIssue = 1, NoConvergence = 2, InvalidInput = 3 };

template <class T>
struct Triplet {
    int r = 0, c = 0;
    T v{};
    Triplet() = default;
    Triplet(int r_, int c_, const T& v_ = T{}) : r(r_), c(c_), v(v_) {}
    int row() const { return r; }
    int col() const { return c; }
    const T& value() const { return v; }
};

template <class T>
inline double abs2_(const T& x) { return (double)x * (double)x; }
template <class T>
inline double abs2_(const std::complex<T>& x) { return std::norm(x); }

template <class T, int R, int C>
class Matrix;
template <class M>
struct Map;

// coefficient-wise view (a.array()): owns its values
template <class T>
struct ArrayX {
    std::vector<T> v;
    long r = 0, c = 0;
    ArrayX array() const { return *this; }
    ArrayX& operator/=(const T& s) { for (auto& x : v) x = x / s; return *this; }
    ArrayX square() const { ArrayX o = *this; for (auto& x : o.v) x = x * x; return o; }
    friend ArrayX operator*(const ArrayX& a, const T& s) { ArrayX o = a; for (auto& x : o.v) x = x * s; return o; }
    ArrayX operator-() const { ArrayX o = *this; for (auto& x : o.v) x = -x; return o; }
    friend ArrayX operator*(const ArrayX& a, const ArrayX& b) { ArrayX o = a; for (size_t i = 0; i < o.v.size(); i++) o.v[i] = a.v[i] * b.v[i]; return o; }
    friend ArrayX operator/(const ArrayX& a, const ArrayX& b) { ArrayX o = a; for (size_t i = 0; i < o.v.size(); i++) o.v[i] = a.v[i] / b.v[i]; return o; }
};

// `m << a, b, c;` -- Eigen's comma initializer: scalars and blocks are placed left to right, a full row of blocks moves the cursor down by the
// block height; vector targets take vector arguments coefficient by coefficient whatever their orientation
template <class T>
class CommaInit {
public:
    CommaInit(std::function<T&(long, long)> at, long rows, long cols) : at_(std::move(at)), rows_(rows), cols_(cols) {}
    CommaInit& operator,(const T& s) { put(s); return *this; }
    template <int R2, int C2>
    CommaInit& operator,(const Matrix<T, R2, C2>& m) { put_block(m); return *this; }
    void put(const T& s) {
        wrap(1);
        check(row_, col_);
        at_(row_, col_++) = s;
    }
    template <int R2, int C2>
    void put_block(const Matrix<T, R2, C2>& m) {
        const bool target_vec = rows_ == 1 || cols_ == 1, arg_vec = m.rows() == 1 || m.cols() == 1;
        if (target_vec && arg_vec) { for (long i = 0; i < m.size(); i++) { if (cols_ == 1) { wrap(1); check(row_, 0); at_(row_, 0) = m(i); col_ = 1; } else put(m(i)); } return; }
        wrap(m.rows());
        for (long i = 0; i < m.rows(); i++) for (long j2 = 0; j2 < m.cols(); j2++) { check(row_ + i, col_ + j2); at_(row_ + i, col_ + j2) = m(i, j2); }
        col_ += m.cols();
    }
private:
    void wrap(long next_block_rows) { if (col_ >= cols_) { row_ += block_rows_; col_ = 0; } if (col_ == 0) block_rows_ = next_block_rows; }
    void check(long i, long j2) const { if (i >= rows_ || j2 >= cols_) { std::fprintf(stderr, "eigen_standin: comma initializer overflow\n"); std::abort(); } }
    std::function<T&(long, long)> at_;
    long rows_, cols_, row_ = 0, col_ = 0, block_rows_ = 1;
};

template <class T, int R, int C>
class Matrix {
public:
    static constexpr bool kIsVector = (R == 1 || C == 1);
    using Scalar = T;
    using Index = long;

    Matrix() { init(R == Dynamic ? 0 : R, C == Dynamic ? 0 : C); }
    // vectors: Matrix(n); matrices: Matrix(rows, cols)
    explicit Matrix(long n) {
        if (C == 1) init(n, 1); else if (R == 1) init(1, n); else init(n, n);
    }
    Matrix(long r, long c) { init(r, c); }
    Matrix(const Matrix&) = default;
    Matrix(Matrix&&) = default;
    Matrix& operator=(const Matrix&) = default;
    Matrix& operator=(Matrix&&) = default;
    template <int R2, int C2>
    Matrix(const Matrix<T, R2, C2>& o) { assign(o); }
    template <int R2, int C2>
    Matrix& operator=(const Matrix<T, R2, C2>& o) { assign(o); return *this; }
    // real -> complex (VecXc connection = VecXd::Zero(n), frf.h:109): Eigen 3.3+ assigns across scalar types that convert
    template <class T2, int R2, int C2, class = std::enable_if_t<!std::is_same<T2, T>::value && std::is_convertible<T2, T>::value>>
    Matrix(const Matrix<T2, R2, C2>& o) { init_like(o.rows(), o.cols()); for (long i = 0; i < o.size(); i++) v_[i] = T(o(i)); }
    Matrix(const ArrayX<T>& a) { init_like(a.r, a.c); v_ = a.v; }
    Matrix(const Map<Matrix>& m);

    long rows() const { return r_; }
    long cols() const { return c_; }
    long size() const { return r_ * c_; }
    T* data() { return v_.data(); }
    const T* data() const { return v_.data(); }
    void resize(long n) { if (R == 1) init(1, n); else init(n, 1); }
    void resize(long r, long c) { init(r, c); }
    void conservativeResize(long n) { if (R == 1) { r_ = 1; c_ = n; } else { r_ = n; c_ = 1; } v_.resize((size_t)n, T{}); }
    Matrix eval() const { return *this; }

    T& operator()(long i) { return v_[i]; }
    const T& operator()(long i) const { return v_[i]; }
    T& operator[](long i) { return v_[i]; }
    const T& operator[](long i) const { return v_[i]; }
    T& operator()(long i, long j) { return v_[(size_t)j * r_ + i]; }
    const T& operator()(long i, long j) const { return v_[(size_t)j * r_ + i]; }

    static Matrix Zero() { Matrix m; std::fill(m.v_.begin(), m.v_.end(), T{}); return m; }
    static Matrix Ones(long n) { Matrix m(n); std::fill(m.v_.begin(), m.v_.end(), T(1)); return m; }
    static Matrix Zero(long n) { Matrix m(n); std::fill(m.v_.begin(), m.v_.end(), T{}); return m; }
    static Matrix Zero(long r, long c) { Matrix m(r, c); std::fill(m.v_.begin(), m.v_.end(), T{}); return m; }
    static Matrix Constant(long n, const T& x) { Matrix m(n); std::fill(m.v_.begin(), m.v_.end(), x); return m; }
    static Matrix Constant(long r, long c, const T& x) { Matrix m(r, c); std::fill(m.v_.begin(), m.v_.end(), x); return m; }
    static Matrix Random(long n) {
        Matrix m(n);
        for (auto& x : m.v_) x = rnd(static_cast<T*>(nullptr));
        return m;
    }

    double squaredNorm() const { double s = 0; for (const auto& x : v_) s += abs2_(x); return s; }
    double norm() const { return std::sqrt(squaredNorm()); }
    template <int P>
    double lpNorm() const {
        static_assert(P == Infinity, "eigen_standin: only lpNorm<Infinity>");
        double m = 0;
        for (const auto& x : v_) m = std::max(m, (double)std::abs(x));
        return m;
    }
    T maxCoeff() const { return *std::max_element(v_.begin(), v_.end()); }
    Matrix normalized() const { const double n = norm(); Matrix o = *this; if (n > 0) for (auto& x : o.v_) x = x / n; return o; }
    void normalize() { *this = normalized(); }
    template <int R2, int C2>
    Matrix cross(const Matrix<T, R2, C2>& b) const {
        Matrix o = *this;
        o.v_[0] = v_[1] * b(2) - v_[2] * b(1);
        o.v_[1] = v_[2] * b(0) - v_[0] * b(2);
        o.v_[2] = v_[0] * b(1) - v_[1] * b(0);
        return o;
    }
    Matrix<T, C, R> adjoint() const {
        Matrix<T, C, R> o(c_, r_);
        for (long i = 0; i < r_; i++) for (long j = 0; j < c_; j++) o(j, i) = conj_((*this)(i, j));
        return o;
    }
    // dense product (only small ones occur: row vector times column vector)
    template <int R2, int C2>
    Matrix<T, R, C2> operator*(const Matrix<T, R2, C2>& b) const {
        Matrix<T, R, C2> o(r_, b.cols());
        for (long i = 0; i < r_; i++) for (long j = 0; j < b.cols(); j++) { T acc{}; for (long k = 0; k < c_; k++) acc += (*this)(i, k) * b(k, j); o(i, j) = acc; }
        return o;
    }
    template <int R2, int C2>
    Matrix& operator+=(const Matrix<T, R2, C2>& b) { for (long i = 0; i < size(); i++) v_[i] += b(i); return *this; }
    template <int R2, int C2>
    Matrix& operator-=(const Matrix<T, R2, C2>& b) { for (long i = 0; i < size(); i++) v_[i] -= b(i); return *this; }
    // x.dot(y) = sum conj(x_i) y_i
    template <int R2, int C2>
    T dot(const Matrix<T, R2, C2>& o) const {
        T s{};
        for (long i = 0; i < size(); i++) s += conj_(v_[i]) * o(i);
        return s;
    }

    using DynVec = Matrix<T, (R == 1 ? 1 : Dynamic), (R == 1 ? Dynamic : 1)>;
    DynVec segment(long o, long n) const { DynVec s(n); for (long i = 0; i < n; i++) s(i) = v_[o + i]; return s; }
    DynVec head(long n) const { return segment(0, n); }
    DynVec tail(long n) const { return segment(size() - n, n); }

    using RowVec = Matrix<T, 1, Dynamic>;
    struct RowRef;
    struct RowArr {  // m.row(i).array()
        Matrix* m;
        long i;
        RowArr& operator+=(const ArrayX<T>& a) { for (long j = 0; j < m->c_; j++) (*m)(i, j) += a.v[j]; return *this; }
        ArrayX<T> operator*(const T& s) const { ArrayX<T> a; a.r = 1; a.c = m->c_; for (long j = 0; j < m->c_; j++) a.v.push_back((*m)(i, j) * s); return a; }
    };
    struct RowRef {
        Matrix* m;
        long i;
        template <int R2, int C2>
        RowRef& operator=(const Matrix<T, R2, C2>& o) { for (long j = 0; j < m->c_; j++) (*m)(i, j) = o(j); return *this; }
        RowVec vec() const { RowVec v(m->c_); for (long j = 0; j < m->c_; j++) v(j) = (*m)(i, j); return v; }
        template <int R2, int C2>
        operator Matrix<T, R2, C2>() const { return Matrix<T, R2, C2>(vec()); }
        T& operator[](long j) { return (*m)(i, j); }
        RowArr array() { return RowArr{m, i}; }
        friend RowVec operator-(const RowRef& a, const RowRef& b) { return a.vec() - b.vec(); }
        friend RowVec operator+(const RowRef& a, const RowRef& b) { return a.vec() + b.vec(); }
        CommaInit<T> comma() { Matrix* mm = m; long ii = i; return CommaInit<T>([mm, ii](long, long k) -> T& { return (*mm)(ii, k); }, 1, m->c_); }
        CommaInit<T> operator<<(const T& s) { CommaInit<T> ci = comma(); ci.put(s); return ci; }
        template <int R2, int C2>
        CommaInit<T> operator<<(const Matrix<T, R2, C2>& o) { CommaInit<T> ci = comma(); ci.put_block(o); return ci; }
        template <int R2, int C2>
        Matrix<T, R2, C2> cast_() const { return Matrix<T, R2, C2>(vec()); }
        RowVec segment(long o, long n) const { return vec().segment(o, n); }
    };
    RowRef row(long i) { return RowRef{this, i}; }
    RowVec row(long i) const { RowVec v(c_); for (long j = 0; j < c_; j++) v(j) = (*this)(i, j); return v; }
    struct RowwiseRef {
        Matrix* m;
        Matrix<T, Dynamic, 1> sum() const { Matrix<T, Dynamic, 1> o(m->r_); for (long i = 0; i < m->r_; i++) { T a{}; for (long j = 0; j < m->c_; j++) a += (*m)(i, j); o(i) = a; } return o; }
        void normalize() {
            for (long i = 0; i < m->r_; i++) {
                double n = 0;
                for (long j = 0; j < m->c_; j++) n += abs2_((*m)(i, j));
                n = std::sqrt(n);
                if (n > 0) for (long j = 0; j < m->c_; j++) (*m)(i, j) = (*m)(i, j) / n;
            }
        }
    };
    RowwiseRef rowwise() { return RowwiseRef{this}; }
    struct RowwiseConst {
        const Matrix* m;
        Matrix<T, Dynamic, 1> sum() const { Matrix<T, Dynamic, 1> o(m->r_); for (long i = 0; i < m->r_; i++) { T a{}; for (long j = 0; j < m->c_; j++) a += (*m)(i, j); o(i) = a; } return o; }
    };
    RowwiseConst rowwise() const { return RowwiseConst{this}; }
    struct ColwiseConst {
        const Matrix* m;
        RowVec maxCoeff() const { RowVec o(m->c_); for (long j = 0; j < m->c_; j++) { T a = (*m)(0, j); for (long i = 1; i < m->r_; i++) a = std::max(a, (*m)(i, j)); o(j) = a; } return o; }
        RowVec minCoeff() const { RowVec o(m->c_); for (long j = 0; j < m->c_; j++) { T a = (*m)(0, j); for (long i = 1; i < m->r_; i++) a = std::min(a, (*m)(i, j)); o(j) = a; } return o; }
    };
    ColwiseConst colwise() const { return ColwiseConst{this}; }
    template <int R2, int C2>
    Matrix cwiseProduct(const Matrix<T, R2, C2>& b) const { Matrix o = *this; for (long j = 0; j < c_; j++) for (long i = 0; i < r_; i++) o(i, j) = (*this)(i, j) * b(i, j); return o; }
    Matrix cwiseAbs() const { Matrix o = *this; for (auto& x : o.v_) x = x < T{} ? -x : x; return o; }
    Matrix<T, C, R> transpose() const { Matrix<T, C, R> o(c_, r_); for (long i = 0; i < r_; i++) for (long j = 0; j < c_; j++) o(j, i) = (*this)(i, j); return o; }
    template <int Order>
    Matrix<T, Dynamic, 1> reshaped() const {  // Order = RowMajor: the coefficients row by row
        Matrix<T, Dynamic, 1> o(size());
        long k = 0;
        if (Order == RowMajor) { for (long i = 0; i < r_; i++) for (long j = 0; j < c_; j++) o(k++) = (*this)(i, j); }
        else { for (long j = 0; j < c_; j++) for (long i = 0; i < r_; i++) o(k++) = (*this)(i, j); }
        return o;
    }
    static Matrix Identity(long r, long c) { Matrix m(r, c); for (long i = 0; i < std::min(r, c); i++) m(i, i) = T(1); return m; }
    void setConstant(const T& x) { std::fill(v_.begin(), v_.end(), x); }
    void setZero() { std::fill(v_.begin(), v_.end(), T{}); }
    T& x() { return v_[0]; }
    T& y() { return v_[1]; }
    T& z() { return v_[2]; }
    const T& x() const { return v_[0]; }
    const T& y() const { return v_[1]; }
    const T& z() const { return v_[2]; }
    using DynMat = Matrix<T, Dynamic, Dynamic>;
    struct BlockRef {
        Matrix* m;
        long i0, j0, nr, nc;
        DynMat mat() const { DynMat o(nr, nc); for (long j = 0; j < nc; j++) for (long i = 0; i < nr; i++) o(i, j) = (*m)(i0 + i, j0 + j); return o; }
        template <int R2, int C2>
        operator Matrix<T, R2, C2>() const { return Matrix<T, R2, C2>(mat()); }
        DynMat operator-() const { return -mat(); }
        BlockRef& operator=(const RowArr& a) { for (long j = 0; j < nc; j++) (*m)(i0, j0 + j) = (*a.m)(a.i, j); return *this; }
        CommaInit<T> comma() { Matrix* mm = m; long a = i0, b = j0; return CommaInit<T>([mm, a, b](long i, long j) -> T& { return (*mm)(a + i, b + j); }, nr, nc); }
        CommaInit<T> operator<<(const T& s) { CommaInit<T> ci = comma(); ci.put(s); return ci; }
        template <int R2, int C2>
        CommaInit<T> operator<<(const Matrix<T, R2, C2>& o) { CommaInit<T> ci = comma(); ci.put_block(o); return ci; }
        template <int R2, int C2>
        BlockRef& operator=(const Matrix<T, R2, C2>& o) {  // a vector fills a 1 x n / n x 1 block in order
            const bool vec = o.rows() == 1 || o.cols() == 1;
            for (long j = 0; j < nc; j++) for (long i = 0; i < nr; i++) (*m)(i0 + i, j0 + j) = vec ? o(j * nr + i) : o(i, j);
            return *this;
        }
    };
    BlockRef block(long i0, long j0, long nr, long nc) { return BlockRef{this, i0, j0, nr, nc}; }
    BlockRef middleCols(long j0, long nc) { return BlockRef{this, 0, j0, r_, nc}; }
    DynMat middleCols(long j0, long nc) const { return block(0, j0, r_, nc); }
    DynMat block(long i0, long j0, long nr, long nc) const {
        DynMat o(nr, nc);
        for (long j = 0; j < nc; j++) for (long i = 0; i < nr; i++) o(i, j) = (*this)(i0 + i, j0 + j);
        return o;
    }

    CommaInit<T> comma() { Matrix* m = this; return CommaInit<T>([m](long i, long j) -> T& { return (*m)(i, j); }, r_, c_); }
    CommaInit<T> operator<<(const T& s) { CommaInit<T> ci = comma(); ci.put(s); return ci; }
    template <int R2, int C2>
    CommaInit<T> operator<<(const Matrix<T, R2, C2>& o) { CommaInit<T> ci = comma(); ci.put_block(o); return ci; }

    ArrayX<T> array() const { ArrayX<T> a; a.v = v_; a.r = r_; a.c = c_; return a; }

    Matrix operator-() const { Matrix o = *this; for (auto& x : o.v_) x = -x; return o; }
    template <int R2, int C2>
    Matrix operator+(const Matrix<T, R2, C2>& b) const { Matrix o = *this; for (long i = 0; i < size(); i++) o.v_[i] = v_[i] + b(i); return o; }
    template <int R2, int C2>
    Matrix operator-(const Matrix<T, R2, C2>& b) const { Matrix o = *this; for (long i = 0; i < size(); i++) o.v_[i] = v_[i] - b(i); return o; }
    Matrix operator*(const T& s) const { Matrix o = *this; for (auto& x : o.v_) x = x * s; return o; }
    friend Matrix operator*(const T& s, const Matrix& a) { Matrix o = a; for (auto& x : o.v_) x = s * x; return o; }
    Matrix operator/(const T& s) const { Matrix o = *this; for (auto& x : o.v_) x = x / s; return o; }
    // complex vector scaled by a real (x /= s, x / sqrt(...))
    template <class S, class = std::enable_if_t<!std::is_same<S, T>::value && std::is_arithmetic<S>::value>>
    Matrix operator/(const S& s) const { Matrix o = *this; for (auto& x : o.v_) x = x / (double)s; return o; }
    template <class S, class = std::enable_if_t<std::is_arithmetic<S>::value>>
    Matrix& operator/=(const S& s) { for (auto& x : v_) x = x / (double)s; return *this; }

private:
    template <class, int, int> friend class Matrix;
    static double conj_(double x) { return x; }
    static int conj_(int x) { return x; }
    static std::complex<double> conj_(const std::complex<double>& x) { return std::conj(x); }
    static double uni() {
        if (refshim::random_source) return refshim::random_source[refshim::random_pos++];
        return 2.0 * std::rand() / (double)RAND_MAX - 1.0;
    }
    static double rnd(double*) { return uni(); }
    static int rnd(int*) { return std::rand(); }
    static bool rnd(bool*) { return std::rand() & 1; }
    static std::complex<double> rnd(std::complex<double>*) { const double a = uni(); const double b = uni(); return {a, b}; }
    void init(long r, long c) { r_ = r; c_ = c; v_.assign((size_t)(r * c), T{}); }
    void init_like(long r, long c) {
        // a vector type keeps its own orientation whatever the source's is (Eigen transposes vectors on assignment)
        if (kIsVector && (r == 1 || c == 1)) { if (R == 1) init(1, r * c); else if (C == 1) init(r * c, 1); else init(r, c); }
        else init(r, c);
    }
    template <int R2, int C2>
    void assign(const Matrix<T, R2, C2>& o) { init_like(o.r_, o.c_); v_ = o.v_; }
    long r_ = 0, c_ = 0;
    std::vector<T> v_;
};

template <class M>
struct Map {
    const typename M::Scalar* p;
    long n;
    Map(const typename M::Scalar* p_, long n_) : p(p_), n(n_) {}
};
template <class T, int R, int C>
Matrix<T, R, C>::Matrix(const Map<Matrix>& m) {
    if (R == 1) init(1, m.n); else init(m.n, 1);
    std::copy(m.p, m.p + m.n, v_.begin());
}

using VectorXd = Matrix<double, Dynamic, 1>;
using VectorXi = Matrix<int, Dynamic, 1>;
using VectorXcd = Matrix<std::complex<double>, Dynamic, 1>;
using MatrixXd = Matrix<double, Dynamic, Dynamic>;
using MatrixXi = Matrix<int, Dynamic, Dynamic>;
using RowVector2i = Matrix<int, 1, 2>;
using RowVector2d = Matrix<double, 1, 2>;
using RowVector3i = Matrix<int, 1, 3>;
using RowVector3d = Matrix<double, 1, 3>;
using Matrix2i = Matrix<int, 2, 2>;
using Matrix2d = Matrix<double, 2, 2>;
using MatrixXcd = Matrix<std::complex<double>, Dynamic, Dynamic>;

// compressed sparse matrix, default (column-major) storage order only -- the only one the cited headers use
template <class T, int Options = ColMajor>
class SparseMatrix {
    static_assert(Options == ColMajor, "eigen_standin: column-major sparse matrices only");
public:
    using Scalar = T;
    SparseMatrix() : outer_(1, 0) {}
    SparseMatrix(long rows, long cols) : rows_((int)rows), cols_((int)cols), outer_(cols + 1, 0) {}

    long rows() const { return rows_; }
    long cols() const { return cols_; }
    long outerSize() const { return cols_; }
    long innerSize() const { return rows_; }
    long nonZeros() const { return (long)inner_.size(); }
    void makeCompressed() {}
    const int* outerIndexPtr() const { return outer_.data(); }
    const int* innerIndexPtr() const { return inner_.data(); }
    const T* valuePtr() const { return val_.data(); }
    void resize(long r, long c) { rows_ = (int)r; cols_ = (int)c; outer_.assign(c + 1, 0); inner_.clear(); val_.clear(); }
    // every caller on this path follows conservativeResize with setFromTriplets, which rebuilds the matrix: entries outside the new
    // shape are dropped, the rest kept, as Eigen does
    void conservativeResize(long r, long c) {
        std::vector<Triplet<T>> keep;
        for (int j = 0; j < cols_ && j < c; j++)
            for (int k = outer_[j]; k < outer_[j + 1]; k++)
                if (inner_[k] < r) keep.emplace_back(inner_[k], j, val_[k]);
        resize(r, c);
        setFromTriplets(keep.begin(), keep.end());
    }

    class InnerIterator {
    public:
        InnerIterator(const SparseMatrix& m, long outer) : m_(&m), outer_((int)outer), k_(m.outer_[outer]), end_(m.outer_[outer + 1]) {}
        explicit operator bool() const { return k_ < end_; }
        InnerIterator& operator++() { ++k_; return *this; }
        const T& value() const { return m_->val_[k_]; }
        int index() const { return m_->inner_[k_]; }
        int row() const { return m_->inner_[k_]; }
        int col() const { return outer_; }
    private:
        const SparseMatrix* m_;
        int outer_, k_, end_;
    };

    // duplicates are summed in the order they were pushed (Eigen: insertion order within a column is preserved by its two transpositions)
    template <class It>
    void setFromTriplets(It b, It e) {
        std::vector<int> cnt(cols_ + 1, 0);
        for (It t = b; t != e; ++t) {
            if (t->row() < 0 || t->row() >= rows_ || t->col() < 0 || t->col() >= cols_) {
                std::fprintf(stderr, "eigen_standin: triplet (%d, %d) outside a %d x %d matrix\n", t->row(), t->col(), rows_, cols_);
                std::abort();
            }
            cnt[t->col() + 1]++;
        }
        for (int j = 0; j < cols_; j++) cnt[j + 1] += cnt[j];
        std::vector<std::pair<int, T>> ent(cnt[cols_]);
        std::vector<int> pos(cnt.begin(), cnt.end() - 1);
        for (It t = b; t != e; ++t) ent[pos[t->col()]++] = {t->row(), t->value()};
        inner_.clear();
        val_.clear();
        outer_.assign(cols_ + 1, 0);
        for (int j = 0; j < cols_; j++) {
            auto first = ent.begin() + cnt[j], last = ent.begin() + cnt[j + 1];
            std::stable_sort(first, last, [](const auto& x, const auto& y) { return x.first < y.first; });
            for (auto it = first; it != last;) {
                const int i = it->first;
                T s{};
                while (it != last && it->first == i) { s += it->second; ++it; }
                inner_.push_back(i);
                val_.push_back(s);
            }
            outer_[j + 1] = (int)inner_.size();
        }
    }

    // Eigen: pruned(reference = 1, epsilon = NumTraits<Scalar>::dummy_precision() = 1e-12) keeps |v| > |reference| * epsilon
    SparseMatrix pruned(const T& ref = T(1), double eps = 1e-12) const {
        SparseMatrix Rm(rows_, cols_);
        for (int j = 0; j < cols_; j++) {
            for (int k = outer_[j]; k < outer_[j + 1]; k++)
                if (std::abs(val_[k]) > std::abs(ref) * eps) { Rm.inner_.push_back(inner_[k]); Rm.val_.push_back(val_[k]); }
            Rm.outer_[j + 1] = (int)Rm.inner_.size();
        }
        return Rm;
    }

    void prune(const T& ref, double eps = 1e-12) { *this = pruned(ref, eps); }

    SparseMatrix transpose() const {
        SparseMatrix Tm(cols_, rows_);
        for (int i : inner_) Tm.outer_[i + 1]++;
        for (int i = 0; i < rows_; i++) Tm.outer_[i + 1] += Tm.outer_[i];
        Tm.inner_.resize(inner_.size());
        Tm.val_.resize(val_.size());
        std::vector<int> pos(Tm.outer_.begin(), Tm.outer_.end() - 1);
        for (int j = 0; j < cols_; j++)
            for (int k = outer_[j]; k < outer_[j + 1]; k++) {
                const int p = pos[inner_[k]]++;
                Tm.inner_[p] = j;
                Tm.val_[p] = val_[k];
            }
        return Tm;
    }

    // column by column with a dense accumulator; inner indices come out sorted
    friend SparseMatrix operator*(const SparseMatrix& A, const SparseMatrix& B) {
        if (A.cols_ != B.rows_) { std::fprintf(stderr, "eigen_standin: sparse product %dx%d * %dx%d\n", A.rows_, A.cols_, B.rows_, B.cols_); std::abort(); }
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
    void setIdentity() {
        outer_.assign(cols_ + 1, 0); inner_.clear(); val_.clear();
        for (int j = 0; j < cols_; j++) { if (j < rows_) { inner_.push_back(j); val_.push_back(T(1)); } outer_[j + 1] = (int)inner_.size(); }
    }
    static SparseMatrix combine(const SparseMatrix& A, const SparseMatrix& B, bool minus) {
        std::vector<Triplet<T>> t;
        for (int j = 0; j < A.cols_; j++) for (int k = A.outer_[j]; k < A.outer_[j + 1]; k++) t.emplace_back(A.inner_[k], j, A.val_[k]);
        for (int j = 0; j < B.cols_; j++) for (int k = B.outer_[j]; k < B.outer_[j + 1]; k++) t.emplace_back(B.inner_[k], j, minus ? -B.val_[k] : B.val_[k]);
        SparseMatrix o(A.rows_, A.cols_);
        o.setFromTriplets(t.begin(), t.end());
        return o;
    }
    friend SparseMatrix operator+(const SparseMatrix& A, const SparseMatrix& B) { return combine(A, B, false); }
    friend SparseMatrix operator-(const SparseMatrix& A, const SparseMatrix& B) { return combine(A, B, true); }
    // real scalar times a complex matrix (1e-9 * I, lambda * M: frf.h:28,75)
    template <class S, class = std::enable_if_t<!std::is_same<S, T>::value && std::is_arithmetic<S>::value>>
    friend SparseMatrix operator*(const S& s, const SparseMatrix& A) { SparseMatrix o = A; for (auto& x : o.val_) x = x * (double)s; return o; }
    // row vector times sparse matrix
    friend Matrix<T, 1, Dynamic> operator*(const Matrix<T, 1, Dynamic>& x, const SparseMatrix& A) {
        Matrix<T, 1, Dynamic> y(A.cols_);
        for (int j = 0; j < A.cols_; j++) { T acc{}; for (int k = A.outer_[j]; k < A.outer_[j + 1]; k++) acc += x(A.inner_[k]) * A.val_[k]; y(j) = acc; }
        return y;
    }
    friend SparseMatrix operator*(const SparseMatrix& A, const T& s) { SparseMatrix o = A; for (auto& x : o.val_) x = x * s; return o; }
    friend SparseMatrix operator*(const T& s, const SparseMatrix& A) { SparseMatrix o = A; for (auto& x : o.val_) x = s * x; return o; }
    template <int R2, int C2>
    friend Matrix<T, Dynamic, 1> operator*(const SparseMatrix& A, const Matrix<T, R2, C2>& x) {
        if (A.cols_ != x.size()) { std::fprintf(stderr, "eigen_standin: sparse %dx%d times a vector of %ld\n", A.rows_, A.cols_, x.size()); std::abort(); }
        Matrix<T, Dynamic, 1> y = Matrix<T, Dynamic, 1>::Zero(A.rows_);
        for (int j = 0; j < A.cols_; j++) {
            const T xj = x(j);
            for (int k = A.outer_[j]; k < A.outer_[j + 1]; k++) y(A.inner_[k]) += A.val_[k] * xj;
        }
        return y;
    }

private:
    int rows_ = 0, cols_ = 0;
    std::vector<int> outer_, inner_;
    std::vector<T> val_;
};

// row-major copy of a column-major matrix (`Eigen::SparseMatrix<double, Eigen::RowMajor> R = A;`): only what a CSR export needs
template <class T>
class SparseMatrix<T, RowMajor> {
public:
    using Scalar = T;
    SparseMatrix() : outer_(1, 0) {}
    SparseMatrix(const SparseMatrix<T, ColMajor>& A) {
        const SparseMatrix<T, ColMajor> At = A.transpose();  // CSC of A^T = CSR of A
        rows_ = (int)A.rows();
        cols_ = (int)A.cols();
        outer_.assign(At.outerIndexPtr(), At.outerIndexPtr() + rows_ + 1);
        inner_.assign(At.innerIndexPtr(), At.innerIndexPtr() + At.nonZeros());
        val_.assign(At.valuePtr(), At.valuePtr() + At.nonZeros());
    }
    long rows() const { return rows_; }
    long cols() const { return cols_; }
    long nonZeros() const { return (long)inner_.size(); }
    void makeCompressed() {}
    const int* outerIndexPtr() const { return outer_.data(); }
    const int* innerIndexPtr() const { return inner_.data(); }
    const T* valuePtr() const { return val_.data(); }
private:
    int rows_ = 0, cols_ = 0;
    std::vector<int> outer_, inner_;
    std::vector<T> val_;
};

// ---- direct solvers: hooks (see the header comment) ----
template <class M, int KIND>
class HookedSolver {
public:
    using T = typename M::Scalar;
    using Vec = Matrix<T, Dynamic, 1>;
    HookedSolver() = default;
    HookedSolver(const HookedSolver&) = delete;
    HookedSolver& operator=(const HookedSolver&) = delete;
    ~HookedSolver() { drop(); }
    void compute(const M& A) {
        drop();
        n_ = (int)A.rows();
        refshim::factor_count++;
        if (!refshim::hooks.factor) { std::fprintf(stderr, "eigen_standin: no solver hooks set\n"); std::abort(); }
        h_ = refshim::hooks.factor(KIND, std::is_same<T, double>::value ? 0 : 1, n_, A.outerIndexPtr(), A.innerIndexPtr(),
                                   reinterpret_cast<const double*>(A.valuePtr()));
        info_ = h_ >= 0 ? Success : NumericalIssue;
    }
    ComputationInfo info() const { return info_; }
    template <int R2, int C2>
    Vec solve(const Matrix<T, R2, C2>& rhs) const {
        Vec x(n_);
        if (h_ < 0 || refshim::hooks.solve(h_, reinterpret_cast<const double*>(rhs.data()), reinterpret_cast<double*>(x.data())) != 0) {
            std::fprintf(stderr, "eigen_standin: solve failed\n");
            std::abort();
        }
        return x;
    }
private:
    void drop() { if (h_ >= 0 && refshim::hooks.release) refshim::hooks.release(h_); h_ = -1; }
    int h_ = -1, n_ = 0;
    ComputationInfo info_ = InvalidInput;
};
template <class M> using SparseLU = HookedSolver<M, 0>;
template <class M> using SimplicialLLT = HookedSolver<M, 1>;
template <class M> using SimplicialLDLT = HookedSolver<M, 1>;

// SparseQR<.., COLAMDOrdering> as reduce_to_linearly_independent uses it (matrix_ops.h:12-30, called by integ(), integ.ipp:6): rank() and a
// column permutation whose first rank() entries are independent columns. NOT Eigen's algorithm (COLAMD + Householder with a pivot threshold):
// columns are taken in their natural order by Gram-Schmidt with the relative tolerance 1e-9, which is what tests/hostglue/igm_glue.hpp's
// independent_rows does. Another valid choice of independent constraint rows spans the same row space: UFull's kernel space and fullx are
// the same, its basis (and x_small) may differ from an Eigen build's. UNPINNED, stated as such in DESIGN.md.
template <class I>
struct COLAMDOrdering {};
template <class M, class O>
class SparseQR {
public:
    struct Perm {
        VectorXi idx;
        const VectorXi& indices() const { return idx; }
    };
    void compute(const M& A) {
        const long n = A.cols(), m = A.rows();
        std::vector<std::vector<double>> basis;
        std::vector<int> keep, drop;
        for (long j = 0; j < n; j++) {
            std::vector<double> r(m, 0.0);
            for (typename M::InnerIterator it(A, j); it; ++it) r[it.row()] = it.value();
            double n0 = 0;
            for (double x : r) n0 += x * x;
            n0 = std::sqrt(n0);
            if (n0 == 0) { drop.push_back((int)j); continue; }
            for (int pass = 0; pass < 2; pass++)
                for (const auto& b : basis) {
                    double d = 0;
                    for (long k = 0; k < m; k++) d += r[k] * b[k];
                    for (long k = 0; k < m; k++) r[k] -= d * b[k];
                }
            double n1 = 0;
            for (double x : r) n1 += x * x;
            n1 = std::sqrt(n1);
            if (n1 > 1e-9 * n0) { for (double& x : r) x /= n1; basis.push_back(r); keep.push_back((int)j); }
            else drop.push_back((int)j);
        }
        rank_ = (int)keep.size();
        perm_.idx.resize(n);
        long k = 0;
        for (int c : keep) perm_.idx(k++) = c;
        for (int c : drop) perm_.idx(k++) = c;
    }
    int rank() const { return rank_; }
    const Perm& colsPermutation() const { return perm_; }
private:
    Perm perm_;
    int rank_ = 0;
};

// Eigen 3.4 FullPivLU: complete pivoting, the first maximum in column-major order wins; rank counts |pivot| > eps * diagSize * |maxpivot|
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
        std::vector<int> piv;
        for (int i = 0; i < nonzero_pivots_; i++)
            if (std::fabs(lu_(i, i)) > std::fabs(maxpivot_) * threshold()) piv.push_back(i);
        std::vector<int> colorder = piv;
        {
            std::vector<char> isp(cols_, 0);
            for (int p : piv) isp[p] = 1;
            for (int j = 0; j < cols_; j++) if (!isp[j]) colorder.push_back(j);
        }
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
    double u(int i, int j) const { return j >= i ? lu_(i, j) : 0.0; }
    MatrixXd lu_;
    int rows_, cols_, nonzero_pivots_ = 0;
    double maxpivot_ = 0.0;
    std::vector<int> colperm_;
};

}  // namespace Eigen
