// Minimal host CSR toolkit for the host-glue stages (triplet assembly with duplicate summation,
// Gustavson SpGEMM, transpose). Stands in for the Eigen::SparseMatrix operations the reference's host
// stages use (setFromTriplets / operator* / pruned), which are not available in this image.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <vector>

namespace hg {

struct Trip {
    int r, c;
    double v;
};

struct Csr {
    int rows = 0, cols = 0;
    std::vector<int64_t> ptr;  // rows + 1
    std::vector<int> col;
    std::vector<double> val;
    int64_t nnz() const { return (int64_t)col.size(); }
};

// duplicates are summed (Eigen setFromTriplets semantics); exact zeros are kept unless prune=true
inline Csr csr_from_triplets(int rows, int cols, const std::vector<Trip>& T, bool prune = false, double eps = 0.0) {
    Csr A;
    A.rows = rows;
    A.cols = cols;
    std::vector<int64_t> cnt(rows + 1, 0);
    for (const Trip& t : T) cnt[t.r + 1]++;
    for (int i = 0; i < rows; i++) cnt[i + 1] += cnt[i];
    std::vector<int> c(T.size());
    std::vector<double> v(T.size());
    {
        std::vector<int64_t> pos(cnt.begin(), cnt.end() - 1);
        for (const Trip& t : T) { int64_t p = pos[t.r]++; c[p] = t.c; v[p] = t.v; }
    }
    A.ptr.assign(rows + 1, 0);
    A.col.reserve(T.size());
    A.val.reserve(T.size());
    std::vector<int> order;
    for (int i = 0; i < rows; i++) {
        int64_t b = cnt[i], e = cnt[i + 1];
        order.resize(e - b);
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return c[b + x] < c[b + y]; });
        for (size_t k = 0; k < order.size();) {
            int cc = c[b + order[k]];
            double s = 0;
            while (k < order.size() && c[b + order[k]] == cc) s += v[b + order[k++]];
            if (prune && std::fabs(s) <= eps) continue;
            A.col.push_back(cc);
            A.val.push_back(s);
        }
        A.ptr[i + 1] = (int64_t)A.col.size();
    }
    return A;
}

inline Csr csr_transpose(const Csr& A) {
    Csr B;
    B.rows = A.cols;
    B.cols = A.rows;
    B.ptr.assign(B.rows + 1, 0);
    for (int c : A.col) B.ptr[c + 1]++;
    for (int i = 0; i < B.rows; i++) B.ptr[i + 1] += B.ptr[i];
    B.col.resize(A.col.size());
    B.val.resize(A.val.size());
    std::vector<int64_t> pos(B.ptr.begin(), B.ptr.end() - 1);
    for (int r = 0; r < A.rows; r++)
        for (int64_t k = A.ptr[r]; k < A.ptr[r + 1]; k++) {
            int64_t p = pos[A.col[k]]++;
            B.col[p] = r;
            B.val[p] = A.val[k];
        }
    return B;
}

// C = A * B (row-wise Gustavson with a dense accumulator), entries with |v| <= eps dropped when prune
inline Csr csr_matmul(const Csr& A, const Csr& B, bool prune = false, double eps = 0.0) {
    Csr C;
    C.rows = A.rows;
    C.cols = B.cols;
    C.ptr.assign(A.rows + 1, 0);
    std::vector<double> acc(B.cols, 0.0);
    std::vector<int> mark(B.cols, -1), touched;
    for (int i = 0; i < A.rows; i++) {
        touched.clear();
        for (int64_t k = A.ptr[i]; k < A.ptr[i + 1]; k++) {
            int j = A.col[k];
            double a = A.val[k];
            for (int64_t l = B.ptr[j]; l < B.ptr[j + 1]; l++) {
                int c = B.col[l];
                if (mark[c] != i) { mark[c] = i; acc[c] = 0; touched.push_back(c); }
                acc[c] += a * B.val[l];
            }
        }
        std::sort(touched.begin(), touched.end());
        for (int c : touched) {
            if (prune && std::fabs(acc[c]) <= eps) continue;
            C.col.push_back(c);
            C.val.push_back(acc[c]);
        }
        C.ptr[i + 1] = (int64_t)C.col.size();
    }
    return C;
}

inline std::vector<double> csr_matvec(const Csr& A, const std::vector<double>& x) {
    std::vector<double> y(A.rows, 0.0);
    for (int i = 0; i < A.rows; i++) {
        double s = 0;
        for (int64_t k = A.ptr[i]; k < A.ptr[i + 1]; k++) s += A.val[k] * x[A.col[k]];
        y[i] = s;
    }
    return y;
}

}  // namespace hg
