// stand-in for libigl's igl::slice(X, R, dim, Y) on a sparse matrix (iter_rounding_init.h:180): dim = 1: Y = X(R, :), dim = 2: Y = X(:, R).
#pragma once
#include <vector>
#include "../eigen_standin.hpp"
namespace igl {
template <class T, class VI>
void slice(const Eigen::SparseMatrix<T>& X, const VI& R, const int dim, Eigen::SparseMatrix<T>& Y) {
    std::vector<std::vector<int>> where(dim == 1 ? X.rows() : X.cols());  // source index -> positions in R (R may repeat an index)
    for (long k = 0; k < R.size(); k++) where[R(k)].push_back((int)k);
    std::vector<Eigen::Triplet<T>> t;
    for (long j = 0; j < X.outerSize(); j++)
        for (typename Eigen::SparseMatrix<T>::InnerIterator it(X, j); it; ++it) {
            if (dim == 1) for (int k : where[it.row()]) t.emplace_back(k, it.col(), it.value());
            else for (int k : where[it.col()]) t.emplace_back(it.row(), k, it.value());
        }
    if (dim == 1) Y.resize(R.size(), X.cols()); else Y.resize(X.rows(), R.size());
    Y.setFromTriplets(t.begin(), t.end());
}
}  // namespace igl
