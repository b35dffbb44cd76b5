// stand-in for libigl's igl::speye(n, I) (iter_rounding_init.h:174, iter_rounding_loop.h:54): I = the n x n sparse identity.
#pragma once
#include <vector>
#include "../eigen_standin.hpp"
namespace igl {
template <class T>
void speye(const int n, Eigen::SparseMatrix<T>& I) {
    std::vector<Eigen::Triplet<T>> t;
    for (int i = 0; i < n; i++) t.emplace_back(i, i, T(1));
    I.resize(n, n);
    I.setFromTriplets(t.begin(), t.end());
}
}  // namespace igl
