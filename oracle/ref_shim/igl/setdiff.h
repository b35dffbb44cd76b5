// stand-in for libigl v2.5.0's igl::setdiff (absent from the image) as iter_rounding_init.h:125 calls it: C = the sorted unique values of A
// that are not in B, IA = an index into A for each of them (first occurrence). Test infrastructure, see ../eigen_standin.hpp.
#pragma once
#include <algorithm>
#include <set>
#include <vector>
#include "../eigen_standin.hpp"
namespace igl {
template <class VA, class VB, class VC, class VI>
void setdiff(const VA& A, const VB& B, VC& C, VI& IA) {
    std::set<int> b;
    for (long i = 0; i < B.size(); i++) b.insert(B(i));
    std::vector<std::pair<int, int>> keep;
    std::set<int> seen;
    for (long i = 0; i < A.size(); i++)
        if (!b.count(A(i)) && seen.insert(A(i)).second) keep.push_back({A(i), (int)i});
    std::sort(keep.begin(), keep.end());
    C.resize((long)keep.size());
    IA.resize((long)keep.size());
    for (size_t k = 0; k < keep.size(); k++) { C(k) = keep[k].first; IA(k) = keep[k].second; }
}
}  // namespace igl
