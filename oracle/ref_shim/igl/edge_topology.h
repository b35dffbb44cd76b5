// stand-in for libigl v2.5.0's igl::edge_topology(V, F, EV, FE, EF) (absent from the image; hmesh/hmesh.h:317), written from its documented
// contract as SURVEY.md Appendix A.1 records it: edges are the unique undirected vertex pairs sorted lexicographically by (vmin, vmax);
// EV(e) = (vmin, vmax); EF(e, 0) = the face in which the edge runs vmin -> vmax, EF(e, 1) = the other (or -1 on a boundary); FE(f, j) = the
// edge from corner j to corner j + 1. UNPINNED (the libigl source is not on disk): tests/hostglue/mesh.hpp makes the same assumption.
#pragma once
#include <algorithm>
#include <array>
#include <vector>
#include "../eigen_standin.hpp"
namespace igl {
template <class MV, class MF, class MI>
void edge_topology(const MV& /*V*/, const MF& F, MI& EV, MI& FE, MI& EF) {
    const long nF = F.rows();
    std::vector<std::array<int, 4>> he;  // (vmin, vmax, face, corner)
    for (long f = 0; f < nF; f++)
        for (int j = 0; j < 3; j++) {
            const int a = F(f, j), b = F(f, (j + 1) % 3);
            he.push_back({std::min(a, b), std::max(a, b), (int)f, j});
        }
    std::sort(he.begin(), he.end());
    long nE = 0;
    for (size_t k = 0; k < he.size(); k++)
        if (k == 0 || he[k][0] != he[k - 1][0] || he[k][1] != he[k - 1][1]) nE++;
    EV = MI::Constant(nE, 2, -1);
    EF = MI::Constant(nE, 2, -1);
    FE = MI::Constant(nF, 3, -1);
    long e = -1;
    for (size_t k = 0; k < he.size(); k++) {
        if (k == 0 || he[k][0] != he[k - 1][0] || he[k][1] != he[k - 1][1]) { e++; EV(e, 0) = he[k][0]; EV(e, 1) = he[k][1]; }
        const int f = he[k][2], j = he[k][3];
        FE(f, j) = (int)e;
        const bool forward = F(f, j) == he[k][0];  // the face sees the edge running vmin -> vmax
        EF(e, forward ? 0 : 1) = f;
    }
}
}  // namespace igl
