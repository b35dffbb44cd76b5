// see dijkstra.h: declarations only (libigl's own header pulls in vertex_triangle_adjacency / triangle_triangle_adjacency as well).
#pragma once
#include "dijkstra.h"
namespace igl {
template <class MV, class MF, class MM, class MC>
void cut_mesh_from_singularities(const MV&, const MF&, const MM&, MC&) { not_compiled("cut_mesh_from_singularities"); }
template <class MV, class MF>
void vertex_triangle_adjacency(const MV&, const MF&, std::vector<std::vector<int>>&, std::vector<std::vector<int>>&) { not_compiled("vertex_triangle_adjacency"); }
template <class MF, class MI>
void triangle_triangle_adjacency(const MF&, MI&, MI&) { not_compiled("triangle_triangle_adjacency"); }
}  // namespace igl
