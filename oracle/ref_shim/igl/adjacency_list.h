// see dijkstra.h: declaration only.
#pragma once
#include "dijkstra.h"
namespace igl {
template <class MF>
void adjacency_list(const MF&, std::vector<std::vector<int>>&) { not_compiled("adjacency_list"); }
}  // namespace igl
