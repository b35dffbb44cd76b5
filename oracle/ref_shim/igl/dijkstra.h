// vectorfield/face_rosy_field.h:268-360 (compute_seam: libigl's cut graph + Dijkstra paths) is NOT part of the compiled pin: these
// declarations only let the header go through the compiler; calling them aborts. (The seam is input to compute_combbed_field here.)
#pragma once
#include <cstdio>
#include <cstdlib>
#include <set>
#include <vector>
#include "../eigen_standin.hpp"
namespace igl {
[[noreturn]] inline void not_compiled(const char* what) { std::fprintf(stderr, "oracle/ref_shim: igl::%s is not part of the pin\n", what); std::abort(); }
template <class VD, class VI>
int dijkstra(int, const std::set<int>&, const std::vector<std::vector<int>>&, VD&, VI&) { not_compiled("dijkstra"); }
template <class VI>
void dijkstra(int, const VI&, std::vector<int>&) { not_compiled("dijkstra"); }
}  // namespace igl
