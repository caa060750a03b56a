// STAND-IN for the gmsh SDK calls made by PFEM3D's Mesh::loadFromFile / computeMeshDim (Mesh.cpp:209-235, 762-917).
// TEST INFRASTRUCTURE ONLY.  gmsh is absent from this image; the test driver injects node/element arrays instead of
// a .msh file (see ref_inject.hpp).  The fake "model" holds one element type of the injected dimension and one
// physical group "Fluid" with a single placeholder node; the injected arrays replace it in triangulateAlphaShape*().
#pragma once
#include <cstddef>
#include <string>
#include <utility>
#include <vector>

#include "../ref_inject.hpp"

namespace gmsh {
inline void initialize() {}
inline void finalize() {}
inline void open(const std::string&) {}
namespace option {
inline void setNumber(const std::string&, double) {}
}  // namespace option
namespace model {
inline void getPhysicalGroups(std::vector<std::pair<int, int>>& dimTags, int dim) {
    dimTags.clear();
    if (dim == refinject::current().dim) dimTags.emplace_back(dim, 1);
}
inline void getPhysicalName(int, int, std::string& name) { name = "Fluid"; }
namespace mesh {
inline void getElementTypes(std::vector<int>& types, int dim) {
    types.clear();
    if (dim == refinject::current().dim) types.push_back(dim == 2 ? 2 : 4);
}
inline void getNodesForPhysicalGroup(int, int, std::vector<std::size_t>& tags, std::vector<double>& coord) {
    tags.assign(1, 1);
    coord.assign(3, 0.0);
}
}  // namespace mesh
}  // namespace model
}  // namespace gmsh
