// Arrays handed from the test driver to the stand-in mesh loader (test infrastructure, see ref_driver.cpp).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace refinject {
struct MeshArrays {
    int dim = 0;
    std::size_t nNodes = 0, nElems = 0, nFacets = 0;
    unsigned int nStates = 0;
    const std::int64_t* conn = nullptr;      // nElems x (dim+1), row-major
    const double* x = nullptr;               // x[n + d*nNodes]
    const std::uint8_t* flags = nullptr;     // C-ABI bits: 1 isBound, 2 isFree, 4 isFixed, 8 isOnFreeSurface
    const std::int32_t* tags = nullptr;      // per node index into tagNames
    std::vector<std::string> tagNames;
    const std::int64_t* facets = nullptr;    // nFacets x (dim + 2): dim facet nodes, out node, element index
};
MeshArrays& current();
}  // namespace refinject
