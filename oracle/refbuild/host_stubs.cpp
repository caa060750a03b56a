// Host-side stand-ins for the parts of PFEM3D that cannot be compiled here (CGAL remeshing, the Lua-driven Problem).
// TEST INFRASTRUCTURE ONLY — original code; everything on the hot path is the reference's own source, compiled in
// place from /root/reference by oracle/refbuild/Makefile.
//
//  * Mesh::triangulateAlphaShape2D/3D (reference: Mesh2D.cpp:32-251, Mesh3D.cpp:36-286, CGAL alpha shapes) are
//    replaced by a loader of injected connectivity.  What it reproduces from the reference is the bookkeeping after
//    CGAL returns: Element::m_nodesIndexes + computeJ/DetJ/InvJ (Mesh3D.cpp:166-170), Node::m_elements in element
//    order (:204-205), sorted unique m_neighbourNodes (:208-215), facets with computeJ/DetJ/InvJ/Normal (:241-246).
//    isFree stays the reference's own definition (Node.inl: m_elements.empty()).
//  * Mesh::computeFSNormalCurvature2D/3D (Mesh2D.cpp:253, Mesh3D.cpp:288) are no-ops (curvature is off the path).
//  * Problem (Problem.cpp needs Lua, extractors, gmsh): minimal member definitions; the equations only read
//    getID/getThreadCount/getCurrentSimTime/isOutputVerbose/updateTime from it.
#include <algorithm>
#include <stdexcept>

#include "ref_inject.hpp"

#include "mesh/Mesh.hpp"
#include "simulation/Problem.hpp"
#include "simulation/Solver.hpp"
#include "simulation/Equation.hpp"
#include "simulation/extractors/Extractor.hpp"

namespace refinject {
MeshArrays& current() {
    static MeshArrays a;
    return a;
}
}  // namespace refinject

// ---- Mesh members normally defined in Mesh2D.cpp / Mesh3D.cpp ----------------------------------------------------
static void refLoad(Mesh& mesh);

void Mesh::triangulateAlphaShape2D() { refLoad(*this); }
void Mesh::triangulateAlphaShape3D() { refLoad(*this); }
void Mesh::computeFSNormalCurvature2D() {}
void Mesh::computeFSNormalCurvature3D() {}

static void refLoad(Mesh& mesh) {
    // compiled with -fno-access-control: the reference grants the same access to Mesh members (friend class Mesh)
    const refinject::MeshArrays& in = refinject::current();
    if (in.dim != mesh.m_dim) throw std::runtime_error("refbuild: injected dimension differs");
    const std::size_t N = in.nNodes;
    const unsigned npe = in.dim + 1;

    mesh.m_tagNames = in.tagNames;
    mesh.m_elementsList.clear();
    mesh.m_facetsList.clear();
    mesh.m_nodesList.clear();
    mesh.m_nodesList.reserve(N);
    for (std::size_t n = 0; n < N; ++n) {
        Node node(mesh);
        node.m_position = {0, 0, 0};
        for (int d = 0; d < in.dim; ++d) node.m_position[d] = in.x[n + d * N];
        node.m_states.assign(in.nStates, 0.0);
        node.m_isBound = in.flags[n] & 1;
        node.m_isFixed = in.flags[n] & 4;
        node.m_isOnFreeSurface = in.flags[n] & 8;
        node.m_tag = in.tags[n];
        mesh.m_nodesList.push_back(std::move(node));
    }

    mesh.m_elementsList.resize(in.nElems);
    for (std::size_t e = 0; e < in.nElems; ++e) {
        Element element(mesh);
        element.m_nodesIndexes.resize(npe);
        for (unsigned k = 0; k < npe; ++k) element.m_nodesIndexes[k] = static_cast<std::size_t>(in.conn[e * npe + k]);
        element.computeJ();
        element.computeDetJ();
        element.computeInvJ();
        for (unsigned a = 0; a < npe; ++a)
            for (unsigned b = 0; b < npe; ++b)
                if (a != b) mesh.m_nodesList[element.m_nodesIndexes[a]].m_neighbourNodes.push_back(element.m_nodesIndexes[b]);
        mesh.m_elementsList[e] = std::move(element);
        for (std::size_t index : mesh.m_elementsList[e].m_nodesIndexes) mesh.m_nodesList[index].m_elements.push_back(e);
    }
    for (std::size_t n = 0; n < N; ++n) {
        auto& nb = mesh.m_nodesList[n].m_neighbourNodes;
        std::sort(nb.begin(), nb.end());
        nb.erase(std::unique(nb.begin(), nb.end()), nb.end());
        const bool freeFlag = in.flags[n] & 2;
        if (freeFlag != mesh.m_nodesList[n].isFree())
            throw std::runtime_error("refbuild: isFree flag of node " + std::to_string(n) + " disagrees with the connectivity");
    }

    const unsigned npf = in.dim;
    for (std::size_t f = 0; f < in.nFacets; ++f) {
        const std::int64_t* row = in.facets + f * (npf + 2);
        Facet facet(mesh);
        facet.m_nodesIndexes.resize(npf);
        for (unsigned k = 0; k < npf; ++k) facet.m_nodesIndexes[k] = static_cast<std::size_t>(row[k]);
        facet.m_outNodeIndex = static_cast<std::size_t>(row[npf]);
        facet.m_elementIndex = static_cast<std::size_t>(row[npf + 1]);
        facet.computeJ();
        facet.computeDetJ();
        facet.computeInvJ();
        if (mesh.m_computeNormalCurvature) facet.computeNormal();
        mesh.m_facetsList.push_back(std::move(facet));
        for (std::size_t index : mesh.m_facetsList.back().m_nodesIndexes) mesh.m_nodesList[index].m_facets.push_back(mesh.m_facetsList.size() - 1);
    }
}

// ---- Problem (reference: Problem.cpp; needs Lua + extractors) ----------------------------------------------------
Problem::Problem(const std::string& id)
    : m_id(id), m_time(0), m_maxTime(0), m_step(0), m_statesNumber(0), m_verboseOutput(false), m_nThreads(1) {}
Problem::~Problem() {}
void Problem::displayParams() const {}
void Problem::displayTimeStats() const {}
std::string Problem::getID() const noexcept { return m_id; }
std::vector<std::string> Problem::getWrittableDataName() const { return {}; }
std::vector<double> Problem::getWrittableData(const std::string&, std::size_t) const { return {}; }
std::vector<std::string> Problem::getGlobalWrittableDataName() const { return {}; }
double Problem::getGlobalWrittableData(const std::string&) const { return 0; }
std::vector<std::string> Problem::getMeshWrittableDataName() const { return {}; }
std::vector<double> Problem::getMeshWrittableData(const std::string&, std::size_t) const { return {}; }
std::vector<std::string> Problem::getBoundaryWrittableDataName() const { return {}; }
std::vector<double> Problem::getBoundaryWrittableData(const std::string&, const std::string&) const { return {}; }
void Problem::dump() {}
void Problem::simulate() {}
void Problem::addExtractors() {}
void Problem::setInitialCondition() {}
void Problem::updateTime(double timeStep) {  // Problem.cpp:297-301
    m_time += timeStep;
    m_step++;
}
