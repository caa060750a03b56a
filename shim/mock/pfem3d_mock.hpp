// Minimal stand-ins for the PFEM3D classes the shim touches, with the reference's signatures
// (srcs/mesh/{Mesh,Node,Element}.hpp, srcs/simulation/{Equation,Solver,Problem}.hpp, utility/SolTable.hpp).
// Compile-check only: gmsh/CGAL/Lua/Eigen are not available in this image, so the real headers cannot be included.
#pragma once
#include <array>
#include <cstddef>
#include <map>
#include <string>
#include <vector>

class Node {
public:
    std::array<double, 3> getPosition() const noexcept { return m_position; }
    double getCoordinate(unsigned int xyz) const noexcept { return m_position[xyz]; }
    double getState(unsigned int s) const noexcept { return m_states[s]; }
    int getTag() const noexcept { return m_tag; }
    bool isBound() const noexcept { return m_isBound; }
    bool isFixed() const noexcept { return m_isFixed; }
    bool isFree() const noexcept { return m_free; }
    bool isOnFreeSurface() const noexcept { return m_isOnFreeSurface; }
    std::array<double, 3> m_position{};
    std::vector<double> m_states;
    bool m_isBound = false, m_isOnFreeSurface = false, m_isFixed = false, m_free = false;
    int m_tag = -1;
};
class Element {
public:
    std::size_t getNodeIndex(unsigned int k) const noexcept { return m_nodesIndexes[k]; }
    std::vector<std::size_t> m_nodesIndexes;
};
class Facet {   // Facet.hpp:38-60
public:
    std::size_t getNodeIndex(unsigned int k) const noexcept { return m_nodesIndexes[k]; }
    std::size_t getOutNodeIndex() const noexcept { return m_outNodeIndex; }
    std::size_t getElementIndex() const noexcept { return m_elementIndex; }
    std::vector<std::size_t> m_nodesIndexes;
    std::size_t m_outNodeIndex = 0, m_elementIndex = 0;
};
class Mesh {
public:
    std::size_t getFacetsCount() const noexcept { return m_facetsList.size(); }
    const Facet& getFacet(std::size_t f) const noexcept { return m_facetsList[f]; }
    std::vector<Facet> m_facetsList;
    unsigned short getDim() const noexcept { return m_dim; }
    std::size_t getNodesCount() const noexcept { return m_nodesList.size(); }
    std::size_t getElementsCount() const noexcept { return m_elementsList.size(); }
    const Node& getNode(std::size_t n) const noexcept { return m_nodesList[n]; }
    const Element& getElement(std::size_t e) const noexcept { return m_elementsList[e]; }
    std::string getNodeType(std::size_t) const noexcept { return "Boundary"; }
    void setNodeState(std::size_t n, unsigned int s, double v) noexcept { m_nodesList[n].m_states[s] = v; }
    void saveNodesList() { m_save = m_nodesList; }
    void restoreNodesList() { m_nodesList = m_save; }
    void updateNodesPosition(const std::vector<double>& d) { move(d, m_nodesList); }
    void updateNodesPositionFromSave(const std::vector<double>& d) { move(d, m_save); }
    unsigned short m_dim = 3;
    std::vector<Node> m_nodesList, m_save;
    std::vector<Element> m_elementsList;
private:
    void move(const std::vector<double>& d, const std::vector<Node>& base) {
        for (std::size_t n = 0; n < m_nodesList.size(); ++n)
            if (!m_nodesList[n].m_isFixed)
                for (unsigned short k = 0; k < m_dim; ++k) m_nodesList[n].m_position[k] = base[n].m_position[k] + d[n + k * m_nodesList.size()];
    }
};
class SolTable {
public:
    template <class T> T checkAndGet(const std::string&) const { return T{}; }
    bool doesVarExist(const std::string&) const { return false; }
    template <class R, class... Args> R call(const std::string&, Args...) const { return R{}; }
};
class Problem {
public:
    std::string getID() const noexcept { return "IncompNewtonNoT"; }
    double getCurrentSimTime() const noexcept { return 0.0; }
};
class Solver {
public:
    std::string getID() const noexcept { return "PSPG"; }
    double getTimeStep() const noexcept { return 1e-3; }
    bool getBcTagFlags(int, unsigned short) const noexcept { return true; }
};
class Equation {
public:
    Equation(Problem* pProblem, Solver* pSolver, Mesh* pMesh, std::vector<SolTable> solverParams, std::vector<SolTable> materialParams,
             const std::vector<unsigned short>& bcFlags, const std::vector<unsigned int>& statesIndex, const std::string& id)
        : m_id(id), m_materialParams(materialParams), m_equationParams(solverParams), m_bcParams(solverParams), m_bcFlags(bcFlags),
          m_statesIndex(statesIndex), m_pProblem(pProblem), m_pSolver(pSolver), m_pMesh(pMesh) {}
    virtual ~Equation() = default;
    std::string getID() const noexcept { return m_id; }
    virtual bool solve() { return false; }
protected:
    bool m_needNormalCurv = false;
    std::string m_id;
    std::vector<SolTable> m_materialParams, m_equationParams, m_bcParams;
    std::vector<unsigned short> m_bcFlags;
    std::vector<unsigned int> m_statesIndex;
    Problem* m_pProblem;
    Solver* m_pSolver;
    Mesh* m_pMesh;
};
