// pfem_b200_equations.hpp -- host-side shim that plugs the B200 path into a PFEM3D checkout.
//
// Header-only.  It is compiled INSIDE PFEM3D (it includes the reference's own Equation/Solver/Mesh headers) and links
// against libpfem_b200.so through the C ABI of include/pfem_b200.h.  The only edits to the reference are at the
// REGISTER_EQ sites and one call in SolverWCompNewton.  In this repository it is compile-checked
// against shim/mock/*.hpp, which mirror the reference signatures it uses (Mesh.hpp:64-194, Node.hpp:33-83,
// Element.hpp:41-111, Equation.hpp:27-87, Solver.hpp:38-61, Problem.hpp:47-81, SolTable.hpp:10-108).
//
//   MomContEqIncompNewtonB200<dim>  replaces MomContEqIncompNewton<dim> for Solver id "PSPG"
//                                   (physics/IncompNewton/MomContEquation.hpp:20-101): same ctor signature, same
//                                   solve() semantics (Picard on the mesh position, false -> dt is divided).
//   WCompNewtonStepB200<dim>        replaces the body of SolverWCompNewton::m_solveWCompNewtonNoT + computeNextDT
//                                   (physics/WCompNewton/Solver.cpp:192-276) with device-resident states between remeshes.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

#include "pfem_b200.h"

namespace pfem_b200_shim {

inline int deviceFromEnv() {
    const char* e = std::getenv("PFEM_DEVICE");
    return e ? std::atoi(e) : 0;
}
inline void check(pfem_ctx* ctx, int rc, const char* what) {
    if (rc < 0) throw std::runtime_error(std::string(what) + ": " + pfem_last_error(ctx));  // fatal, reference style (MomContEquation.inl:240-241)
}

// Spatial renumbering applied by the shim between the reference's Mesh and the device (PFEM_RENUMBER=0 switches it off).
// PFEM3D's node numbering drifts towards random as nodes are added and removed (Mesh.cpp:27-147, 928-1017); the gather
// kernels are ~3x slower on a randomly numbered 20 M-tet mesh (profiles/).  Nothing in the C ABI
// depends on the numbering, so nodes are uploaded in Morton (Z-order) order of their coordinates and elements by their
// smallest new node index; every nodal array crossing the ABI is mapped with toNew/toOld.
struct Renumbering {
    std::vector<std::size_t> newOfOld, oldOfNew, elemOldOfNew, elemNewOfOld;
    bool active = false;
    static bool enabled() {
        const char* e = std::getenv("PFEM_RENUMBER");
        return !(e && e[0] == '0');
    }
    std::size_t nodeNew(std::size_t n) const { return active ? newOfOld[n] : n; }
    std::size_t nodeOld(std::size_t n) const { return active ? oldOfNew[n] : n; }
    // nodal array in the ABI layout q[n + s*nN]
    std::vector<double> toNew(const std::vector<double>& q, std::size_t nN) const {
        if (!active) return q;
        std::vector<double> o(q.size());
        for (std::size_t s = 0; s < q.size() / nN; ++s)
            for (std::size_t n = 0; n < nN; ++n) o[newOfOld[n] + s * nN] = q[n + s * nN];
        return o;
    }
    std::vector<double> toOld(const std::vector<double>& q, std::size_t nN) const {
        if (!active) return q;
        std::vector<double> o(q.size());
        for (std::size_t s = 0; s < q.size() / nN; ++s)
            for (std::size_t n = 0; n < nN; ++n) o[n + s * nN] = q[newOfOld[n] + s * nN];
        return o;
    }
    template <unsigned short dim, class MeshT> void build(MeshT* pMesh) {
        const std::size_t nN = pMesh->getNodesCount(), nE = pMesh->getElementsCount();
        active = enabled() && nN > 1;
        if (!active) return;
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (std::size_t n = 0; n < nN; ++n)
            for (unsigned short d = 0; d < dim; ++d) {
                const double c = pMesh->getNode(n).getCoordinate(d);
                lo[d] = std::min(lo[d], c);
                hi[d] = std::max(hi[d], c);
            }
        auto spread = [](std::uint64_t v) {  // 21 bits -> every third bit
            v &= 0x1FFFFFull;
            v = (v | (v << 32)) & 0x1F00000000FFFFull;
            v = (v | (v << 16)) & 0x1F0000FF0000FFull;
            v = (v | (v << 8)) & 0x100F00F00F00F00Full;
            v = (v | (v << 4)) & 0x10C30C30C30C30C3ull;
            v = (v | (v << 2)) & 0x1249249249249249ull;
            return v;
        };
        std::vector<std::pair<std::uint64_t, std::size_t>> keys(nN);
        for (std::size_t n = 0; n < nN; ++n) {
            std::uint64_t key = 0;
            for (unsigned short d = 0; d < dim; ++d) {
                const double span = hi[d] > lo[d] ? hi[d] - lo[d] : 1.0;
                const double t = (pMesh->getNode(n).getCoordinate(d) - lo[d]) / span * 2097151.0;
                key |= spread(static_cast<std::uint64_t>(t < 0 ? 0 : (t > 2097151.0 ? 2097151.0 : t))) << d;
            }
            keys[n] = {key, n};
        }
        std::sort(keys.begin(), keys.end());
        oldOfNew.resize(nN);
        newOfOld.resize(nN);
        for (std::size_t k = 0; k < nN; ++k) {
            oldOfNew[k] = keys[k].second;
            newOfOld[keys[k].second] = k;
        }
        std::vector<std::pair<std::size_t, std::size_t>> ek(nE);
        for (std::size_t e = 0; e < nE; ++e) {
            std::size_t m = newOfOld[pMesh->getElement(e).getNodeIndex(0)];
            for (unsigned short k = 1; k <= dim; ++k) m = std::min(m, newOfOld[pMesh->getElement(e).getNodeIndex(k)]);
            ek[e] = {m, e};
        }
        std::sort(ek.begin(), ek.end());
        elemOldOfNew.resize(nE);
        elemNewOfOld.resize(nE);
        for (std::size_t k = 0; k < nE; ++k) {
            elemOldOfNew[k] = ek[k].second;
            elemNewOfOld[ek[k].second] = k;
        }
    }
};

// Mesh -> device: connectivity, flags, positions, states [first, first+count), Dirichlet mask/values (Lua evaluated here,
// serially, exactly where the reference evaluates it: PSPG.inl:206-214, WCompNewton/MomEquation.inl:355-364).
template <unsigned short dim, class MeshT, class SolverT, class ProblemT, class BcTable>
void uploadMesh(pfem_ctx* ctx, MeshT* pMesh, SolverT* pSolver, ProblemT* pProblem, BcTable& bc, unsigned short bcFlag,
                unsigned int firstState, unsigned int stateCount, bool topologyChanged, Renumbering& rn,
                bool withFacets = false) {
    const std::size_t nN = pMesh->getNodesCount(), nE = pMesh->getElementsCount();
    if (topologyChanged) {
        rn.template build<dim>(pMesh);
        std::vector<uint64_t> conn(nE * (dim + 1));
        for (std::size_t e = 0; e < nE; ++e) {
            const auto& element = pMesh->getElement(rn.active ? rn.elemOldOfNew[e] : e);
            for (unsigned short k = 0; k <= dim; ++k) conn[e * (dim + 1) + k] = rn.nodeNew(element.getNodeIndex(k));
        }
        std::vector<uint8_t> flags(nN);
        for (std::size_t n = 0; n < nN; ++n) {
            const auto& node = pMesh->getNode(rn.nodeOld(n));
            flags[n] = (node.isBound() ? PFEM_NODE_BOUND : 0u) | (node.isFree() ? PFEM_NODE_FREE : 0u) |
                       (node.isFixed() ? PFEM_NODE_FIXED : 0u) | (node.isOnFreeSurface() ? PFEM_NODE_FREE_SURFACE : 0u);
        }
        check(ctx, pfem_set_topology(ctx, (int64_t)nN, (int64_t)nE, conn.data(), flags.data()), "pfem_set_topology");
        if (withFacets) {  // gamma > 0: Mesh::m_facetsList for the surface-tension facet loops (PSPG.inl:155-187, MomEquation.inl:312-336)
            const std::size_t nF = pMesh->getFacetsCount();
            std::vector<uint64_t> fNodes(nF * dim), fOut(nF), fElem(nF);
            for (std::size_t f = 0; f < nF; ++f) {
                const auto& facet = pMesh->getFacet(f);
                for (unsigned short k = 0; k < dim; ++k) fNodes[f * dim + k] = rn.nodeNew(facet.getNodeIndex(k));
                fOut[f] = rn.nodeNew(facet.getOutNodeIndex());
                fElem[f] = rn.active ? rn.elemNewOfOld[facet.getElementIndex()] : facet.getElementIndex();
            }
            check(ctx, pfem_set_facets(ctx, (int64_t)nF, fNodes.data(), fOut.data(), fElem.data()), "pfem_set_facets");
        }
    }
    std::vector<double> x(dim * nN), q(stateCount * nN), dval(dim * nN, 0.0);
    std::vector<uint8_t> dmask(nN, 0);
    const double tNext = pProblem->getCurrentSimTime() + pSolver->getTimeStep();
    for (std::size_t nOld = 0; nOld < nN; ++nOld) {  // old order: the Lua BC functions are called as the reference calls them
        const auto& node = pMesh->getNode(nOld);
        const std::size_t n = rn.nodeNew(nOld);
        for (unsigned short d = 0; d < dim; ++d) x[n + d * nN] = node.getCoordinate(d);
        for (unsigned int s = 0; s < stateCount; ++s) q[n + s * nN] = node.getState(firstState + s);
        if (node.isBound() && pSolver->getBcTagFlags(node.getTag(), bcFlag)) {
            const std::array<double, dim> r = bc.template call<std::array<double, dim>>(pMesh->getNodeType(nOld) + "V", node.getPosition(), tNext);
            dmask[n] = 1;
            for (unsigned short d = 0; d < dim; ++d) dval[n + d * nN] = r[d];
        }
    }
    check(ctx, pfem_set_positions(ctx, x.data()), "pfem_set_positions");
    check(ctx, pfem_set_states(ctx, (int)firstState, (int)stateCount, q.data()), "pfem_set_states");
    check(ctx, pfem_set_dirichlet(ctx, dmask.data(), dval.data()), "pfem_set_dirichlet");
}

}  // namespace pfem_b200_shim

// ======================================================================================================================
template <unsigned short dim>
class MomContEqIncompNewtonB200 : public Equation {
public:
    MomContEqIncompNewtonB200(Problem* pProblem, Solver* pSolver, Mesh* pMesh, std::vector<SolTable> solverParams,
                              std::vector<SolTable> materialParams, const std::vector<unsigned short>& bcFlags,
                              const std::vector<unsigned int>& statesIndex)
        : Equation(pProblem, pSolver, pMesh, solverParams, materialParams, bcFlags, statesIndex, "MomContEq") {
        if (m_pSolver->getID() != "PSPG") throw std::runtime_error("the B200 path implements the PSPG solver only");
        if (m_pProblem->getID() != "IncompNewtonNoT") throw std::runtime_error("the B200 path implements IncompNewtonNoT only");
        if (bcFlags.size() != 1 || statesIndex.size() != 1)
            throw std::runtime_error("the " + getID() + " equation requires one BC flag and one statesIndex");  // MomContEquation.inl:64-69
        m_par.rho = m_materialParams[0].template checkAndGet<double>("rho");
        m_par.mu = m_materialParams[0].template checkAndGet<double>("mu");
        m_gamma = m_materialParams[0].template checkAndGet<double>("gamma");
        m_maxIter = m_equationParams[0].template checkAndGet<unsigned int>("maxIter");
        m_minRes = m_equationParams[0].template checkAndGet<double>("minRes");
        m_residual = m_equationParams[0].template checkAndGet<std::string>("residual");
        if (m_residual != "U" && m_residual != "U_P" && m_residual != "Ax_f") throw std::runtime_error("unknown residual type: " + m_residual);
        auto bodyForce = m_equationParams[0].template checkAndGet<std::vector<double>>("bodyForce");
        if (bodyForce.size() != m_pMesh->getDim()) throw std::runtime_error("the body force vector has not the right dimension!");
        for (unsigned short d = 0; d < 3; ++d) m_par.bodyForce[d] = d < dim ? bodyForce[d] : 0.0;
        m_relTol = m_equationParams[0].doesVarExist("krylovTol") ? m_equationParams[0].template checkAndGet<double>("krylovTol") : 1e-12;
        m_needNormalCurv = false;  // facet normals are recomputed on the device (MomContEquation.inl:257 asks the host mesh for them)
        const int rc = pfem_create(&m_ctx, dim, pfem_b200_shim::deviceFromEnv());
        if (rc != PFEM_OK) throw std::runtime_error(std::string("pfem_create: ") + pfem_last_error(nullptr));
        pfem_b200_shim::check(m_ctx, pfem_set_surface_tension(m_ctx, m_gamma), "pfem_set_surface_tension");
        // optional new key: preconditioner = "auto" | "point" | "block" | "mg"   (default auto: multigrid with hand-over)
        if (m_equationParams[0].doesVarExist("preconditioner")) {
            const std::string pre = m_equationParams[0].template checkAndGet<std::string>("preconditioner");
            const int kind = pre == "auto" ? PFEM_PRECOND_AUTO : pre == "point" ? PFEM_PRECOND_POINT : pre == "block" ? PFEM_PRECOND_BLOCK
                           : pre == "mg" ? PFEM_PRECOND_MG : -1;
            if (kind < 0) throw std::runtime_error("unknown preconditioner: " + pre);
            pfem_b200_shim::check(m_ctx, pfem_pspg_set_preconditioner(m_ctx, kind, 0, 0.0), "pfem_pspg_set_preconditioner");
        }
    }
    ~MomContEqIncompNewtonB200() override { pfem_destroy(m_ctx); }

    // MomContEqIncompNewton::solve + PicardAlgo::solve (MomContEquation.inl:287-298, PicardAlgo.cpp:31-94)
    bool solve() override {
        using namespace pfem_b200_shim;
        const std::size_t nN = m_pMesh->getNodesCount();
        m_par.dt = m_pSolver->getTimeStep();
        // the incompressible solver remeshes after every successful step (IncompNewton/Solver.cpp:241-242): new topology
        uploadMesh<dim>(m_ctx, m_pMesh, m_pSolver, m_pProblem, m_bcParams[0], m_bcFlags[0], m_statesIndex[0], dim + 1, true,
                        m_rn, m_gamma >= 1e-15);
        std::vector<double> qPrev((dim + 1) * nN), qIter((dim + 1) * nN, 0.0), qIterPrev;
        for (std::size_t n = 0; n < nN; ++n)
            for (unsigned int s = 0; s <= dim; ++s) qPrev[n + s * nN] = m_pMesh->getNode(n).getState(m_statesIndex[0] + s);
        qIterPrev = qPrev;
        check(m_ctx, pfem_snapshot_positions(m_ctx), "pfem_snapshot_positions");         // m_prepare: saveNodesList
        check(m_ctx, pfem_pspg_assemble(m_ctx, &m_par, m_rn.toNew(qPrev, nN).data()), "pfem_pspg_assemble");
        unsigned int iterCount = 0;
        double res = std::numeric_limits<double>::max();
        while (res > m_minRes) {
            if (iterCount > m_maxIter) return false;                                      // PicardAlgo.cpp:55-65 (host mesh untouched)
            double resAxf = 0;
            int iters = 0;
            const int rc = pfem_pspg_picard_iter(m_ctx, &m_par, nullptr, m_relTol, 100000, qIter.data(), &resAxf, &iters);
            check(m_ctx, rc, "pfem_pspg_picard_iter");
            if (rc == PFEM_NOT_CONVERGED || rc == PFEM_NAN) return false;                 // like a failed factorisation, PSPG.inl:304-312
            qIter = m_rn.toOld(qIter, nN);                                                // device numbering -> the Mesh's
            res = (m_residual == "Ax_f") ? resAxf : relativeChange(qIter, qIterPrev, nN);
            qIterPrev = qIter;
            if (std::isnan(res)) return false;                                            // PicardAlgo.cpp:79-86
            iterCount++;
        }
        // publish the converged iterate to the host mesh: states (PSPG.inl:293) and positions (PSPG.inl:294-295)
        for (std::size_t n = 0; n < nN; ++n)
            for (unsigned int s = 0; s <= dim; ++s) m_pMesh->setNodeState(n, m_statesIndex[0] + s, qIter[n + s * nN]);
        m_pMesh->saveNodesList();
        m_pMesh->updateNodesPositionFromSave(scaled(qIter, m_par.dt, dim * nN));  // the std::vector overload wants exactly dim*nNodes (Mesh.cpp:1197-1203)
        return true;
    }

private:
    // Res::U / Res::U_P (PSPG.inl:319-364): relative change over the nodes that are not free
    double relativeChange(const std::vector<double>& q, const std::vector<double>& qp, std::size_t nN) const {
        auto rel = [&](unsigned first, unsigned last) {
            double num = 0, den = 0;
            for (std::size_t n = 0; n < nN; ++n) {
                if (m_pMesh->getNode(n).isFree()) continue;
                for (unsigned d = first; d <= last; ++d) {
                    const double a = q[n + d * nN], b = qp[n + d * nN];
                    num += (a - b) * (a - b);
                    den += b * b;
                }
            }
            return den == 0 ? std::numeric_limits<double>::max() : std::sqrt(num / den);
        };
        const double resV = rel(0, dim - 1);
        return m_residual == "U_P" ? std::max(resV, rel(dim, dim)) : resV;
    }
    static std::vector<double> scaled(const std::vector<double>& q, double dt, std::size_t count) {
        std::vector<double> d(count);
        for (std::size_t i = 0; i < count; ++i) d[i] = q[i] * dt;
        return d;
    }
    pfem_ctx* m_ctx = nullptr;
    pfem_pspg_params m_par{};
    pfem_b200_shim::Renumbering m_rn;
    double m_gamma = 0.0;
    unsigned int m_maxIter = 10;
    double m_minRes = 1e-6, m_relTol = 1e-12;
    std::string m_residual;
};

// ======================================================================================================================
template <unsigned short dim>
class WCompNewtonStepB200 {
public:
    WCompNewtonStepB200(Problem* pProblem, Solver* pSolver, Mesh* pMesh, SolTable& material, SolTable& contParams,
                        SolTable& momParams, SolTable bcParams, double securityCoeff)
        : m_pProblem(pProblem), m_pSolver(pSolver), m_pMesh(pMesh), m_bc(bcParams), m_securityCoeff(securityCoeff) {
        m_par.mu = material.template checkAndGet<double>("mu");
        m_par.K0 = material.template checkAndGet<double>("K0");
        m_par.K0p = material.template checkAndGet<double>("K0p");
        m_par.rhoStar = material.template checkAndGet<double>("rhoStar");
        m_gamma = material.template checkAndGet<double>("gamma");
        const std::string stab = contParams.template checkAndGet<std::string>("stabilization");
        if (stab != "None" && stab != "Meduri") throw std::runtime_error("unknown stabilization: " + stab);  // ContEquation.inl:32-37
        m_par.meduri = stab == "Meduri";
        const std::string id = m_pSolver->getID();  // ContEquation.inl:38-43
        if (id == "CDS_dpdt") m_par.eqType = PFEM_WC_CDS_DPDT;
        else if (id == "CDS_drhodt") m_par.eqType = PFEM_WC_CDS_DRHODT;
        else if (id == "CDS_rho") m_par.eqType = PFEM_WC_CDS_RHO;
        else throw std::runtime_error("unknown WCompNewton solver id: " + id);
        auto bodyForce = momParams.template checkAndGet<std::vector<double>>("bodyForce");
        for (unsigned short d = 0; d < 3; ++d) m_par.bodyForce[d] = d < dim ? bodyForce[d] : 0.0;
        const int rc = pfem_create(&m_ctx, dim, pfem_b200_shim::deviceFromEnv());
        if (rc != PFEM_OK) throw std::runtime_error(std::string("pfem_create: ") + pfem_last_error(nullptr));
        pfem_b200_shim::check(m_ctx, pfem_set_surface_tension(m_ctx, m_gamma), "pfem_set_surface_tension");
    }
    ~WCompNewtonStepB200() { pfem_destroy(m_ctx); }

    void markRemeshed() { m_dirty = true; }  // call after Mesh::remesh (WCompNewton/Solver.cpp:266-269)

    // m_solveWCompNewtonNoT up to the remesh test (Solver.cpp:236-263); states stay on the device between remeshes
    bool step() {
        using namespace pfem_b200_shim;
        if (m_dirty) {
            uploadMesh<dim>(m_ctx, m_pMesh, m_pSolver, m_pProblem, m_bc, 0, 0, 2 * dim + 2, true, m_rn, m_gamma >= 1e-15);
            m_dirty = false;
        }
        check(m_ctx, pfem_wc_step(m_ctx, &m_par, m_pSolver->getTimeStep()), "pfem_wc_step");
        return true;
    }
    // computeNextDT (Solver.cpp:192-234); throws on NaN like the reference (:231-232)
    double nextDT(double maxDT) {
        double dt = 0;
        const int rc = pfem_wc_next_dt(m_ctx, &m_par, m_securityCoeff, maxDT, &dt);
        pfem_b200_shim::check(m_ctx, rc, "pfem_wc_next_dt");
        if (rc == PFEM_NAN) throw std::runtime_error("NaN time step!");
        return dt;
    }
    // device -> host mesh (before remeshing or extractor output): states and positions
    void download() {
        const std::size_t nN = m_pMesh->getNodesCount();
        std::vector<double> q((2 * dim + 2) * nN), x(dim * nN), xOld(dim * nN);
        pfem_b200_shim::check(m_ctx, pfem_get_states(m_ctx, 0, 2 * dim + 2, q.data()), "pfem_get_states");
        pfem_b200_shim::check(m_ctx, pfem_get_positions(m_ctx, x.data()), "pfem_get_positions");
        q = m_rn.toOld(q, nN);
        x = m_rn.toOld(x, nN);
        for (std::size_t n = 0; n < nN; ++n) {
            for (unsigned int s = 0; s < 2u * dim + 2u; ++s) m_pMesh->setNodeState(n, s, q[n + s * nN]);
            for (unsigned short d = 0; d < dim; ++d) xOld[n + d * nN] = x[n + d * nN] - m_pMesh->getNode(n).getCoordinate(d);
        }
        m_pMesh->updateNodesPosition(xOld);  // delta to the device positions (fixed nodes have delta 0)
    }

private:
    Problem* m_pProblem;
    Solver* m_pSolver;
    Mesh* m_pMesh;
    SolTable m_bc;
    double m_securityCoeff;
    pfem_ctx* m_ctx = nullptr;
    pfem_wc_params m_par{};
    pfem_b200_shim::Renumbering m_rn;
    double m_gamma = 0.0;
    bool m_dirty = true;
};
