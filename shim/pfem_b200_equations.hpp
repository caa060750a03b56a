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
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
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

// Devices of this process: PFEM_DEVICES="0,1,2,3" (default: PFEM_DEVICE or 0).  With more than one entry the mesh is
// RCB-partitioned by the library (pfem_partition_*) and every device gets one context driven by its own host thread
// (pfem_comm_local_*): this is how the reference's single host process reaches several GPUs.  An id may repeat (several
// ranks on one GPU: how the drop-in test exercises the path on a single-GPU box).
inline std::vector<int> devicesFromEnv() {
    std::vector<int> d;
    if (const char* e = std::getenv("PFEM_DEVICES")) {
        std::stringstream ss(e);
        std::string tok;
        while (std::getline(ss, tok, ','))
            if (!tok.empty()) d.push_back(std::atoi(tok.c_str()));
    }
    if (d.empty()) d.push_back(deviceFromEnv());
    return d;
}
struct RankSet {
    std::vector<pfem_ctx*> ctx;
    void* group = nullptr;
    // local mesh of every rank (multi-device only): local -> global node ids, owned-node count
    std::vector<std::vector<int64_t>> l2g;
    std::vector<int64_t> nOwned;
    int n() const { return (int)ctx.size(); }
    bool multi() const { return ctx.size() > 1; }
    RankSet(int dim) {
        const std::vector<int> dev = devicesFromEnv();
        ctx.assign(dev.size(), nullptr);
        for (std::size_t r = 0; r < dev.size(); ++r)
            if (pfem_create(&ctx[r], dim, dev[r]) != PFEM_OK) {
                const std::string msg = std::string("pfem_create: ") + pfem_last_error(nullptr);
                destroy();
                throw std::runtime_error(msg);
            }
        if (multi()) {
            if (pfem_comm_local_create(n(), &group) != PFEM_OK) {
                destroy();
                throw std::runtime_error(std::string("pfem_comm_local_create: ") + pfem_last_error(nullptr));
            }
            for (int r = 0; r < n(); ++r) check(ctx[r], pfem_comm_init_local(ctx[r], group, r), "pfem_comm_init_local");
            l2g.resize(n());
            nOwned.assign(n(), 0);
        }
    }
    RankSet(const RankSet&) = delete;
    RankSet& operator=(const RankSet&) = delete;
    void destroy() {
        for (auto* c : ctx)
            if (c) pfem_destroy(c);
        ctx.clear();
        if (group) pfem_comm_local_destroy(group);
        group = nullptr;
    }
    ~RankSet() { destroy(); }
    // run f(rank) on every rank concurrently (collective library calls meet at internal barriers); rethrows the first failure
    template <class F> void each(F f) {
        if (!multi()) {
            f(0);
            return;
        }
        std::vector<std::string> err(n());
        std::vector<std::thread> th;
        for (int r = 0; r < n(); ++r)
            th.emplace_back([&, r] {
                try {
                    f(r);
                } catch (const std::exception& e) {
                    err[r] = e.what();
                    pfem_comm_abort(ctx[r]);
                }
            });
        for (auto& t : th) t.join();
        for (auto& e : err)
            if (!e.empty() && e.find("aborted by another rank") == std::string::npos) throw std::runtime_error(e);
        for (auto& e : err)
            if (!e.empty()) throw std::runtime_error(e);
    }
    // nodal SoA array of the whole mesh (nComp x nN) -> the local array of rank r
    std::vector<double> scatter(int r, const std::vector<double>& q, std::size_t nN) const {
        const auto& m = l2g[r];
        const std::size_t nComp = q.size() / nN, nl = m.size();
        std::vector<double> o(nComp * nl);
        for (std::size_t s = 0; s < nComp; ++s)
            for (std::size_t k = 0; k < nl; ++k) o[k + s * nl] = q[(std::size_t)m[k] + s * nN];
        return o;
    }
    // owned entries of the local array of rank r -> the whole-mesh array (ranks write disjoint entries)
    void gatherOwned(int r, const std::vector<double>& loc, std::vector<double>& q, std::size_t nN) const {
        const auto& m = l2g[r];
        const std::size_t nl = m.size(), nComp = loc.size() / nl;
        for (std::size_t s = 0; s < nComp; ++s)
            for (int64_t k = 0; k < nOwned[r]; ++k) q[(std::size_t)m[k] + s * nN] = loc[(std::size_t)k + s * nl];
    }
};

// whole-mesh download in the device numbering: states [first, first+count) and positions
inline void downloadAll(RankSet& R, std::size_t nN, int dim, int first, int count, std::vector<double>* q, std::vector<double>* x) {
    if (q) q->assign((std::size_t)count * nN, 0.0);
    if (x) x->assign((std::size_t)dim * nN, 0.0);
    R.each([&](int r) {
        const std::size_t nl = R.multi() ? R.l2g[r].size() : nN;
        std::vector<double> ql((std::size_t)count * nl), xl((std::size_t)dim * nl);
        if (q) check(R.ctx[r], pfem_get_states(R.ctx[r], first, count, ql.data()), "pfem_get_states");
        if (x) check(R.ctx[r], pfem_get_positions(R.ctx[r], xl.data()), "pfem_get_positions");
        if (!R.multi()) {
            if (q) *q = ql;
            if (x) *x = xl;
        } else {
            if (q) R.gatherOwned(r, ql, *q, nN);
            if (x) R.gatherOwned(r, xl, *x, nN);
        }
    });
}

// Mesh -> device: connectivity, flags, positions, states [first, first+count), Dirichlet mask/values (Lua evaluated here,
// serially, exactly where the reference evaluates it: PSPG.inl:206-214, WCompNewton/MomEquation.inl:355-364).
// Velocity Dirichlet data of the bound nodes whose tag carries a "<type>V" function, evaluated at time tNext and at the
// given positions (device numbering, null: the host mesh's coordinates) -- what the reference evaluates inside
// m_applyBCPSPG (PSPG.inl:206-214) and, on EVERY explicit step, MomEqWCompNewton::m_applyBC (MomEquation.inl:355-371).
template <unsigned short dim, class MeshT, class SolverT, class BcTable>
void evalDirichlet(MeshT* pMesh, SolverT* pSolver, BcTable& bc, unsigned short bcFlag, double tNext, const Renumbering& rn,
                   const std::vector<double>* xDevice, std::vector<uint8_t>& dmask, std::vector<double>& dval, bool* anyMovingBcNode) {
    const std::size_t nN = pMesh->getNodesCount();
    dmask.assign(nN, 0);
    dval.assign(dim * nN, 0.0);
    if (anyMovingBcNode) *anyMovingBcNode = false;
    for (std::size_t nOld = 0; nOld < nN; ++nOld) {  // old order: the Lua BC functions are called as the reference calls them
        const auto& node = pMesh->getNode(nOld);
        if (!(node.isBound() && pSolver->getBcTagFlags(node.getTag(), bcFlag))) continue;
        const std::size_t n = rn.nodeNew(nOld);
        std::array<double, 3> pos = node.getPosition();
        if (xDevice && !node.isFixed())
            for (unsigned short d = 0; d < dim; ++d) pos[d] = (*xDevice)[n + d * nN];
        if (anyMovingBcNode && !node.isFixed()) *anyMovingBcNode = true;
        const std::array<double, dim> r = bc.template call<std::array<double, dim>>(pMesh->getNodeType(nOld) + "V", pos, tNext);
        dmask[n] = 1;
        for (unsigned short d = 0; d < dim; ++d) dval[n + d * nN] = r[d];
    }
}
inline void pushDirichlet(RankSet& R, std::size_t nN, const std::vector<uint8_t>& dmask, const std::vector<double>& dval) {
    R.each([&](int r) {
        if (!R.multi()) {
            check(R.ctx[r], pfem_set_dirichlet(R.ctx[r], dmask.data(), dval.data()), "pfem_set_dirichlet");
            return;
        }
        const auto& m = R.l2g[r];
        std::vector<uint8_t> ml(m.size());
        for (std::size_t k = 0; k < m.size(); ++k) ml[k] = dmask[(std::size_t)m[k]];
        const std::vector<double> vl = R.scatter(r, dval, nN);
        check(R.ctx[r], pfem_set_dirichlet(R.ctx[r], ml.data(), vl.data()), "pfem_set_dirichlet");
    });
}

// "<type>T"(pos, t + dt) of the nodes whose tag carries flag `bcFlag` (IncompNewton/HeatEquation.inl:383-408,
// WCompNewton/HeatEquation.inl:226-244), device numbering
template <class MeshT, class SolverT, class BcTable>
void evalTemperatureBc(MeshT* pMesh, SolverT* pSolver, BcTable& bc, unsigned short bcFlag, double tNext, const Renumbering& rn,
                       std::vector<uint8_t>& tmask, std::vector<double>& tval) {
    const std::size_t nN = pMesh->getNodesCount();
    tmask.assign(nN, 0);
    tval.assign(nN, 0.0);
    for (std::size_t nOld = 0; nOld < nN; ++nOld) {
        const auto& node = pMesh->getNode(nOld);
        if (!pSolver->getBcTagFlags(node.getTag(), bcFlag)) continue;
        const std::array<double, 1> res = bc.template call<std::array<double, 1>>(pMesh->getNodeType(nOld) + "T", node.getPosition(), tNext);
        tmask[rn.nodeNew(nOld)] = 1;
        tval[rn.nodeNew(nOld)] = res[0];
    }
}
constexpr unsigned short kNoVelocityBc = 0xffff;  // bcFlag value: do not evaluate / upload velocity Dirichlet data
// Mesh -> device(s): connectivity, flags, positions, states [first, first+count), Dirichlet mask/values.
template <unsigned short dim, class MeshT, class SolverT, class ProblemT, class BcTable>
void uploadMesh(RankSet& R, MeshT* pMesh, SolverT* pSolver, ProblemT* pProblem, BcTable& bc, unsigned short bcFlag,
                unsigned int firstState, unsigned int stateCount, bool topologyChanged, Renumbering& rn,
                bool withFacets = false, bool* anyMovingBcNode = nullptr) {
    const std::size_t nN = pMesh->getNodesCount(), nE = pMesh->getElementsCount();
    std::vector<double> x(dim * nN), q(stateCount * nN), dval;
    std::vector<uint8_t> dmask;
    if (topologyChanged) rn.template build<dim>(pMesh);
    for (std::size_t nOld = 0; nOld < nN; ++nOld) {
        const auto& node = pMesh->getNode(nOld);
        const std::size_t n = rn.nodeNew(nOld);
        for (unsigned short d = 0; d < dim; ++d) x[n + d * nN] = node.getCoordinate(d);
        for (unsigned int s = 0; s < stateCount; ++s) q[n + s * nN] = node.getState(firstState + s);
    }
    const double tNext = pProblem->getCurrentSimTime() + pSolver->getTimeStep();
    const bool withVelocityBc = bcFlag != kNoVelocityBc;  // the heat equation's context carries no velocity data
    if (withVelocityBc) evalDirichlet<dim>(pMesh, pSolver, bc, bcFlag, tNext, rn, nullptr, dmask, dval, anyMovingBcNode);
    if (topologyChanged) {
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
        if (!R.multi()) {
            pfem_ctx* ctx = R.ctx[0];
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
        } else {
            if (withFacets) throw std::runtime_error("the B200 path supports surface tension on single-device runs only");
            // RCB partition by the library; every rank gets its local mesh (owned nodes + one ghost-element layer) and halo plan
            pfem_partition* part = nullptr;
            if (pfem_partition_create(&part, dim, (int64_t)nN, (int64_t)nE, conn.data(), x.data(), R.n()) != PFEM_OK)
                throw std::runtime_error("pfem_partition_create failed");
            std::shared_ptr<pfem_partition> guard(part, [](pfem_partition* p) { pfem_partition_destroy(p); });
            for (int r = 0; r < R.n(); ++r) {  // (builds the local parts serially: they share host scratch inside the handle)
                int64_t nl = 0, no = 0, ne = 0, ns = 0;
                int32_t np = 0;
                if (pfem_partition_local_sizes(part, r, &nl, &no, &ne, &np, &ns) != PFEM_OK) throw std::runtime_error("pfem_partition_local_sizes failed");
                R.l2g[r].assign((std::size_t)nl, 0);
                R.nOwned[r] = no;
            }
            R.each([&](int r) {
                int64_t nl = 0, no = 0, ne = 0, ns = 0;
                int32_t np = 0;
                pfem_partition_local_sizes(part, r, &nl, &no, &ne, &np, &ns);
                std::vector<int64_t> l2gElems((std::size_t)ne), sendOff((std::size_t)np + 1), recvStart((std::size_t)np), recvCount((std::size_t)np);
                std::vector<uint64_t> lconn((std::size_t)ne * (dim + 1));
                std::vector<int32_t> peers((std::size_t)np), sendIdx((std::size_t)ns);
                if (pfem_partition_local_get(part, r, R.l2g[r].data(), l2gElems.data(), lconn.data(), peers.data(), sendOff.data(),
                                             sendIdx.data(), recvStart.data(), recvCount.data()) != PFEM_OK)
                    throw std::runtime_error("pfem_partition_local_get failed");
                std::vector<uint8_t> lflags((std::size_t)nl);
                for (std::size_t k = 0; k < (std::size_t)nl; ++k) lflags[k] = flags[(std::size_t)R.l2g[r][k]];
                pfem_ctx* ctx = R.ctx[r];
                check(ctx, pfem_set_topology(ctx, nl, ne, lconn.data(), lflags.data()), "pfem_set_topology");
                check(ctx, pfem_set_partition(ctx, no, np, peers.data(), sendOff.data(), sendIdx.data(), recvStart.data(), recvCount.data()),
                      "pfem_set_partition");
            });
        }
    }
    R.each([&](int r) {
        pfem_ctx* ctx = R.ctx[r];
        if (!R.multi()) {
            check(ctx, pfem_set_positions(ctx, x.data()), "pfem_set_positions");
            if (stateCount > 0) check(ctx, pfem_set_states(ctx, (int)firstState, (int)stateCount, q.data()), "pfem_set_states");
        } else {
            check(ctx, pfem_set_positions(ctx, R.scatter(r, x, nN).data()), "pfem_set_positions");
            if (stateCount > 0)
                check(ctx, pfem_set_states(ctx, (int)firstState, (int)stateCount, R.scatter(r, q, nN).data()), "pfem_set_states");
        }
    });
    if (withVelocityBc) pushDirichlet(R, nN, dmask, dval);
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
        const std::string problemId = m_pProblem->getID();
        if (problemId != "IncompNewtonNoT" && problemId != "Bingham" && problemId != "Boussinesq")
            throw std::runtime_error("the B200 PSPG equation implements the IncompNewtonNoT, Bingham and Boussinesq problems");
        m_isBoussinesq = problemId == "Boussinesq";
        if (bcFlags.size() != 1 || statesIndex.size() != (m_isBoussinesq ? 2u : 1u))  // MomContEquation.inl:64-69; IN/Solver.cpp:63-64
            throw std::runtime_error("the " + getID() + " equation requires one BC flag and one statesIndex (two for Boussinesq)");
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
        m_ranks.reset(new pfem_b200_shim::RankSet(dim));
        if (m_ranks->multi()) throw std::runtime_error("the B200 PSPG equation drives one device from the shim (PFEM_DEVICES lists several)");
        m_ctx = m_ranks->ctx[0];
        pfem_b200_shim::check(m_ctx, pfem_set_surface_tension(m_ctx, m_gamma), "pfem_set_surface_tension");
        if (m_isBoussinesq) {  // MomContEquation.inl:32-38, 166-199: F and H carry (1 - alpha (T - Tr)); T is state statesIndex[1]
            pfem_thermal_params th{};
            th.alpha = m_materialParams[0].template checkAndGet<double>("alpha");
            th.Tr = m_materialParams[0].template checkAndGet<double>("Tr");
            th.k = m_materialParams[0].template checkAndGet<double>("k");
            th.cv = m_materialParams[0].template checkAndGet<double>("cv");
            pfem_b200_shim::check(m_ctx, pfem_set_thermal(m_ctx, &th), "pfem_set_thermal");
        }
        if (problemId == "Bingham") {  // MomContEquation.inl:54-58, 102-119: regularised yield-stress viscosity
            const double tau0 = m_materialParams[0].template checkAndGet<double>("tau0");
            const double mReg = m_materialParams[0].template checkAndGet<double>("mReg");
            pfem_b200_shim::check(m_ctx, pfem_set_bingham(m_ctx, 1, tau0, mReg), "pfem_set_bingham");
        }
        // optional new key: preconditioner = "auto" | "point" | "block" | "mg"   (default auto: multigrid with hand-over)
        if (m_equationParams[0].doesVarExist("preconditioner")) {
            const std::string pre = m_equationParams[0].template checkAndGet<std::string>("preconditioner");
            const int kind = pre == "auto" ? PFEM_PRECOND_AUTO : pre == "point" ? PFEM_PRECOND_POINT : pre == "block" ? PFEM_PRECOND_BLOCK
                           : pre == "mg" ? PFEM_PRECOND_MG : -1;
            if (kind < 0) throw std::runtime_error("unknown preconditioner: " + pre);
            pfem_b200_shim::check(m_ctx, pfem_pspg_set_preconditioner(m_ctx, kind, 0, 0.0), "pfem_pspg_set_preconditioner");
        }
    }
    ~MomContEqIncompNewtonB200() override = default;  // the rank set destroys its context

    // MomContEqIncompNewton::solve + PicardAlgo::solve (MomContEquation.inl:287-298, PicardAlgo.cpp:31-94)
    bool solve() override {
        using namespace pfem_b200_shim;
        const std::size_t nN = m_pMesh->getNodesCount();
        m_par.dt = m_pSolver->getTimeStep();
        // the incompressible solver remeshes after every successful step (IncompNewton/Solver.cpp:241-242): new topology
        // (Dirichlet data: evaluated here at t + dt and the start-of-step positions; the reference re-evaluates it on every
        //  Picard iterate at the moved positions, PSPG.inl:206-214 -- identical for fixed boundary nodes and for
        //  position-independent BC functions, which is every example of the reference)
        uploadMesh<dim>(*m_ranks, m_pMesh, m_pSolver, m_pProblem, m_bcParams[0], m_bcFlags[0], m_statesIndex[0], dim + 1, true,
                        m_rn, m_gamma >= 1e-15);
        if (m_isBoussinesq) {  // the temperature the buoyancy factors read: the host mesh's current T state
            std::vector<double> T(nN);
            for (std::size_t nOld = 0; nOld < nN; ++nOld) T[m_rn.nodeNew(nOld)] = m_pMesh->getNode(nOld).getState(m_statesIndex[1]);
            check(m_ctx, pfem_set_temperature(m_ctx, T.data()), "pfem_set_temperature");
        }
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
    std::unique_ptr<pfem_b200_shim::RankSet> m_ranks;
    pfem_ctx* m_ctx = nullptr;
    pfem_pspg_params m_par{};
    pfem_b200_shim::Renumbering m_rn;
    double m_gamma = 0.0;
    bool m_isBoussinesq = false;
    unsigned int m_maxIter = 10;
    double m_minRes = 1e-6, m_relTol = 1e-12;
    std::string m_residual;
};

// ======================================================================================================================
// HeatEqIncompNewton (IncompNewton/HeatEquation.inl) without phase change and without flux facet terms: the implicit heat
// system M(cv rho) + dt L(k), b = M theta_prev with Dirichlet / free-node rows, solved by Jacobi-CG started from the current
// temperature (the reference: Eigen::ConjugateGradient::solveWithGuess, :129-135; the Picard loop runs once, :206-207).
// Registered where IN/Solver.cpp:71-74 registers HeatEqIncompNewton (problem ids "Boussinesq" and "Conduction").
// Its device context is its own (the momentum-continuity equation of the same solver keeps another one): the two
// equations solve on different meshes anyway when solveHeatFirst is false (the momentum solve moves the nodes).
template <unsigned short dim>
class HeatEqIncompNewtonB200 : public Equation {
public:
    HeatEqIncompNewtonB200(Problem* pProblem, Solver* pSolver, Mesh* pMesh, std::vector<SolTable> solverParams,
                           std::vector<SolTable> materialParams, const std::vector<unsigned short>& bcFlags,
                           const std::vector<unsigned int>& statesIndex)
        : Equation(pProblem, pSolver, pMesh, solverParams, materialParams, bcFlags, statesIndex, "HeatEq") {
        m_k = m_materialParams[0].template checkAndGet<double>("k");
        m_rho = m_materialParams[0].template checkAndGet<double>("rho");
        m_cv = m_materialParams[0].template checkAndGet<double>("cv");
        const double h = m_materialParams[0].template checkAndGet<double>("h");
        m_materialParams[0].template checkAndGet<double>("Tinf");
        const double epsRad = m_materialParams[0].template checkAndGet<double>("epsRad");
        if (h != 0.0 || epsRad != 0.0) throw std::runtime_error("the B200 heat equation has no convection / radiation facet terms (h, epsRad must be 0)");
        for (const char* key : {"Tm", "C", "eps", "DT", "Lm"})
            if (m_materialParams[0].doesVarExist(key)) throw std::runtime_error("the B200 heat equation has no phase change");
        if (bcFlags.size() != 4) throw std::runtime_error("the " + getID() + " equation require two flags for four possible boundary conditions!");
        if (statesIndex.size() != 1) throw std::runtime_error("the " + getID() + " equation require one state index describing the T state !");
        m_equationParams[0].template checkAndGet<unsigned int>("maxIter");
        m_equationParams[0].template checkAndGet<double>("minRes");
        const std::string residual = m_equationParams[0].template checkAndGet<std::string>("residual");
        if (residual != "T" && residual != "Ax_f") throw std::runtime_error("unknown residual type: " + residual);
        m_relTol = m_equationParams[0].doesVarExist("krylovTol") ? m_equationParams[0].template checkAndGet<double>("krylovTol") : 1e-12;
        m_needNormalCurv = false;
        m_ranks.reset(new pfem_b200_shim::RankSet(dim));
        if (m_ranks->multi()) throw std::runtime_error("the B200 heat equation drives one device from the shim (PFEM_DEVICES lists several)");
        m_ctx = m_ranks->ctx[0];
    }
    ~HeatEqIncompNewtonB200() override = default;

    bool solve() override {
        using namespace pfem_b200_shim;
        const std::size_t nN = m_pMesh->getNodesCount();
        const double dt = m_pSolver->getTimeStep();
        // flux BCs (flags 2-4) are not taken over: refuse them instead of dropping them silently
        for (std::size_t n = 0; n < nN; ++n) {
            const auto& node = m_pMesh->getNode(n);
            for (unsigned short f = 1; f < 4; ++f)
                if (m_pSolver->getBcTagFlags(node.getTag(), m_bcFlags[f])) throw std::runtime_error("the B200 heat equation has no flux boundary conditions (Q, Qh, Qr)");
        }
        uploadMesh<dim>(*m_ranks, m_pMesh, m_pSolver, m_pProblem, m_bcParams[0], kNoVelocityBc, 0, 0, true, m_rn, false);
        std::vector<double> T(nN);
        for (std::size_t nOld = 0; nOld < nN; ++nOld) T[m_rn.nodeNew(nOld)] = m_pMesh->getNode(nOld).getState(m_statesIndex[0]);
        std::vector<uint8_t> tmask;
        std::vector<double> tval;
        evalTemperatureBc(m_pMesh, m_pSolver, m_bcParams[0], m_bcFlags[0], m_pProblem->getCurrentSimTime() + dt, m_rn, tmask, tval);
        check(m_ctx, pfem_set_temperature(m_ctx, T.data()), "pfem_set_temperature");           // solveWithGuess: start from theta_prev
        check(m_ctx, pfem_set_temperature_bc(m_ctx, tmask.data(), tval.data()), "pfem_set_temperature_bc");
        check(m_ctx, pfem_heat_assemble(m_ctx, m_rho, m_cv, m_k, dt, T.data()), "pfem_heat_assemble");
        std::vector<double> Tn(nN);
        int iters = 0;
        double relRes = 0;
        const int rc = pfem_heat_solve(m_ctx, m_relTol, 100000, Tn.data(), &iters, &relRes);
        check(m_ctx, rc, "pfem_heat_solve");
        if (rc == PFEM_NOT_CONVERGED || rc == PFEM_NAN) return false;
        for (std::size_t nOld = 0; nOld < nN; ++nOld) m_pMesh->setNodeState(nOld, m_statesIndex[0], Tn[m_rn.nodeNew(nOld)]);
        return true;
    }

private:
    std::unique_ptr<pfem_b200_shim::RankSet> m_ranks;
    pfem_ctx* m_ctx = nullptr;
    pfem_b200_shim::Renumbering m_rn;
    double m_k = 0, m_rho = 0, m_cv = 0, m_relTol = 1e-12;
};

// ======================================================================================================================
template <unsigned short dim>
class WCompNewtonStepB200 {
public:
    // heatBc: the BC table of the heat equation (m_pEquations[2]->getBCParam(0)), required for problem id "BoussinesqWC"
    WCompNewtonStepB200(Problem* pProblem, Solver* pSolver, Mesh* pMesh, SolTable& material, SolTable& contParams,
                        SolTable& momParams, SolTable bcParams, double securityCoeff, const SolTable* heatBc = nullptr)
        : m_pProblem(pProblem), m_pSolver(pSolver), m_pMesh(pMesh), m_bc(bcParams), m_securityCoeff(securityCoeff) {
        if (m_pProblem->getID() == "BoussinesqWC") {  // WCompNewton/Solver.cpp:72-130, HeatEquation.inl:26-31, MomEquation.inl:35-40
            if (!heatBc) throw std::runtime_error("BoussinesqWC needs the heat equation's BC table");
            m_heatBc.reset(new SolTable(*heatBc));
            m_thermal.k = material.template checkAndGet<double>("k");
            m_thermal.cv = material.template checkAndGet<double>("cv");
            m_thermal.alpha = material.template checkAndGet<double>("alpha");
            m_thermal.Tr = material.template checkAndGet<double>("Tr");
            m_isThermal = true;
        }
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
        m_ranks.reset(new pfem_b200_shim::RankSet(dim));
        for (auto* c : m_ranks->ctx) {
            pfem_b200_shim::check(c, pfem_set_surface_tension(c, m_gamma), "pfem_set_surface_tension");
            if (m_isThermal) pfem_b200_shim::check(c, pfem_set_thermal(c, &m_thermal), "pfem_set_thermal");
        }
    }
    ~WCompNewtonStepB200() = default;

    void markRemeshed() { m_dirty = true; }  // call after Mesh::remesh (WCompNewton/Solver.cpp:266-269)

    // m_solveWCompNewtonNoT up to the remesh test (Solver.cpp:236-263); states stay on the device between remeshes
    bool step() {
        using namespace pfem_b200_shim;
        RankSet& R = *m_ranks;
        const std::size_t nN = m_pMesh->getNodesCount();
        if (m_dirty) {
            uploadMesh<dim>(R, m_pMesh, m_pSolver, m_pProblem, m_bc, 0, 0, 2 * dim + 2, true, m_rn, m_gamma >= 1e-15, &m_movingBcNodes);
            m_dirty = false;
            m_dmask.clear();  // the upload evaluated the BC table at t + dt: the cache below restarts
            if (m_isThermal) {  // the temperature is the extra node state 2 dim + 2 (WCompNewton/Solver.cpp:91-92)
                std::vector<double> T(nN);
                for (std::size_t nOld = 0; nOld < nN; ++nOld) T[m_rn.nodeNew(nOld)] = m_pMesh->getNode(nOld).getState(2 * dim + 2);
                R.each([&](int r) {
                    const std::vector<double> Tl = R.multi() ? R.scatter(r, T, nN) : T;
                    check(R.ctx[r], pfem_set_temperature(R.ctx[r], Tl.data()), "pfem_set_temperature");
                });
                m_tmask.clear();
            }
        } else {
            // the reference calls "<type>V"(pos, t + dt) for every bound node on EVERY explicit step (MomEquation.inl:355-371):
            // re-evaluate the table at the new time (and, for boundary nodes that move, at the device positions) and upload
            // it when it changed -- a constant table costs the Lua calls only, like on the host
            std::vector<uint8_t> dmask;
            std::vector<double> dval, xDev;
            if (m_movingBcNodes) downloadAll(R, nN, dim, 0, 1, nullptr, &xDev);
            const double tNext = m_pProblem->getCurrentSimTime() + m_pSolver->getTimeStep();
            evalDirichlet<dim>(m_pMesh, m_pSolver, m_bc, 0, tNext, m_rn, m_movingBcNodes ? &xDev : nullptr, dmask, dval, &m_movingBcNodes);
            if (dmask != m_dmask || dval != m_dval) pushDirichlet(R, nN, dmask, dval);
            m_dmask.swap(dmask);
            m_dval.swap(dval);
        }
        if (m_isThermal) {  // "<type>T"(pos, t + dt) of the nodes whose tag carries flag 1 (HeatEquation.inl:226-244), every step
            std::vector<uint8_t> tmask(nN, 0);
            std::vector<double> tval(nN, 0.0);
            const double tNext = m_pProblem->getCurrentSimTime() + m_pSolver->getTimeStep();
            for (std::size_t nOld = 0; nOld < nN; ++nOld) {
                const auto& node = m_pMesh->getNode(nOld);
                if (!m_pSolver->getBcTagFlags(node.getTag(), 1)) continue;
                const std::array<double, 1> res = m_heatBc->template call<std::array<double, 1>>(m_pMesh->getNodeType(nOld) + "T", node.getPosition(), tNext);
                tmask[m_rn.nodeNew(nOld)] = 1;
                tval[m_rn.nodeNew(nOld)] = res[0];
            }
            if (tmask != m_tmask || tval != m_tval) {
                R.each([&](int r) {
                    if (!R.multi()) {
                        check(R.ctx[r], pfem_set_temperature_bc(R.ctx[r], tmask.data(), tval.data()), "pfem_set_temperature_bc");
                        return;
                    }
                    const auto& m = R.l2g[r];
                    std::vector<uint8_t> ml(m.size());
                    for (std::size_t k = 0; k < m.size(); ++k) ml[k] = tmask[(std::size_t)m[k]];
                    const std::vector<double> vl = R.scatter(r, tval, nN);
                    check(R.ctx[r], pfem_set_temperature_bc(R.ctx[r], ml.data(), vl.data()), "pfem_set_temperature_bc");
                });
                m_tmask.swap(tmask);
                m_tval.swap(tval);
            }
        }
        const double dt = m_pSolver->getTimeStep();
        R.each([&](int r) { check(R.ctx[r], pfem_wc_step(R.ctx[r], &m_par, dt), "pfem_wc_step"); });
        return true;
    }
    // computeNextDT (Solver.cpp:192-234); throws on NaN like the reference (:231-232)
    double nextDT(double maxDT) {
        pfem_b200_shim::RankSet& R = *m_ranks;
        std::vector<double> dt(R.n(), 0.0);
        std::vector<int> rcs(R.n(), 0);
        R.each([&](int r) {
            rcs[r] = pfem_wc_next_dt(R.ctx[r], &m_par, m_securityCoeff, maxDT, &dt[r]);
            pfem_b200_shim::check(R.ctx[r], rcs[r], "pfem_wc_next_dt");
        });
        if (rcs[0] == PFEM_NAN) throw std::runtime_error("NaN time step!");
        return dt[0];  // the minimum is all-reduced: every rank holds the same value
    }
    // device -> host mesh (before remeshing or extractor output): states and positions
    void download() {
        const std::size_t nN = m_pMesh->getNodesCount();
        std::vector<double> q, x, xOld(dim * nN);
        pfem_b200_shim::downloadAll(*m_ranks, nN, dim, 0, 2 * dim + 2, &q, &x);
        q = m_rn.toOld(q, nN);
        x = m_rn.toOld(x, nN);
        for (std::size_t n = 0; n < nN; ++n) {
            for (unsigned int s = 0; s < 2u * dim + 2u; ++s) m_pMesh->setNodeState(n, s, q[n + s * nN]);
            for (unsigned short d = 0; d < dim; ++d) xOld[n + d * nN] = x[n + d * nN] - m_pMesh->getNode(n).getCoordinate(d);
        }
        m_pMesh->updateNodesPosition(xOld);  // delta to the device positions (fixed nodes have delta 0)
        if (m_isThermal) {
            pfem_b200_shim::RankSet& R = *m_ranks;
            std::vector<double> T(nN, 0.0);
            R.each([&](int r) {
                const std::size_t nl = R.multi() ? R.l2g[r].size() : nN;
                std::vector<double> Tl(nl);
                pfem_b200_shim::check(R.ctx[r], pfem_get_temperature(R.ctx[r], Tl.data()), "pfem_get_temperature");
                if (R.multi()) R.gatherOwned(r, Tl, T, nN);
                else T = Tl;
            });
            for (std::size_t nOld = 0; nOld < nN; ++nOld) m_pMesh->setNodeState(nOld, 2 * dim + 2, T[m_rn.nodeNew(nOld)]);
        }
    }

private:
    Problem* m_pProblem;
    Solver* m_pSolver;
    Mesh* m_pMesh;
    SolTable m_bc;
    double m_securityCoeff;
    std::unique_ptr<pfem_b200_shim::RankSet> m_ranks;
    pfem_wc_params m_par{};
    pfem_b200_shim::Renumbering m_rn;
    double m_gamma = 0.0;
    bool m_dirty = true;
    bool m_isThermal = false;              // problem id "BoussinesqWC"
    pfem_thermal_params m_thermal{};
    std::unique_ptr<SolTable> m_heatBc;
    std::vector<uint8_t> m_tmask;
    std::vector<double> m_tval;
    bool m_movingBcNodes = false;          // a bound node with a velocity BC that is not fixed: its BC sees the device position
    std::vector<uint8_t> m_dmask;          // Dirichlet table last uploaded outside a mesh upload
    std::vector<double> m_dval;
};
