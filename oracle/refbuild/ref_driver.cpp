// C entry points that drive the REFERENCE'S OWN hot-path code (compiled in place from /root/reference against the
// stand-in Eigen / sol2 / gmsh headers of this directory).  TEST INFRASTRUCTURE ONLY: used by tests/ and by
// tests/golden/make_golden.py to pin oracle/pfem_oracle.cpp; never linked into or called by the product library.
//
// What runs here is reference code, unmodified:
//   Mesh(MeshCreateInfo) -> loadFromFile (Mesh.cpp:762-917)            [gmsh calls answered by standin/gmsh.h]
//   SolverIncompNewton / SolverWCompNewton constructors (IN/Solver.cpp:12-209, WC/Solver.cpp:15-160), which build
//   MomContEqIncompNewton<dim> / ContEqWCompNewton<dim> / MomEqWCompNewton<dim> through REGISTER_EQ and set the BC tag
//   flags through Solver::checkBC (Solver.cpp:104-135);
//   MomContEqIncompNewton::m_buildAbPSPG / m_applyBCPSPG / m_computeTauPSPG (PSPG.inl:7-259), ::solve -> PicardAlgo
//   (MomContEquation.inl:274-300, PicardAlgo.cpp:31-94, PSPG.inl:262-373);
//   SolverWCompNewton::m_solveWCompNewtonNoT and computeNextDT (WC/Solver.cpp:192-276);
//   MatrixBuilder<dim> (MatricesBuilder.inl), Element/Facet geometry (Element.cpp, Facet.cpp), Mesh tables.
// What is NOT reference code: the dense/sparse arithmetic library underneath (standin/Eigen), the parameter tables
// (standin/sol, filled below instead of a Lua file), the connectivity loader replacing CGAL (host_stubs.cpp).
// Private members are reached with g++ -fno-access-control; no reference source is edited or copied.
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include <omp.h>

#include "ref_inject.hpp"

#include "mesh/Mesh.hpp"
#include "simulation/Problem.hpp"
#include "simulation/Solver.hpp"
#include "simulation/physics/IncompNewton/Solver.hpp"
#include "simulation/physics/IncompNewton/MomContEquation.hpp"
#include "simulation/physics/IncompNewton/HeatEquation.hpp"
#include "simulation/physics/WCompNewton/Solver.hpp"
#include "simulation/physics/WCompNewton/ContEquation.hpp"
#include "simulation/physics/WCompNewton/MomEquation.hpp"
#include "simulation/utility/StatesFromToQ.hpp"

#ifdef PFEM_REF_WITH_B200
// Drop-in build (libpfem_ref_dropin.so): the reference's host code above drives libpfem_b200.so through the shim that a
// PFEM3D maintainer would add (shim/pfem_b200_equations.hpp).
#include "pfem_b200_equations.hpp"
#endif

namespace {

struct PosKey {
    std::uint64_t a, b, c;
    bool operator==(const PosKey& o) const { return a == o.a && b == o.b && c == o.c; }
};
struct PosHash {
    std::size_t operator()(const PosKey& k) const { return std::hash<std::uint64_t>()(k.a * 0x9E3779B97F4A7C15ull ^ (k.b << 1) ^ (k.c * 31)); }
};
PosKey keyOf(const std::array<double, 3>& p) {
    PosKey k;
    std::memcpy(&k.a, &p[0], 8);
    std::memcpy(&k.b, &p[1], 8);
    std::memcpy(&k.c, &p[2], 8);
    return k;
}

struct RefCase {
    int dim = 0;
    std::size_t N = 0, E = 0;
    std::string problemId, solverId;
    std::vector<std::int64_t> conn, facets;
    std::vector<double> x;
    std::vector<std::uint8_t> flags, dirMask;
    std::vector<std::int32_t> tags;
    std::vector<double> dirVal;
    double bcRamp = 0.0;              // g_D(pos, t) = dirVal (1 + bcRamp t)
    std::vector<std::uint8_t> tMask;  // BoussinesqWC: node has a "<type>T" boundary condition
    std::vector<double> tVal;
    sol::table root;
    std::unique_ptr<Problem> problem;
    Mesh* mesh = nullptr;
    Solver* solver = nullptr;
    std::unordered_map<PosKey, std::size_t, PosHash> byPos;
    std::string error;
    std::shared_ptr<void> wcShim;  // WCompNewtonStepB200<dim> of the drop-in build

    void rebuildPositions() {
        byPos.clear();
        for (std::size_t n = 0; n < mesh->getNodesCount(); ++n) byPos[keyOf(mesh->getNode(n).getPosition())] = n;
    }
    std::mutex posMutex;
    std::size_t nodeAt(const std::array<double, 3>& pos) {
        std::lock_guard<std::mutex> lock(posMutex);  // MomEqWCompNewton::m_applyBC calls the BC table from an omp loop
        auto it = byPos.find(keyOf(pos));
        if (it == byPos.end()) {
            rebuildPositions();
            it = byPos.find(keyOf(pos));
            if (it == byPos.end()) throw std::runtime_error("refbuild: BC callback position matches no node");
        }
        return it->second;
    }
};

template <typename M> void copyOut(const M& m, double* dst) {  // row-major out
    for (Eigen::Index i = 0; i < m.rows(); ++i)
        for (Eigen::Index j = 0; j < m.cols(); ++j) dst[i * m.cols() + j] = m(i, j);
}

template <unsigned short dim> int elementsPSPG(RefCase& rc, const double* qPrevPtr, double* Ae_out, double* be_out, double* tau_out) {
    auto* eq = dynamic_cast<MomContEqIncompNewton<dim>*>(rc.solver->m_pEquations[0].get());
    if (!eq) return -1;
    constexpr unsigned short npe = dim + 1;
    constexpr int nt = (dim + 1) * npe;
    const std::size_t nNodes = rc.N;
    Eigen::VectorXd qPrev((dim + 1) * nNodes);
    for (std::size_t i = 0; i < (dim + 1) * nNodes; ++i) qPrev[i] = qPrevPtr[i];
    const double dt = rc.solver->getTimeStep();
    for (std::size_t elm = 0; elm < rc.E; ++elm) {
        // same calls, in the same order, as the body of the element loop PSPG.inl:26-53 (that loop keeps Ae/be local,
        // so the per-element view has to repeat its composition; the assembled A and b below come from the loop itself)
        const Element& element = rc.mesh->getElement(elm);
        Eigen::Matrix<double, nt, nt> Ae;
        Eigen::Matrix<double, nt, 1> be;
        double tau = eq->m_computeTauPSPG(element);
        GradNmatType<dim> gradNe = eq->m_pMatBuilder->getGradN(element);
        BmatType<dim> Be = eq->m_pMatBuilder->getB(gradNe);
        Eigen::Matrix<double, npe, npe> Me_dt_s = (1 / dt) * eq->m_pMatBuilder->getM(element);
        Eigen::Matrix<double, dim * npe, dim * npe> Me_dt = MatrixBuilder<dim>::diagBlock(Me_dt_s);
        Eigen::Matrix<double, dim * npe, dim * npe> Ke = eq->m_pMatBuilder->getK(element, Be);
        Eigen::Matrix<double, npe, dim * npe> De = eq->m_pMatBuilder->getD(element, Be);
        Eigen::Matrix<double, npe, dim * npe> Ce_dt = (tau / dt) * eq->m_pMatBuilder->getC(element, Be, gradNe);
        Eigen::Matrix<double, npe, npe> Le = tau * eq->m_pMatBuilder->getL(element, Be, gradNe);
        Eigen::Matrix<double, dim * npe, 1> Fe = eq->m_pMatBuilder->getF(element, eq->m_bodyForce, Be);
        Eigen::Matrix<double, npe, 1> He = tau * eq->m_pMatBuilder->getH(element, eq->m_bodyForce, Be, gradNe);
        Ae << Me_dt + Ke, -De.transpose(), Ce_dt + De, Le;
        Eigen::Matrix<double, dim * npe, 1> vPrev = getElementVecState<dim>(qPrev, element, 0, nNodes);
        be << Fe + Me_dt * vPrev, He + Ce_dt * vPrev;
        copyOut(Ae, Ae_out + elm * nt * nt);
        copyOut(be, be_out + elm * nt);
        tau_out[elm] = tau;
    }
    return 0;
}

template <unsigned short dim> int buildPSPG(RefCase& rc, const double* qPrevPtr, int applyBC) {
    auto* eq = dynamic_cast<MomContEqIncompNewton<dim>*>(rc.solver->m_pEquations[0].get());
    if (!eq) return -1;
    const std::size_t n = (dim + 1) * rc.N;
    Eigen::VectorXd qPrev(n);
    for (std::size_t i = 0; i < n; ++i) qPrev[i] = qPrevPtr[i];
    eq->m_A.resize(n, n);  // PSPG.inl:266-267
    eq->m_b.resize(n);
    eq->m_b.setZero();
    eq->m_buildAbPSPG(qPrev);
    if (applyBC) eq->m_applyBCPSPG(qPrev);
    return 0;
}

template <unsigned short dim> const Eigen::SparseMatrix<double>* matrixOf(RefCase& rc, const Eigen::VectorXd** b) {
    auto* eq = dynamic_cast<MomContEqIncompNewton<dim>*>(rc.solver->m_pEquations[0].get());
    if (!eq) return nullptr;
    *b = &eq->m_b;
    return &eq->m_A;
}

template <unsigned short dim> const Eigen::SparseMatrix<double>* heatBuildIN(RefCase& rc, const double* thetaPrev, int applyBC, const Eigen::VectorXd** b) {
    auto* eq = dynamic_cast<HeatEqIncompNewton<dim>*>(rc.solver->m_pEquations.back().get());
    if (!eq) return nullptr;
    Eigen::VectorXd q(rc.N);
    for (std::size_t i = 0; i < rc.N; ++i) q[i] = thetaPrev[i];
    eq->m_A.resize(rc.N, rc.N);  // HeatEquation.inl:113-115
    eq->m_b.resize(rc.N);
    eq->m_b.setZero();
    eq->m_buildAb(q);
    if (applyBC) eq->m_applyBC(q);
    *b = &eq->m_b;
    return &eq->m_A;
}

// FracStep sub-systems of MomContEquationFracStep.inl on the current mesh, run one at a time with given inputs:
//   which 0: m_buildMatFracStep({a = v_prev (dim N), b = p_prev (N)}) + m_applyBCVAppStep(v_prev)  -> m_MK_dt, m_bVAppStep
//   which 1: m_buildMatPcorrStep(a = vTilde, b = p_prev) + m_applyBCPCorrStep()                      -> m_L, m_bPcorrStep
//            (uses m_DTelm / m_Lelm / m_L of the last which-0 build, as the Picard body does: :487-507)
//   which 2: m_buildMatVStep(a = deltaP) + m_applyBCVStep()                                         -> m_M, m_bVStep
template <unsigned short dim> struct FsView {
    const Eigen::SparseMatrix<double>* A = nullptr;
    const Eigen::VectorXd* b = nullptr;
};
template <unsigned short dim> FsView<dim> fsSystem(RefCase& rc, int which) {
    FsView<dim> v;
    auto* eq = dynamic_cast<MomContEqIncompNewton<dim>*>(rc.solver->m_pEquations[0].get());
    if (!eq) return v;
    if (which == 0) v.A = &eq->m_MK_dt, v.b = &eq->m_bVAppStep;
    else if (which == 1) v.A = &eq->m_L, v.b = &eq->m_bPcorrStep;
    else v.A = &eq->m_M, v.b = &eq->m_bVStep;
    return v;
}
template <unsigned short dim> std::int64_t fsBuild(RefCase& rc, int which, const double* a, const double* b) {
    auto* eq = dynamic_cast<MomContEqIncompNewton<dim>*>(rc.solver->m_pEquations[0].get());
    if (!eq) return -1;
    const Eigen::Index nV = static_cast<Eigen::Index>(dim * rc.N), nP = static_cast<Eigen::Index>(rc.N);
    auto vec = [](const double* src, Eigen::Index n) {
        Eigen::VectorXd q(n);
        for (Eigen::Index i = 0; i < n; ++i) q[i] = src[i];
        return q;
    };
    if (which == 0) {  // the "prepare" lambda of m_setupPicardFracStep (:454-462), then the head of its body (:468-470)
        eq->m_M.resize(nV, nV);
        eq->m_MK_dt.resize(nV, nV);
        eq->m_L.resize(nP, nP);
        eq->m_bVAppStep.resize(nV); eq->m_bVAppStep.setZero();
        eq->m_bPcorrStep.resize(nP); eq->m_bPcorrStep.setZero();
        eq->m_bVStep.resize(nV); eq->m_bVStep.setZero();
        std::vector<Eigen::VectorXd> qPrev = {vec(a, nV), vec(b, nP)};
        eq->m_buildMatFracStep(qPrev);
        eq->m_applyBCVAppStep(qPrev[0]);
    } else if (which == 1) {
        eq->m_buildMatPcorrStep(vec(a, nV), vec(b, nP));
        eq->m_applyBCPCorrStep();
    } else {
        eq->m_buildMatVStep(vec(a, nP));
        eq->m_applyBCVStep();
    }
    return fsSystem<dim>(rc, which).A->nonZeros();
}
template <unsigned short dim> int fsSolve(RefCase& rc, int which, double* x, double* itersErrInfo) {
    auto* eq = dynamic_cast<MomContEqIncompNewton<dim>*>(rc.solver->m_pEquations[0].get());
    if (!eq) return -1;
    const FsView<dim> v = fsSystem<dim>(rc, which);
    eq->m_solverIt.compute(*v.A);                       // :472, :491, :516
    const Eigen::VectorXd sol = eq->m_solverIt.solve(*v.b);
    for (Eigen::Index i = 0; i < sol.rows(); ++i) x[i] = sol[i];
    itersErrInfo[0] = static_cast<double>(eq->m_solverIt.iterations());
    itersErrInfo[1] = static_cast<double>(eq->m_solverIt.error());
    itersErrInfo[2] = static_cast<double>(eq->m_solverIt.info());
    return 0;
}

template <unsigned short dim> int elementMatrices(RefCase& rc, double* M, double* K, double* D, double* L, double* C, double* F, double* H) {
    auto* eq = dynamic_cast<MomContEqIncompNewton<dim>*>(rc.solver->m_pEquations[0].get());
    if (!eq) return -1;
    constexpr unsigned short npe = dim + 1;
    for (std::size_t elm = 0; elm < rc.E; ++elm) {
        const Element& element = rc.mesh->getElement(elm);
        GradNmatType<dim> gradNe = eq->m_pMatBuilder->getGradN(element);
        BmatType<dim> Be = eq->m_pMatBuilder->getB(gradNe);
        copyOut(eq->m_pMatBuilder->getM(element), M + elm * npe * npe);
        copyOut(eq->m_pMatBuilder->getK(element, Be), K + elm * dim * npe * dim * npe);
        copyOut(eq->m_pMatBuilder->getD(element, Be), D + elm * npe * dim * npe);
        copyOut(eq->m_pMatBuilder->getL(element, Be, gradNe), L + elm * npe * npe);
        copyOut(eq->m_pMatBuilder->getC(element, Be, gradNe), C + elm * npe * dim * npe);
        copyOut(eq->m_pMatBuilder->getF(element, eq->m_bodyForce, Be), F + elm * dim * npe);
        copyOut(eq->m_pMatBuilder->getH(element, eq->m_bodyForce, Be, gradNe), H + elm * npe);
    }
    return 0;
}

thread_local std::string g_lastError;
int g_threads = 1;
// temperature Dirichlet data of the next BoussinesqWC case (pfem_ref_set_thermal_bc)
std::vector<std::uint8_t> g_tMask;
std::vector<double> g_tVal;  // Problem::m_nThreads of the next case = number of "Lua states" = OpenMP threads (Problem.cpp:30-45)

}  // namespace

extern "C" {

const char* pfem_ref_last_error() { return g_lastError.c_str(); }
// BoussinesqWC: per-node temperature Dirichlet mask / values consumed by the next pfem_ref_create (n = 0 clears)
void pfem_ref_set_thermal_bc(std::int64_t n, const std::uint8_t* tMask, const double* tVal) {
    g_tMask.assign(tMask, tMask + (n > 0 ? n : 0));
    g_tVal.assign(tVal, tVal + (n > 0 ? n : 0));
}
void pfem_ref_set_threads(int n) { g_threads = n > 0 ? n : omp_get_num_procs(); }
// time dependence of the velocity Dirichlet table of the cases created from now on: g_D(pos, t) = dirVal (1 + ramp t)
// (the "<type>V"(pos, t) signature of the reference's Lua BC functions, e.g. examples/2D/cylinder/cylinderComp.lua)
double g_bcRamp = 0.0;
void pfem_ref_set_bc_ramp(double ramp) { g_bcRamp = ramp; }
int pfem_ref_get_threads() { return g_threads; }

// Direct solver used by the stand-in Eigen::SparseLU (nullptr -> built-in dense LU).
void pfem_ref_set_direct_solver(Eigen::standin::DirectSolverFn fn) { Eigen::standin::directSolverHook() = fn; }
long pfem_ref_direct_solves() { return Eigen::standin::directSolveCount(); }
// log of the stand-in ConjugateGradient solves since the last call: rows of (n, iterations, error, info)
long pfem_ref_cg_log(double* out, long maxRows) {
    auto& log = Eigen::standin::cgLog();
    long k = 0;
    for (; k < static_cast<long>(log.size()) && k < maxRows; ++k) {
        out[4 * k] = log[k].n; out[4 * k + 1] = log[k].iterations; out[4 * k + 2] = log[k].error; out[4 * k + 3] = log[k].info;
    }
    log.clear();
    return k;
}

// params: IncompNewtonNoT|Bingham / PSPG|FracStep -> [rho, mu, dt, bx, by, bz, gamma, maxIter, minRes, gammaFS,
//                                                       residual (0 Ax_f, 1 U, 2 U_P), tau0, mReg]   (the last two for Bingham)
//         WCompNewtonNoT|BoussinesqWC/CDS_* -> [mu, K0, K0p, rhoStar, bx, by, bz, meduri, gamma, initialDT, maxDT, securityCoeff,
//                                               k, cv, alpha, Tr]   (the last four for BoussinesqWC only)
// facets: nFacets x (dim+2) = facet nodes, out node, element index (may be null / 0)
void* pfem_ref_create(int dim, std::int64_t nNodes, std::int64_t nElems, const std::int64_t* conn, const double* x,
                      const std::uint8_t* flags, const std::uint8_t* dirMask, const double* dirVal,
                      std::int64_t nFacets, const std::int64_t* facets, const char* problemId, const char* solverId,
                      const double* p) {
    try {
        omp_set_num_threads(g_threads);  // one parameter table ("Lua state") per thread, as in Problem.cpp:30-45
        auto rc = std::make_unique<RefCase>();
        rc->dim = dim;
        rc->N = static_cast<std::size_t>(nNodes);
        rc->E = static_cast<std::size_t>(nElems);
        rc->problemId = problemId;
        rc->solverId = solverId;
        rc->conn.assign(conn, conn + nElems * (dim + 1));
        rc->x.assign(x, x + dim * nNodes);
        rc->flags.assign(flags, flags + nNodes);
        rc->dirMask.assign(dirMask, dirMask + nNodes);
        rc->dirVal.assign(dirVal, dirVal + dim * nNodes);
        rc->bcRamp = g_bcRamp;
        if (nFacets > 0) rc->facets.assign(facets, facets + nFacets * (dim + 2));
        rc->tags.resize(rc->N);
        const bool boussinesqWC = rc->problemId == "BoussinesqWC";
        const bool boussinesqIN = rc->problemId == "Boussinesq";
        if ((boussinesqWC || boussinesqIN) && g_tMask.size() == rc->N) {
            rc->tMask = g_tMask;
            rc->tVal = g_tVal;
        } else {
            rc->tMask.assign(rc->N, 0);
            rc->tVal.assign(rc->N, 0.0);
        }
        // tags of bound nodes by the boundary conditions they carry: Dir (velocity), BT (temperature), BVT (both), Wall (none)
        for (std::size_t n = 0; n < rc->N; ++n) {
            const bool vbc = dirMask[n] != 0, tbc = rc->tMask[n] != 0;
            rc->tags[n] = (flags[n] & 1) ? (vbc ? (tbc ? 4 : 1) : (tbc ? 3 : 2)) : 0;
        }

        const bool wc = rc->problemId == "WCompNewtonNoT" || boussinesqWC;
        refinject::MeshArrays& in = refinject::current();
        in = refinject::MeshArrays();
        in.dim = dim;
        in.nNodes = rc->N;
        in.nElems = rc->E;
        in.nFacets = static_cast<std::size_t>(nFacets > 0 ? nFacets : 0);
        in.nStates = boussinesqWC ? 2 * dim + 3 : (wc ? 2 * dim + 2 : (boussinesqIN ? dim + 2 : dim + 1));  // WC/Problem.cpp:17-18, 130-131; IN/Problem.cpp:13-14
        in.conn = rc->conn.data();
        in.x = rc->x.data();
        in.flags = rc->flags.data();
        in.tags = rc->tags.data();
        in.tagNames = {"Fluid", "Dir", "Wall", "BT", "BVT"};
        in.facets = rc->facets.empty() ? nullptr : rc->facets.data();

        rc->problem.reset(new Problem(rc->problemId));
        rc->problem->m_nThreads = static_cast<unsigned int>(g_threads);
        MeshCreateInfo info;
        info.hchar = 1;
        info.boundingBox.assign(2 * dim, 0.0);
        info.mshFile = "/dev/null";
        rc->problem->m_pMesh = std::make_unique<Mesh>(info);  // -> loadFromFile -> triangulateAlphaShape (injected)
        rc->mesh = rc->problem->m_pMesh.get();

        // the "Lua" tables (examples/3D/damBreakKoshizuka/*.lua layout: Problem.Solver.<EqID>.BC, Problem.Material)
        RefCase* self = rc.get();
        sol::function dirV = [self](const sol::Args& a) -> std::any {
            const auto pos = std::any_cast<std::array<double, 3>>(a.at(1));
            const std::size_t n = self->nodeAt(pos);
            double t = 0.0;
            if (a.size() > 2) {
                if (const double* pt = std::any_cast<double>(&a.at(2))) t = *pt;
                else if (self->bcRamp != 0.0) throw std::runtime_error(std::string("BC time argument of type ") + a.at(2).type().name());
            }
            const double f = 1.0 + self->bcRamp * t;
            std::vector<double> g(self->dim);
            for (int d = 0; d < self->dim; ++d) g[d] = self->dirVal[n + d * self->N] * f;
            return g;
        };
        sol::function dirT = [self](const sol::Args& a) -> std::any {
            const auto pos = std::any_cast<std::array<double, 3>>(a.at(1));
            return std::vector<double>{self->tVal[self->nodeAt(pos)]};
        };
        sol::table material, solverT, bc;
        bc.set_function("DirV", dirV);
        bc.set_function("BVTV", dirV);
        if (!wc) {
            material.set("rho", p[0]);
            material.set("mu", p[1]);
            material.set("gamma", p[6]);
            if (rc->problemId == "Bingham") {
                material.set("tau0", p[11]);
                material.set("mReg", p[12]);
            }
            if (boussinesqIN) {  // examples/2D/thermalConv/thermalConvIncomp.lua layout; params [11..14] = alpha, Tr, k, cv
                material.set("alpha", p[11]);
                material.set("Tr", p[12]);
                material.set("k", p[13]);
                material.set("cv", p[14]);
                material.set("DgammaDT", 0.0);
                material.set("h", 0.0);
                material.set("Tinf", 0.0);
                material.set("epsRad", 0.0);
                bc.set_function("BVTV", dirV);
                sol::table heatT, heatBC;
                heatBC.set_function("BTT", dirT);
                heatBC.set_function("BVTT", dirT);
                heatT.set("maxIter", p[7]);
                heatT.set("minRes", p[8]);
                heatT.set("residual", std::string("Ax_f"));
                heatT.set("BC", heatBC);
                solverT.set("HeatEq", heatT);
                solverT.set("solveHeatFirst", true);
            }
            sol::table eqT;
            eqT.set("maxIter", p[7]);
            eqT.set("minRes", p[8]);
            eqT.set("residual", std::string(p[10] == 1 ? "U" : p[10] == 2 ? "U_P" : "Ax_f"));
            eqT.set("gammaFS", p[9]);
            eqT.set("bodyForce", std::vector<double>(p + 3, p + 3 + dim));
            eqT.set("BC", bc);
            solverT.set("id", rc->solverId);
            solverT.set("adaptDT", false);
            solverT.set("maxDT", p[2]);
            solverT.set("initialDT", p[2]);
            solverT.set("coeffDTDecrease", 2.0);
            solverT.set("coeffDTincrease", 1.0);
            solverT.set("MomContEq", eqT);
        } else {
            material.set("mu", p[0]);
            material.set("K0", p[1]);
            material.set("K0p", p[2]);
            material.set("rhoStar", p[3]);
            material.set("gamma", p[8]);
            sol::table contT, momT, emptyBC;
            contT.set("stabilization", std::string(p[7] != 0 ? "Meduri" : "None"));
            contT.set("BC", emptyBC);
            momT.set("bodyForce", std::vector<double>(p + 4, p + 4 + dim));
            momT.set("BC", bc);
            solverT.set("id", rc->solverId);
            solverT.set("adaptDT", true);
            solverT.set("maxDT", p[10]);
            solverT.set("initialDT", p[9]);
            solverT.set("securityCoeff", p[11]);
            solverT.set("ContEq", contT);
            solverT.set("MomEq", momT);
            if (boussinesqWC) {  // examples/2D/thermalConv/thermalConvComp.lua layout
                material.set("k", p[12]);
                material.set("cv", p[13]);
                material.set("alpha", p[14]);
                material.set("Tr", p[15]);
                material.set("DgammaDT", 0.0);
                material.set("h", 0.0);
                material.set("Tinf", 0.0);
                material.set("epsRad", 0.0);
                sol::table heatT, heatBC;
                heatBC.set_function("BTT", dirT);
                heatBC.set_function("BVTT", dirT);
                heatT.set("BC", heatBC);
                solverT.set("HeatEq", heatT);
            }
        }
        rc->root.set("Solver", solverT);
        rc->root.set("Material", material);
        rc->problem->m_problemParams.assign(static_cast<std::size_t>(g_threads), SolTable(rc->root));
        rc->problem->m_statesNumber = in.nStates;

        if (!wc)
            rc->problem->m_pSolver = std::make_unique<SolverIncompNewton>(rc->problem.get(), rc->mesh, rc->problem->m_problemParams);
        else
            rc->problem->m_pSolver = std::make_unique<SolverWCompNewton>(rc->problem.get(), rc->mesh, rc->problem->m_problemParams);
        rc->solver = rc->problem->m_pSolver.get();
        rc->solver->m_nextTimeToRemesh = std::numeric_limits<double>::max();  // left uninitialised by Solver.cpp:12-66; no remesh here
        rc->solver->m_maxRemeshDT = std::numeric_limits<double>::max();
        // the solver constructors switch the facet normals off when gamma == 0 (IN/Solver.cpp:196-205); recompute them if needed
        return rc.release();
    } catch (const std::exception& e) {
        g_lastError = e.what();
        return nullptr;
    }
}

void pfem_ref_destroy(void* h) { delete static_cast<RefCase*>(h); }

// all node states, layout q[n + s*nNodes] (StatesFromToQ.hpp:9-34)
int pfem_ref_set_states(void* h, const double* q, int s0, int ns) {
    auto& rc = *static_cast<RefCase*>(h);
    for (std::size_t n = 0; n < rc.N; ++n)
        for (int s = 0; s < ns; ++s) rc.mesh->setNodeState(n, s0 + s, q[n + s * rc.N]);
    return 0;
}
int pfem_ref_get_states(void* h, double* q, int s0, int ns) {
    auto& rc = *static_cast<RefCase*>(h);
    for (std::size_t n = 0; n < rc.N; ++n)
        for (int s = 0; s < ns; ++s) q[n + s * rc.N] = rc.mesh->getNode(n).getState(s0 + s);
    return 0;
}
int pfem_ref_get_positions(void* h, double* x) {
    auto& rc = *static_cast<RefCase*>(h);
    for (std::size_t n = 0; n < rc.N; ++n)
        for (int d = 0; d < rc.dim; ++d) x[n + d * rc.N] = rc.mesh->getNode(n).getCoordinate(d);
    return 0;
}
int pfem_ref_set_dirichlet_values(void* h, const double* dirVal) {
    auto& rc = *static_cast<RefCase*>(h);
    rc.dirVal.assign(dirVal, dirVal + rc.dim * rc.N);
    return 0;
}
int pfem_ref_set_time_step(void* h, double dt) {
    static_cast<RefCase*>(h)->solver->m_timeStep = dt;
    return 0;
}

// Element.cpp:15-135 (J, detJ, invJ as stored by the reference, row-major 3x3 each) and Element::getRin (:226-294)
int pfem_ref_element_geometry(void* h, double* detJ, double* J, double* invJ, double* rin) {
    auto& rc = *static_cast<RefCase*>(h);
    for (std::size_t e = 0; e < rc.E; ++e) {
        const Element& el = rc.mesh->getElement(e);
        detJ[e] = el.getDetJ();
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                const bool in = i < rc.dim && j < rc.dim;
                J[e * 9 + i * 3 + j] = in ? el.getJ(i, j) : 0.0;
                invJ[e * 9 + i * 3 + j] = in ? el.getInvJ(i, j) : 0.0;
            }
        rin[e] = el.getRin();
    }
    return 0;
}

// Mesh::getGaussPoints/getGaussWeight/getShapeFunctions/getRefElementSize (Mesh.cpp:342-530)
int pfem_ref_tables(void* h, int dimension, int nGP, double* gp, double* w, double* sf, double* refSize) {
    auto& rc = *static_cast<RefCase*>(h);
    try {
        auto gps = rc.mesh->getGaussPoints(dimension, nGP);
        auto ws = rc.mesh->getGaussWeight(dimension, nGP);
        auto sfs = rc.mesh->getShapeFunctions(dimension, nGP);
        for (int g = 0; g < nGP; ++g) {
            for (int k = 0; k < 3; ++k) gp[g * 3 + k] = gps[g][k];
            w[g] = ws[g];
            for (int k = 0; k <= dimension; ++k) sf[g * (dimension + 1) + k] = sfs[g][k];
        }
        *refSize = rc.mesh->getRefElementSize(dimension);
        return 0;
    } catch (const std::exception& e) {
        g_lastError = e.what();
        return -1;
    }
}

// MatrixBuilder getM/getK/getD/getL/getC/getF/getH with the factors set by the real MomContEqIncompNewton constructor
int pfem_ref_element_matrices(void* h, double* M, double* K, double* D, double* L, double* C, double* F, double* H) {
    auto& rc = *static_cast<RefCase*>(h);
    try {
        return rc.dim == 2 ? elementMatrices<2>(rc, M, K, D, L, C, F, H) : elementMatrices<3>(rc, M, K, D, L, C, F, H);
    } catch (const std::exception& e) {
        g_lastError = e.what();
        return -1;
    }
}

int pfem_ref_pspg_elements(void* h, const double* qPrev, double* Ae, double* be, double* tau) {
    auto& rc = *static_cast<RefCase*>(h);
    try {
        return rc.dim == 2 ? elementsPSPG<2>(rc, qPrev, Ae, be, tau) : elementsPSPG<3>(rc, qPrev, Ae, be, tau);
    } catch (const std::exception& e) {
        g_lastError = e.what();
        return -1;
    }
}

// m_buildAbPSPG (+ m_applyBCPSPG) on the current mesh / node states; returns nnz of m_A or < 0
std::int64_t pfem_ref_pspg_build(void* h, const double* qPrev, int applyBC) {
    auto& rc = *static_cast<RefCase*>(h);
    try {
        rc.rebuildPositions();
        const int r = rc.dim == 2 ? buildPSPG<2>(rc, qPrev, applyBC) : buildPSPG<3>(rc, qPrev, applyBC);
        if (r) return r;
        const Eigen::VectorXd* b;
        const auto* A = rc.dim == 2 ? matrixOf<2>(rc, &b) : matrixOf<3>(rc, &b);
        return A->nonZeros();
    } catch (const std::exception& e) {
        g_lastError = e.what();
        return -1;
    }
}
int pfem_ref_csc_copy(void* h, std::int64_t* colPtr, std::int32_t* rowIdx, double* val, double* bOut) {
    auto& rc = *static_cast<RefCase*>(h);
    const Eigen::VectorXd* b;
    const auto* A = rc.dim == 2 ? matrixOf<2>(rc, &b) : matrixOf<3>(rc, &b);
    if (!A) return -1;
    for (Eigen::Index j = 0; j <= A->cols(); ++j) colPtr[j] = A->outerIndexPtr()[j];
    for (Eigen::Index k = 0; k < A->nonZeros(); ++k) {
        rowIdx[k] = A->innerIndexPtr()[k];
        val[k] = A->valuePtr()[k];
    }
    for (Eigen::Index i = 0; i < b->rows(); ++i) bOut[i] = (*b)[i];
    return 0;
}

// HeatEqIncompNewton::m_buildAb (+ m_applyBC) on the current mesh; returns nnz or < 0; copy out with pfem_ref_in_heat_copy
std::int64_t pfem_ref_in_heat_build(void* h, const double* thetaPrev, int applyBC) {
    auto& rc = *static_cast<RefCase*>(h);
    try {
        rc.rebuildPositions();
        const Eigen::VectorXd* b;
        const auto* A = rc.dim == 2 ? heatBuildIN<2>(rc, thetaPrev, applyBC, &b) : heatBuildIN<3>(rc, thetaPrev, applyBC, &b);
        return A ? A->nonZeros() : -1;
    } catch (const std::exception& e) {
        g_lastError = e.what();
        return -1;
    }
}
int pfem_ref_in_heat_copy(void* h, std::int64_t* colPtr, std::int32_t* rowIdx, double* val, double* bOut) {
    auto& rc = *static_cast<RefCase*>(h);
    const Eigen::SparseMatrix<double>* A;
    const Eigen::VectorXd* b;
    if (rc.dim == 2) {
        auto* eq = dynamic_cast<HeatEqIncompNewton<2>*>(rc.solver->m_pEquations.back().get());
        if (!eq) return -1;
        A = &eq->m_A, b = &eq->m_b;
    } else {
        auto* eq = dynamic_cast<HeatEqIncompNewton<3>*>(rc.solver->m_pEquations.back().get());
        if (!eq) return -1;
        A = &eq->m_A, b = &eq->m_b;
    }
    for (Eigen::Index j = 0; j <= A->cols(); ++j) colPtr[j] = A->outerIndexPtr()[j];
    for (Eigen::Index k = 0; k < A->nonZeros(); ++k) {
        rowIdx[k] = A->innerIndexPtr()[k];
        val[k] = A->valuePtr()[k];
    }
    for (Eigen::Index i = 0; i < b->rows(); ++i) bOut[i] = (*b)[i];
    return 0;
}

// FracStep sub-systems (see fsBuild): returns nnz or < 0; copy out with pfem_ref_fs_copy; solve with the equation's own
// ConjugateGradient object (m_solverIt) through pfem_ref_fs_solve: out = (iterations, error, info)
std::int64_t pfem_ref_fs_build(void* h, int which, const double* a, const double* b) {
    auto& rc = *static_cast<RefCase*>(h);
    try {
        if (which == 0) rc.rebuildPositions();
        return rc.dim == 2 ? fsBuild<2>(rc, which, a, b) : fsBuild<3>(rc, which, a, b);
    } catch (const std::exception& e) {
        g_lastError = e.what();
        return -1;
    }
}
int pfem_ref_fs_copy(void* h, int which, std::int64_t* colPtr, std::int32_t* rowIdx, double* val, double* bOut) {
    auto& rc = *static_cast<RefCase*>(h);
    const Eigen::SparseMatrix<double>* A;
    const Eigen::VectorXd* b;
    if (rc.dim == 2) {
        const auto v = fsSystem<2>(rc, which);
        A = v.A, b = v.b;
    } else {
        const auto v = fsSystem<3>(rc, which);
        A = v.A, b = v.b;
    }
    if (!A) return -1;
    for (Eigen::Index j = 0; j <= A->cols(); ++j) colPtr[j] = A->outerIndexPtr()[j];
    for (Eigen::Index k = 0; k < A->nonZeros(); ++k) {
        rowIdx[k] = A->innerIndexPtr()[k];
        val[k] = A->valuePtr()[k];
    }
    for (Eigen::Index i = 0; i < b->rows(); ++i) bOut[i] = (*b)[i];
    return 0;
}
int pfem_ref_fs_solve(void* h, int which, double* x, double* itersErrInfo) {
    auto& rc = *static_cast<RefCase*>(h);
    try {
        return rc.dim == 2 ? fsSolve<2>(rc, which, x, itersErrInfo) : fsSolve<3>(rc, which, x, itersErrInfo);
    } catch (const std::exception& e) {
        g_lastError = e.what();
        return -1;
    }
}

// One PSPG time step of the equation: MomContEqIncompNewton::solve() = the Picard loop.  Returns 1 ok / 0 failed.
int pfem_ref_pspg_solve(void* h) {
    auto& rc = *static_cast<RefCase*>(h);
    try {
        rc.rebuildPositions();
        return rc.solver->m_pEquations[0]->solve() ? 1 : 0;
    } catch (const std::exception& e) {
        g_lastError = e.what();
        return -1;
    }
}

// The heat equation of the Boussinesq problem: m_pEquations[1]->solve() (IN/Solver.cpp:249-258).  Returns 1 ok / 0 failed.
int pfem_ref_heat_solve(void* h) {
    auto& rc = *static_cast<RefCase*>(h);
    try {
        rc.rebuildPositions();
        if (rc.solver->m_pEquations.size() < 2) return -1;
        return rc.solver->m_pEquations[1]->solve() ? 1 : 0;
    } catch (const std::exception& e) {
        g_lastError = e.what();
        return -1;
    }
}

// One explicit step: SolverWCompNewton::m_solveWCompNewtonNoT with the given dt (WC/Solver.cpp:236-276)
int pfem_ref_wc_step(void* h, double dt) {
    auto& rc = *static_cast<RefCase*>(h);
    try {
        rc.rebuildPositions();
        auto* s = dynamic_cast<SolverWCompNewton*>(rc.solver);
        if (!s) return -1;
        s->m_timeStep = dt;
        return s->solveOneTimeStep() ? 1 : 0;  // m_solveFunc: m_solveWCompNewtonNoT or m_solveBoussinesqWC
    } catch (const std::exception& e) {
        g_lastError = e.what();
        return -1;
    }
}
// SolverWCompNewton::computeNextDT (WC/Solver.cpp:192-234)
double pfem_ref_wc_next_dt(void* h) {
    auto& rc = *static_cast<RefCase*>(h);
    try {
        rc.solver->computeNextDT();
        return rc.solver->getTimeStep();
    } catch (const std::exception& e) {
        g_lastError = e.what();
        return -1.0;
    }
}

#ifdef PFEM_REF_WITH_B200
// ---- drop-in build: the reference's solver objects with the B200 equation classes swapped in ---------------------------
// Equivalent of editing the REGISTER_EQ site (IN/Solver.cpp:38-40): m_pEquations[0] becomes MomContEqIncompNewtonB200<dim>,
// constructed with exactly the arguments REGISTER_EQ passes (IN/Solver.cpp:6-10, 22-37).
int pfem_ref_use_b200_equation(void* h) {
    auto& rc = *static_cast<RefCase*>(h);
    try {
        Solver* s = rc.solver;
        std::vector<SolTable> materialParams(rc.problem->m_problemParams.size());
        for (std::size_t i = 0; i < materialParams.size(); ++i) materialParams[i] = SolTable("Material", rc.problem->m_problemParams[i]);
        const bool boussinesq = rc.problemId == "Boussinesq";
        const std::vector<unsigned short> bcFlags = {0};
        std::vector<unsigned int> statesIndex = {0};
        if (boussinesq) statesIndex.push_back(static_cast<unsigned int>(rc.dim) + 1);  // IN/Solver.cpp:63-64
        if (rc.dim == 2)
            s->m_pEquations[0] = std::make_unique<MomContEqIncompNewtonB200<2>>(rc.problem.get(), s, rc.mesh, s->m_solverParams, materialParams, bcFlags, statesIndex);
        else
            s->m_pEquations[0] = std::make_unique<MomContEqIncompNewtonB200<3>>(rc.problem.get(), s, rc.mesh, s->m_solverParams, materialParams, bcFlags, statesIndex);
        if (boussinesq) {  // IN/Solver.cpp:70-75: the heat equation, flags {1,2,3,4}, state dim + 1
            const std::vector<unsigned short> hFlags = {1, 2, 3, 4};
            const std::vector<unsigned int> hStates = {static_cast<unsigned int>(rc.dim) + 1};
            if (rc.dim == 2)
                s->m_pEquations[1] = std::make_unique<HeatEqIncompNewtonB200<2>>(rc.problem.get(), s, rc.mesh, s->m_solverParams, materialParams, hFlags, hStates);
            else
                s->m_pEquations[1] = std::make_unique<HeatEqIncompNewtonB200<3>>(rc.problem.get(), s, rc.mesh, s->m_solverParams, materialParams, hFlags, hStates);
        }
        return 0;
    } catch (const std::exception& e) {
        g_lastError = e.what();
        return -1;
    }
}

}  // extern "C"
template <unsigned short dim> int wcStepB200(RefCase& rc, double dt, int download) {
    using Shim = WCompNewtonStepB200<dim>;
    Solver* s = rc.solver;
    if (!rc.wcShim) {  // the object SolverWCompNewton would own; tables as in WC/Solver.cpp:26-30 and Equation.cpp:21-27
        SolTable material("Material", rc.problem->m_problemParams[0]);
        SolTable cont("ContEq", s->m_solverParams[0]), mom("MomEq", s->m_solverParams[0]);
        SolTable bc("BC", mom);
        if (rc.problemId == "BoussinesqWC") {  // the heat equation's BC table: m_pEquations[2]->getBCParam(0)
            SolTable heat("HeatEq", s->m_solverParams[0]);
            SolTable heatBc("BC", heat);
            rc.wcShim = std::make_shared<Shim>(rc.problem.get(), s, rc.mesh, material, cont, mom, bc,
                                               static_cast<SolverWCompNewton*>(s)->m_securityCoeff, &heatBc);
        } else
        rc.wcShim = std::make_shared<Shim>(rc.problem.get(), s, rc.mesh, material, cont, mom, bc, static_cast<SolverWCompNewton*>(s)->m_securityCoeff);
    }
    auto* shim = static_cast<Shim*>(rc.wcShim.get());
    s->m_timeStep = dt;
    shim->step();                       // replaces kick + move + continuity + momentum of m_solveWCompNewtonNoT
    rc.problem->updateTime(dt);         // WC/Solver.cpp:265
    if (download) shim->download();     // what the host does before a remesh / an extractor write
    return 1;
}
extern "C" {
int pfem_ref_wc_step_b200(void* h, double dt, int download) {
    auto& rc = *static_cast<RefCase*>(h);
    try {
        rc.rebuildPositions();
        return rc.dim == 2 ? wcStepB200<2>(rc, dt, download) : wcStepB200<3>(rc, dt, download);
    } catch (const std::exception& e) {
        g_lastError = e.what();
        return -1;
    }
}
double pfem_ref_wc_next_dt_b200(void* h) {
    auto& rc = *static_cast<RefCase*>(h);
    try {
        if (!rc.wcShim) return -1.0;
        const double maxDT = rc.solver->m_maxDT;
        return rc.dim == 2 ? static_cast<WCompNewtonStepB200<2>*>(rc.wcShim.get())->nextDT(maxDT)
                           : static_cast<WCompNewtonStepB200<3>*>(rc.wcShim.get())->nextDT(maxDT);
    } catch (const std::exception& e) {
        g_lastError = e.what();
        return -1.0;
    }
}
#endif

}  // extern "C"
