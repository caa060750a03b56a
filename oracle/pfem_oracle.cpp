// =====================================================================================
// oracle/pfem_oracle.cpp  --  TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT PATH.
//
// CPU restatement (plain C++17 + OpenMP, no Eigen) of the per-time-step finite-element
// hot path of PFEM3D (ImperatorS79/PFEM).  Every function cites the reference file:line
// it restates (paths relative to /root/reference/).  The arithmetic deliberately follows
// the reference's *literal* structure -- Gauss-point loops calling a factor functor,
// dense B^T*ddev*B products, triplets -> duplicate-summing CSC compression, serial nodal
// scatters -- and NOT the closed forms the CUDA kernels use, so that a CUDA-vs-oracle
// comparison is a comparison of two independent derivations.
//
// PARITY PINNED AGAINST THE REFERENCE'S OWN CODE (with one stated caveat).  The reference ships no tests, golden
// vectors or fixtures for this path (CMakeLists.txt:88 enable_testing() with zero add_test) and its real build needs
// Eigen, CGAL, gmsh and Lua/sol2 -- all absent here, no network.  oracle/refbuild/ therefore compiles the reference's
// UNMODIFIED hot-path sources where they lie under /root/reference (Element.cpp, Mesh.cpp, MatricesBuilder.inl,
// MomContEquation*.inl, PicardAlgo.cpp, WCompNewton/{Cont,Mom}Equation.inl, both Solver.cpp, ...) against original
// stand-in headers for the Eigen/sol2/gmsh calls they make, into oracle/_ref/libpfem_ref.so; its outputs are committed
// as tests/golden/*.npz (generator: tests/golden/make_golden.py).  tests/test_oracle_vs_reference.py checks this file
// against those fixtures and, where the library is present, against live runs: assembled CSC pattern identical, values,
// RHS, tau, Picard fields/iteration counts, explicit steps and CFL dt all agree (in practice bit for bit).
// Caveat: the dense/sparse arithmetic library underneath that build is the stand-in, not Eigen itself -- the element
// and assembly LOGIC is the reference's, last-bit rounding of Eigen's own kernels is not observable here.
// Additional pins: a second independent numpy restatement (oracle/literal_numpy.py), algebraic invariants
// (tests/test_oracle_invariants.py), scipy SuperLU as the stand-in for Eigen::SparseLU.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library.
// =====================================================================================
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <numeric>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace {

// ---------------------------------------------------------------- node flag bits (ABI)
constexpr uint8_t F_BOUND = 1, F_FREE = 2, F_FIXED = 4, F_FS = 8;

// ------------------------------------------------------------------------------------
// Quadrature tables.  srcs/mesh/Mesh.cpp:342-414 (points), :416-470 (weights),
// :472-488 (reference size), :490-530 (shape functions); 2-D 3-point, 3-D 4-point rules
// selected by MomContEquation.inl:12-23 / WCompNewton/MomEquation.inl:13-24.
// ------------------------------------------------------------------------------------
template <int DIM> struct Quad;
template <> struct Quad<2> {
    static constexpr int NGP = 3;
    static void gp(double g[3][3]) {
        const double p[3][3] = {{1.0 / 6.0, 1.0 / 6.0, 0.0}, {1.0 / 6.0, 2.0 / 3.0, 0.0}, {2.0 / 3.0, 1.0 / 6.0, 0.0}};
        std::memcpy(g, p, sizeof(p));
    }
    static double w(int) { return 1.0 / 3.0; }
    static double ref() { return 0.5; }
};
template <> struct Quad<3> {
    static constexpr int NGP = 4;
    static void gp(double g[4][3]) {
        const double p[4][3] = {{0.585410196624968, 0.138196601125011, 0.138196601125011},
                                {0.138196601125011, 0.585410196624968, 0.138196601125011},
                                {0.138196601125011, 0.138196601125011, 0.585410196624968},
                                {0.138196601125011, 0.138196601125011, 0.138196601125011}};
        std::memcpy(g, p, sizeof(p));
    }
    static double w(int) { return 0.25; }
    static double ref() { return 0.16666666666666666666666666666667; }
};

// ------------------------------------------------------------------------------------
// Element geometry.  srcs/mesh/Element.cpp:15-69 (J), :71-86 (detJ, Sarrus), :88-135
// (cofactor inverse, each entry divided by detJ).
// ------------------------------------------------------------------------------------
template <int DIM> struct Geo {
    double J[3][3];
    double invJ[3][3];
    double detJ;
};

template <int DIM>
void computeGeo(const double* x, int64_t nNodes, const int64_t* en, Geo<DIM>& G) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) G.J[i][j] = G.invJ[i][j] = 0.0;
    auto X = [&](int node, int d) { return x[en[node] + (int64_t)d * nNodes]; };
    for (int d = 0; d < DIM; ++d)
        for (int k = 0; k < DIM; ++k) G.J[d][k] = X(k + 1, d) - X(0, d);
    auto& J = G.J;
    if constexpr (DIM == 2) {
        G.detJ = J[0][0] * J[1][1] - J[1][0] * J[0][1];
        G.invJ[0][0] = J[1][1] / G.detJ;
        G.invJ[0][1] = -J[0][1] / G.detJ;
        G.invJ[1][0] = -J[1][0] / G.detJ;
        G.invJ[1][1] = J[0][0] / G.detJ;
    } else {
        G.detJ = J[0][0] * J[1][1] * J[2][2] + J[0][1] * J[1][2] * J[2][0] + J[0][2] * J[1][0] * J[2][1] -
                 J[2][0] * J[1][1] * J[0][2] - J[2][1] * J[1][2] * J[0][0] - J[2][2] * J[1][0] * J[0][1];
        const double d = G.detJ;
        G.invJ[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) / d;
        G.invJ[0][1] = (J[2][1] * J[0][2] - J[2][2] * J[0][1]) / d;
        G.invJ[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / d;
        G.invJ[1][0] = (J[2][0] * J[1][2] - J[1][0] * J[2][2]) / d;
        G.invJ[1][1] = (J[0][0] * J[2][2] - J[2][0] * J[0][2]) / d;
        G.invJ[1][2] = (J[1][0] * J[0][2] - J[0][0] * J[1][2]) / d;
        G.invJ[2][0] = (J[1][0] * J[2][1] - J[2][0] * J[1][1]) / d;
        G.invJ[2][1] = (J[2][0] * J[0][1] - J[0][0] * J[2][1]) / d;
        G.invJ[2][2] = (J[0][0] * J[1][1] - J[1][0] * J[0][1]) / d;
    }
}

// ------------------------------------------------------------------------------------
// MatrixBuilder restatement.  srcs/simulation/matricesBuilder/MatricesBuilder.inl.
// NPE = dim+1 nodes, NS = Voigt size (3 | 6), ND = dim*NPE velocity dofs per element.
// Factor functors take the shape-function row N_g, exactly where the reference calls its
// std::function (MB.inl:223, 284, 300, 316, 332, 349, 380).
// ------------------------------------------------------------------------------------
template <int DIM> struct MB {
    static constexpr int NPE = DIM + 1;
    static constexpr int NS = DIM * DIM - 2 * DIM + 3;
    static constexpr int ND = DIM * NPE;
    static constexpr int NGP = Quad<DIM>::NGP;

    double N[NGP][NPE];          // m_NHD        (MB.inl:26-45)
    double NtN[NGP][NPE][NPE];   // m_NhdTNhd    (MB.inl:72-76)
    double Nt[NGP][DIM][ND];     // m_NHDtilde   (MB.inl:28-42)
    double w[NGP];
    double ref;
    double ddev[NS][NS];
    double m[NS];

    MB() {
        double gp[NGP][3];
        Quad<DIM>::gp(gp);
        for (int g = 0; g < NGP; ++g) {
            w[g] = Quad<DIM>::w(g);
            // Mesh.cpp:509-519
            double s = 1.0;
            for (int d = 0; d < DIM; ++d) s -= gp[g][d];
            N[g][0] = s;
            for (int d = 0; d < DIM; ++d) N[g][d + 1] = gp[g][d];
            for (int r = 0; r < DIM; ++r)
                for (int c = 0; c < ND; ++c) Nt[g][r][c] = 0.0;
            for (int r = 0; r < DIM; ++r)
                for (int k = 0; k < NPE; ++k) Nt[g][r][r * NPE + k] = N[g][k];
            for (int i = 0; i < NPE; ++i)
                for (int j = 0; j < NPE; ++j) NtN[g][i][j] = N[g][i] * N[g][j];
        }
        ref = Quad<DIM>::ref();
        for (int i = 0; i < NS; ++i) {
            m[i] = 0;
            for (int j = 0; j < NS; ++j) ddev[i][j] = 0;
        }
    }

    // MB.inl:93-127
    static void gradN(const Geo<DIM>& G, double g[DIM][NPE]) {
        for (int d = 0; d < DIM; ++d) {
            double s = -G.invJ[0][d];
            for (int k = 1; k < DIM; ++k) s = s - G.invJ[k][d];
            g[d][0] = s;
            for (int k = 0; k < DIM; ++k) g[d][k + 1] = G.invJ[k][d];
        }
    }
    // MB.inl:130-164
    static void Bmat(const double g[DIM][NPE], double B[NS][ND]) {
        for (int i = 0; i < NS; ++i)
            for (int j = 0; j < ND; ++j) B[i][j] = 0.0;
        if constexpr (DIM == 2) {
            for (int k = 0; k < 3; ++k) {
                B[0][k] = B[2][3 + k] = g[0][k];
                B[1][3 + k] = B[2][k] = g[1][k];
            }
        } else {
            for (int k = 0; k < 4; ++k) {
                B[0][k] = B[3][4 + k] = B[4][8 + k] = g[0][k];
                B[1][4 + k] = B[3][k] = B[5][8 + k] = g[1][k];
                B[2][8 + k] = B[4][k] = B[5][4 + k] = g[2][k];
            }
        }
    }
    // MB.inl:217-229
    template <class F> void getM(const Geo<DIM>& G, F&& f, double M[NPE][NPE]) const {
        for (int i = 0; i < NPE; ++i)
            for (int j = 0; j < NPE; ++j) M[i][j] = 0.0;
        for (int g = 0; g < NGP; ++g) {
            const double fg = f(N[g]);
            for (int i = 0; i < NPE; ++i)
                for (int j = 0; j < NPE; ++j) M[i][j] += fg * NtN[g][i][j] * w[g];
        }
        const double s = G.detJ * ref;
        for (int i = 0; i < NPE; ++i)
            for (int j = 0; j < NPE; ++j) M[i][j] *= s;
    }
    // MB.inl:277-290   K = detJ*ref*fact * B^T * ddev * B  (evaluated left to right)
    template <class F> void getK(const Geo<DIM>& G, const double B[NS][ND], F&& f, double K[ND][ND]) const {
        double fact = 0;
        for (int g = 0; g < NGP; ++g) fact += f(N[g]) * w[g];
        const double s = G.detJ * ref * fact;
        double T1[ND][NS], T2[ND][NS];
        for (int i = 0; i < ND; ++i)
            for (int k = 0; k < NS; ++k) T1[i][k] = s * B[k][i];
        for (int i = 0; i < ND; ++i)
            for (int k = 0; k < NS; ++k) {
                double a = 0;
                for (int l = 0; l < NS; ++l) a += T1[i][l] * ddev[l][k];
                T2[i][k] = a;
            }
        for (int i = 0; i < ND; ++i)
            for (int j = 0; j < ND; ++j) {
                double a = 0;
                for (int l = 0; l < NS; ++l) a += T2[i][l] * B[l][j];
                K[i][j] = a;
            }
    }
    // MB.inl:293-306   D = detJ*ref * sumWNT * m^T * B
    template <class F> void getD(const Geo<DIM>& G, const double B[NS][ND], F&& f, double D[NPE][ND]) const {
        double sumWNT[NPE];
        for (int i = 0; i < NPE; ++i) sumWNT[i] = 0;
        for (int g = 0; g < NGP; ++g) {
            const double fg = f(N[g]);
            for (int i = 0; i < NPE; ++i) sumWNT[i] += fg * N[g][i] * w[g];
        }
        const double s = G.detJ * ref;
        double mB[ND];
        for (int j = 0; j < ND; ++j) {
            double a = 0;
            for (int l = 0; l < NS; ++l) a += m[l] * B[l][j];
            mB[j] = a;
        }
        for (int i = 0; i < NPE; ++i)
            for (int j = 0; j < ND; ++j) D[i][j] = (s * sumWNT[i]) * mB[j];
    }
    // MB.inl:309-322
    template <class F> void getL(const Geo<DIM>& G, const double g[DIM][NPE], F&& f, double L[NPE][NPE]) const {
        double fact = 0;
        for (int q = 0; q < NGP; ++q) fact += f(N[q]) * w[q];
        const double s = G.detJ * ref * fact;
        for (int i = 0; i < NPE; ++i)
            for (int j = 0; j < NPE; ++j) {
                double a = 0;
                for (int d = 0; d < DIM; ++d) a += (s * g[d][i]) * g[d][j];
                L[i][j] = a;
            }
    }
    // MB.inl:325-338
    template <class F> void getC(const Geo<DIM>& G, const double g[DIM][NPE], F&& f, double C[NPE][ND]) const {
        double sumNW[DIM][ND];
        for (int r = 0; r < DIM; ++r)
            for (int c = 0; c < ND; ++c) sumNW[r][c] = 0;
        for (int q = 0; q < NGP; ++q) {
            const double fg = f(N[q]);
            for (int r = 0; r < DIM; ++r)
                for (int c = 0; c < ND; ++c) sumNW[r][c] += fg * Nt[q][r][c] * w[q];
        }
        const double s = G.detJ * ref;
        for (int i = 0; i < NPE; ++i)
            for (int c = 0; c < ND; ++c) {
                double a = 0;
                for (int d = 0; d < DIM; ++d) a += (s * g[d][i]) * sumNW[d][c];
                C[i][c] = a;
            }
    }
    // MB.inl:341-355
    template <class F> void getF(const Geo<DIM>& G, const double vec[DIM], F&& f, double Fv[ND]) const {
        for (int c = 0; c < ND; ++c) Fv[c] = 0;
        for (int q = 0; q < NGP; ++q) {
            const double fg = f(N[q]);
            for (int c = 0; c < ND; ++c) {
                double a = 0;
                for (int r = 0; r < DIM; ++r) a += (fg * Nt[q][r][c]) * vec[r];
                Fv[c] += a * w[q];
            }
        }
        const double s = G.detJ * ref;
        for (int c = 0; c < ND; ++c) Fv[c] *= s;
    }
    // MB.inl:373-386
    template <class F>
    void getH(const Geo<DIM>& G, const double vec[DIM], const double g[DIM][NPE], F&& f, double H[NPE]) const {
        for (int i = 0; i < NPE; ++i) H[i] = 0;
        for (int q = 0; q < NGP; ++q) {
            const double fg = f(N[q]);
            for (int i = 0; i < NPE; ++i) {
                double a = 0;
                for (int d = 0; d < DIM; ++d) a += (fg * g[d][i]) * vec[d];
                H[i] += a * w[q];
            }
        }
        const double s = G.detJ * ref;
        for (int i = 0; i < NPE; ++i) H[i] *= s;
    }
};

// ddev / m of the incompressible equation.  MomContEquation.inl:73-95.
template <int DIM> void setIncomp(MB<DIM>& mb) {
    constexpr int NS = MB<DIM>::NS;
    for (int i = 0; i < NS; ++i) {
        mb.m[i] = (i < DIM) ? 1.0 : 0.0;
        for (int j = 0; j < NS; ++j) mb.ddev[i][j] = (i == j) ? ((i < DIM) ? 2.0 : 1.0) : 0.0;
    }
}

template <int DIM> double dotN(const double* N, const double* s) {
    double a = 0;
    for (int i = 0; i < DIM + 1; ++i) a += N[i] * s[i];
    return a;
}

struct Triplet {
    int32_t r, c;
    double v;
};  // Eigen::Triplet<double> default-constructs to (0,0,0.0)

// ------------------------------------------------------------------------------------
// Eigen::SparseMatrix::setFromTriplets semantics (PSPG.inl:136; Eigen 3.3.x, not in tree):
// column-major result, inner (row) indices sorted, duplicates summed in triplet order,
// explicit zeros kept.  Implemented as stable bucket-by-column + stable sort by row.
// ------------------------------------------------------------------------------------
void tripletsToCSC(int64_t n, const std::vector<Triplet>& T, std::vector<int64_t>& colPtr, std::vector<int32_t>& rowIdx,
                   std::vector<double>& val) {
    std::vector<int64_t> cnt(n + 1, 0);
    for (const auto& t : T) cnt[t.c + 1]++;
    for (int64_t c = 0; c < n; ++c) cnt[c + 1] += cnt[c];
    std::vector<Triplet> S(T.size());
    {
        std::vector<int64_t> cur(cnt.begin(), cnt.end() - 1);
        for (const auto& t : T) S[cur[t.c]++] = t;
    }
    std::vector<int64_t> nnzCol(n, 0);
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t c = 0; c < n; ++c) {
        auto b = S.begin() + cnt[c], e = S.begin() + cnt[c + 1];
        std::stable_sort(b, e, [](const Triplet& a, const Triplet& bb) { return a.r < bb.r; });
        int64_t k = 0;
        int32_t last = -1;
        for (auto it = b; it != e; ++it)
            if (it->r != last) {
                ++k;
                last = it->r;
            }
        nnzCol[c] = k;
    }
    colPtr.assign(n + 1, 0);
    for (int64_t c = 0; c < n; ++c) colPtr[c + 1] = colPtr[c] + nnzCol[c];
    rowIdx.resize(colPtr[n]);
    val.resize(colPtr[n]);
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t c = 0; c < n; ++c) {
        int64_t o = colPtr[c] - 1;
        int32_t last = -1;
        for (int64_t k = cnt[c]; k < cnt[c + 1]; ++k) {
            if (S[k].r != last) {
                ++o;
                last = S[k].r;
                rowIdx[o] = S[k].r;
                val[o] = S[k].v;
            } else
                val[o] += S[k].v;
        }
    }
}

struct PspgParams {
    double rho, mu, dt, bodyForce[3];
};

// ------------------------------------------------------------------------------------
// tau_PSPG.  MomContEquationPSPG.inl:238-259 (h uses the 2-D formula in 3-D as well;
// U from the CURRENT node states, not qPrev).
// ------------------------------------------------------------------------------------
template <int DIM>
double tauPSPG(const Geo<DIM>& G, const int64_t* en, const double* vcur, int64_t nNodes, const PspgParams& P) {
    const double h = std::sqrt(Quad<DIM>::ref() * G.detJ / M_PI);
    double U = 0;
    for (int n = 0; n < DIM + 1; ++n) {
        double nodeU = 0;
        for (int d = 0; d < DIM; ++d) {
            const double s = vcur[en[n] + (int64_t)d * nNodes];
            nodeU += s * s;
        }
        U += std::sqrt(nodeU);
    }
    U /= (DIM + 1);
    return 1 / std::sqrt((2 / P.dt) * (2 / P.dt) + (2 * U / h) * (2 * U / h) +
                         9 * (4 * P.mu / (h * h * P.rho)) * (4 * P.mu / (h * h * P.rho)));
}

// ------------------------------------------------------------------------------------
// Per-element PSPG system.  MomContEquationPSPG.inl:26-53; factors MomContEquation.inl:
// 73-93 (ddev, m), :97-100 (M: rho), :123-128 (K: mu), :131-135 (D: 1), :139-149 (C: 1,
// L: 1/rho), :203-207 (F: rho), :218-222 (H: 1).
// ------------------------------------------------------------------------------------
// Bingham regularised viscosity (Problem id "Bingham"): the K factor of MomContEquation.inl:102-119,
// mu + tau0 (1 - exp(-mReg gammaDot))/gammaDot with gammaDot = sqrt(V^T B^T ddev B V) of the element's CURRENT velocities.
// Incompressible Boussinesq (Problem id "Boussinesq"): F factor rho (1 - alpha (T - Tr)) and H factor (1 - alpha (T - Tr))
// with T = N . T_e of the CURRENT node states (MomContEquation.inl:166-199), gamma = DgammaDT = 0, no phase change.
struct PspgThermalCtx {
    bool on = false;
    double alpha = 0, Tr = 0;
    const double* T = nullptr;
};
PspgThermalCtx g_pspgThermal;

struct BinghamCtx {
    bool on = false;
    double tau0 = 0, mReg = 0;
};
BinghamCtx g_bingham;

template <int DIM>
void pspgElement(const MB<DIM>& mb, const double* x, const double* vcur, const double* qPrev, int64_t nNodes,
                 const int64_t* en, const PspgParams& P, double* Ae /*[(DIM+1)*NPE]^2 row-major*/, double* be,
                 double* tauOut) {
    constexpr int NPE = DIM + 1, ND = DIM * NPE, NT = (DIM + 1) * NPE, NS = MB<DIM>::NS;
    Geo<DIM> G;
    computeGeo<DIM>(x, nNodes, en, G);
    const double tau = tauPSPG<DIM>(G, en, vcur, nNodes, P);
    if (tauOut) *tauOut = tau;
    double g[DIM][NPE], B[NS][ND];
    MB<DIM>::gradN(G, g);
    MB<DIM>::Bmat(g, B);
    double M[NPE][NPE], K[ND][ND], D[NPE][ND], C[NPE][ND], L[NPE][NPE], F[ND], H[NPE];
    mb.getM(G, [&](const double*) { return P.rho; }, M);
    for (int i = 0; i < NPE; ++i)
        for (int j = 0; j < NPE; ++j) M[i][j] = (1 / P.dt) * M[i][j];
    if (g_bingham.on) {
        double V[ND];
        for (int d = 0; d < DIM; ++d)
            for (int k = 0; k < NPE; ++k) V[d * NPE + k] = vcur[en[k] + (int64_t)d * nNodes];
        mb.getK(G, B, [&](const double*) {
            double t1[NS], t2[NS], t3[ND];  // ((V^T B^T) ddev) B, left to right
            for (int k = 0; k < NS; ++k) {
                double a = 0;
                for (int r = 0; r < ND; ++r) a += V[r] * B[k][r];
                t1[k] = a;
            }
            for (int c = 0; c < NS; ++c) {
                double a = 0;
                for (int k = 0; k < NS; ++k) a += t1[k] * mb.ddev[k][c];
                t2[c] = a;
            }
            for (int c = 0; c < ND; ++c) {
                double a = 0;
                for (int k = 0; k < NS; ++k) a += t2[k] * B[k][c];
                t3[c] = a;
            }
            double gd2 = 0;
            for (int r = 0; r < ND; ++r) gd2 += t3[r] * V[r];
            const double gammaDot = std::sqrt(gd2);
            double muEq = g_bingham.tau0;
            if (gammaDot < 1e-15)
                muEq *= g_bingham.mReg;
            else
                muEq *= (1 - std::exp(-g_bingham.mReg * gammaDot)) / gammaDot;
            return P.mu + muEq;
        }, K);
    } else
        mb.getK(G, B, [&](const double*) { return P.mu; }, K);
    mb.getD(G, B, [&](const double*) { return 1.0; }, D);
    mb.getC(G, g, [&](const double*) { return 1.0; }, C);
    for (int i = 0; i < NPE; ++i)
        for (int j = 0; j < ND; ++j) C[i][j] = (tau / P.dt) * C[i][j];
    mb.getL(G, g, [&](const double*) { return 1 / P.rho; }, L);
    for (int i = 0; i < NPE; ++i)
        for (int j = 0; j < NPE; ++j) L[i][j] = tau * L[i][j];
    if (g_pspgThermal.on) {
        double Te[NPE];
        for (int k = 0; k < NPE; ++k) Te[k] = g_pspgThermal.T[en[k]];
        const double al = g_pspgThermal.alpha, Tr = g_pspgThermal.Tr;
        mb.getF(G, P.bodyForce, [&](const double* N) { return P.rho * (1 - al * (dotN<DIM>(N, Te) - Tr)); }, F);
        mb.getH(G, P.bodyForce, g, [&](const double* N) { return (1 - al * (dotN<DIM>(N, Te) - Tr)); }, H);
    } else {
        mb.getF(G, P.bodyForce, [&](const double*) { return P.rho; }, F);
        mb.getH(G, P.bodyForce, g, [&](const double*) { return 1.0; }, H);
    }
    for (int i = 0; i < NPE; ++i) H[i] = tau * H[i];

    auto A = [&](int r, int c) -> double& { return Ae[r * NT + c]; };
    // Ae << Me_dt + Ke, -De^T, Ce_dt + De, Le   (PSPG.inl:42), Me_dt = diagBlock (MB.hpp:120-129)
    for (int r = 0; r < ND; ++r)
        for (int c = 0; c < ND; ++c) {
            double mv = 0;
            if (r / NPE == c / NPE) mv = M[r % NPE][c % NPE];
            A(r, c) = mv + K[r][c];
        }
    for (int r = 0; r < ND; ++r)
        for (int c = 0; c < NPE; ++c) A(r, ND + c) = -D[c][r];
    for (int r = 0; r < NPE; ++r)
        for (int c = 0; c < ND; ++c) A(ND + r, c) = C[r][c] + D[r][c];
    for (int r = 0; r < NPE; ++r)
        for (int c = 0; c < NPE; ++c) A(ND + r, ND + c) = L[r][c];

    // vPrev: StatesFromToQ.hpp:136-170 ; be << Fe + Me_dt*vPrev, He + Ce_dt*vPrev (PSPG.inl:53)
    double vPrev[ND];
    for (int d = 0; d < DIM; ++d)
        for (int k = 0; k < NPE; ++k) vPrev[d * NPE + k] = qPrev[en[k] + (int64_t)d * nNodes];
    for (int r = 0; r < ND; ++r) {
        double a = 0;
        const int blk = r / NPE;
        for (int k = 0; k < NPE; ++k) a += M[r % NPE][k] * vPrev[blk * NPE + k];
        be[r] = F[r] + a;
    }
    for (int r = 0; r < NPE; ++r) {
        double a = 0;
        for (int c = 0; c < ND; ++c) a += C[r][c] * vPrev[c];
        be[ND + r] = H[r] + a;
    }
}

// ------------------------------------------------------------------------------------
// m_buildAbPSPG.  MomContEquationPSPG.inl:7-146: omp element loop -> row-masked triplets
// (default (0,0,0) triplets for masked slots), identity triplets appended in node order
// (pressure first, then velocity components, :108-128), setFromTriplets, serial RHS.
// ------------------------------------------------------------------------------------
template <int DIM>
void pspgBuild(int64_t nNodes, int64_t nElm, const int64_t* conn, const double* x, const double* vcur,
               const double* qPrev, const uint8_t* flags, const PspgParams& P, std::vector<int64_t>& colPtr,
               std::vector<int32_t>& rowIdx, std::vector<double>& val, double* b, double* phaseSec) {
    constexpr int NPE = DIM + 1, NT = (DIM + 1) * NPE;
    const int64_t tripletPerElm = (int64_t)NT * NT, doubletPerElm = NT;
    MB<DIM> mb;
    setIncomp<DIM>(mb);
    double t0 = omp_get_wtime();
    std::vector<Triplet> indexA(tripletPerElm * nElm, Triplet{0, 0, 0.0});
    std::vector<std::pair<int64_t, double>> indexb(doubletPerElm * nElm);
    const int64_t nDof = (DIM + 1) * nNodes;
    for (int64_t i = 0; i < nDof; ++i) b[i] = 0;
    double t1 = omp_get_wtime();
    if (phaseSec) phaseSec[0] += t1 - t0;  // "Prepare matrix assembly"
#pragma omp parallel for default(shared)
    for (int64_t elm = 0; elm < nElm; ++elm) {
        const int64_t* en = conn + elm * NPE;
        double Ae[NT * NT], be[NT];
        pspgElement<DIM>(mb, x, vcur, qPrev, nNodes, en, P, Ae, be, nullptr);
        int64_t countA = 0, countb = 0;
        for (int i = 0; i < NPE; ++i) {
            const uint8_t fi = flags[en[i]];
            const bool bound = fi & F_BOUND, free_ = fi & F_FREE;
            for (int j = 0; j < NPE; ++j) {
                for (int d1 = 0; d1 < DIM; ++d1)
                    for (int d2 = 0; d2 <= DIM; ++d2) {
                        if (!(bound || free_))
                            indexA[tripletPerElm * elm + countA] =
                                Triplet{(int32_t)(en[i] + d1 * nNodes), (int32_t)(en[j] + d2 * nNodes),
                                        Ae[(i + d1 * NPE) * NT + (j + d2 * NPE)]};
                        countA++;
                    }
                for (int d2 = 0; d2 <= DIM; ++d2) {
                    if (!free_)
                        indexA[tripletPerElm * elm + countA] =
                            Triplet{(int32_t)(en[i] + DIM * nNodes), (int32_t)(en[j] + d2 * nNodes),
                                    Ae[(i + DIM * NPE) * NT + (j + d2 * NPE)]};
                    countA++;
                }
            }
            for (int d = 0; d <= DIM; ++d) {
                indexb[doubletPerElm * elm + countb] = std::make_pair(en[i] + d * nNodes, be[i + d * NPE]);
                countb++;
            }
        }
    }
    double t2 = omp_get_wtime();
    if (phaseSec) phaseSec[1] += t2 - t1;  // "Compute triplets"
    for (int64_t n = 0; n < nNodes; ++n) {
        const bool bound = flags[n] & F_BOUND, free_ = flags[n] & F_FREE;
        if (free_) indexA.push_back(Triplet{(int32_t)(n + DIM * nNodes), (int32_t)(n + DIM * nNodes), 1.0});
        if (bound || free_)
            for (int d = 0; d < DIM; ++d)
                indexA.push_back(Triplet{(int32_t)(n + d * nNodes), (int32_t)(n + d * nNodes), 1.0});
    }
    double t3 = omp_get_wtime();
    if (phaseSec) phaseSec[2] += t3 - t2;  // "Push back (n, n, 1)"
    tripletsToCSC(nDof, indexA, colPtr, rowIdx, val);
    double t4 = omp_get_wtime();
    if (phaseSec) phaseSec[3] += t4 - t3;  // "Assemble matrix"
    for (const auto& d : indexb) b[d.first] += d.second;
    double t5 = omp_get_wtime();
    if (phaseSec) phaseSec[4] += t5 - t4;  // "Assemble vector"
}

// ------------------------------------------------------------------------------------
// HeatEqIncompNewton (IncompNewton/HeatEquation.inl): m_buildAb (:227-319) -- A = M + dt L with f_M = cv rho, f_L = k,
// rows of nodes with a temperature BC flag or free nodes skipped (identity appended in node order, :300-308), b = M
// theta_prev through the reference's `indexb[noPerEl*elm + countb]` indexing (a smaller stride than the vector was sized
// for: the tail keeps default pairs (0, 0.0), harmless) -- and m_applyBC (:321-412) without the flux facet terms:
// free nodes b = theta_prev, Dirichlet nodes b = g_T with the column elimination keeping explicit zeros.
// params: [rho, cv, k, dt].  tMask/tVal as in the BoussinesqWC block.
// ------------------------------------------------------------------------------------
template <int DIM>
void inHeatBuild(int64_t nNodes, int64_t nElm, const int64_t* conn, const double* x, const uint8_t* flags,
                 const uint8_t* tMask, const double* tVal, const double* thetaPrev, const double* par, int applyBC,
                 std::vector<int64_t>& colPtr, std::vector<int32_t>& rowIdx, std::vector<double>& val, double* b) {
    constexpr int NPE = DIM + 1;
    const double rho = par[0], cv = par[1], k = par[2], dt = par[3];
    const int64_t tripletPerElm = DIM * NPE * NPE + DIM * NPE * DIM * NPE + 3 * NPE * DIM * NPE + NPE * NPE;  // :230
    const int64_t doubletPerElm = 2 * DIM * NPE + 2 * NPE;
    MB<DIM> mb;
    std::vector<Triplet> indexA(tripletPerElm * nElm, Triplet{0, 0, 0.0});
    std::vector<std::pair<int64_t, double>> indexb(doubletPerElm * nElm, std::make_pair((int64_t)0, 0.0));
    for (int64_t i = 0; i < nNodes; ++i) b[i] = 0;
#pragma omp parallel for default(shared)
    for (int64_t elm = 0; elm < nElm; ++elm) {
        const int64_t* en = conn + elm * NPE;
        Geo<DIM> G;
        computeGeo<DIM>(x, nNodes, en, G);
        double g[DIM][NPE];
        MB<DIM>::gradN(G, g);
        double Me[NPE][NPE], Le[NPE][NPE];
        mb.getM(G, [&](const double*) { return cv * rho; }, Me);
        mb.getL(G, g, [&](const double*) { return k; }, Le);
        for (int i = 0; i < NPE; ++i)
            for (int j = 0; j < NPE; ++j) Le[i][j] = dt * Le[i][j];
        double Mth[NPE];
        for (int i = 0; i < NPE; ++i) {
            double a = 0;
            for (int j = 0; j < NPE; ++j) a += Me[i][j] * thetaPrev[en[j]];
            Mth[i] = a;
        }
        int64_t countA = 0, countb = 0;
        for (int i = 0; i < NPE; ++i) {
            const bool free_ = flags[en[i]] & F_FREE, tbc = tMask[en[i]] != 0;
            for (int j = 0; j < NPE; ++j) {
                if (!tbc) {
                    if (!free_) indexA[tripletPerElm * elm + countA] = Triplet{(int32_t)en[i], (int32_t)en[j], Me[i][j]};
                    countA++;
                    if (!free_) indexA[tripletPerElm * elm + countA] = Triplet{(int32_t)en[i], (int32_t)en[j], Le[i][j]};
                    countA++;
                }
            }
            indexb[NPE * elm + countb] = std::make_pair(en[i], Mth[i]);  // the reference's stride (:283)
            countb++;
        }
    }
    for (int64_t n = 0; n < nNodes; ++n)
        if (tMask[n] || (flags[n] & F_FREE)) indexA.push_back(Triplet{(int32_t)n, (int32_t)n, 1.0});
    tripletsToCSC(nNodes, indexA, colPtr, rowIdx, val);
    for (const auto& d : indexb) b[d.first] += d.second;
    if (!applyBC) return;
    for (int64_t n = 0; n < nNodes; ++n) {
        const bool free_ = flags[n] & F_FREE, tbc = tMask[n] != 0;
        if (free_ && !tbc) {
            b[n] = thetaPrev[n];
        } else if (tbc) {
            const double r = tVal[n];
            b[n] = r;
            for (int64_t kk = colPtr[n]; kk < colPtr[n + 1]; ++kk) {
                const int64_t row = rowIdx[kk];
                if (row == n) continue;
                const double v = val[kk];
                b[row] -= v * r;
                val[kk] = 0;
            }
        }
    }
}

// ------------------------------------------------------------------------------------
// Surface tension on free-surface facets.  Facet::computeJ/computeDetJ/computeNormal (srcs/mesh/Facet.cpp:16-77,
// 130-210), MatrixBuilder::getP/getT/getFST (MB.inl:167-214, 389-403; factor = gamma, MomContEquation.inl:213-218,
// WC/MomEquation.inl:141-146), the facet loops of m_applyBCPSPG (PSPG.inl:155-187: a facet counts when ANY of its
// nodes is on the free surface) and of MomEqWCompNewton::m_applyBC (WC/MomEquation.inl:312-336: Facet::
// isOnFreeSurface = ALL nodes, Facet.cpp:249-255).  Facet quadrature: 3 points in both dimensions
// (MomContEquation.inl:13-23) with weights summing to one (Mesh.cpp:428, 443-444), reference sizes 2 and 1/2 (:476-480).
// facets: nF x (DIM+2) = DIM facet nodes, the node in front of the facet (m_outNodeIndex), the element index.
// ------------------------------------------------------------------------------------
struct FacetCtx {
    int64_t nF = 0;
    std::vector<int64_t> facets;
    double gamma = 0;
};
FacetCtx g_facets;

template <int DIM>
void addFST(int64_t nNodes, const int64_t* conn, const double* x, const uint8_t* flags, bool allNodesRule, double* target) {
    constexpr int NPE = DIM + 1, ND = DIM * NPE, NS = MB<DIM>::NS, NPF = DIM;
    const FacetCtx& C = g_facets;
    if (C.gamma < 1e-15) return;                       // PSPG.inl:157 / MomEquation.inl:314 (DgammaDT = 0 outside Boussinesq)
    const double wLD[3] = {DIM == 2 ? 5.0 / 18.0 : 1.0 / 3.0, DIM == 2 ? 8.0 / 18.0 : 1.0 / 3.0, DIM == 2 ? 5.0 / 18.0 : 1.0 / 3.0};
    const double refLD = (DIM == 2) ? 2.0 : 0.5;
    auto X = [&](int64_t node, int d) { return x[node + (int64_t)d * nNodes]; };
    for (int64_t f = 0; f < C.nF; ++f) {
        const int64_t* row = C.facets.data() + f * (DIM + 2);
        int nOnFS = 0;
        for (int k = 0; k < NPF; ++k) nOnFS += (flags[row[k]] & F_FS) ? 1 : 0;
        if (allNodesRule ? (nOnFS != NPF) : (nOnFS == 0)) continue;
        const int64_t out = row[NPF], elm = row[NPF + 1];
        // Facet::computeJ / computeDetJ / computeNormal
        double detJf, nrm[3] = {0, 0, 0};
        if constexpr (DIM == 2) {
            const double x0 = X(row[0], 0), x1 = X(row[1], 0), y0 = X(row[0], 1), y1 = X(row[1], 1);
            const double J00 = (x1 - x0) / 2, J10 = (y1 - y0) / 2;
            detJf = std::sqrt(J00 * J00 + J10 * J10);
            nrm[0] = y1 - y0;
            nrm[1] = x0 - x1;
            const double vo[2] = {X(out, 0) - x0, X(out, 1) - y0};
            if (nrm[0] * vo[0] + nrm[1] * vo[1] > 0) {
                nrm[0] *= -1;
                nrm[1] *= -1;
            }
            const double norm = std::sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1]);
            nrm[0] /= norm;
            nrm[1] /= norm;
        } else {
            double J[3][2];
            for (int d = 0; d < 3; ++d) {
                J[d][0] = X(row[1], d) - X(row[0], d);
                J[d][1] = X(row[2], d) - X(row[0], d);
            }
            const double dG[3] = {J[1][0] * J[2][1] - J[1][1] * J[2][0], J[2][0] * J[0][1] - J[2][1] * J[0][0],
                                  J[0][0] * J[1][1] - J[1][0] * J[0][1]};
            detJf = std::sqrt(dG[0] * dG[0] + dG[1] * dG[1] + dG[2] * dG[2]);
            const double a[3] = {J[0][0], J[1][0], J[2][0]}, b[3] = {J[0][1], J[1][1], J[2][1]};
            nrm[0] = a[1] * b[2] - a[2] * b[1];
            nrm[1] = a[2] * b[0] - a[0] * b[2];
            nrm[2] = a[0] * b[1] - a[1] * b[0];
            const double vo[3] = {X(out, 0) - X(row[0], 0), X(out, 1) - X(row[0], 1), X(out, 2) - X(row[0], 2)};
            if (vo[0] * nrm[0] + vo[1] * nrm[1] + vo[2] * nrm[2] > 0)
                for (int d = 0; d < 3; ++d) nrm[d] *= -1.0;
            const double norm = std::sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
            for (int d = 0; d < 3; ++d) nrm[d] /= norm;
        }
        // getP, getT
        double P[NS], T[NS][NS];
        if constexpr (DIM == 2) {
            P[0] = 1 - nrm[0] * nrm[0];
            P[1] = 1 - nrm[1] * nrm[1];
            P[2] = -nrm[0] * nrm[1];
            T[0][0] = P[0] * P[0]; T[0][1] = P[2] * P[2]; T[0][2] = 2 * P[0] * P[2];
            T[1][0] = T[0][1];     T[1][1] = P[1] * P[1]; T[1][2] = 2 * P[2] * P[1];
            T[2][0] = P[0] * P[2]; T[2][1] = P[2] * P[1]; T[2][2] = P[0] * P[1] + P[2] * P[2];
        } else {
            P[0] = 1 - nrm[0] * nrm[0];
            P[1] = 1 - nrm[1] * nrm[1];
            P[2] = 1 - nrm[2] * nrm[2];
            P[3] = -nrm[0] * nrm[1];
            P[4] = -nrm[0] * nrm[2];
            P[5] = -nrm[1] * nrm[2];
            T[0][0] = P[0] * P[0]; T[0][1] = P[3] * P[3]; T[0][2] = P[4] * P[4]; T[0][3] = 2 * P[0] * P[3]; T[0][5] = 2 * P[4] * P[3]; T[0][4] = 2 * P[0] * P[4];
            T[1][0] = T[0][1];     T[1][1] = P[1] * P[1]; T[1][2] = P[5] * P[5]; T[1][3] = 2 * P[3] * P[1]; T[1][5] = 2 * P[5] * P[1]; T[1][4] = 2 * P[3] * P[5];
            T[2][0] = T[0][2];     T[2][1] = T[1][2];     T[2][2] = P[2] * P[2]; T[2][3] = 2 * P[5] * P[4]; T[2][5] = 2 * P[5] * P[2]; T[2][4] = 2 * P[2] * P[4];
            T[3][0] = P[0] * P[3]; T[3][1] = P[3] * P[1]; T[3][2] = P[4] * P[5]; T[3][3] = P[0] * P[1] + P[3] * P[3]; T[3][5] = P[4] * P[1] + P[5] * P[3]; T[3][4] = P[4] * P[3] + P[0] * P[5];
            T[5][0] = P[3] * P[4]; T[5][1] = P[5] * P[1]; T[5][2] = P[5] * P[2]; T[5][3] = P[4] * P[1] + P[5] * P[3]; T[5][5] = P[1] * P[2] + P[5] * P[5]; T[5][4] = P[5] * P[4] + P[3] * P[2];
            T[4][0] = P[0] * P[4]; T[4][1] = P[3] * P[5]; T[4][2] = P[2] * P[4]; T[4][3] = P[4] * P[3] + P[0] * P[5]; T[4][5] = P[5] * P[4] + P[3] * P[2]; T[4][4] = P[2] * P[0] + P[4] * P[4];
        }
        // element gradN, B
        const int64_t* en = conn + elm * NPE;
        Geo<DIM> G;
        computeGeo<DIM>(x, nNodes, en, G);
        double g[DIM][NPE], B[NS][ND];
        MB<DIM>::gradN(G, g);
        MB<DIM>::Bmat(g, B);
        // getFST: FST -= fact*Be^T*T*P*w per Gauss point (left to right), then *= detJ*refSize
        double FST[ND];
        for (int r = 0; r < ND; ++r) FST[r] = 0.0;
        for (int gp = 0; gp < 3; ++gp) {
            double gBt[ND][NS], gBtT[ND][NS];
            for (int r = 0; r < ND; ++r)
                for (int k = 0; k < NS; ++k) gBt[r][k] = C.gamma * B[k][r];
            for (int r = 0; r < ND; ++r)
                for (int c = 0; c < NS; ++c) {
                    double sum = 0;
                    for (int k = 0; k < NS; ++k) sum += gBt[r][k] * T[k][c];
                    gBtT[r][c] = sum;
                }
            for (int r = 0; r < ND; ++r) {
                double sum = 0;
                for (int k = 0; k < NS; ++k) sum += gBtT[r][k] * P[k];
                FST[r] -= sum * wLD[gp];
            }
        }
        for (int r = 0; r < ND; ++r) FST[r] *= detJf * refLD;
        for (int n = 0; n < NPE; ++n)
            for (int d = 0; d < DIM; ++d) target[en[n] + (int64_t)d * nNodes] += FST[n + d * NPE];
    }
}

// ------------------------------------------------------------------------------------
// m_applyBCPSPG.  MomContEquationPSPG.inl:149-235 (gamma = 0: facet loop skipped, :157).
// dirMask[n] != 0  <=>  node.isBound() && getBcTagFlags(tag, flag0) (:206-208);
// dirVal[n + d*nNodes] = the Lua "<type>V" result (:211-214), host-evaluated.
// Column walk + explicit zeros: :219-228.
// ------------------------------------------------------------------------------------
template <int DIM>
void pspgApplyBC(int64_t nNodes, const uint8_t* flags, const uint8_t* dirMask, const double* dirVal,
                 const double* qPrev, const PspgParams& P, const int64_t* colPtr, const int32_t* rowIdx, double* val,
                 double* b) {
    for (int64_t n = 0; n < nNodes; ++n) {
        const bool bound = flags[n] & F_BOUND, free_ = flags[n] & F_FREE;
        if (free_) {
            b[n + DIM * nNodes] = 0;
            if (!bound)
                for (int d = 0; d < DIM; ++d) b[n + d * nNodes] = qPrev[n + d * nNodes] + P.dt * P.bodyForce[d];
        }
        if (bound && dirMask[n]) {
            for (int d = 0; d < DIM; ++d) {
                const int64_t col = n + d * nNodes;
                const double r = dirVal[col];
                b[col] = r;
                for (int64_t k = colPtr[col]; k < colPtr[col + 1]; ++k) {
                    const int64_t row = rowIdx[k];
                    if (row == col) continue;
                    const double v = val[k];
                    b[row] -= v * r;
                    val[k] = 0;
                }
            }
        }
    }
}

struct WcParams {
    double mu, K0, K0p, rhoStar, bodyForce[3];
    int meduri;
    int eqType;  // 0 CDS_dpdt, 1 CDS_drhodt, 2 CDS_rho  (ContEquation.inl:38-43)
};

inline WcParams wcParamsFromArray(const double* a) {
    return WcParams{a[0], a[1], a[2], a[3], {a[4], a[5], a[6]}, (int)a[7], (int)a[8]};
}

// ------------------------------------------------------------------------------------
// Mesh::updateNodesPosition.  srcs/mesh/Mesh.cpp:1101-1137: x += delta unless m_isFixed.
// (Jacobians are recomputed from coordinates on use in this restatement.)
// ------------------------------------------------------------------------------------
void movePositions(int dim, int64_t nNodes, const uint8_t* flags, const double* delta, const double* base, double* x) {
#pragma omp parallel for
    for (int64_t n = 0; n < nNodes; ++n)
        if (!(flags[n] & F_FIXED))
            for (int d = 0; d < dim; ++d) x[n + d * nNodes] = base[n + d * nNodes] + delta[n + d * nNodes];
}

// ------------------------------------------------------------------------------------
// ContEqWCompNewton (CDS_dpdt).  WCompNewton/ContEquation.inl:353-413 (build), :334-350
// (BC), :123-148 (solve), :319-331 (Tait-Murnaghan); factors :75-86.
// ------------------------------------------------------------------------------------
template <int DIM>
void wcCont(int64_t nNodes, int64_t nElm, const int64_t* conn, const double* x, const uint8_t* flags, const double* v,
            double* p, double* rho, const WcParams& P, double dt) {
    constexpr int NPE = DIM + 1, ND = DIM * NPE, NS = MB<DIM>::NS;
    MB<DIM> mb;
    for (int i = 0; i < NS; ++i) mb.m[i] = (i < DIM) ? 1.0 : 0.0;
    std::vector<double> MeL((size_t)nElm * NPE), F0e((size_t)nElm * NPE);
#pragma omp parallel for default(shared)
    for (int64_t elm = 0; elm < nElm; ++elm) {
        const int64_t* en = conn + elm * NPE;
        Geo<DIM> G;
        computeGeo<DIM>(x, nNodes, en, G);
        double Me[NPE][NPE];
        mb.getM(G, [](const double*) { return 1.0; }, Me);
        double lumped[NPE];  // lump2, MB.hpp:87-102
        for (int i = 0; i < NPE; ++i) {
            lumped[i] = 0;
            for (int j = 0; j < NPE; ++j) lumped[i] += Me[i][j];
        }
        double Pe[NPE], V[ND];
        for (int k = 0; k < NPE; ++k) Pe[k] = p[en[k]];
        for (int d = 0; d < DIM; ++d)
            for (int k = 0; k < NPE; ++k) V[d * NPE + k] = v[en[k] + (int64_t)d * nNodes];
        double g[DIM][NPE], B[NS][ND], D[NPE][ND];
        MB<DIM>::gradN(G, g);
        MB<DIM>::Bmat(g, B);
        mb.getD(G, B, [&](const double* N) { return P.K0 + P.K0p * dotN<DIM>(N, Pe); }, D);
        for (int i = 0; i < NPE; ++i) {
            double a = 0;
            for (int c = 0; c < ND; ++c) a += (-dt * D[i][c]) * V[c];
            double s = 0;
            if (P.meduri)
                for (int j = 0; j < NPE; ++j) s += Me[i][j] * Pe[j];
            else
                s = lumped[i] * Pe[i];
            F0e[elm * NPE + i] = a + s;
            MeL[elm * NPE + i] = lumped[i];
        }
    }
    std::vector<double> invM(nNodes, 0.0), F0(nNodes, 0.0);
    for (int64_t elm = 0; elm < nElm; ++elm)  // serial scatter, :398-408
        for (int i = 0; i < NPE; ++i) {
            invM[conn[elm * NPE + i]] += MeL[elm * NPE + i];
            F0[conn[elm * NPE + i]] += F0e[elm * NPE + i];
        }
    for (int64_t n = 0; n < nNodes; ++n) invM[n] = 1 / invM[n];
    for (int64_t n = 0; n < nNodes; ++n)
        if (flags[n] & F_FREE) {
            F0[n] = 0;
            invM[n] = 1;
        }
    for (int64_t n = 0; n < nNodes; ++n) {
        p[n] = invM[n] * F0[n];
        rho[n] = std::pow((P.K0p / P.K0) * p[n] + 1, 1 / P.K0p) * P.rhoStar;
    }
}

// ------------------------------------------------------------------------------------
// ContEqWCompNewton (CDS_rho / CDS_drhodt).  WCompNewton/ContEquation.inl:196-232
// (m_buildF0, called by preCompute :171-175 on the configuration *before* the move),
// :234-301 (m_buildSystem), :177-194 (BC: free or free-surface nodes get rho*),
// :149-165 (solve), :303-316 (Tait-Murnaghan p(rho)); factors :75-97 (M: 1, D: N.rho_e).
// ------------------------------------------------------------------------------------
template <int DIM>
void wcBuildF0(int64_t nNodes, int64_t nElm, const int64_t* conn, const double* x, const double* rho, double* F0) {
    constexpr int NPE = DIM + 1, NS = MB<DIM>::NS;
    MB<DIM> mb;
    for (int i = 0; i < NS; ++i) mb.m[i] = (i < DIM) ? 1.0 : 0.0;
    std::vector<double> F0e((size_t)nElm * NPE);
#pragma omp parallel for default(shared)
    for (int64_t elm = 0; elm < nElm; ++elm) {
        const int64_t* en = conn + elm * NPE;
        Geo<DIM> G;
        computeGeo<DIM>(x, nNodes, en, G);
        double Me[NPE][NPE];
        mb.getM(G, [](const double*) { return 1.0; }, Me);
        for (int i = 0; i < NPE; ++i) {
            double s = 0;
            for (int j = 0; j < NPE; ++j) s += Me[i][j] * rho[en[j]];
            F0e[elm * NPE + i] = s;
        }
    }
    for (int64_t n = 0; n < nNodes; ++n) F0[n] = 0;
    for (int64_t elm = 0; elm < nElm; ++elm)
        for (int i = 0; i < NPE; ++i) F0[conn[elm * NPE + i]] += F0e[elm * NPE + i];
}

template <int DIM>
void wcContRho(int64_t nNodes, int64_t nElm, const int64_t* conn, const double* x, const uint8_t* flags,
               const double* v, double* p, double* rho, const WcParams& P, double dt, std::vector<double>& F0) {
    constexpr int NPE = DIM + 1, ND = DIM * NPE, NS = MB<DIM>::NS;
    MB<DIM> mb;
    for (int i = 0; i < NS; ++i) mb.m[i] = (i < DIM) ? 1.0 : 0.0;
    const bool drhodt = P.eqType == 1;
    std::vector<double> MeL((size_t)nElm * NPE), F0e(drhodt ? (size_t)nElm * NPE : 0);
#pragma omp parallel for default(shared)
    for (int64_t elm = 0; elm < nElm; ++elm) {
        const int64_t* en = conn + elm * NPE;
        Geo<DIM> G;
        computeGeo<DIM>(x, nNodes, en, G);
        double Me[NPE][NPE];
        mb.getM(G, [](const double*) { return 1.0; }, Me);
        double lumped[NPE];
        for (int i = 0; i < NPE; ++i) {
            lumped[i] = 0;
            for (int j = 0; j < NPE; ++j) lumped[i] += Me[i][j];
            MeL[elm * NPE + i] = lumped[i];
        }
        if (!drhodt) continue;
        double Re[NPE], V[ND];
        for (int k = 0; k < NPE; ++k) Re[k] = rho[en[k]];
        for (int d = 0; d < DIM; ++d)
            for (int k = 0; k < NPE; ++k) V[d * NPE + k] = v[en[k] + (int64_t)d * nNodes];
        double g[DIM][NPE], B[NS][ND], D[NPE][ND];
        MB<DIM>::gradN(G, g);
        MB<DIM>::Bmat(g, B);
        mb.getD(G, B, [&](const double* N) { return dotN<DIM>(N, Re); }, D);
        for (int i = 0; i < NPE; ++i) {
            double a = 0;
            for (int c = 0; c < ND; ++c) a += (-dt * D[i][c]) * V[c];
            double s = 0;
            if (P.meduri)  // "!= Stab::None", :268-271
                for (int j = 0; j < NPE; ++j) s += Me[i][j] * Re[j];
            else
                s = lumped[i] * Re[i];
            F0e[elm * NPE + i] = a + s;
        }
    }
    std::vector<double> invM(nNodes, 0.0);
    if (drhodt) F0.assign(nNodes, 0.0);
    for (int64_t elm = 0; elm < nElm; ++elm)
        for (int i = 0; i < NPE; ++i) {
            invM[conn[elm * NPE + i]] += MeL[elm * NPE + i];
            if (drhodt) F0[conn[elm * NPE + i]] += F0e[elm * NPE + i];
        }
    for (int64_t n = 0; n < nNodes; ++n) invM[n] = 1 / invM[n];
    for (int64_t n = 0; n < nNodes; ++n)
        if (flags[n] & (F_FREE | F_FS)) {
            F0[n] = P.rhoStar;
            invM[n] = 1;
        }
    for (int64_t n = 0; n < nNodes; ++n) {
        rho[n] = invM[n] * F0[n];
        p[n] = (P.K0 / P.K0p) * (std::pow(rho[n] / P.rhoStar, P.K0p) - 1);
    }
}

// ------------------------------------------------------------------------------------
// MomEqWCompNewton.  WCompNewton/MomEquation.inl:229-302 (build), :305-374 (BC, gamma=0),
// :201-226 (solve); factors :63-103, :135-139.  The "acceleration" of a Dirichlet node is
// set to the Dirichlet *velocity* value (:361-369), reproduced as is.
// ------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------
// BoussinesqWC additions.  HeatEqWCompNewton (WCompNewton/HeatEquation.inl:154-298: lumped
// M with f = cv (N.rho), L with f = k, FTot = -dt L T + M T, serial scatter, BC :171-226 without the flux terms
// Q/Qh/Qr, T = invM F), the buoyancy factor rho (1 - alpha (T - Tr)) of the momentum body force
// (MomEquation.inl:105-112), the thermal diffusivity k/(cv rho) in computeNextDT (Solver.cpp:214-216,
// HeatEquation.inl:124-127) and the order heat -> continuity -> momentum of m_solveBoussinesqWC (Solver.cpp:278-320).
// tMask[n] != 0  <=>  getBcTagFlags(node.getTag(), flag 1) (a "<type>T" Lua function exists); tVal[n] its value.
// ------------------------------------------------------------------------------------
struct ThermalCtx {
    bool on = false;
    double k = 0, cv = 1, alpha = 0, Tr = 0;
    double* T = nullptr;            // nodal temperature, updated in place by oracle_wc_step
    const uint8_t* tMask = nullptr;
    const double* tVal = nullptr;
};
ThermalCtx g_thermal;

template <int DIM>
void wcHeat(int64_t nNodes, int64_t nElm, const int64_t* conn, const double* x, const uint8_t* flags, const double* rho,
            double dt) {
    constexpr int NPE = DIM + 1;
    const ThermalCtx& C = g_thermal;
    MB<DIM> mb;
    std::vector<double> Mdiag((size_t)nElm * NPE), FTot((size_t)nElm * NPE);
#pragma omp parallel for default(shared)
    for (int64_t elm = 0; elm < nElm; ++elm) {
        const int64_t* en = conn + elm * NPE;
        Geo<DIM> G;
        computeGeo<DIM>(x, nNodes, en, G);
        double Te[NPE], Re[NPE];
        for (int k = 0; k < NPE; ++k) {
            Te[k] = C.T[en[k]];
            Re[k] = rho[en[k]];
        }
        double g[DIM][NPE];
        MB<DIM>::gradN(G, g);
        double Me[NPE][NPE], Le[NPE][NPE];
        mb.getM(G, [&](const double* N) { return C.cv * dotN<DIM>(N, Re); }, Me);
        for (int i = 0; i < NPE; ++i)  // lump, MB.hpp:104-118
            for (int j = 0; j < NPE; ++j)
                if (i != j) {
                    Me[i][i] += Me[i][j];
                    Me[i][j] = 0;
                }
        mb.getL(G, g, [&](const double*) { return C.k; }, Le);
        for (int i = 0; i < NPE; ++i) {
            double lt = 0, mt = 0;
            for (int j = 0; j < NPE; ++j) lt += ((-dt) * Le[i][j]) * Te[j];
            for (int j = 0; j < NPE; ++j) mt += Me[i][j] * Te[j];
            FTot[elm * NPE + i] = lt + mt;
            Mdiag[elm * NPE + i] = Me[i][i];
        }
    }
    std::vector<double> invM((size_t)nNodes, 0.0), F((size_t)nNodes, 0.0);
    for (int64_t elm = 0; elm < nElm; ++elm)
        for (int i = 0; i < NPE; ++i) {
            invM[conn[elm * NPE + i]] += Mdiag[elm * NPE + i];
            F[conn[elm * NPE + i]] += FTot[elm * NPE + i];
        }
    for (auto& m : invM) m = 1 / m;
    for (int64_t n = 0; n < nNodes; ++n) {
        const bool free_ = flags[n] & F_FREE, tbc = C.tMask && C.tMask[n];
        if (free_ && !tbc) {
            F[n] = C.T[n];
            invM[n] = 1;
        } else if (tbc) {
            F[n] = C.tVal[n];
            invM[n] = 1;
        }
    }
    for (int64_t n = 0; n < nNodes; ++n) C.T[n] = invM[n] * F[n];
}

template <int DIM>
void wcMom(int64_t nNodes, int64_t nElm, const int64_t* conn, const double* x, const uint8_t* flags,
           const uint8_t* dirMask, const double* dirVal, double* v /*in: v_half, out: v*/, double* acc, const double* p,
           const double* rho, const WcParams& P, double dt) {
    constexpr int NPE = DIM + 1, ND = DIM * NPE, NS = MB<DIM>::NS;
    MB<DIM> mb;
    for (int i = 0; i < NS; ++i) mb.m[i] = (i < DIM) ? 1.0 : 0.0;
    for (int i = 0; i < NS; ++i)
        for (int j = 0; j < NS; ++j) {
            if (i < DIM && j < DIM)
                mb.ddev[i][j] = (i == j) ? 4.0 / 3 : -2.0 / 3;
            else
                mb.ddev[i][j] = (i == j) ? 1.0 : 0.0;
        }
    std::vector<double> Mdiag((size_t)nElm * ND), FTot((size_t)nElm * ND);
#pragma omp parallel for default(shared)
    for (int64_t elm = 0; elm < nElm; ++elm) {
        const int64_t* en = conn + elm * NPE;
        Geo<DIM> G;
        computeGeo<DIM>(x, nNodes, en, G);
        double V[ND], Pe[NPE], Re[NPE];
        for (int d = 0; d < DIM; ++d)
            for (int k = 0; k < NPE; ++k) V[d * NPE + k] = v[en[k] + (int64_t)d * nNodes];
        for (int k = 0; k < NPE; ++k) {
            Pe[k] = p[en[k]];
            Re[k] = rho[en[k]];
        }
        double g[DIM][NPE], B[NS][ND];
        MB<DIM>::gradN(G, g);
        MB<DIM>::Bmat(g, B);
        double Mt[NPE][NPE];
        mb.getM(G, [&](const double* N) { return dotN<DIM>(N, Re); }, Mt);
        // diagBlock then lump (MB.hpp:104-129): diagonal entry accumulates its row in column order
        double Me[ND][ND];
        for (int r = 0; r < ND; ++r)
            for (int c = 0; c < ND; ++c) Me[r][c] = (r / NPE == c / NPE) ? Mt[r % NPE][c % NPE] : 0.0;
        for (int i = 0; i < ND; ++i)
            for (int j = 0; j < ND; ++j)
                if (i != j) {
                    Me[i][i] += Me[i][j];
                    Me[i][j] = 0;
                }
        double K[ND][ND], D[NPE][ND], F[ND];
        mb.getK(G, B, [&](const double*) { return P.mu; }, K);
        mb.getD(G, B, [](const double*) { return 1.0; }, D);
        if (g_thermal.on) {  // MomEquation.inl:105-112
            double Te[NPE];
            for (int k = 0; k < NPE; ++k) Te[k] = g_thermal.T[en[k]];
            mb.getF(G, P.bodyForce, [&](const double* N) {
                const double r = dotN<DIM>(N, Re), T = dotN<DIM>(N, Te);
                return r * (1 - g_thermal.alpha * (T - g_thermal.Tr));
            }, F);
        } else
            mb.getF(G, P.bodyForce, [&](const double* N) { return dotN<DIM>(N, Re); }, F);
        for (int r = 0; r < ND; ++r) {
            double kv = 0;
            for (int c = 0; c < ND; ++c) kv += (-K[r][c]) * V[c];
            double dp = 0;
            for (int k = 0; k < NPE; ++k) dp += D[k][r] * Pe[k];
            FTot[elm * ND + r] = (kv + dp) + F[r];
            Mdiag[elm * ND + r] = Me[r][r];
        }
    }
    std::vector<double> invM((size_t)DIM * nNodes, 0.0), F((size_t)DIM * nNodes, 0.0);
    for (int64_t elm = 0; elm < nElm; ++elm)  // serial scatter, :279-298
        for (int i = 0; i < NPE; ++i)
            for (int d = 0; d < DIM; ++d) {
                invM[conn[elm * NPE + i] + (int64_t)d * nNodes] += Mdiag[elm * ND + i + d * NPE];
                F[conn[elm * NPE + i] + (int64_t)d * nNodes] += FTot[elm * ND + i + d * NPE];
            }
    for (size_t i = 0; i < invM.size(); ++i) invM[i] = 1 / invM[i];
    addFST<DIM>(nNodes, conn, x, flags, true, F.data());  // m_applyBC facet loop, MomEquation.inl:312-336
    for (int64_t n = 0; n < nNodes; ++n) {
        const bool bound = flags[n] & F_BOUND, free_ = flags[n] & F_FREE;
        if (free_ && !bound) {
            for (int d = 0; d < DIM; ++d) {
                F[n + (int64_t)d * nNodes] = P.bodyForce[d];
                invM[n + (int64_t)d * nNodes] = 1;
            }
        } else if (bound && dirMask[n]) {
            for (int d = 0; d < DIM; ++d) {
                F[n + (int64_t)d * nNodes] = dirVal[n + (int64_t)d * nNodes];
                invM[n + (int64_t)d * nNodes] = 1;
            }
        }
    }
    for (int64_t i = 0; i < (int64_t)DIM * nNodes; ++i) {
        const double a = invM[i] * F[i];
        acc[i] = a;
        v[i] = v[i] + 0.5 * dt * a;
    }
}

// ------------------------------------------------------------------------------------
// Element::getRin (srcs/mesh/Element.cpp:226-294) and SolverWCompNewton::computeNextDT
// (WCompNewton/Solver.cpp:192-234) with c^2 = (K0 + K0' p)/rho (ContEquation.inl:117-120),
// u^2 (MomEquation.inl:184-192), alpha = mu/rho (:195-198).
// ------------------------------------------------------------------------------------
template <int DIM> double rin(const double* x, int64_t nNodes, const int64_t* en) {
    Geo<DIM> G;
    computeGeo<DIM>(x, nNodes, en, G);
    const double size = G.detJ * Quad<DIM>::ref();
    auto X = [&](int node, int d) { return x[en[node] + (int64_t)d * nNodes]; };
    if constexpr (DIM == 2) {
        auto dist = [&](int a, int b) {
            // Node::distance (srcs/mesh/Node.cpp): Euclidean over the 3 stored coordinates (z = 0 in 2-D)
            double s = 0;
            for (int d = 0; d < 2; ++d) s += (X(a, d) - X(b, d)) * (X(a, d) - X(b, d));
            return std::sqrt(s);
        };
        const double a = dist(0, 1), b = dist(1, 2), c = dist(0, 2);
        const double s = (a + b + c) / 2;
        return size / s;
    } else {
        const double x0 = X(0, 0), x1 = X(1, 0), x2 = X(2, 0), x3 = X(3, 0);
        const double y0 = X(0, 1), y1 = X(1, 1), y2 = X(2, 1), y3 = X(3, 1);
        const double z0 = X(0, 2), z1 = X(1, 2), z2 = X(2, 2), z3 = X(3, 2);
        auto nrm = [](double a, double b, double c) { return std::sqrt(a * a + b * b + c * c); };
        const double n1 = nrm((y1 - y0) * (z2 - z0) - (z1 - z0) * (y2 - y0), (z1 - z0) * (x2 - x0) - (x1 - x0) * (z2 - z0),
                              (x1 - x0) * (y2 - y0) - (y1 - y0) * (x2 - x0));
        const double n2 = nrm((y3 - y0) * (z2 - z0) - (z3 - z0) * (y2 - y0), (z3 - z0) * (x2 - x0) - (x3 - x0) * (z2 - z0),
                              (x3 - x0) * (y2 - y0) - (y3 - y0) * (x2 - x0));
        const double n3 = nrm((y1 - y0) * (z3 - z0) - (z1 - z0) * (y3 - y0), (z1 - z0) * (x3 - x0) - (x1 - x0) * (z3 - z0),
                              (x1 - x0) * (y3 - y0) - (y1 - y0) * (x3 - x0));
        const double n4 = nrm((y1 - y3) * (z2 - z3) - (z1 - z3) * (y2 - y3), (z1 - z3) * (x2 - x3) - (x1 - x3) * (z2 - z3),
                              (x1 - x3) * (y2 - y3) - (y1 - y3) * (x2 - x3));
        return 6 * size / (n1 + n2 + n3 + n4);
    }
}

template <int DIM>
double wcNextDt(int64_t nNodes, int64_t nElm, const int64_t* conn, const double* x, const double* v, const double* p,
                const double* rho, const WcParams& P, double securityCoeff, double maxDT) {
    double ts = std::numeric_limits<double>::max();
#pragma omp parallel for reduction(min : ts)
    for (int64_t elm = 0; elm < nElm; ++elm) {
        const int64_t* en = conn + elm * (DIM + 1);
        const double he = 2 * rin<DIM>(x, nNodes, en);
        double mx = 0;
        for (int n = 0; n < DIM + 1; ++n) {
            const int64_t nd = en[n];
            const double c2 = (P.K0 + P.K0p * p[nd]) / rho[nd];
            double u2 = 0;
            for (int d = 0; d < DIM; ++d) u2 += v[nd + (int64_t)d * nNodes] * v[nd + (int64_t)d * nNodes];
            double alpha = P.mu / rho[nd];
            if (g_thermal.on) alpha = std::max(alpha, g_thermal.k / (g_thermal.cv * rho[nd]));  // Solver.cpp:214-216
            mx = std::max(std::max(u2, c2), mx);
            mx = std::max(mx, 4 * alpha * alpha / (he * he));
        }
        ts = std::min(ts, securityCoeff * securityCoeff * he * he / mx);
    }
    return std::min(std::sqrt(ts), maxDT);
}

// handle for CSC results that outlive one call
struct CscHandle {
    std::vector<int64_t> colPtr;
    std::vector<int32_t> rowIdx;
    std::vector<double> val;
};

}  // namespace

extern "C" {

int oracle_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#endif
}

// Per-element Ae (row-major (dim+1)*npe square) / be / tau for every element: test hook for MB.inl + PSPG.inl:26-53.
int oracle_pspg_elements(int dim, int64_t nNodes, int64_t nElm, const int64_t* conn, const double* x,
                         const double* vcur, const double* qPrev, const double* params /*rho,mu,dt,b[3]*/, double* Ae,
                         double* be, double* tau) {
    PspgParams P{params[0], params[1], params[2], {params[3], params[4], params[5]}};
    if (dim == 2) {
        MB<2> mb;
        setIncomp<2>(mb);
        constexpr int NT = 9;
#pragma omp parallel for
        for (int64_t e = 0; e < nElm; ++e)
            pspgElement<2>(mb, x, vcur, qPrev, nNodes, conn + e * 3, P, Ae + e * NT * NT, be + e * NT, tau + e);
    } else if (dim == 3) {
        MB<3> mb;
        setIncomp<3>(mb);
        constexpr int NT = 16;
#pragma omp parallel for
        for (int64_t e = 0; e < nElm; ++e)
            pspgElement<3>(mb, x, vcur, qPrev, nNodes, conn + e * 4, P, Ae + e * NT * NT, be + e * NT, tau + e);
    } else
        return -1;
    return 0;
}

// m_buildAbPSPG (+ optionally m_applyBCPSPG).  Returns a handle; query nnz, then copy out.
// phaseSec[5]: Prepare / Compute triplets / Push back / Assemble matrix / Assemble vector; [5] = Apply BC.
void* oracle_pspg_build(int dim, int64_t nNodes, int64_t nElm, const int64_t* conn, const double* x,
                        const double* vcur, const double* qPrev, const uint8_t* flags, const double* params,
                        int applyBC, const uint8_t* dirMask, const double* dirVal, double* b, double* phaseSec) {
    PspgParams P{params[0], params[1], params[2], {params[3], params[4], params[5]}};
    auto* H = new CscHandle;
    if (dim == 2)
        pspgBuild<2>(nNodes, nElm, conn, x, vcur, qPrev, flags, P, H->colPtr, H->rowIdx, H->val, b, phaseSec);
    else if (dim == 3)
        pspgBuild<3>(nNodes, nElm, conn, x, vcur, qPrev, flags, P, H->colPtr, H->rowIdx, H->val, b, phaseSec);
    else {
        delete H;
        return nullptr;
    }
    if (applyBC) {
        double t0 = omp_get_wtime();
        if (dim == 2)   // facet loop of m_applyBCPSPG (PSPG.inl:155-187) comes before the nodal BC pass
            addFST<2>(nNodes, conn, x, flags, false, b);
        else
            addFST<3>(nNodes, conn, x, flags, false, b);
        if (dim == 2)
            pspgApplyBC<2>(nNodes, flags, dirMask, dirVal, qPrev, P, H->colPtr.data(), H->rowIdx.data(), H->val.data(), b);
        else
            pspgApplyBC<3>(nNodes, flags, dirMask, dirVal, qPrev, P, H->colPtr.data(), H->rowIdx.data(), H->val.data(), b);
        if (phaseSec) phaseSec[5] += omp_get_wtime() - t0;
    }
    return H;
}
// Boundary facets + surface-tension coefficient used by the next oracle_pspg_build(applyBC) / oracle_wc_step calls
// (gamma < 1e-15 or nF = 0: no facet terms, the default).  facets: nF x (dim+2), see addFST.
void oracle_set_facets(int dim, int64_t nF, const int64_t* facets, double gamma) {
    g_facets.nF = nF;
    g_facets.facets.assign(facets, facets + (nF > 0 ? nF * (dim + 2) : 0));
    g_facets.gamma = (nF > 0) ? gamma : 0.0;
}
// Incompressible Boussinesq: buoyancy factors for the following oracle_pspg_* calls (on = 0: off); T = current node temperatures
void oracle_set_pspg_thermal(int on, double alpha, double Tr, const double* T) {
    g_pspgThermal.on = on != 0;
    g_pspgThermal.alpha = alpha, g_pspgThermal.Tr = Tr, g_pspgThermal.T = T;
}
// HeatEqIncompNewton::m_buildAb (+ m_applyBC): scalar system of nNodes unknowns; returns a CSC handle like oracle_pspg_build
void* oracle_in_heat_build(int dim, int64_t nNodes, int64_t nElm, const int64_t* conn, const double* x, const uint8_t* flags,
                           const uint8_t* tMask, const double* tVal, const double* thetaPrev, const double* params, int applyBC,
                           double* b) {
    auto* H = new CscHandle;
    if (dim == 2)
        inHeatBuild<2>(nNodes, nElm, conn, x, flags, tMask, tVal, thetaPrev, params, applyBC, H->colPtr, H->rowIdx, H->val, b);
    else if (dim == 3)
        inHeatBuild<3>(nNodes, nElm, conn, x, flags, tMask, tVal, thetaPrev, params, applyBC, H->colPtr, H->rowIdx, H->val, b);
    else {
        delete H;
        return nullptr;
    }
    return H;
}
// Bingham viscosity for the following oracle_pspg_elements / oracle_pspg_build calls (on = 0: Newtonian, the default)
void oracle_set_bingham(int on, double tau0, double mReg) {
    g_bingham.on = on != 0;
    g_bingham.tau0 = tau0, g_bingham.mReg = mReg;
}
// BoussinesqWC: thermal constants and nodal arrays used by the next oracle_wc_step / oracle_wc_next_dt calls (on = 0: off).
// T is updated in place by oracle_wc_step; the caller keeps the arrays alive.
void oracle_set_thermal(int on, double k, double cv, double alpha, double Tr, double* T, const uint8_t* tMask, const double* tVal) {
    g_thermal.on = on != 0;
    g_thermal.k = k, g_thermal.cv = cv, g_thermal.alpha = alpha, g_thermal.Tr = Tr;
    g_thermal.T = T, g_thermal.tMask = tMask, g_thermal.tVal = tVal;
}
int64_t oracle_csc_nnz(void* h) { return (int64_t) static_cast<CscHandle*>(h)->val.size(); }
void oracle_csc_copy(void* h, int64_t* colPtr, int32_t* rowIdx, double* val) {
    auto* H = static_cast<CscHandle*>(h);
    std::copy(H->colPtr.begin(), H->colPtr.end(), colPtr);
    std::copy(H->rowIdx.begin(), H->rowIdx.end(), rowIdx);
    std::copy(H->val.begin(), H->val.end(), val);
}
void oracle_csc_free(void* h) { delete static_cast<CscHandle*>(h); }

// Mesh::updateNodesPosition (base = x) / updateNodesPositionFromSave (base = saved), Mesh.cpp:1101-1137, 1238-1277
void oracle_move_positions(int dim, int64_t nNodes, const uint8_t* flags, const double* delta, const double* base,
                           double* x) {
    movePositions(dim, nNodes, flags, delta, base, x);
}

// One explicit step: WCompNewton/Solver.cpp:236-263.  State arrays are SoA: v[dim*nNodes], acc[dim*nNodes], p, rho.
// params: mu, K0, K0p, rhoStar, b[3], meduri
int oracle_wc_step(int dim, int64_t nNodes, int64_t nElm, const int64_t* conn, double* x, const uint8_t* flags,
                   const uint8_t* dirMask, const double* dirVal, double* v, double* acc, double* p, double* rho,
                   const double* params, double dt) {
    WcParams P = wcParamsFromArray(params);
    const int64_t nv = (int64_t)dim * nNodes;
    std::vector<double> delta(nv), F0;
    if (dim != 2 && dim != 3) return -1;
    // preCompute() before the move (Solver.cpp:244-246); only CDS_rho does anything there.
    if (P.eqType == 2) {
        F0.resize(nNodes);
        if (dim == 2)
            wcBuildF0<2>(nNodes, nElm, conn, x, rho, F0.data());
        else
            wcBuildF0<3>(nNodes, nElm, conn, x, rho, F0.data());
    }
    // qV1half = qVPrev + 0.5*dt*qAccPrev ; states <- v_half ; x += v_half*dt   (Solver.cpp:249-253)
    for (int64_t i = 0; i < nv; ++i) {
        v[i] = v[i] + 0.5 * dt * acc[i];
        delta[i] = v[i] * dt;
    }
    movePositions(dim, nNodes, flags, delta.data(), x, x);
    if (g_thermal.on) {  // m_solveBoussinesqWC: the heat equation comes first, on the moved mesh with the old rho (Solver.cpp:301-303)
        if (dim == 2)
            wcHeat<2>(nNodes, nElm, conn, x, flags, rho, dt);
        else
            wcHeat<3>(nNodes, nElm, conn, x, flags, rho, dt);
    }
    if (dim == 2) {
        if (P.eqType == 0)
            wcCont<2>(nNodes, nElm, conn, x, flags, v, p, rho, P, dt);
        else
            wcContRho<2>(nNodes, nElm, conn, x, flags, v, p, rho, P, dt, F0);
        wcMom<2>(nNodes, nElm, conn, x, flags, dirMask, dirVal, v, acc, p, rho, P, dt);
    } else {
        if (P.eqType == 0)
            wcCont<3>(nNodes, nElm, conn, x, flags, v, p, rho, P, dt);
        else
            wcContRho<3>(nNodes, nElm, conn, x, flags, v, p, rho, P, dt, F0);
        wcMom<3>(nNodes, nElm, conn, x, flags, dirMask, dirVal, v, acc, p, rho, P, dt);
    }
    return 0;
}

double oracle_wc_next_dt(int dim, int64_t nNodes, int64_t nElm, const int64_t* conn, const double* x, const double* v,
                         const double* p, const double* rho, const double* params, double securityCoeff, double maxDT) {
    WcParams P = wcParamsFromArray(params);
    if (dim == 2) return wcNextDt<2>(nNodes, nElm, conn, x, v, p, rho, P, securityCoeff, maxDT);
    return wcNextDt<3>(nNodes, nElm, conn, x, v, p, rho, P, securityCoeff, maxDT);
}

// y = A x for a CSC matrix, serial column sweep (what Eigen 3.3 does for a column-major sparse * dense vector;
// PSPG.inl:368).  ompRows != 0: transpose-free row-parallel variant is not possible on CSC, so the caller passes CSR.
void oracle_csc_matvec(int64_t n, const int64_t* colPtr, const int32_t* rowIdx, const double* val, const double* xv,
                       double* y) {
    for (int64_t i = 0; i < n; ++i) y[i] = 0;
    for (int64_t c = 0; c < n; ++c) {
        const double xc = xv[c];
        for (int64_t k = colPtr[c]; k < colPtr[c + 1]; ++k) y[rowIdx[k]] += val[k] * xc;
    }
}

// Jacobi-preconditioned BiCGSTAB on CSR (omp row-parallel SpMV), Eigen 3.3 formulation (IterativeLinearSolvers/BiCGSTAB.h,
// not in tree: restarts when |rho| < eps^2*|r0|^2, tol on ||r||/||b||).  CPU baseline for the Krylov solve only; the
// reference itself uses SparseLU on this system (PSPG.inl:281-290).
int oracle_bicgstab_csr(int64_t n, const int64_t* rowPtr, const int32_t* colIdx, const double* val, const double* b,
                        double* xsol, double tol, int maxIter, double* relRes) {
    std::vector<double> dinv(n), r(n), r0(n), p(n, 0.0), v(n, 0.0), s(n), t(n), y(n), z(n);
    auto spmv = [&](const double* in, double* out) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            double a = 0;
            for (int64_t k = rowPtr[i]; k < rowPtr[i + 1]; ++k) a += val[k] * in[colIdx[k]];
            out[i] = a;
        }
    };
    auto dot = [&](const double* a, const double* c) {
        double sum = 0;
#pragma omp parallel for reduction(+ : sum) schedule(static)
        for (int64_t i = 0; i < n; ++i) sum += a[i] * c[i];
        return sum;
    };
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double d = 0;
        for (int64_t k = rowPtr[i]; k < rowPtr[i + 1]; ++k)
            if (colIdx[k] == i) d = val[k];
        dinv[i] = (d != 0) ? 1 / d : 1;
    }
    spmv(xsol, r.data());
    for (int64_t i = 0; i < n; ++i) r0[i] = r[i] = b[i] - r[i];
    const double r0sq = dot(r0.data(), r0.data()), rhsSq = dot(b, b);
    if (rhsSq == 0) {
        for (int64_t i = 0; i < n; ++i) xsol[i] = 0;
        *relRes = 0;
        return 0;
    }
    double rho = 1, alpha = 1, w = 1;
    const double tol2 = tol * tol * rhsSq, eps2 = std::numeric_limits<double>::epsilon() * std::numeric_limits<double>::epsilon();
    int it = 0, restarts = 0;
    double rsq = r0sq;
    while (rsq > tol2 && it < maxIter) {
        const double rhoOld = rho;
        rho = dot(r0.data(), r.data());
        if (std::fabs(rho) < eps2 * r0sq) {
            spmv(xsol, r.data());
            for (int64_t i = 0; i < n; ++i) r0[i] = r[i] = b[i] - r[i];
            rho = dot(r.data(), r.data());
            if (restarts++ == 0) it = 0;
        }
        const double beta = (rho / rhoOld) * (alpha / w);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            p[i] = r[i] + beta * (p[i] - w * v[i]);
            y[i] = dinv[i] * p[i];
        }
        spmv(y.data(), v.data());
        alpha = rho / dot(r0.data(), v.data());
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            s[i] = r[i] - alpha * v[i];
            z[i] = dinv[i] * s[i];
        }
        spmv(z.data(), t.data());
        const double tt = dot(t.data(), t.data());
        w = (tt > 0) ? dot(t.data(), s.data()) / tt : 0;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            xsol[i] += alpha * y[i] + w * z[i];
            r[i] = s[i] - w * t[i];
        }
        rsq = dot(r.data(), r.data());
        ++it;
    }
    *relRes = std::sqrt(rsq / rhsSq);
    return it;
}

}  // extern "C"
