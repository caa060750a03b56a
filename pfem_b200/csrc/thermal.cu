// thermal.cu -- the SURVEY 8(f) rank-3 rows on the device: temperature-dependent and shear-rate-dependent factors and the
// heat equations.
//   * Bingham regularised viscosity      (IncompNewton/MomContEquation.inl:102-119): K factor mu + tau0 (1 - exp(-m gd))/gd,
//     gd = sqrt(V^T B^T ddev B V) of the element's CURRENT velocities;
//   * incompressible Boussinesq          (MomContEquation.inl:166-199): F factor rho (1 - alpha (T - Tr)), H factor
//     (1 - alpha (T - Tr)), T = N.T_e at the Gauss points;
//   * implicit heat equation + CG        (IncompNewton/HeatEquation.inl:227-419, solveWithGuess :129-135): scalar system
//     A = M(cv rho) + dt L(k), b = M theta_prev, Dirichlet / free-node rows, Jacobi-preconditioned conjugate gradients;
//   * explicit heat equation (BoussinesqWC, WCompNewton/HeatEquation.inl:154-298), the buoyancy factor of the explicit
//     momentum equation (WCompNewton/MomEquation.inl:105-112) and the thermal diffusivity in the CFL step
//     (WCompNewton/Solver.cpp:214-216) live in wc.cu next to the kernels they extend; the nodal pass is here.
// These are the GENERAL assembly kernels: an element pass writes one record per element (grad N, V, tau, viscosity), one
// thread per node block (i, j) gathers the elements around edge (i, j) in ascending element index and evaluates the closed
// forms of SURVEY appendix A with the per-element factors; a last pass per node sums the right-hand side and applies the
// boundary conditions.  The fractional-step systems and the device-driven CG are at the end of the file.  Simple and deterministic, not tuned: the tuned
// kernel (pspg.cu) covers the constant-factor problem of the named configurations; problems with these factors take
// this path (about 4x slower at C4).
#include "common.cuh"
#include "spmv.cuh"

namespace {

__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }

template <int DIM> struct GElem {
    double g[DIM + 1][DIM];  // grad N_m
    double V;
};
// geometry of element e from the nodal records (Element.cpp:15-135, MatricesBuilder.inl:93-127)
template <int DIM>
__device__ __forceinline__ void elemGeo(const double* __restrict__ X4, const int (&nd)[DIM + 1], GElem<DIM>& G) {
    constexpr int NPE = DIM + 1;
    constexpr double REF = (DIM == 2) ? 0.5 : 0.16666666666666666666666666666667;
    double px[NPE][DIM];
#pragma unroll
    for (int m = 0; m < NPE; ++m)
#pragma unroll
        for (int d = 0; d < DIM; ++d) px[m][d] = X4[(size_t)nd[m] * 4 + d];
    double J[DIM][DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d)
#pragma unroll
        for (int m = 0; m < DIM; ++m) J[d][m] = px[m + 1][d] - px[0][d];
    double det, inv[DIM][DIM];
    if constexpr (DIM == 2) {
        det = J[0][0] * J[1][1] - J[1][0] * J[0][1];
        const double rd = 1.0 / det;
        inv[0][0] = J[1][1] * rd, inv[0][1] = -J[0][1] * rd, inv[1][0] = -J[1][0] * rd, inv[1][1] = J[0][0] * rd;
    } else {
        det = J[0][0] * J[1][1] * J[2][2] + J[0][1] * J[1][2] * J[2][0] + J[0][2] * J[1][0] * J[2][1] -
              J[2][0] * J[1][1] * J[0][2] - J[2][1] * J[1][2] * J[0][0] - J[2][2] * J[1][0] * J[0][1];
        const double rd = 1.0 / det;
        inv[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) * rd;
        inv[0][1] = (J[2][1] * J[0][2] - J[2][2] * J[0][1]) * rd;
        inv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * rd;
        inv[1][0] = (J[2][0] * J[1][2] - J[1][0] * J[2][2]) * rd;
        inv[1][1] = (J[0][0] * J[2][2] - J[2][0] * J[0][2]) * rd;
        inv[1][2] = (J[1][0] * J[0][2] - J[0][0] * J[1][2]) * rd;
        inv[2][0] = (J[1][0] * J[2][1] - J[2][0] * J[1][1]) * rd;
        inv[2][1] = (J[2][0] * J[0][1] - J[0][0] * J[2][1]) * rd;
        inv[2][2] = (J[0][0] * J[1][1] - J[1][0] * J[0][1]) * rd;
    }
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
        double s = -inv[0][d];
#pragma unroll
        for (int m = 1; m < DIM; ++m) s -= inv[m][d];
        G.g[0][d] = s;
#pragma unroll
        for (int m = 0; m < DIM; ++m) G.g[m + 1][d] = inv[m][d];
    }
    G.V = det * REF;
}
template <int DIM> __device__ __forceinline__ void loadConn(const int* __restrict__ conn, int e, int (&nd)[DIM + 1]) {
#pragma unroll
    for (int m = 0; m < DIM + 1; ++m) nd[m] = conn[(size_t)e * (DIM + 1) + m];
}
// Gauss rule of the element integrals (Mesh.cpp:375-399, 443-459): at point g the shape functions are GA for node g and GB
// for the others, weight 1/(dim+1)
template <int DIM> struct Gauss {
    static constexpr double GA = (DIM == 3) ? 0.585410196624968 : 0.66666666666666666667;
    static constexpr double GB = (DIM == 3) ? 0.138196601125011 : 0.16666666666666666667;
    static constexpr double W = 1.0 / (DIM + 1);
};

struct GenArgs {
    const int* conn;
    const int* n2ePtr;
    const int* n2e;
    const unsigned* blkMask;
    const int* nbrPtr;
    const int* nbr;
    const int* diagSlot;
    const uint8_t* flags;
    const uint8_t* dirMask;
    const double* dirVal4;
    const double* X4;
    const double* V4;    // current velocities (tau, Bingham shear rate)
    const double* VP4;   // previous velocities
    const double* T;     // nodal temperature or null
    const double* fst4;
    double* Aval;
    double* b;
    double* dinv;
    int nRows, CH;
    double rho, mu, dt, body[3];
    double tau0, mReg;   // Bingham (bingham != 0)
    double alpha, Tr;    // Boussinesq (T != null)
    int bingham;
    // mode 1: velocity-prediction system of the fractional-step solver (MomContEquationFracStep.inl:8-217): M/dt + K on the
    // velocity rows, identity pressure rows, b = F + M/dt v_prev + gammaFS D^T p_prev
    int mode = 0;
    double gammaFS = 0.0;
    const double* pPrev = nullptr;
    // per-element record written once by k_gen_elem (grad N, V, tau, viscosity): the block / right-hand-side passes read it
    // instead of re-evaluating the element for every node block it touches (16x at C4)
    double* rec = nullptr;
    int nElems = 0;
};
template <int DIM> struct GenRec {
    static constexpr int STRIDE = (DIM == 3) ? 16 : 10;  // doubles: (DIM+1) DIM gradients, V, tau, mu (+ pad: 16-byte multiples)
};
template <int DIM>
__device__ __forceinline__ void loadRec(const double* __restrict__ rec, int e, GElem<DIM>& G, double& tau, double& muE) {
    const double2* r = reinterpret_cast<const double2*>(rec + (size_t)e * GenRec<DIM>::STRIDE);
    double v[GenRec<DIM>::STRIDE];
#pragma unroll
    for (int k = 0; k < GenRec<DIM>::STRIDE / 2; ++k) {
        const double2 t = __ldg(r + k);
        v[2 * k] = t.x, v[2 * k + 1] = t.y;
    }
#pragma unroll
    for (int m = 0; m < DIM + 1; ++m)
#pragma unroll
        for (int d = 0; d < DIM; ++d) G.g[m][d] = v[m * DIM + d];
    G.V = v[(DIM + 1) * DIM], tau = v[(DIM + 1) * DIM + 1], muE = v[(DIM + 1) * DIM + 2];
}

// tau (PSPG.inl:238-259) and the element viscosity
template <int DIM>
__device__ __forceinline__ void elemTauMu(const GenArgs& a, const int (&nd)[DIM + 1], const GElem<DIM>& G, double& tau, double& muE) {
    constexpr int NPE = DIM + 1;
    constexpr double REF = (DIM == 2) ? 0.5 : 0.16666666666666666666666666666667;
    double usum = 0, L[DIM][DIM];  // L[a][c] = d v_a / d x_c
#pragma unroll
    for (int aa = 0; aa < DIM; ++aa)
#pragma unroll
        for (int c = 0; c < DIM; ++c) L[aa][c] = 0;
#pragma unroll
    for (int m = 0; m < NPE; ++m) {
        double v[DIM], s = 0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
            v[d] = a.V4[(size_t)nd[m] * 4 + d];
            s += v[d] * v[d];
        }
        usum += sqrt(s);
#pragma unroll
        for (int aa = 0; aa < DIM; ++aa)
#pragma unroll
            for (int c = 0; c < DIM; ++c) L[aa][c] += v[aa] * G.g[m][c];
    }
    const double det = G.V / REF;
    const double h2 = REF * det / 3.14159265358979323846;  // reference hazard 7: the 2-D formula also in 3-D
    const double U = usum / NPE;
    const double t1 = 2.0 / a.dt, t3 = 4.0 * a.mu / (h2 * a.rho);
    tau = 1.0 / sqrt(t1 * t1 + 4.0 * U * U / h2 + 9.0 * t3 * t3);
    muE = a.mu;
    if (a.bingham) {
        double gd2 = 0;  // V^T B^T ddev B V with ddev = diag(2,..,1,..): 2 sum eps_dd^2 + sum engineering shear^2
#pragma unroll
        for (int d = 0; d < DIM; ++d) gd2 += 2.0 * L[d][d] * L[d][d];
#pragma unroll
        for (int p = 0; p < DIM; ++p)
#pragma unroll
            for (int q = p + 1; q < DIM; ++q) {
                const double sh = L[p][q] + L[q][p];
                gd2 += sh * sh;
            }
        const double gd = sqrt(gd2);
        muE += (gd < 1e-15) ? a.tau0 * a.mReg : a.tau0 * (1.0 - exp(-a.mReg * gd)) / gd;
    }
}

// element pass: geometry, tau and the element viscosity once per element
template <int DIM>
__global__ void __launch_bounds__(128) k_gen_elem(const GenArgs a) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= a.nElems) return;
    int nd[DIM + 1];
    loadConn<DIM>(a.conn, e, nd);
    GElem<DIM> G;
    elemGeo<DIM>(a.X4, nd, G);
    double tau, muE;
    elemTauMu<DIM>(a, nd, G, tau, muE);
    double v[GenRec<DIM>::STRIDE];
#pragma unroll
    for (int k = 0; k < GenRec<DIM>::STRIDE; ++k) v[k] = 0.0;
#pragma unroll
    for (int m = 0; m < DIM + 1; ++m)
#pragma unroll
        for (int d = 0; d < DIM; ++d) v[m * DIM + d] = G.g[m][d];
    v[(DIM + 1) * DIM] = G.V, v[(DIM + 1) * DIM + 1] = tau, v[(DIM + 1) * DIM + 2] = muE;
    double2* r = reinterpret_cast<double2*>(a.rec + (size_t)e * GenRec<DIM>::STRIDE);
#pragma unroll
    for (int k = 0; k < GenRec<DIM>::STRIDE / 2; ++k) r[k] = make_double2(v[2 * k], v[2 * k + 1]);
}

// block (i, slot) of the PSPG matrix: every element around edge (i, j), ascending element index, row masks applied
template <int DIM>
__global__ void __launch_bounds__(128) k_gen_blocks(const GenArgs a) {
    constexpr int NPE = DIM + 1, BS = DIM + 1;
    constexpr double PHI = 1.0 / ((DIM + 1) * (DIM + 2));
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    // thread -> (node, slot) by a search in nbrPtr restricted to the owned rows
    const int64_t nBlk = a.nbrPtr[a.nRows];
    if (t >= nBlk) return;
    int lo = 0, hi = a.nRows - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (a.nbrPtr[mid] <= t) lo = mid;
        else hi = mid - 1;
    }
    const int i = lo, nb0 = a.nbrPtr[i], s = (int)(t - nb0);
    const int j = a.nbr[nb0 + s], eb = a.n2ePtr[i], ne = a.n2ePtr[i + 1] - eb;
    const uint8_t fl = a.flags[i];
    const bool isBound = fl & PFEM_NODE_BOUND, isFree = fl & PFEM_NODE_FREE;
    const bool maskV = isBound || isFree, maskP = isFree;
    double acc[BS][BS];
#pragma unroll
    for (int r = 0; r < BS; ++r)
#pragma unroll
        for (int c = 0; c < BS; ++c) acc[r][c] = 0;
    for (int k = 0; k < ne; ++k) {
        if (!((a.blkMask[(size_t)(nb0 + s) * a.CH + (k >> 5)] >> (k & 31)) & 1u)) continue;
        const int e = a.n2e[eb + k];
        int nd[NPE];
        loadConn<DIM>(a.conn, e, nd);
        int li = 0, lj = 0;
#pragma unroll
        for (int m = 0; m < NPE; ++m) {
            li = (nd[m] == i) ? m : li;
            lj = (nd[m] == j) ? m : lj;
        }
        GElem<DIM> G;
        double tau, muE;
        loadRec<DIM>(a.rec, e, G, tau, muE);
        double gi[DIM], gj[DIM], dot = 0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
            gi[d] = G.g[0][d], gj[d] = G.g[0][d];
#pragma unroll
            for (int m = 1; m < NPE; ++m) {
                gi[d] = (li == m) ? G.g[m][d] : gi[d];
                gj[d] = (lj == m) ? G.g[m][d] : gj[d];
            }
            dot += gi[d] * gj[d];
        }
        const double V = G.V;
        const double cm = a.rho * V * PHI * (i == j ? 2.0 : 1.0) / a.dt;
#pragma unroll
        for (int aa = 0; aa < DIM; ++aa) {
#pragma unroll
            for (int c = 0; c < DIM; ++c) acc[aa][c] += muE * V * gi[c] * gj[aa];
            acc[aa][aa] += cm + muE * V * dot;
            if (a.mode == 0) {
                acc[aa][DIM] -= (V / NPE) * gi[aa];
                acc[DIM][aa] += (V / NPE) * ((tau / a.dt) * gi[aa] + gj[aa]);
            }
        }
        if (a.mode == 0) acc[DIM][DIM] += tau * (V / a.rho) * dot;
    }
    const bool diag = (s == a.diagSlot[i]);
#pragma unroll
    for (int r = 0; r < BS; ++r) {
        const bool rowMasked = (r < DIM) ? maskV : (maskP || a.mode == 1);
#pragma unroll
        for (int c = 0; c < BS; ++c) {
            double v = acc[r][c];
            if (rowMasked) v = (diag && r == c) ? 1.0 : 0.0;
            a.Aval[((size_t)nb0 + s) * BS * BS + r * BS + c] = v;
        }
    }
}

// right-hand side rows of node i + m_applyBCPSPG (PSPG.inl:140-144, 149-235) on the blocks k_gen_blocks wrote
template <int DIM>
__global__ void __launch_bounds__(128) k_gen_rhs_bc(const GenArgs a) {
    constexpr int NPE = DIM + 1, BS = DIM + 1;
    constexpr double PHI = 1.0 / ((DIM + 1) * (DIM + 2));
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.nRows) return;
    const int eb = a.n2ePtr[i], ne = a.n2ePtr[i + 1] - eb;
    double bi[BS];
#pragma unroll
    for (int r = 0; r < BS; ++r) bi[r] = 0;
    for (int k = 0; k < ne; ++k) {
        const int e = a.n2e[eb + k];
        int nd[NPE];
        loadConn<DIM>(a.conn, e, nd);
        int li = 0;
#pragma unroll
        for (int m = 1; m < NPE; ++m) li = (nd[m] == i) ? m : li;
        GElem<DIM> G;
        double tau, muE;
        loadRec<DIM>(a.rec, e, G, tau, muE);
        double gi[DIM], sv[DIM], vpi[DIM];
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
            gi[d] = G.g[0][d];
#pragma unroll
            for (int m = 1; m < NPE; ++m) gi[d] = (li == m) ? G.g[m][d] : gi[d];
            sv[d] = 0;
#pragma unroll
            for (int m = 0; m < NPE; ++m) sv[d] += a.VP4[(size_t)nd[m] * 4 + d];
            vpi[d] = a.VP4[(size_t)i * 4 + d];
        }
        // Gauss sums of the F and H factors (MomContEquation.inl:166-207, 218-222)
        double fF = a.rho / NPE, fH = 1.0;  // sum_g w f N_g[i]  |  sum_g w f
        if (a.T) {
            double Te[NPE], sT = 0;
#pragma unroll
            for (int m = 0; m < NPE; ++m) {
                Te[m] = a.T[nd[m]];
                sT += Te[m];
            }
            fF = 0, fH = 0;
#pragma unroll
            for (int g = 0; g < NPE; ++g) {
                const double Tg = Gauss<DIM>::GB * sT + (Gauss<DIM>::GA - Gauss<DIM>::GB) * Te[g];
                const double f = 1.0 - a.alpha * (Tg - a.Tr);
                fF += Gauss<DIM>::W * (a.rho * f) * (g == li ? Gauss<DIM>::GA : Gauss<DIM>::GB);
                fH += Gauss<DIM>::W * f;
            }
        }
        const double V = G.V, cmass = a.rho * V * PHI / a.dt;
        double gb = 0, gs = 0;
        double sumP = 0;
        if (a.mode == 1) {
#pragma unroll
            for (int m = 0; m < NPE; ++m) sumP += a.pPrev[nd[m]];
        }
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
            gb += gi[d] * a.body[d];
            gs += gi[d] * sv[d];
            bi[d] += V * a.body[d] * fF + cmass * (vpi[d] + sv[d]);
            if (a.mode == 1) bi[d] += a.gammaFS * (V / NPE) * gi[d] * sumP;  // gammaFS D^T p_prev (FracStep.inl:61, 146-147)
        }
        if (a.mode == 0) bi[DIM] += tau * V * gb * fH + (tau / a.dt) * (V / NPE) * gs;
    }
    const uint8_t fl = a.flags[i];
    const bool isBound = fl & PFEM_NODE_BOUND, isFree = fl & PFEM_NODE_FREE;
    if (a.mode == 1 && (isBound || isFree)) {  // FracStep.inl:138-150: nothing is assembled into the rows of bound / free nodes
#pragma unroll
        for (int r = 0; r < BS; ++r) bi[r] = 0;
    }
    if (a.fst4)
#pragma unroll
        for (int d = 0; d < DIM; ++d) bi[d] += a.fst4[(size_t)i * 4 + d];
    // Dirichlet column elimination (PSPG.inl:216-228): b_row -= A(row, col) g_D ; A(row, col) = 0 for row != col
    const int nb0 = a.nbrPtr[i], nb = a.nbrPtr[i + 1] - nb0, si = a.diagSlot[i];
    for (int s = 0; s < nb; ++s) {
        const int j = a.nbr[nb0 + s];
        if (!(a.dirMask[j] && (a.flags[j] & PFEM_NODE_BOUND))) continue;
        double* blk = a.Aval + ((size_t)nb0 + s) * BS * BS;
#pragma unroll
        for (int cc = 0; cc < DIM; ++cc) {
            const double gv = a.dirVal4[(size_t)j * 4 + cc];
#pragma unroll
            for (int r = 0; r < BS; ++r) {
                if (s == si && r == cc) continue;
                bi[r] -= blk[r * BS + cc] * gv;
                blk[r * BS + cc] = 0.0;
            }
        }
    }
#pragma unroll
    for (int r = 0; r < BS; ++r) {
        double bv = bi[r];
        if (isFree) {
            if (r == DIM) bv = 0.0;
            else if (!isBound) bv = a.VP4[(size_t)i * 4 + r] + a.dt * a.body[r];
        }
        if (a.mode == 1 && r == DIM) bv = 0.0;
        if (isBound && a.dirMask[i] && r < DIM) bv = a.dirVal4[(size_t)i * 4 + r];
        a.b[(size_t)i * BS + r] = bv;
        const double d = a.Aval[((size_t)nb0 + si) * BS * BS + r * BS + r];
        a.dinv[(size_t)i * BS + r] = (d != 0.0) ? rsqrt(fabs(d)) : 1.0;
    }
}

// ---- implicit heat equation (scalar system on the node pattern) -------------------------------------------------------------
struct HeatArgs {
    const int* conn;
    const int* n2ePtr;
    const int* n2e;
    const unsigned* blkMask;
    const int* nbrPtr;
    const int* nbr;
    const int* diagSlot;
    const uint8_t* flags;
    const uint8_t* tMask;
    const double* tVal;
    const double* X4;
    const double* thetaPrev;
    double* A;     // one value per node block
    double* b;
    int nRows, CH;
    double rhoCv, k, dt;
};
template <int DIM> __global__ void __launch_bounds__(128) k_heat_blocks(const HeatArgs a) {
    constexpr int NPE = DIM + 1;
    constexpr double PHI = 1.0 / ((DIM + 1) * (DIM + 2));
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t nBlk = a.nbrPtr[a.nRows];
    if (t >= nBlk) return;
    int lo = 0, hi = a.nRows - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (a.nbrPtr[mid] <= t) lo = mid;
        else hi = mid - 1;
    }
    const int i = lo, nb0 = a.nbrPtr[i], s = (int)(t - nb0);
    const int j = a.nbr[nb0 + s], eb = a.n2ePtr[i], ne = a.n2ePtr[i + 1] - eb;
    const bool masked = a.tMask[i] || (a.flags[i] & PFEM_NODE_FREE);  // HeatEquation.inl:268-280, 300-308
    double acc = 0;
    if (!masked)
        for (int k = 0; k < ne; ++k) {
            if (!((a.blkMask[(size_t)(nb0 + s) * a.CH + (k >> 5)] >> (k & 31)) & 1u)) continue;
            const int e = a.n2e[eb + k];
            int nd[NPE];
            loadConn<DIM>(a.conn, e, nd);
            int li = 0, lj = 0;
#pragma unroll
            for (int m = 0; m < NPE; ++m) {
                li = (nd[m] == i) ? m : li;
                lj = (nd[m] == j) ? m : lj;
            }
            GElem<DIM> G;
            elemGeo<DIM>(a.X4, nd, G);
            double dot = 0;
#pragma unroll
            for (int d = 0; d < DIM; ++d) {
                double gi = G.g[0][d], gj = G.g[0][d];
#pragma unroll
                for (int m = 1; m < NPE; ++m) {
                    gi = (li == m) ? G.g[m][d] : gi;
                    gj = (lj == m) ? G.g[m][d] : gj;
                }
                dot += gi * gj;
            }
            acc += a.rhoCv * G.V * PHI * (i == j ? 2.0 : 1.0) + a.dt * (a.k * G.V * dot);  // M + dt L
        }
    else
        acc = (s == a.diagSlot[i]) ? 1.0 : 0.0;
    a.A[(size_t)nb0 + s] = acc;
}
template <int DIM> __global__ void __launch_bounds__(128) k_heat_rhs_bc(const HeatArgs a) {
    constexpr int NPE = DIM + 1;
    constexpr double PHI = 1.0 / ((DIM + 1) * (DIM + 2));
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.nRows) return;
    const int eb = a.n2ePtr[i], ne = a.n2ePtr[i + 1] - eb;
    double bi = 0;
    for (int k = 0; k < ne; ++k) {  // b = M theta_prev: all rows, masked or not (HeatEquation.inl:283-288, 313-317)
        const int e = a.n2e[eb + k];
        int nd[NPE];
        loadConn<DIM>(a.conn, e, nd);
        GElem<DIM> G;
        elemGeo<DIM>(a.X4, nd, G);
        double st = 0;
#pragma unroll
        for (int m = 0; m < NPE; ++m) st += a.thetaPrev[nd[m]];
        bi += a.rhoCv * G.V * PHI * (a.thetaPrev[i] + st);
    }
    const bool free_ = a.flags[i] & PFEM_NODE_FREE, tbc = a.tMask[i] != 0;
    const int nb0 = a.nbrPtr[i], nb = a.nbrPtr[i + 1] - nb0, si = a.diagSlot[i];
    for (int s = 0; s < nb; ++s) {  // Dirichlet columns (HeatEquation.inl:396-408)
        const int j = a.nbr[nb0 + s];
        if (s == si || !a.tMask[j]) continue;
        bi -= a.A[(size_t)nb0 + s] * a.tVal[j];
        a.A[(size_t)nb0 + s] = 0.0;
    }
    if (free_ && !tbc) bi = a.thetaPrev[i];
    else if (tbc) bi = a.tVal[i];
    a.b[i] = bi;
}

__global__ void k_nodal_copy(int n, const double* __restrict__ src, double* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}

// ---- device-driven conjugate gradients (round 2) ---------------------------------------------------------------------------------
// Jacobi-preconditioned CG for the scalar node-pattern systems (heat equation, fractional-step pressure / velocity correction)
// and the node-block velocity-prediction system.  The first version launched 11 kernels and synchronised the host once per
// iteration (121 us per iteration on a 343 k-row scalar system whose SpMV takes 15 us).  Here an iteration is 3 launches (4 for the node-block system), every dot product
// goes through the ordered partial-sum bank (each consumer block re-reduces it: same bits everywhere, no atomics), the
// convergence test and the iteration count live on the device (Eigen's semantics: the pass that reaches the threshold is not
// counted) and the host polls every 8 iterations.
enum { CGD_RZ0 = 0, CGD_RZ1, CGD_RR, CGD_BB, CGD_THR, CGD_DONE, CGD_ITERS, CGD_BAD, CGD_COUNT = 8 };
enum { CGB_PAP = 0, CGB_RR, CGB_RZ, CGB_COUNT };
// y = A x on the scalar node-pattern matrix + partial dot (x, y) into bank slot CGB_PAP at [blockOffset + blockIdx]
__global__ void __launch_bounds__(RB_THREADS) k_cgd_spmv(int n, const int* __restrict__ ptr, const int* __restrict__ col,
                                                         const double* __restrict__ A, const double* __restrict__ x,
                                                         double* __restrict__ y, double* bank, int stride, int blockOffset,
                                                         const double* __restrict__ scal) {
    if (scal[CGD_DONE] != 0.0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double d = 0;
    if (i < n) {
        double s = 0;
        for (int k = ptr[i]; k < ptr[i + 1]; ++k) s += A[k] * x[col[k]];
        y[i] = s;
        d = s * x[i];
    }
    double v[1] = {d};
    __shared__ double sh[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const double t = warpSum(v[0]);
    if (lane == 0) sh[w] = t;
    __syncthreads();
    if (w == 0) {
        double u = lane < nw ? sh[lane] : 0.0;
        u = warpSum(u);
        if (lane == 0) bank[(size_t)CGB_PAP * stride + blockOffset + blockIdx.x] = u;
    }
}
// partial dot (a, b) into bank slot CGB_PAP (node-block system: the SpMV is the Krylov kernel, the dot follows)
__global__ void __launch_bounds__(RB_THREADS) k_cgd_dot(int n, const double* __restrict__ a, const double* __restrict__ b, double* bank,
                                                        int stride, const double* __restrict__ scal) {
    if (scal[CGD_DONE] != 0.0) return;
    double s = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s += a[i] * b[i];
    double v[1] = {s};
    const int slots[1] = {CGB_PAP};
    blockSumStore<1>(v, bank, stride, slots);
}
// r = b - Ax (Ax may be null: x = 0) ; z = dinv r ; p = z ; partials (r, r), (r, z), and (b, b) into CGB_PAP
__global__ void __launch_bounds__(RB_THREADS) k_cgd_init(int n, const double* __restrict__ b, const double* __restrict__ Ax,
                                                         const double* __restrict__ dinv, double* __restrict__ r,
                                                         double* __restrict__ z, double* __restrict__ p, double* bank, int stride) {
    double rr = 0, rz = 0, bb = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double bi = b[i], ri = Ax ? bi - Ax[i] : bi, zi = dinv[i] * ri;
        r[i] = ri, z[i] = zi, p[i] = zi;
        rr += ri * ri, rz += ri * zi, bb += bi * bi;
    }
    double v[3] = {bb, rr, rz};
    const int slots[3] = {CGB_PAP, CGB_RR, CGB_RZ};
    blockSumStore<3>(v, bank, stride, slots);
}
__global__ void __launch_bounds__(RB_THREADS) k_cgd_init_scal(double* scal, const double* bank, int stride, int nPart, double relTol) {
    const double bb = bankSum(bank, stride, CGB_PAP, nPart);
    const double rr = bankSum(bank, stride, CGB_RR, nPart);
    const double rz = bankSum(bank, stride, CGB_RZ, nPart);
    if (threadIdx.x == 0) {
        const double thr = fmax(relTol * relTol * bb, 2.2250738585072014e-308);
        scal[CGD_RZ0] = rz, scal[CGD_RZ1] = rz, scal[CGD_RR] = rr, scal[CGD_BB] = bb, scal[CGD_THR] = thr;
        scal[CGD_ITERS] = 0.0, scal[CGD_BAD] = (rr == rr) ? 0.0 : 1.0;
        scal[CGD_DONE] = (bb == 0.0 || rr < thr || !(rr == rr)) ? 1.0 : 0.0;
    }
}
// alpha = (r, z) / (p, Ap) ; x += alpha p ; r -= alpha Ap ; z = dinv r ; partials (r, r), (r, z)
__global__ void __launch_bounds__(RB_THREADS) k_cgd_xr(int n, int nPartPAp, const double* __restrict__ scal, int parity,
                                                       const double* __restrict__ p, const double* __restrict__ Ap,
                                                       const double* __restrict__ dinv, double* __restrict__ x, double* __restrict__ r,
                                                       double* __restrict__ z, double* bank, int stride) {
    if (scal[CGD_DONE] != 0.0) return;
    const double pAp = bankSum(bank, stride, CGB_PAP, nPartPAp);
    const double alpha = scal[CGD_RZ0 + parity] / pAp;
    double rr = 0, rz = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        x[i] += alpha * p[i];
        const double ri = r[i] - alpha * Ap[i], zi = dinv[i] * ri;
        r[i] = ri, z[i] = zi;
        rr += ri * ri, rz += ri * zi;
    }
    double v[2] = {rr, rz};
    const int slots[2] = {CGB_RR, CGB_RZ};
    blockSumStore<2>(v, bank, stride, slots);
}
// convergence test (Eigen: stop BEFORE counting the pass), beta = (r, z)_new / (r, z) ; p = z + beta p
__global__ void __launch_bounds__(RB_THREADS) k_cgd_p(int n, int nPart, double* scal, int parity, const double* __restrict__ z,
                                                      double* __restrict__ p, const double* bank, int stride) {
    if (scal[CGD_DONE] != 0.0) return;
    const double rr = bankSum(bank, stride, CGB_RR, nPart);
    const double rzNew = bankSum(bank, stride, CGB_RZ, nPart);
    const double rzOld = scal[CGD_RZ0 + parity];
    const bool bad = !(rr == rr), conv = rr < scal[CGD_THR];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        scal[CGD_RR] = rr;
        if (bad) scal[CGD_BAD] = 1.0;
        if (bad || conv) scal[CGD_DONE] = 1.0;
        else {
            scal[CGD_RZ0 + (parity ^ 1)] = rzNew;
            scal[CGD_ITERS] += 1.0;
        }
    }
    if (bad || conv) return;
    const double beta = rzNew / rzOld;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = z[i] + beta * p[i];
}

// explicit heat equation of BoussinesqWC, nodal gather (WCompNewton/HeatEquation.inl:154-298): lumped M with f = cv (N.rho),
// L with f = k, F = -dt L T + M T; free nodes keep T, Dirichlet nodes take g_T
template <int DIM>
__global__ void __launch_bounds__(128) k_wc_heat(int nRows, const int* __restrict__ conn, const int* __restrict__ n2ePtr,
                                                 const int* __restrict__ n2e, const uint8_t* __restrict__ flags,
                                                 const uint8_t* __restrict__ tMask, const double* __restrict__ tVal,
                                                 const double* __restrict__ X4, const double* __restrict__ V4, const double* __restrict__ T,
                                                 double* __restrict__ Tnew, double k, double cv, double dtVal, const double* __restrict__ dtPtr) {
    constexpr int NPE = DIM + 1;
    constexpr double PHI = 1.0 / ((DIM + 1) * (DIM + 2));
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nRows) return;
    const double dt = dtPtr ? *dtPtr : dtVal;
    double M = 0, F = 0;
    for (int kk = n2ePtr[i]; kk < n2ePtr[i + 1]; ++kk) {
        const int e = n2e[kk];
        int nd[NPE];
        loadConn<DIM>(conn, e, nd);
        int li = 0;
#pragma unroll
        for (int m = 1; m < NPE; ++m) li = (nd[m] == i) ? m : li;
        GElem<DIM> G;
        elemGeo<DIM>(X4, nd, G);
        double sumR = 0, gT[DIM], gi[DIM], lt = 0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) gT[d] = 0;
#pragma unroll
        for (int m = 0; m < NPE; ++m) {
            sumR += V4[(size_t)nd[m] * 4 + 3];
            const double Tm = T[nd[m]];
#pragma unroll
            for (int d = 0; d < DIM; ++d) gT[d] += G.g[m][d] * Tm;
        }
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
            gi[d] = G.g[0][d];
#pragma unroll
            for (int m = 1; m < NPE; ++m) gi[d] = (li == m) ? G.g[m][d] : gi[d];
            lt += gi[d] * gT[d];
        }
        const double lump = cv * G.V * PHI * (V4[(size_t)i * 4 + 3] + sumR);  // lumped cv (N.rho) mass
        F += -dt * (k * G.V * lt) + lump * T[i];
        M += lump;
    }
    const bool free_ = flags[i] & PFEM_NODE_FREE, tbc = tMask[i] != 0;
    double inv = 1.0 / M;
    if (free_ && !tbc) {
        F = T[i];
        inv = 1.0;
    } else if (tbc) {
        F = tVal[i];
        inv = 1.0;
    }
    Tnew[i] = inv * F;
}

// ---- fractional-step solver: pressure and velocity-correction right-hand sides ----------------------------------------------
// b_p (MomContEquationFracStep.inl:300-347 + m_applyBCPCorrStep :349-376): rows of nodes that are neither free nor on the free
// surface get -(rho/dt) (D vTilde)_i + gammaFS (L p_prev)_i summed over their elements in ascending order, the others 0.
// vT: vTilde as [d][nNodes]
template <int DIM>
__global__ void __launch_bounds__(128) k_fs_pcorr_rhs(int nRows, int nNodes, const int* __restrict__ conn, const int* __restrict__ n2ePtr,
                                                       const int* __restrict__ n2e, const uint8_t* __restrict__ flags,
                                                       const double* __restrict__ X4, const double* __restrict__ vT,
                                                       const double* __restrict__ pPrev, double rhoOverDt, double gammaFS,
                                                       double* __restrict__ b) {
    constexpr int NPE = DIM + 1;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nRows) return;
    const uint8_t fl = flags[i];
    double bi = 0;
    if (!(fl & (PFEM_NODE_FREE | PFEM_NODE_FREE_SURFACE))) {
        const int eb = n2ePtr[i], ne = n2ePtr[i + 1] - eb;
        for (int k = 0; k < ne; ++k) {
            const int e = n2e[eb + k];
            int nd[NPE];
            loadConn<DIM>(conn, e, nd);
            int li = 0;
#pragma unroll
            for (int m = 1; m < NPE; ++m) li = (nd[m] == i) ? m : li;
            GElem<DIM> G;
            elemGeo<DIM>(X4, nd, G);
            double gi[DIM], div = 0, lp = 0;
#pragma unroll
            for (int d = 0; d < DIM; ++d) {
                gi[d] = G.g[0][d];
#pragma unroll
                for (int m = 1; m < NPE; ++m) gi[d] = (li == m) ? G.g[m][d] : gi[d];
            }
#pragma unroll
            for (int m = 0; m < NPE; ++m) {
                double dot = 0;
#pragma unroll
                for (int d = 0; d < DIM; ++d) {
                    div += G.g[m][d] * vT[(size_t)d * nNodes + nd[m]];
                    dot += gi[d] * G.g[m][d];
                }
                lp += dot * pPrev[nd[m]];
            }
            bi += -rhoOverDt * (G.V / NPE) * div + gammaFS * G.V * lp;
        }
    }
    b[i] = bi;
}
// b_v (FracStep.inl:378-452): rows of nodes that are neither free nor bound get dt (D^T deltaP)_(i,d), the others 0; b as [d][nNodes]
template <int DIM>
__global__ void __launch_bounds__(128) k_fs_vcorr_rhs(int nRows, int nNodes, const int* __restrict__ conn, const int* __restrict__ n2ePtr,
                                                       const int* __restrict__ n2e, const uint8_t* __restrict__ flags,
                                                       const double* __restrict__ X4, const double* __restrict__ dP, double dt,
                                                       double* __restrict__ b) {
    constexpr int NPE = DIM + 1;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nRows) return;
    double bi[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) bi[d] = 0;
    if (!(flags[i] & (PFEM_NODE_FREE | PFEM_NODE_BOUND))) {
        const int eb = n2ePtr[i], ne = n2ePtr[i + 1] - eb;
        for (int k = 0; k < ne; ++k) {
            const int e = n2e[eb + k];
            int nd[NPE];
            loadConn<DIM>(conn, e, nd);
            int li = 0;
#pragma unroll
            for (int m = 1; m < NPE; ++m) li = (nd[m] == i) ? m : li;
            GElem<DIM> G;
            elemGeo<DIM>(X4, nd, G);
            double sp = 0;
#pragma unroll
            for (int m = 0; m < NPE; ++m) sp += dP[nd[m]];
#pragma unroll
            for (int d = 0; d < DIM; ++d) {
                double gi = G.g[0][d];
#pragma unroll
                for (int m = 1; m < NPE; ++m) gi = (li == m) ? G.g[m][d] : gi;
                bi[d] += dt * (G.V / NPE) * gi * sp;
            }
        }
    }
#pragma unroll
    for (int d = 0; d < DIM; ++d) b[(size_t)d * nNodes + i] = bi[d];
}
// m_applyBCPCorrStep / m_applyBCVStep walk the COLUMN of every masked node (Eigen's InnerIterator on a column-major matrix):
// diagonal 1, the other entries of the column 0 (FracStep.inl:359-371, 436-447); the rows were never assembled (identity)
__global__ void k_fs_zero_cols(int nRows, const int* __restrict__ nbrPtr, const int* __restrict__ nbr, const int* __restrict__ diagSlot,
                               const uint8_t* __restrict__ mask, const uint8_t* __restrict__ flags, double* __restrict__ A) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nRows) return;
    const int nb0 = nbrPtr[i], nb = nbrPtr[i + 1] - nb0, si = diagSlot[i];
    for (int s = 0; s < nb; ++s) {
        if (s == si) continue;
        const int j = nbr[nb0 + s];
        if (mask[j] || (flags[j] & PFEM_NODE_FREE)) A[(size_t)nb0 + s] = 0.0;
    }
}
__global__ void k_fs_bound_mask(int n, const uint8_t* __restrict__ flags, uint8_t* __restrict__ mask) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) mask[i] = (flags[i] & PFEM_NODE_BOUND) ? 1 : 0;
}
// 1 / diagonal of the node-block system (Eigen::DiagonalPreconditioner), internal dof order node*BS + r
__global__ void k_fs_block_dinv(int nNodes, int BS, const int* __restrict__ nbrPtr, const int* __restrict__ diagSlot,
                                const double* __restrict__ Aval, double* __restrict__ dinv) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nNodes * BS) return;
    const int i = t / BS, r = t % BS;
    const double d = Aval[((size_t)nbrPtr[i] + diagSlot[i]) * BS * BS + r * BS + r];
    dinv[t] = d != 0.0 ? 1.0 / d : 1.0;
}
__global__ void k_fs_scalar_dinv(int n, int comps, const int* __restrict__ ptr, const int* __restrict__ diag, const double* __restrict__ A,
                                 double* __restrict__ dinv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double d = A[ptr[i] + diag[i]];
    const double v = d != 0.0 ? 1.0 / d : 1.0;
    for (int cc = 0; cc < comps; ++cc) dinv[(size_t)cc * n + i] = v;
}
GenArgs makeGenArgs(pfem_ctx* c, const pfem_pspg_params& p) {
    GenArgs a;
    a.conn = c->conn.p, a.n2ePtr = c->n2ePtr.p, a.n2e = c->n2e.p, a.blkMask = c->blkMask.p, a.nbrPtr = c->nbrPtr.p, a.nbr = c->nbr.p;
    a.diagSlot = c->diagSlot.p, a.flags = c->flags.p, a.dirMask = c->dirMask.p, a.dirVal4 = c->dirVal4.p;
    a.X4 = c->X4.p, a.V4 = c->V4.p, a.VP4 = c->VP4.p, a.T = c->thermalOn ? c->Tn.p : nullptr;
    a.Aval = c->Aval.p, a.b = c->bvec.p, a.dinv = c->dinv.p, a.nRows = c->nRows, a.CH = c->maskWords;
    a.rho = p.rho, a.mu = p.mu, a.dt = p.dt;
    for (int d = 0; d < 3; ++d) a.body[d] = p.bodyForce[d];
    a.bingham = c->binghamOn ? 1 : 0, a.tau0 = c->binghamTau0, a.mReg = c->binghamM;
    a.alpha = c->thAlpha, a.Tr = c->thTr;
    a.fst4 = nullptr;
    c->genRec.reserve((size_t)std::max(c->nElems, 1) * 16 + 8);
    a.rec = c->genRec.p, a.nElems = c->nElems;
    return a;
}

}  // namespace

bool pspgNeedsGeneralPath(const pfem_ctx* c) { return c->binghamOn || c->thermalOn; }

// m_buildAbPSPG + m_applyBCPSPG with per-element factors (called by pspgAssemble when Bingham / Boussinesq are on)
void pspgAssembleGeneral(pfem_ctx* c, const pfem_pspg_params& p, const double* fst4) {
    PFEM_REQUIRE(!c->thermalOn || c->haveTemperature, PFEM_ERR_STATE, "pspg_assemble: Boussinesq factors need pfem_set_temperature");
    GenArgs a = makeGenArgs(c, p);
    a.fst4 = fst4;
    PhaseScope ph(c, "Assemble system");
    const int64_t nBlkRows = c->nBlocks;  // (partitioned mesh: blocks of the owned rows come first; the kernel bounds itself)
    if (c->dim == 2) {
        k_gen_elem<2><<<divUp(std::max(c->nElems, 1), 128), 128, 0, c->stream>>>(a);
        LAUNCH_CHECK(c);
        k_gen_blocks<2><<<divUp(nBlkRows, 128), 128, 0, c->stream>>>(a);
        LAUNCH_CHECK(c);
        k_gen_rhs_bc<2><<<divUp(c->nRows, 128), 128, 0, c->stream>>>(a);
    } else {
        k_gen_elem<3><<<divUp(std::max(c->nElems, 1), 128), 128, 0, c->stream>>>(a);
        LAUNCH_CHECK(c);
        k_gen_blocks<3><<<divUp(nBlkRows, 128), 128, 0, c->stream>>>(a);
        LAUNCH_CHECK(c);
        k_gen_rhs_bc<3><<<divUp(c->nRows, 128), 128, 0, c->stream>>>(a);
    }
    LAUNCH_CHECK(c);
}

void thermalSetTemperature(pfem_ctx* c, const double* T) {
    PFEM_REQUIRE(c->haveTopology && T, PFEM_ERR_STATE, "set_temperature: topology missing or null");
    c->Tn.reserve((size_t)c->nNodes + 4);
    c->Tnb.reserve((size_t)c->nNodes + 4);
    CUDA_CHECK(cudaMemcpyAsync(c->Tn.p, T, (size_t)c->nNodes * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->haveTemperature = true;
}
void thermalGetTemperature(pfem_ctx* c, double* T) {
    PFEM_REQUIRE(c->haveTemperature && T, PFEM_ERR_STATE, "get_temperature: no temperature on the device");
    CUDA_CHECK(cudaMemcpyAsync(T, c->Tn.p, (size_t)c->nNodes * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
}
void thermalSetBc(pfem_ctx* c, const uint8_t* mask, const double* values) {
    PFEM_REQUIRE(c->haveTopology && mask && values, PFEM_ERR_STATE, "set_temperature_bc: topology missing or null");
    c->tMask.reserve((size_t)c->nNodes + 4);
    c->tVal.reserve((size_t)c->nNodes + 4);
    CUDA_CHECK(cudaMemcpyAsync(c->tMask.p, mask, (size_t)c->nNodes, cudaMemcpyHostToDevice, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(c->tVal.p, values, (size_t)c->nNodes * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->haveTemperatureBc = true;
}
static void needThermalBc(pfem_ctx* c) {
    if (c->haveTemperatureBc) return;  // no temperature BC given: none anywhere
    c->tMask.reserve((size_t)c->nNodes + 4);
    c->tVal.reserve((size_t)c->nNodes + 4);
    CUDA_CHECK(cudaMemsetAsync(c->tMask.p, 0, (size_t)c->nNodes, c->stream));
    CUDA_CHECK(cudaMemsetAsync(c->tVal.p, 0, (size_t)c->nNodes * sizeof(double), c->stream));
    c->haveTemperatureBc = true;
}

void thermalPrepare(pfem_ctx* c) {  // allocations / defaults outside any graph capture
    PFEM_REQUIRE(c->haveTemperature, PFEM_ERR_STATE, "wc_step (Boussinesq): call pfem_set_temperature first");
    needThermalBc(c);
    c->Tnb.reserve((size_t)c->nNodes + 4);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

// explicit heat step of m_solveBoussinesqWC (Solver.cpp:301-303): on the moved mesh, with the density of the previous step
void thermalWcHeat(pfem_ctx* c, double dt, const double* dtPtr) {
    PFEM_REQUIRE(c->haveTemperature, PFEM_ERR_STATE, "wc_step (Boussinesq): call pfem_set_temperature first");
    needThermalBc(c);
    PhaseScope ph(c, "Solving heat eq");
    if (c->dim == 2)
        k_wc_heat<2><<<divUp(c->nRows, 128), 128, 0, c->stream>>>(c->nRows, c->conn.p, c->n2ePtr.p, c->n2e.p, c->flags.p, c->tMask.p,
                                                                c->tVal.p, c->X4.p, c->V4.p, c->Tn.p, c->Tnb.p, c->thK, c->thCv, dt, dtPtr);
    else
        k_wc_heat<3><<<divUp(c->nRows, 128), 128, 0, c->stream>>>(c->nRows, c->conn.p, c->n2ePtr.p, c->n2e.p, c->flags.p, c->tMask.p,
                                                                c->tVal.p, c->X4.p, c->V4.p, c->Tn.p, c->Tnb.p, c->thK, c->thCv, dt, dtPtr);
    LAUNCH_CHECK(c);
    if (c->nRanks > 1) commHalo(c, c->Tnb.p, nullptr, 1);
    k_nodal_copy<<<divUp(c->nNodes, 256), 256, 0, c->stream>>>(c->nNodes, c->Tnb.p, c->Tn.p);
    LAUNCH_CHECK(c);
}

namespace {
// bank row length: the vector kernels write <= smCount * 4 partials, a scalar SpMV over `comps` components comps * ceil(N / 256)
int cgdStride(pfem_ctx* c, int n) { return divUp(n, RB_THREADS) + c->smCount * 4 + 16; }
// Jacobi-preconditioned CG driven from the device (kernels k_cgd_*).  spmvDot(p, Ap) must compute Ap = A p AND leave the
// partial sums of (p, Ap) in bank slot CGB_PAP, returning how many partials it wrote; x holds the initial guess on entry when
// withGuess (then Ax0 is computed through spmvDot first).  Returns the status; iterations as Eigen counts them.
template <class SpmvDot>
int cgDevice(pfem_ctx* c, int n, SpmvDot spmvDot, const double* dinv, const double* b, double* x, bool withGuess, double relTol,
             int maxIter, int* itersOut, double* relResOut) {
    const int grid = std::max(1, std::min(c->smCount * 4, divUp(n, RB_THREADS)));
    for (auto* v : {&c->cgR, &c->cgZ, &c->cgP, &c->cgAp}) v->reserve((size_t)n + 8);
    const int stride = cgdStride(c, n);
    c->gmBank.reserve((size_t)CGB_COUNT * stride);
    c->gmS.reserve(CGD_COUNT + 8);
    if (!c->hScal) CUDA_CHECK(cudaMallocHost(&c->hScal, SC_COUNT * sizeof(double)));
    double* S = c->gmS.p;
    double* bank = c->gmBank.p;
    CUDA_CHECK(cudaMemsetAsync(S, 0, CGD_COUNT * sizeof(double), c->stream));
    const double* Ax = nullptr;
    if (withGuess) {
        spmvDot(x, c->cgAp.p);
        Ax = c->cgAp.p;
    } else
        CUDA_CHECK(cudaMemsetAsync(x, 0, (size_t)n * sizeof(double), c->stream));
    k_cgd_init<<<grid, RB_THREADS, 0, c->stream>>>(n, b, Ax, dinv, c->cgR.p, c->cgZ.p, c->cgP.p, bank, stride);
    LAUNCH_CHECK(c);
    k_cgd_init_scal<<<1, RB_THREADS, 0, c->stream>>>(S, bank, stride, grid, relTol);
    LAUNCH_CHECK(c);
    auto poll = [&]() {
        CUDA_CHECK(cudaMemcpyAsync(c->hScal, S, CGD_COUNT * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        return c->hScal[CGD_DONE] != 0.0;
    };
    bool done = poll();
    int launched = 0;
    while (!done && launched < maxIter) {
        const int batch = std::min(8, maxIter - launched);
        for (int k = 0; k < batch; ++k, ++launched) {
            const int parity = launched & 1;
            const int nPartPAp = spmvDot(c->cgP.p, c->cgAp.p);
            k_cgd_xr<<<grid, RB_THREADS, 0, c->stream>>>(n, nPartPAp, S, parity, c->cgP.p, c->cgAp.p, dinv, x, c->cgR.p, c->cgZ.p, bank, stride);
            LAUNCH_CHECK(c);
            k_cgd_p<<<grid, RB_THREADS, 0, c->stream>>>(n, grid, S, parity, c->cgZ.p, c->cgP.p, bank, stride);
            LAUNCH_CHECK(c);
        }
        done = poll();
    }
    const double bb = c->hScal[CGD_BB], rr = c->hScal[CGD_RR];
    if (itersOut) *itersOut = (int)c->hScal[CGD_ITERS];
    if (relResOut) *relResOut = bb > 0 ? sqrt(rr / bb) : 0.0;
    if (c->hScal[CGD_BAD] != 0.0) return PFEM_NAN;
    if (bb == 0.0) {  // Eigen: x = 0 for b = 0
        CUDA_CHECK(cudaMemsetAsync(x, 0, (size_t)n * sizeof(double), c->stream));
        return PFEM_OK;
    }
    return rr < c->hScal[CGD_THR] ? PFEM_OK : PFEM_NOT_CONVERGED;
}
// scalar node-pattern matrix applied to `comps` components stored one after the other
struct ScalarSpmvDot {
    pfem_ctx* c;
    int N, comps, stride;
    int operator()(const double* x, double* y) const {
        const int g = divUp(N, RB_THREADS);
        for (int cc = 0; cc < comps; ++cc) {
            k_cgd_spmv<<<g, RB_THREADS, 0, c->stream>>>(N, c->nbrPtr.p, c->nbr.p, c->hA.p, x + (size_t)cc * N, y + (size_t)cc * N,
                                                      c->gmBank.p, stride, cc * g, c->gmS.p);
            LAUNCH_CHECK(c);
        }
        return comps * g;
    }
};
}  // namespace

// HeatEqIncompNewton::m_buildAb + m_applyBC (IncompNewton/HeatEquation.inl:227-412, no flux facet terms)
void heatAssemble(pfem_ctx* c, double rho, double cv, double k, double dt, const double* thetaPrevHost) {
    PFEM_REQUIRE(c->haveTopology && c->havePositions, PFEM_ERR_STATE, "heat_assemble: topology/positions missing");
    PFEM_REQUIRE(c->nRanks == 1, PFEM_ERR_STATE, "heat_assemble: single-GPU contexts");
    PFEM_REQUIRE(thetaPrevHost && dt > 0 && rho > 0 && cv > 0, PFEM_ERR_INVALID, "heat_assemble: bad arguments");
    needThermalBc(c);
    const int n = c->nNodes;
    c->hA.reserve((size_t)c->nBlocks + 4);
    c->hb.reserve((size_t)n + 4);
    c->hTheta.reserve((size_t)n + 4);
    CUDA_CHECK(cudaMemcpyAsync(c->hTheta.p, thetaPrevHost, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    HeatArgs a;
    a.conn = c->conn.p, a.n2ePtr = c->n2ePtr.p, a.n2e = c->n2e.p, a.blkMask = c->blkMask.p, a.nbrPtr = c->nbrPtr.p, a.nbr = c->nbr.p;
    a.diagSlot = c->diagSlot.p, a.flags = c->flags.p, a.tMask = c->tMask.p, a.tVal = c->tVal.p, a.X4 = c->X4.p;
    a.thetaPrev = c->hTheta.p, a.A = c->hA.p, a.b = c->hb.p, a.nRows = n, a.CH = c->maskWords;
    a.rhoCv = cv * rho, a.k = k, a.dt = dt;
    c->heatMask = c->tMask.p;
    PhaseScope ph(c, "Assemble heat system");
    if (c->dim == 2) {
        k_heat_blocks<2><<<divUp(c->nBlocks, 128), 128, 0, c->stream>>>(a);
        LAUNCH_CHECK(c);
        k_heat_rhs_bc<2><<<divUp(n, 128), 128, 0, c->stream>>>(a);
    } else {
        k_heat_blocks<3><<<divUp(c->nBlocks, 128), 128, 0, c->stream>>>(a);
        LAUNCH_CHECK(c);
        k_heat_rhs_bc<3><<<divUp(n, 128), 128, 0, c->stream>>>(a);
    }
    LAUNCH_CHECK(c);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->haveHeatSystem = true;
}

// Eigen::ConjugateGradient<..., Lower|Upper, DiagonalPreconditioner>::solveWithGuess (IncompNewton/HeatEquation.inl:129-135):
// x0 = the temperature on the device (pfem_set_temperature) or zero; stops at ||r|| <= relTol ||b||
int heatSolve(pfem_ctx* c, double relTol, int maxIter, double* Tout, int* itersOut, double* relResOut) {
    PFEM_REQUIRE(c->haveHeatSystem, PFEM_ERR_STATE, "heat_solve: no assembled heat system (pfem_heat_assemble)");
    PFEM_REQUIRE(relTol > 0 && maxIter > 0, PFEM_ERR_INVALID, "heat_solve: relTol and maxIter must be positive");
    PhaseScope ph(c, "Solve heat system");
    const int n = c->nNodes;
    c->cgD.reserve((size_t)n + 4);
    c->Tn.reserve((size_t)n + 4);
    c->Tnb.reserve((size_t)n + 4);
    const bool guess = c->haveTemperature;
    c->haveTemperature = true;
    k_fs_scalar_dinv<<<divUp(n, 256), 256, 0, c->stream>>>(n, 1, c->nbrPtr.p, c->diagSlot.p, c->hA.p, c->cgD.p);  // Eigen::DiagonalPreconditioner
    LAUNCH_CHECK(c);
    c->gmBank.reserve((size_t)CGB_COUNT * cgdStride(c, n));
    c->gmS.reserve(CGD_COUNT + 8);
    const ScalarSpmvDot spmv{c, n, 1, cgdStride(c, n)};
    const int status = cgDevice(c, n, spmv, c->cgD.p, c->hb.p, c->Tn.p, guess, relTol, maxIter, itersOut, relResOut);
    if (Tout) {
        CUDA_CHECK(cudaMemcpyAsync(Tout, c->Tn.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
    }
    return status;
}

// the heat system in the reference's format: column-major compressed, rows of masked nodes reduced to their diagonal
void heatExport(pfem_ctx* c, int64_t* nnz, int32_t* colPtr, int32_t* rowIdx, double* val, double* b) {
    PFEM_REQUIRE(c->haveHeatSystem && nnz, PFEM_ERR_STATE, "heat_export: no assembled heat system");
    const int n = c->nNodes;
    std::vector<int> ptr((size_t)n + 1), nbr((size_t)c->nBlocks), diag((size_t)n);
    std::vector<double> A((size_t)c->nBlocks), bb((size_t)n);
    std::vector<uint8_t> fl((size_t)n), tm((size_t)n);
    CUDA_CHECK(cudaMemcpyAsync(ptr.data(), c->nbrPtr.p, ((size_t)n + 1) * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(nbr.data(), c->nbr.p, (size_t)c->nBlocks * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(A.data(), c->hA.p, (size_t)c->nBlocks * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(bb.data(), c->hb.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(fl.data(), c->flags.p, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(tm.data(), c->heatMask ? c->heatMask : c->tMask.p, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    // entry (i, j) exists when row i is unmasked, or i == j; the pattern is symmetric in (i, j), so column j lists the
    // neighbours i of j that qualify (ascending: nbr lists are sorted)
    auto masked = [&](int i) { return tm[i] || (fl[i] & PFEM_NODE_FREE); };
    int64_t count = 0;
    for (int j = 0; j < n; ++j)
        for (int k = ptr[j]; k < ptr[j + 1]; ++k) {
            const int i = nbr[k];
            if (!masked(i) || i == j) ++count;
        }
    *nnz = count;
    if (!colPtr) return;
    PFEM_REQUIRE(rowIdx && val && b, PFEM_ERR_INVALID, "heat_export: null output array");
    int64_t o = 0;
    for (int j = 0; j < n; ++j) {
        colPtr[j] = (int32_t)o;
        for (int k = ptr[j]; k < ptr[j + 1]; ++k) {
            const int i = nbr[k];
            if (!(!masked(i) || i == j)) continue;
            // value A(i, j): slot of j in row i
            const int* lo = std::lower_bound(nbr.data() + ptr[i], nbr.data() + ptr[i + 1], j);
            rowIdx[o] = i;
            val[o] = A[(size_t)(lo - nbr.data())];
            ++o;
        }
    }
    colPtr[n] = (int32_t)o;
    for (int i = 0; i < n; ++i) b[i] = bb[i];
}

// ---- fractional-step solver id "FracStep" (MomContEquationFracStep.inl) ---------------------------------------
// The three linear systems of one Picard body (:464-548), each assembled on the device from given inputs, and Eigen's
// Jacobi-preconditioned conjugate gradients for them.  Single-GPU contexts, gamma = 0 (no surface-tension facet term).
namespace {
void fsRequire(pfem_ctx* c, const char* what) {
    PFEM_REQUIRE(c->haveTopology && c->havePositions, PFEM_ERR_STATE, std::string(what) + ": topology/positions missing");
    PFEM_REQUIRE(c->nRanks == 1, PFEM_ERR_STATE, std::string(what) + ": single-GPU contexts");
    PFEM_REQUIRE(c->gammaST < 1e-15, PFEM_ERR_STATE, std::string(what) + ": the fractional-step systems carry no surface-tension term (gamma must be 0)");
}
HeatArgs fsHeatArgs(pfem_ctx* c, const uint8_t* mask) {
    HeatArgs a;
    a.conn = c->conn.p, a.n2ePtr = c->n2ePtr.p, a.n2e = c->n2e.p, a.blkMask = c->blkMask.p, a.nbrPtr = c->nbrPtr.p, a.nbr = c->nbr.p;
    a.diagSlot = c->diagSlot.p, a.flags = c->flags.p, a.tMask = mask, a.tVal = nullptr, a.X4 = c->X4.p;
    a.thetaPrev = nullptr, a.A = c->hA.p, a.b = nullptr, a.nRows = c->nNodes, a.CH = c->maskWords;
    a.rhoCv = 0, a.k = 0, a.dt = 1;
    return a;
}
}  // namespace

// system 1: (M/dt + K) vTilde = F + M/dt v_prev + gammaFS D^T p_prev with m_applyBCVAppStep (FracStep.inl:8-298), held in the
// node-block storage of the PSPG system with identity pressure rows: pfem_pspg_export_csc / pfem_pspg_matvec see it
void fsAssembleVapp(pfem_ctx* c, const pfem_pspg_params& p, double gammaFS, const double* qPrevHost) {
    fsRequire(c, "fs_assemble_vapp");
    PFEM_REQUIRE(qPrevHost && p.dt > 0 && p.rho > 0, PFEM_ERR_INVALID, "fs_assemble_vapp: bad arguments");
    PFEM_REQUIRE(!c->thermalOn, PFEM_ERR_STATE, "fs_assemble_vapp: Boussinesq factors are not taken over for FracStep");
    const int BS = c->dim + 1, n = c->nNodes;
    c->Aval.reserve((size_t)c->nBlocks * BS * BS);
    c->bvec.reserve((size_t)n * BS);
    c->dinv.reserve((size_t)n * BS);
    c->hTheta.reserve((size_t)n + 4);
    fieldsSetQprev(c, qPrevHost);  // velocity part -> VP4
    CUDA_CHECK(cudaMemcpyAsync(c->hTheta.p, qPrevHost + (size_t)c->dim * n, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    GenArgs a = makeGenArgs(c, p);
    a.T = nullptr;
    a.mode = 1, a.gammaFS = gammaFS, a.pPrev = c->hTheta.p;
    PhaseScope ph(c, "Assemble system");
    if (c->dim == 2) {
        k_gen_elem<2><<<divUp(std::max(c->nElems, 1), 128), 128, 0, c->stream>>>(a);
        LAUNCH_CHECK(c);
        k_gen_blocks<2><<<divUp(c->nBlocks, 128), 128, 0, c->stream>>>(a);
        LAUNCH_CHECK(c);
        k_gen_rhs_bc<2><<<divUp(n, 128), 128, 0, c->stream>>>(a);
    } else {
        k_gen_elem<3><<<divUp(std::max(c->nElems, 1), 128), 128, 0, c->stream>>>(a);
        LAUNCH_CHECK(c);
        k_gen_blocks<3><<<divUp(c->nBlocks, 128), 128, 0, c->stream>>>(a);
        LAUNCH_CHECK(c);
        k_gen_rhs_bc<3><<<divUp(n, 128), 128, 0, c->stream>>>(a);
    }
    LAUNCH_CHECK(c);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->haveSystem = true;
    c->haveSolution = false;
    mgInvalidate(c, false);
    c->asmStamp = p.dt;
    c->fsWhich = 0;
}
// system 2: L p = -(rho/dt) D vTilde + gammaFS L p_prev with m_applyBCPCorrStep (FracStep.inl:124-130, 158-166, 300-376), scalar
// system in the storage of the heat equation: pfem_heat_export_csc sees it
void fsAssemblePcorr(pfem_ctx* c, double rho, double dt, double gammaFS, const double* vTildeHost, const double* pPrevHost) {
    fsRequire(c, "fs_assemble_pcorr");
    PFEM_REQUIRE(vTildeHost && pPrevHost && dt > 0 && rho > 0, PFEM_ERR_INVALID, "fs_assemble_pcorr: bad arguments");
    const int n = c->nNodes, dim = c->dim;
    c->hA.reserve((size_t)c->nBlocks + 4);
    c->hb.reserve((size_t)dim * n + 4);
    c->hTheta.reserve((size_t)n + 4);
    c->fsVec.reserve((size_t)dim * n + 4);
    c->fsMask.reserve((size_t)n + 4);
    CUDA_CHECK(cudaMemsetAsync(c->fsMask.p, 0, (size_t)n, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(c->fsVec.p, vTildeHost, (size_t)dim * n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(c->hTheta.p, pPrevHost, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    HeatArgs a = fsHeatArgs(c, c->fsMask.p);
    a.k = 1.0;  // L with factor 1 (MomContEquation.inl:154-158)
    PhaseScope ph(c, "Assemble system");
    if (dim == 2) {
        k_heat_blocks<2><<<divUp(c->nBlocks, 128), 128, 0, c->stream>>>(a);
        LAUNCH_CHECK(c);
        k_fs_zero_cols<<<divUp(n, 128), 128, 0, c->stream>>>(n, c->nbrPtr.p, c->nbr.p, c->diagSlot.p, c->fsMask.p, c->flags.p, c->hA.p);
        LAUNCH_CHECK(c);
        k_fs_pcorr_rhs<2><<<divUp(n, 128), 128, 0, c->stream>>>(n, n, c->conn.p, c->n2ePtr.p, c->n2e.p, c->flags.p, c->X4.p, c->fsVec.p,
                                                               c->hTheta.p, rho / dt, gammaFS, c->hb.p);
    } else {
        k_heat_blocks<3><<<divUp(c->nBlocks, 128), 128, 0, c->stream>>>(a);
        LAUNCH_CHECK(c);
        k_fs_zero_cols<<<divUp(n, 128), 128, 0, c->stream>>>(n, c->nbrPtr.p, c->nbr.p, c->diagSlot.p, c->fsMask.p, c->flags.p, c->hA.p);
        LAUNCH_CHECK(c);
        k_fs_pcorr_rhs<3><<<divUp(n, 128), 128, 0, c->stream>>>(n, n, c->conn.p, c->n2ePtr.p, c->n2e.p, c->flags.p, c->X4.p, c->fsVec.p,
                                                               c->hTheta.p, rho / dt, gammaFS, c->hb.p);
    }
    LAUNCH_CHECK(c);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->heatMask = c->fsMask.p;
    c->haveHeatSystem = true;
    c->fsWhich = 1;
}
// system 3: M deltaV = dt D^T deltaP with m_applyBCVStep (FracStep.inl:76-83, 167-176, 378-452): the scalar consistent mass
// matrix (factor rho, identity rows for bound / free nodes) for every velocity component, right-hand side [d][nNodes]
void fsAssembleVcorr(pfem_ctx* c, double rho, double dt, const double* deltaPHost) {
    fsRequire(c, "fs_assemble_vcorr");
    PFEM_REQUIRE(deltaPHost && dt > 0 && rho > 0, PFEM_ERR_INVALID, "fs_assemble_vcorr: bad arguments");
    const int n = c->nNodes, dim = c->dim;
    c->hA.reserve((size_t)c->nBlocks + 4);
    c->hb.reserve((size_t)dim * n + 4);
    c->hTheta.reserve((size_t)n + 4);
    c->fsMask.reserve((size_t)n + 4);
    k_fs_bound_mask<<<divUp(n, 256), 256, 0, c->stream>>>(n, c->flags.p, c->fsMask.p);
    LAUNCH_CHECK(c);
    CUDA_CHECK(cudaMemcpyAsync(c->hTheta.p, deltaPHost, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    HeatArgs a = fsHeatArgs(c, c->fsMask.p);
    a.rhoCv = rho;  // M with factor rho (MomContEquation.inl:96-99)
    PhaseScope ph(c, "Assemble system");
    if (dim == 2) {
        k_heat_blocks<2><<<divUp(c->nBlocks, 128), 128, 0, c->stream>>>(a);
        LAUNCH_CHECK(c);
        k_fs_zero_cols<<<divUp(n, 128), 128, 0, c->stream>>>(n, c->nbrPtr.p, c->nbr.p, c->diagSlot.p, c->fsMask.p, c->flags.p, c->hA.p);
        LAUNCH_CHECK(c);
        k_fs_vcorr_rhs<2><<<divUp(n, 128), 128, 0, c->stream>>>(n, n, c->conn.p, c->n2ePtr.p, c->n2e.p, c->flags.p, c->X4.p, c->hTheta.p, dt, c->hb.p);
    } else {
        k_heat_blocks<3><<<divUp(c->nBlocks, 128), 128, 0, c->stream>>>(a);
        LAUNCH_CHECK(c);
        k_fs_zero_cols<<<divUp(n, 128), 128, 0, c->stream>>>(n, c->nbrPtr.p, c->nbr.p, c->diagSlot.p, c->fsMask.p, c->flags.p, c->hA.p);
        LAUNCH_CHECK(c);
        k_fs_vcorr_rhs<3><<<divUp(n, 128), 128, 0, c->stream>>>(n, n, c->conn.p, c->n2ePtr.p, c->n2e.p, c->flags.p, c->X4.p, c->hTheta.p, dt, c->hb.p);
    }
    LAUNCH_CHECK(c);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->heatMask = c->fsMask.p;
    c->haveHeatSystem = true;
    c->fsWhich = 2;
}
// right-hand side of the scalar-storage systems (2: nNodes values, 3: dim * nNodes as [d][nNodes])
void fsGetRhs(pfem_ctx* c, double* b) {
    PFEM_REQUIRE(b && (c->fsWhich == 1 || c->fsWhich == 2) && c->haveHeatSystem, PFEM_ERR_STATE, "fs_get_rhs: assemble system 2 or 3 first");
    const size_t n = (size_t)(c->fsWhich == 2 ? c->dim : 1) * c->nNodes;
    CUDA_CHECK(cudaMemcpyAsync(b, c->hb.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
}
// m_solverIt.compute(A); x = m_solverIt.solve(b) on the system assembled last (FracStep.inl:472-478, 491-497, 516-522)
int fsSolve(pfem_ctx* c, double relTol, int maxIter, double* xHost, int* itersOut, double* relResOut) {
    PFEM_REQUIRE(c->fsWhich >= 0, PFEM_ERR_STATE, "fs_solve: no fractional-step system assembled");
    PFEM_REQUIRE(relTol > 0 && maxIter > 0, PFEM_ERR_INVALID, "fs_solve: relTol and maxIter must be positive");
    PhaseScope ph(c, "Solve system");
    const int N = c->nNodes, dim = c->dim;
    int status;
    if (c->fsWhich == 0) {  // node-block system, internal dof order node*BS + d (pressure entries stay 0: identity rows, b = 0)
        PFEM_REQUIRE(c->haveSystem, PFEM_ERR_STATE, "fs_solve: system 1 is gone (another assembly ran)");
        const int BS = dim + 1, n = N * BS;
        c->cgD.reserve((size_t)n + 8);
        c->kx.reserve((size_t)n + 8);
        k_fs_block_dinv<<<divUp(n, 256), 256, 0, c->stream>>>(N, BS, c->nbrPtr.p, c->diagSlot.p, c->Aval.p, c->cgD.p);
        LAUNCH_CHECK(c);
        c->gmBank.reserve((size_t)CGB_COUNT * cgdStride(c, n));
        c->gmS.reserve(CGD_COUNT + 8);
        const int stride = cgdStride(c, n), dotGrid = std::max(1, std::min(c->smCount * 4, divUp(n, RB_THREADS)));
        auto spmvDot = [&](const double* x, double* y) {
            krylovMatvec(c, const_cast<double*>(x), y);
            k_cgd_dot<<<dotGrid, RB_THREADS, 0, c->stream>>>(n, x, y, c->gmBank.p, stride, c->gmS.p);
            LAUNCH_CHECK(c);
            return dotGrid;
        };
        status = cgDevice(c, n, spmvDot, c->cgD.p, c->bvec.p, c->kx.p, false, relTol, maxIter, itersOut, relResOut);
        c->haveSolution = true;
        if (xHost) {  // velocity part, [d][nNodes]
            std::vector<double> q((size_t)n);
            krylovStoreVector(c, c->kx.p, q.data());
            std::copy(q.begin(), q.begin() + (size_t)dim * N, xHost);
        }
        return status;
    }
    PFEM_REQUIRE(c->haveHeatSystem, PFEM_ERR_STATE, "fs_solve: the scalar system is gone");
    const int comps = c->fsWhich == 2 ? dim : 1, n = N * comps;
    c->cgD.reserve((size_t)n + 8);
    c->fsVec.reserve((size_t)dim * N + 4);
    k_fs_scalar_dinv<<<divUp(N, 256), 256, 0, c->stream>>>(N, comps, c->nbrPtr.p, c->diagSlot.p, c->hA.p, c->cgD.p);
    LAUNCH_CHECK(c);
    c->gmBank.reserve((size_t)CGB_COUNT * cgdStride(c, n));
    c->gmS.reserve(CGD_COUNT + 8);
    const ScalarSpmvDot spmv{c, N, comps, cgdStride(c, n)};
    status = cgDevice(c, n, spmv, c->cgD.p, c->hb.p, c->fsVec.p, false, relTol, maxIter, itersOut, relResOut);
    if (xHost) {
        CUDA_CHECK(cudaMemcpyAsync(xHost, c->fsVec.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
    }
    return status;
}
