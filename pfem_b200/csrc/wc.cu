// wc.cu -- explicit weakly-compressible step (SolverWCompNewton::m_solveWCompNewtonNoT, WCompNewton/Solver.cpp:236-263)
// and the CFL time step (computeNextDT, :192-234) on the device.
//
// Reference structure: omp element loop -> per-element temporaries -> SERIAL nodal scatter (ContEquation.inl:398-408,
// MomEquation.inl:279-298).  B200 design: node GATHER.  LPN lanes own one node, stride over its incident elements
// (ascending element index == the reference's serial scatter order), recompute the element geometry from coordinates,
// keep only the row of node i, reduce over the LPN lanes and apply the nodal epilogue (1/M, BC, EOS, kick) in the same
// kernel.  No atomics, no per-element temporaries (the reference stores a 12x12 matrix per element just to read its
// diagonal, MomEquation.inl:237), deterministic.  Three launches per step:
//   k_wc_kick_move : v_half = v + dt/2 a ; x += dt v_half (non-fixed)                       (Solver.cpp:249-253)
//   k_wc_cont      : F0, lumped M -> p = F0/M (free: 0) -> rho = Tait-Murnaghan(p)          (ContEquation.inl:123-148)
//   k_wc_mom       : F = -K v + D^T p + F_b, lumped rho-mass -> a -> v = v_half + dt/2 a    (MomEquation.inl:201-226)
// Device layout: X4=(x,y,z,p), V4=(u,v,w,rho), A4=(ax,ay,az,-); cont writes the ping-pong copies X4b/V4b so that the
// gathers of other nodes never see half-updated p/rho.
#include <cstdlib>

#include "common.cuh"

namespace {

__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
// 32-byte records move with ONE 256-bit instruction (sm_100: LDG.E.ENL2.256 / STG.E.ENL2.256); p must be 32-byte aligned
__device__ __forceinline__ void st4(double* p, double a, double b, double c, double d) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
struct D4 {
    double x, y, z, w;
};
__device__ __forceinline__ D4 ld4(const double* p) {
    D4 r;
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}

__global__ void k_wc_kick_move(int nNodes, int dim, double dtVal, const double* __restrict__ dtPtr,
                               const uint8_t* __restrict__ flags, double* __restrict__ X4, double* __restrict__ V4,
                               const double* __restrict__ A4) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nNodes) return;
    const double dt = dtPtr ? *dtPtr : dtVal;  // device-resident dt when steps are chained without a host round trip
    const bool fixed = flags[n] & PFEM_NODE_FIXED;
    for (int d = 0; d < dim; ++d) {
        const double vh = V4[(size_t)n * 4 + d] + 0.5 * dt * A4[(size_t)n * 4 + d];
        V4[(size_t)n * 4 + d] = vh;
        if (!fixed) X4[(size_t)n * 4 + d] += vh * dt;
    }
}

template <int DIM> struct ElemGeo {
    double g[DIM][DIM + 1];
    double V;
};

// grad N and the element size from the node coordinates (Element.cpp:15-135, MatricesBuilder.inl:93-127)
template <int DIM> __device__ __forceinline__ void buildGeo(const double (&px)[DIM + 1][DIM], ElemGeo<DIM>& G) {
    constexpr double REF = (DIM == 2) ? 0.5 : 0.16666666666666666666666666666667;
    double J[DIM][DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d)
#pragma unroll
        for (int m = 0; m < DIM; ++m) J[d][m] = px[m + 1][d] - px[0][d];
    double det, inv[DIM][DIM];
    if constexpr (DIM == 2) {
        det = J[0][0] * J[1][1] - J[1][0] * J[0][1];
        const double rd = 1.0 / det;
        inv[0][0] = J[1][1] * rd;
        inv[0][1] = -J[0][1] * rd;
        inv[1][0] = -J[1][0] * rd;
        inv[1][1] = J[0][0] * rd;
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
        G.g[d][0] = s;
#pragma unroll
        for (int m = 0; m < DIM; ++m) G.g[d][m + 1] = inv[m][d];
    }
    G.V = det * REF;
}

// gather the 4-double records of the element nodes and build grad N and the volume (Element.cpp:15-135, MB.inl:93-127)
template <int DIM>
__device__ __forceinline__ void loadElem(const double* __restrict__ XA, const double* __restrict__ VA,
                                         const int (&nd)[DIM + 1], double (&xw)[DIM + 1], double (&vel)[DIM + 1][DIM],
                                         double (&vw)[DIM + 1], ElemGeo<DIM>& G) {
    constexpr int NPE = DIM + 1;
    double px[NPE][DIM];
#pragma unroll
    for (int m = 0; m < NPE; ++m) {
        const D4 xr = ld4(XA + (size_t)nd[m] * 4), vr = ld4(VA + (size_t)nd[m] * 4);
        px[m][0] = xr.x, px[m][1] = xr.y;
        vel[m][0] = vr.x, vel[m][1] = vr.y;
        if constexpr (DIM == 3) {
            px[m][2] = xr.z;
            vel[m][2] = vr.z;
        }
        xw[m] = xr.w;
        vw[m] = vr.w;
    }
    buildGeo<DIM>(px, G);
}

// he = 2 r_in of Element::getRin (Element.cpp:226-294) from the node positions
template <int DIM> __device__ __forceinline__ double elemHe(const double (&px)[DIM + 1][3]) {
    constexpr double REF = (DIM == 2) ? 0.5 : 0.16666666666666666666666666666667;
    if constexpr (DIM == 2) {
        const double J00 = px[1][0] - px[0][0], J01 = px[2][0] - px[0][0], J10 = px[1][1] - px[0][1], J11 = px[2][1] - px[0][1];
        const double A = (J00 * J11 - J10 * J01) * REF;
        auto dist = [&](int p, int q) {
            const double dx = px[p][0] - px[q][0], dy = px[p][1] - px[q][1];
            return sqrt(dx * dx + dy * dy);
        };
        const double s = (dist(0, 1) + dist(1, 2) + dist(0, 2)) / 2;
        return 2 * (A / s);
    } else {
        const double x0 = px[0][0], x1 = px[1][0], x2 = px[2][0], x3 = px[3][0];
        const double y0 = px[0][1], y1 = px[1][1], y2 = px[2][1], y3 = px[3][1];
        const double z0 = px[0][2], z1 = px[1][2], z2 = px[2][2], z3 = px[3][2];
        const double J00 = x1 - x0, J01 = x2 - x0, J02 = x3 - x0, J10 = y1 - y0, J11 = y2 - y0, J12 = y3 - y0,
                     J20 = z1 - z0, J21 = z2 - z0, J22 = z3 - z0;
        const double det = J00 * J11 * J22 + J01 * J12 * J20 + J02 * J10 * J21 - J20 * J11 * J02 - J21 * J12 * J00 -
                           J22 * J10 * J01;
        auto nrm = [](double p, double q, double r) { return sqrt(p * p + q * q + r * r); };
        const double n1 = nrm(J10 * J21 - J20 * J11, J20 * J01 - J00 * J21, J00 * J11 - J10 * J01);
        const double n2 = nrm(J12 * J21 - J22 * J11, J22 * J01 - J02 * J21, J02 * J11 - J12 * J01);
        const double n3 = nrm(J10 * J22 - J20 * J12, J20 * J02 - J00 * J22, J00 * J12 - J10 * J02);
        const double n4 = nrm((y1 - y3) * (z2 - z3) - (z1 - z3) * (y2 - y3), (z1 - z3) * (x2 - x3) - (x1 - x3) * (z2 - z3),
                              (x1 - x3) * (y2 - y3) - (y1 - y3) * (x2 - x3));
        return 2 * (6 * (det * REF) / (n1 + n2 + n3 + n4));
    }
}

// Incident-element stream of one lane with look-ahead: element ids are fetched three trips ahead, connectivity two trips
// ahead, and the nodal records of the NEXT element are prefetched into L1 while the current element is computed, so the
// n2e -> conn -> record dependent chain is paid once per node instead of once per element.
template <int DIM, int LPN> struct ElemStream {
    static constexpr int NPE = DIM + 1;
    const int* conn;
    const int* n2e;
    const double* XA;
    const double* VA;
    int pos, end;          // current position in n2e, one past the last
    int nd[NPE], nd1[NPE], nd2[NPE], e3;
    __device__ __forceinline__ void ldConn(int e, int (&out)[NPE]) const {
        if constexpr (DIM == 3) {
            const int4 q = __ldg(reinterpret_cast<const int4*>(conn + (size_t)e * 4));
            out[0] = q.x, out[1] = q.y, out[2] = q.z, out[3] = q.w;
        } else {
#pragma unroll
            for (int m = 0; m < NPE; ++m) out[m] = __ldg(conn + (size_t)e * NPE + m);
        }
    }
    __device__ __forceinline__ void init(const int* c, const int* l, const double* xa, const double* va, int first, int last) {
        conn = c, n2e = l, XA = xa, VA = va, pos = first, end = last;
#pragma unroll
        for (int m = 0; m < NPE; ++m) nd[m] = nd1[m] = nd2[m] = 0;
        e3 = 0;
        if (pos < end) ldConn(__ldg(n2e + pos), nd);
        if (pos + LPN < end) ldConn(__ldg(n2e + pos + LPN), nd1);
        if (pos + 2 * LPN < end) ldConn(__ldg(n2e + pos + 2 * LPN), nd2);
        if (pos + 3 * LPN < end) e3 = __ldg(n2e + pos + 3 * LPN);
    }
    __device__ __forceinline__ bool valid() const { return pos < end; }
    // call at the top of a trip: warm L1 with the records of the next element
    __device__ __forceinline__ void prefetchNext() const {
        if (pos + LPN < end) {
#pragma unroll
            for (int m = 0; m < NPE; ++m) {
                asm volatile("prefetch.global.L1 [%0];" ::"l"(XA + (size_t)nd1[m] * 4));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(VA + (size_t)nd1[m] * 4));
            }
        }
    }
    __device__ __forceinline__ void advance() {
#pragma unroll
        for (int m = 0; m < NPE; ++m) nd[m] = nd1[m], nd1[m] = nd2[m];
        pos += LPN;
        if (pos + 2 * LPN < end) ldConn(e3, nd2);
        if (pos + 3 * LPN < end) e3 = __ldg(n2e + pos + 3 * LPN);
    }
};

// max that keeps a NaN operand (fmax drops it): a NaN velocity/density must reach the CFL minimum (Solver.cpp:228-232)
__device__ __forceinline__ double nanMax(double a, double b) { return (a > b || a != a) ? a : b; }
template <int N> __device__ __forceinline__ double pick(const double (&a)[N], int idx) {
    double t = a[0];
#pragma unroll
    for (int m = 1; m < N; ++m) t = (idx == m) ? a[m] : t;
    return t;
}
// min that keeps a NaN operand
__device__ __forceinline__ double nanMin(double a, double b) { return (a < b || a != a) ? a : b; }
template <int LPN> __device__ __forceinline__ double groupMin(double v) {
#pragma unroll
    for (int o = LPN / 2; o > 0; o >>= 1) v = nanMin(__shfl_xor_sync(0xffffffffu, v, o), v);
    return v;
}
template <int LPN> __device__ __forceinline__ double groupSum(double v) {
#pragma unroll
    for (int o = LPN / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct WcArgs {
    const int* conn;
    const int* n2ePtr;
    const int* n2e;
    const uint8_t* flags;
    const uint8_t* dirMask;
    const double* dirVal4;
    const double* fst4 = nullptr;  // nodal surface-tension force (facets.cu) or null
    int nNodes;
    double dt, mu, K0, K0p, rhoStar, body[3];
    const double* dtPtr;  // if non-null the time step is read from the device (pfem_wc_run)
    int meduri;
    // node passes of the two-pass formulation may cover a slice of an ORDER of the nodes (interface nodes first, so that
    // their halo exchange overlaps the rest): node = order[k0 + k], k < nNodes; order == null: node = k0 + k
    const int* order = nullptr;
    int k0 = 0;
    // BoussinesqWC: nodal temperature (null: off) and the buoyancy constants (WCompNewton/MomEquation.inl:105-112)
    const double* T = nullptr;
    double thAlpha = 0, thTr = 0;
};
__device__ __forceinline__ int wcNodeOf(const WcArgs& a, int k) { return a.order ? __ldg(a.order + a.k0 + k) : a.k0 + k; }

// continuity, CDS_dpdt (ContEquation.inl:353-413, 334-350, 139-146, 319-331)
template <int DIM, int LPN, int MINB>
__global__ void __launch_bounds__(256, MINB) k_wc_cont(const WcArgs a, const double* __restrict__ X4, const double* __restrict__ V4,
                                                 double* __restrict__ X4n, double* __restrict__ V4n) {
    constexpr int NPE = DIM + 1;
    constexpr double PHI = 1.0 / ((DIM + 1) * (DIM + 2));
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = t / LPN, sub = t % LPN;
    const bool valid = i < a.nNodes;
    const double dtStep = a.dtPtr ? *a.dtPtr : a.dt;
    double m = 0, F0 = 0;
    if (valid) {
        const int eb = a.n2ePtr[i], ne = a.n2ePtr[i + 1] - eb;
        ElemStream<DIM, LPN> es;
        es.init(a.conn, a.n2e, X4, V4, eb + sub, eb + ne);
        for (; es.valid(); es.advance()) {
            es.prefetchNext();
            const int(&nd)[NPE] = es.nd;
            double P[NPE], vel[NPE][DIM], rho[NPE];
            ElemGeo<DIM> G;
            loadElem<DIM>(X4, V4, nd, P, vel, rho, G);
            int li = 0;
            double sumP = 0, divv = 0;
#pragma unroll
            for (int q = 0; q < NPE; ++q) {
                li = (nd[q] == i) ? q : li;
                sumP += P[q];
#pragma unroll
                for (int c = 0; c < DIM; ++c) divv += G.g[c][q] * vel[q][c];
            }
            const double pi = pick<NPE>(P, li);
            const double si = a.K0 / NPE + a.K0p * PHI * (pi + sumP);       // sum_g w (K0 + K0' N.p) N_i
            const double stab = a.meduri ? G.V * PHI * (pi + sumP) : (G.V / NPE) * pi;  // Me*P | MeLumped*P
            F0 += -dtStep * G.V * si * divv + stab;
            m += G.V / NPE;                                                  // lump2(M), f = 1
        }
    }
    m = groupSum<LPN>(m);
    F0 = groupSum<LPN>(F0);
    if (valid && sub == 0) {
        const bool isFree = a.flags[i] & PFEM_NODE_FREE;
        double inv = 1.0 / m;
        if (isFree) {
            F0 = 0.0;
            inv = 1.0;
        }
        const double p = inv * F0;
        const double rho = pow((a.K0p / a.K0) * p + 1.0, 1.0 / a.K0p) * a.rhoStar;
        const double* xp = X4 + (size_t)i * 4;
        const double* vp = V4 + (size_t)i * 4;
        st4(X4n + (size_t)i * 4, xp[0], xp[1], xp[2], p);
        st4(V4n + (size_t)i * 4, vp[0], vp[1], vp[2], rho);
    }
}

// continuity in the density forms.  MODE 1: CDS_drhodt (ContEquation.inl:234-301: F0e = -dt D_rho V + M rho | M_lumped rho),
// MODE 2: CDS_rho (lumped mass on the moved mesh, F0 from the pre-pass), MODE 3: the CDS_rho pre-pass
// F0 = sum_e M_e rho_e on the configuration before the move (m_buildF0, :196-232).  BC (:177-194): free or
// free-surface nodes get rho*; p(rho) by Tait-Murnaghan (:303-316).
template <int DIM, int LPN, int MINB, int MODE>
__global__ void __launch_bounds__(256, MINB) k_wc_cont_rho(const WcArgs a, const double* __restrict__ X4, const double* __restrict__ V4,
                                                     double* __restrict__ X4n, double* __restrict__ V4n, double* __restrict__ F0pre) {
    constexpr int NPE = DIM + 1;
    constexpr double PHI = 1.0 / ((DIM + 1) * (DIM + 2));
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = t / LPN, sub = t % LPN;
    const bool valid = i < a.nNodes;
    const double dtStep = a.dtPtr ? *a.dtPtr : a.dt;
    double m = 0, F0 = 0;
    if (valid) {
        const int eb = a.n2ePtr[i], ne = a.n2ePtr[i + 1] - eb;
        ElemStream<DIM, LPN> es;
        es.init(a.conn, a.n2e, X4, V4, eb + sub, eb + ne);
        for (; es.valid(); es.advance()) {
            es.prefetchNext();
            const int(&nd)[NPE] = es.nd;
            double P[NPE], vel[NPE][DIM], rho[NPE];
            ElemGeo<DIM> G;
            loadElem<DIM>(X4, V4, nd, P, vel, rho, G);
            int li = 0;
            double sumR = 0, divv = 0;
#pragma unroll
            for (int q = 0; q < NPE; ++q) {
                li = (nd[q] == i) ? q : li;
                sumR += rho[q];
#pragma unroll
                for (int c = 0; c < DIM; ++c) divv += G.g[c][q] * vel[q][c];
            }
            const double ri = pick<NPE>(rho, li);
            const double cons = G.V * PHI * (ri + sumR);  // (M_e rho_e)_i == V sum_g w (N.rho) N_i
            if (MODE == 1) F0 += -dtStep * cons * divv + (a.meduri ? cons : (G.V / NPE) * ri);
            if (MODE == 3) F0 += cons;
            m += G.V / NPE;
        }
    }
    m = groupSum<LPN>(m);
    F0 = groupSum<LPN>(F0);
    if (valid && sub == 0) {
        if (MODE == 3) {
            F0pre[i] = F0;
            return;
        }
        if (MODE == 2) F0 = F0pre[i];
        double inv = 1.0 / m;
        if (a.flags[i] & (PFEM_NODE_FREE | PFEM_NODE_FREE_SURFACE)) {
            F0 = a.rhoStar;
            inv = 1.0;
        }
        const double rho = inv * F0;
        const double p = (a.K0 / a.K0p) * (pow(rho / a.rhoStar, a.K0p) - 1.0);
        const double* xp = X4 + (size_t)i * 4;
        const double* vp = V4 + (size_t)i * 4;
        st4(X4n + (size_t)i * 4, xp[0], xp[1], xp[2], p);
        st4(V4n + (size_t)i * 4, vp[0], vp[1], vp[2], rho);
    }
}

// momentum (MomEquation.inl:229-302, 305-374, 216-222)
template <int DIM, int LPN, int MINB>
__global__ void __launch_bounds__(256, MINB) k_wc_mom(const WcArgs a, const double* __restrict__ X4, const double* __restrict__ V4,
                                                double* __restrict__ V4out, double* __restrict__ A4out, double* __restrict__ X4out,
                                                double* __restrict__ cfl2 = nullptr) {
    constexpr int NPE = DIM + 1;
    constexpr double PHI = 1.0 / ((DIM + 1) * (DIM + 2));
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = t / LPN, sub = t % LPN;
    const bool valid = i < a.nNodes;
    const double dtStep = a.dtPtr ? *a.dtPtr : a.dt;
    double M = 0, F[DIM];
#pragma unroll
    for (int c = 0; c < DIM; ++c) F[c] = 0;
    if (valid) {
        const int eb = a.n2ePtr[i], ne = a.n2ePtr[i + 1] - eb;
        ElemStream<DIM, LPN> es;
        es.init(a.conn, a.n2e, X4, V4, eb + sub, eb + ne);
        for (; es.valid(); es.advance()) {
            es.prefetchNext();
            const int(&nd)[NPE] = es.nd;
            double P[NPE], vel[NPE][DIM], rho[NPE];
            ElemGeo<DIM> G;
            loadElem<DIM>(X4, V4, nd, P, vel, rho, G);
            int li = 0;
            double sumP = 0, sumR = 0;
            double Gm[DIM][DIM];  // G_ac = sum_j v_{j,a} g[c][j]
#pragma unroll
            for (int aa = 0; aa < DIM; ++aa)
#pragma unroll
                for (int c = 0; c < DIM; ++c) Gm[aa][c] = 0;
#pragma unroll
            for (int q = 0; q < NPE; ++q) {
                li = (nd[q] == i) ? q : li;
                sumP += P[q];
                sumR += rho[q];
#pragma unroll
                for (int aa = 0; aa < DIM; ++aa)
#pragma unroll
                    for (int c = 0; c < DIM; ++c) Gm[aa][c] += vel[q][aa] * G.g[c][q];
            }
            double tr = 0;
#pragma unroll
            for (int aa = 0; aa < DIM; ++aa) tr += Gm[aa][aa];
            double gi[DIM];
#pragma unroll
            for (int c = 0; c < DIM; ++c) gi[c] = pick<NPE>(G.g[c], li);
            const double pbar = sumP / NPE;
            const double li_mass = G.V * PHI * (pick<NPE>(rho, li) + sumR);  // lumped rho-mass == sum_g w (N.rho) N_i
            double bodyMass = li_mass;
            if (a.T) {  // F factor (N.rho)(1 - alpha (N.T - Tr)): Gauss sum, shape functions GA at "their" point, GB elsewhere
                constexpr double GA = (DIM == 3) ? 0.585410196624968 : 0.66666666666666666667;
                constexpr double GB = (DIM == 3) ? 0.138196601125011 : 0.16666666666666666667;
                double sumT = 0, Te[NPE];
#pragma unroll
                for (int q = 0; q < NPE; ++q) {
                    Te[q] = a.T[nd[q]];
                    sumT += Te[q];
                }
                double fs = 0;
#pragma unroll
                for (int g = 0; g < NPE; ++g) {
                    const double rg = GB * sumR + (GA - GB) * rho[g], Tg = GB * sumT + (GA - GB) * Te[g];
                    fs += (1.0 / NPE) * (rg * (1.0 - a.thAlpha * (Tg - a.thTr))) * (g == li ? GA : GB);
                }
                bodyMass = G.V * fs;
            }
#pragma unroll
            for (int aa = 0; aa < DIM; ++aa) {
                double sg = 0;  // sum_c sigma_ac g[c][i],  sigma = mu (G + G^T - 2/3 tr I)
#pragma unroll
                for (int c = 0; c < DIM; ++c) {
                    double sig = Gm[aa][c] + Gm[c][aa];
                    if (c == aa) sig -= (2.0 / 3.0) * tr;
                    sg += a.mu * sig * gi[c];
                }
                F[aa] += -G.V * sg + G.V * pbar * gi[aa] + a.body[aa] * bodyMass;
            }
            M += li_mass;
        }
    }
    M = groupSum<LPN>(M);
#pragma unroll
    for (int c = 0; c < DIM; ++c) F[c] = groupSum<LPN>(F[c]);
    if (valid && sub == 0) {
        const uint8_t fl = a.flags[i];
        const bool isFree = fl & PFEM_NODE_FREE, isBound = fl & PFEM_NODE_BOUND;
        const double* vp = V4 + (size_t)i * 4;
        double inv = 1.0 / M;
        double acc[3] = {0, 0, 0}, vn[3] = {0, 0, 0};
#pragma unroll
        for (int c = 0; c < DIM; ++c) {
            double f = F[c], iv = inv;
            if (a.fst4) f += a.fst4[(size_t)i * 4 + c];  // facet loop of m_applyBC (MomEquation.inl:312-336)
            if (isFree && !isBound) {
                f = a.body[c];
                iv = 1.0;
            } else if (isBound && a.dirMask[i]) {
                f = a.dirVal4[(size_t)i * 4 + c];  // reference hazard 10: the Dirichlet *velocity* becomes the acceleration
                iv = 1.0;
            }
            acc[c] = iv * f;
            vn[c] = vp[c] + 0.5 * dtStep * acc[c];
        }
        st4(V4out + (size_t)i * 4, vn[0], vn[1], vn[2], vp[3]);
        st4(A4out + (size_t)i * 4, acc[0], acc[1], acc[2], 0.0);
        const double* xq = X4 + (size_t)i * 4;
        st4(X4out + (size_t)i * 4, xq[0], xq[1], xq[2], xq[3]);  // (x, p_new) becomes the current record: no buffer swap
        if (cfl2) {  // nodal CFL quantities for k_wc_dt_fast (see k_wc_mom_node)
            double u2 = vn[0] * vn[0] + vn[1] * vn[1];
            if (DIM == 3) u2 += vn[2] * vn[2];
            const double c2 = (a.K0 + a.K0p * xq[3]) / vp[3];
            const double alpha = a.mu / vp[3];
            *reinterpret_cast<double2*>(cfl2 + (size_t)i * 2) = make_double2(nanMax(u2, c2), alpha * alpha);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Two-pass ("element record") variants of the CDS_dpdt continuity and of the momentum pass.  The gather kernels above
// recompute the geometry of every element once per incident node (4x in 3-D) behind a dependent n2e -> conn -> record
// chain.  Here pass E runs one thread per ELEMENT (coalesced connectivity, four independent record gathers, geometry
// once) and writes what each of its nodes will need; pass N runs the usual LPN lanes per NODE over its incidence list,
// in the same order as the gather kernels (bit-reproducible, partition-independent), reads one 32-byte sector per
// incident element and applies the nodal epilogue.  Still no atomics.
//  continuity: the nodal contribution is affine in the node's own pressure, F0_i = alpha_e + beta_e p_i, m_i = V/NPE, so
//              the record is per ELEMENT (alpha, beta, V/NPE, -) and is re-read from L2 by the element's other nodes;
//  momentum:   the record is per (element, local node): (F_x, F_y, F_z, lumped rho-mass).
template <int DIM, int MINB>
__global__ void __launch_bounds__(256, MINB) k_wc_cont_elem(int nElems, const int* __restrict__ conn, const double* __restrict__ X4,
                                                      const double* __restrict__ V4, double dtVal, const double* __restrict__ dtPtr,
                                                      double K0, double K0p, int meduri, double* __restrict__ rec) {
    constexpr int NPE = DIM + 1;
    constexpr double PHI = 1.0 / ((DIM + 1) * (DIM + 2));
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nElems) return;
    const double dtStep = dtPtr ? *dtPtr : dtVal;
    int nd[NPE];
    if constexpr (DIM == 3) {
        const int4 q = __ldg(reinterpret_cast<const int4*>(conn + (size_t)e * 4));
        nd[0] = q.x, nd[1] = q.y, nd[2] = q.z, nd[3] = q.w;
    } else {
#pragma unroll
        for (int m = 0; m < NPE; ++m) nd[m] = __ldg(conn + (size_t)e * NPE + m);
    }
    double P[NPE], vel[NPE][DIM], rho[NPE];
    ElemGeo<DIM> G;
    loadElem<DIM>(X4, V4, nd, P, vel, rho, G);
    double sumP = 0, divv = 0;
#pragma unroll
    for (int q = 0; q < NPE; ++q) {
        sumP += P[q];
#pragma unroll
        for (int c = 0; c < DIM; ++c) divv += G.g[c][q] * vel[q][c];
    }
    // F0_i = -dt V divv (K0/NPE + K0' PHI (p_i + sumP)) + (meduri ? V PHI (p_i + sumP) : (V/NPE) p_i)
    const double adv = -dtStep * G.V * divv;
    const double alpha = adv * (K0 / NPE + K0p * PHI * sumP) + (meduri ? G.V * PHI * sumP : 0.0);
    const double beta = adv * (K0p * PHI) + (meduri ? G.V * PHI : G.V / NPE);
    double px[NPE][3];
#pragma unroll
    for (int m = 0; m < NPE; ++m) {
        const double* xp = X4 + (size_t)nd[m] * 4;  // L1 hits: loadElem just read these records
        px[m][0] = xp[0], px[m][1] = xp[1], px[m][2] = xp[2];
    }
    st4(rec + (size_t)e * 4, alpha, beta, G.V / NPE, elemHe<DIM>(px));  // he for the CFL pass (k_wc_dt_fast)
}

template <int DIM, int LPN>
__global__ void __launch_bounds__(256) k_wc_cont_node(const WcArgs a, const double* __restrict__ rec, const double* __restrict__ X4,
                                                      const double* __restrict__ V4, double* __restrict__ X4n,
                                                      double* __restrict__ V4n, double* __restrict__ hminOut) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int kn = t / LPN, sub = t % LPN;
    const bool valid = kn < a.nNodes;
    const int i = valid ? wcNodeOf(a, kn) : 0;
    double m = 0, F0 = 0, hmin = 1.7976931348623157e308;
    if (valid) {
        const int eb = a.n2ePtr[i], end = a.n2ePtr[i + 1];
        const double pi = X4[(size_t)i * 4 + 3];
        for (int pos = eb + sub; pos < end; pos += LPN) {
            const D4 r = ld4(rec + (size_t)__ldg(a.n2e + pos) * 4);
            F0 += r.x + r.y * pi;
            m += r.z;
            hmin = nanMin(r.w, hmin);  // smallest he = 2 r_in among the incident elements: the nodal CFL pass (k_wc_dt_nodal)
        }
    }
    m = groupSum<LPN>(m);
    F0 = groupSum<LPN>(F0);
    hmin = groupMin<LPN>(hmin);
    if (valid && sub == 0) {
        hminOut[i] = hmin;
        const bool isFree = a.flags[i] & PFEM_NODE_FREE;
        double inv = 1.0 / m;
        if (isFree) {
            F0 = 0.0;
            inv = 1.0;
        }
        const double p = inv * F0;
        const double rho = pow((a.K0p / a.K0) * p + 1.0, 1.0 / a.K0p) * a.rhoStar;
        const double* xp = X4 + (size_t)i * 4;
        const double* vp = V4 + (size_t)i * 4;
        st4(X4n + (size_t)i * 4, xp[0], xp[1], xp[2], p);
        st4(V4n + (size_t)i * 4, vp[0], vp[1], vp[2], rho);
    }
}

// TH: BoussinesqWC -- the body-force term takes the buoyancy-weighted mass sum_g w (N.rho)(1 - alpha (N.T - Tr)) N_q
// (WCompNewton/MomEquation.inl:105-112), the same Gauss sum as in the gather kernel k_wc_mom
template <int DIM, int MINB, bool TH = false>
__global__ void __launch_bounds__(256, MINB) k_wc_mom_elem(int nElems, const int* __restrict__ conn, const double* __restrict__ X4,
                                                     const double* __restrict__ V4, double mu, double bx, double by, double bz,
                                                     double* __restrict__ rec, const double* __restrict__ T = nullptr,
                                                     double thAlpha = 0.0, double thTr = 0.0) {
    constexpr int NPE = DIM + 1;
    constexpr double PHI = 1.0 / ((DIM + 1) * (DIM + 2));
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nElems) return;
    int nd[NPE];
    if constexpr (DIM == 3) {
        const int4 q = __ldg(reinterpret_cast<const int4*>(conn + (size_t)e * 4));
        nd[0] = q.x, nd[1] = q.y, nd[2] = q.z, nd[3] = q.w;
    } else {
#pragma unroll
        for (int m = 0; m < NPE; ++m) nd[m] = __ldg(conn + (size_t)e * NPE + m);
    }
    double P[NPE], vel[NPE][DIM], rho[NPE];
    ElemGeo<DIM> G;
    loadElem<DIM>(X4, V4, nd, P, vel, rho, G);
    const double body[3] = {bx, by, bz};
    double sumP = 0, sumR = 0;
    double Gm[DIM][DIM];  // G_ac = sum_j v_{j,a} g[c][j]
#pragma unroll
    for (int aa = 0; aa < DIM; ++aa)
#pragma unroll
        for (int c = 0; c < DIM; ++c) Gm[aa][c] = 0;
#pragma unroll
    for (int q = 0; q < NPE; ++q) {
        sumP += P[q];
        sumR += rho[q];
#pragma unroll
        for (int aa = 0; aa < DIM; ++aa)
#pragma unroll
            for (int c = 0; c < DIM; ++c) Gm[aa][c] += vel[q][aa] * G.g[c][q];
    }
    double tr = 0;
#pragma unroll
    for (int aa = 0; aa < DIM; ++aa) tr += Gm[aa][aa];
    const double pbar = sumP / NPE;
    double sig[DIM][DIM];  // mu (G + G^T - 2/3 tr I)
#pragma unroll
    for (int aa = 0; aa < DIM; ++aa)
#pragma unroll
        for (int c = 0; c < DIM; ++c) {
            double sv = Gm[aa][c] + Gm[c][aa];
            if (c == aa) sv -= (2.0 / 3.0) * tr;
            sig[aa][c] = mu * sv;
        }
    double Te[NPE], sumT = 0;
    if constexpr (TH) {
#pragma unroll
        for (int q = 0; q < NPE; ++q) {
            Te[q] = T[nd[q]];
            sumT += Te[q];
        }
    }
#pragma unroll
    for (int q = 0; q < NPE; ++q) {
        const double li_mass = G.V * PHI * (rho[q] + sumR);  // lumped rho-mass == sum_g w (N.rho) N_q
        double bodyMass = li_mass;
        if constexpr (TH) {
            constexpr double GA = (DIM == 3) ? 0.585410196624968 : 0.66666666666666666667;
            constexpr double GB = (DIM == 3) ? 0.138196601125011 : 0.16666666666666666667;
            double fs = 0;
#pragma unroll
            for (int g = 0; g < NPE; ++g) {
                const double rg = GB * sumR + (GA - GB) * rho[g], Tg = GB * sumT + (GA - GB) * Te[g];
                fs += (1.0 / NPE) * (rg * (1.0 - thAlpha * (Tg - thTr))) * (g == q ? GA : GB);
            }
            bodyMass = G.V * fs;
        }
        double F[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int aa = 0; aa < DIM; ++aa) {
            double sg = 0;
#pragma unroll
            for (int c = 0; c < DIM; ++c) sg += sig[aa][c] * G.g[c][q];
            F[aa] = -G.V * sg + G.V * pbar * G.g[aa][q] + body[aa] * bodyMass;
        }
        st4(rec + ((size_t)q * nElems + e) * 4, F[0], F[1], F[2], li_mass);  // plane q: consecutive lanes, consecutive sectors
    }
}

template <int DIM, int LPN>
__global__ void __launch_bounds__(256) k_wc_mom_node(const WcArgs a, const double* __restrict__ rec, size_t nElems,
                                                     const unsigned* __restrict__ n2eSlots, const int* __restrict__ diagSlot,
                                                     const double* __restrict__ X4, const double* __restrict__ V4,
                                                     double* __restrict__ V4out, double* __restrict__ A4out,
                                                     double* __restrict__ X4out, double* __restrict__ cfl2) {
    constexpr int NPE = DIM + 1;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int kn = t / LPN, sub = t % LPN;
    const bool valid = kn < a.nNodes;
    const int i = valid ? wcNodeOf(a, kn) : 0;
    const double dtStep = a.dtPtr ? *a.dtPtr : a.dt;
    double M = 0, F[DIM];
#pragma unroll
    for (int c = 0; c < DIM; ++c) F[c] = 0;
    if (valid) {
        const int eb = a.n2ePtr[i], end = a.n2ePtr[i + 1];
        const unsigned mine = (unsigned)diagSlot[i];
        for (int pos = eb + sub; pos < end; pos += LPN) {
            const unsigned sl = __ldg(n2eSlots + pos);  // slots of the element's nodes in this node's neighbour list
            int li = 0;
#pragma unroll
            for (int q = 1; q < NPE; ++q) li = (((sl >> (8 * q)) & 0xffu) == mine) ? q : li;
            const D4 r = ld4(rec + ((size_t)li * nElems + (size_t)__ldg(a.n2e + pos)) * 4);
            F[0] += r.x;
            F[1] += r.y;
            if constexpr (DIM == 3) F[2] += r.z;
            M += r.w;
        }
    }
    M = groupSum<LPN>(M);
#pragma unroll
    for (int c = 0; c < DIM; ++c) F[c] = groupSum<LPN>(F[c]);
    if (valid && sub == 0) {
        const uint8_t fl = a.flags[i];
        const bool isFree = fl & PFEM_NODE_FREE, isBound = fl & PFEM_NODE_BOUND;
        const double* vp = V4 + (size_t)i * 4;
        double inv = 1.0 / M;
        double acc[3] = {0, 0, 0}, vn[3] = {0, 0, 0};
#pragma unroll
        for (int c = 0; c < DIM; ++c) {
            double f = F[c], iv = inv;
            if (a.fst4) f += a.fst4[(size_t)i * 4 + c];  // facet loop of m_applyBC (MomEquation.inl:312-336)
            if (isFree && !isBound) {
                f = a.body[c];
                iv = 1.0;
            } else if (isBound && a.dirMask[i]) {
                f = a.dirVal4[(size_t)i * 4 + c];  // reference hazard 10
                iv = 1.0;
            }
            acc[c] = iv * f;
            vn[c] = vp[c] + 0.5 * dtStep * acc[c];
        }
        st4(V4out + (size_t)i * 4, vn[0], vn[1], vn[2], vp[3]);
        st4(A4out + (size_t)i * 4, acc[0], acc[1], acc[2], 0.0);
        const double* xq = X4 + (size_t)i * 4;
        st4(X4out + (size_t)i * 4, xq[0], xq[1], xq[2], xq[3]);
        // nodal CFL quantities of computeNextDT (Solver.cpp:209-216) on the new state: max(u^2, c^2) and alpha^2
        double u2 = vn[0] * vn[0] + vn[1] * vn[1];
        if (DIM == 3) u2 += vn[2] * vn[2];
        const double c2 = (a.K0 + a.K0p * xq[3]) / vp[3];
        const double alpha = a.mu / vp[3];
        *reinterpret_cast<double2*>(cfl2 + (size_t)i * 2) = make_double2(nanMax(u2, c2), alpha * alpha);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Staged variants: the LPN lanes of a node first copy the 64-byte records (x,y,z,p | u,v,w,rho) of the node's
// neighbours into shared memory ONCE (instead of gathering them once per incident element, ~5x redundantly), then read
// the element nodes by neighbour SLOT (n2eSlots) -- no n2e -> conn -> node dependent chain, ~3x fewer L1 wavefronts.
struct alignas(16) WcRec {
    double x[4];  // x, y, z, p
    double v[4];  // u, v, w, rho
    double pad_[2];
};
struct WcArgsS {
    const int* n2ePtr;
    const unsigned* n2eSlots;
    const int* nbrPtr;
    const int* nbr;
    const int* diagSlot;
    const uint8_t* flags;
    const uint8_t* dirMask;
    const double* dirVal4;
    const double* fst4 = nullptr;
    int nNodes, nbcap;
    double dt, mu, K0, K0p, rhoStar, body[3];
    const double* dtPtr;  // if non-null the time step is read from the device (pfem_wc_run)
    int meduri;
};
template <int DIM>
__device__ __forceinline__ void elemFromRecs(const WcRec* __restrict__ recs, unsigned packed, double (&xw)[DIM + 1],
                                             double (&vel)[DIM + 1][DIM], double (&vw)[DIM + 1], ElemGeo<DIM>& G) {
    constexpr int NPE = DIM + 1;
    constexpr double REF = (DIM == 2) ? 0.5 : 0.16666666666666666666666666666667;
    double px[NPE][DIM];
#pragma unroll
    for (int m = 0; m < NPE; ++m) {
        const WcRec& R = recs[(packed >> (8 * m)) & 0xffu];
        const double2 x01 = ld2(R.x), x23 = ld2(R.x + 2), v01 = ld2(R.v), v23 = ld2(R.v + 2);
        px[m][0] = x01.x, px[m][1] = x01.y;
        vel[m][0] = v01.x, vel[m][1] = v01.y;
        if constexpr (DIM == 3) {
            px[m][2] = x23.x;
            vel[m][2] = v23.x;
        }
        xw[m] = x23.y;
        vw[m] = v23.y;
    }
    double J[DIM][DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d)
#pragma unroll
        for (int m = 0; m < DIM; ++m) J[d][m] = px[m + 1][d] - px[0][d];
    double det, inv[DIM][DIM];
    if constexpr (DIM == 2) {
        det = J[0][0] * J[1][1] - J[1][0] * J[0][1];
        const double rd = 1.0 / det;
        inv[0][0] = J[1][1] * rd;
        inv[0][1] = -J[0][1] * rd;
        inv[1][0] = -J[1][0] * rd;
        inv[1][1] = J[0][0] * rd;
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
        G.g[d][0] = s;
#pragma unroll
        for (int m = 0; m < DIM; ++m) G.g[d][m + 1] = inv[m][d];
    }
    G.V = det * REF;
}
__device__ __forceinline__ int findSlotByte(unsigned packed, int v) {
    const unsigned eq = __vcmpeq4(packed, (unsigned)v * 0x01010101u);
    return (__ffs(eq) - 1) >> 3;
}
template <int LPN>
__device__ __forceinline__ void stageRecords(const WcArgsS& a, int i, int sub, bool valid, const double* __restrict__ X4,
                                             const double* __restrict__ V4, WcRec* recs) {
    if (valid) {
        const int nb0 = a.nbrPtr[i], nb = a.nbrPtr[i + 1] - nb0;
        for (int s = sub; s < nb; s += LPN) {
            const int nd = a.nbr[nb0 + s];
            const double* xp = X4 + (size_t)nd * 4;
            const double* vp = V4 + (size_t)nd * 4;
            const double2 x01 = ld2(xp), x23 = ld2(xp + 2), v01 = ld2(vp), v23 = ld2(vp + 2);
            WcRec& R = recs[s];
            *reinterpret_cast<double2*>(R.x) = x01;
            *reinterpret_cast<double2*>(R.x + 2) = x23;
            *reinterpret_cast<double2*>(R.v) = v01;
            *reinterpret_cast<double2*>(R.v + 2) = v23;
        }
    }
    __syncwarp();
}

template <int DIM, int LPN, int MINB>
__global__ void __launch_bounds__(256, MINB) k_wc_cont_s(const WcArgsS a, const double* __restrict__ X4,
                                                         const double* __restrict__ V4, double* __restrict__ X4n,
                                                         double* __restrict__ V4n) {
    constexpr int NPE = DIM + 1;
    constexpr double PHI = 1.0 / ((DIM + 1) * (DIM + 2));
    extern __shared__ __align__(16) unsigned char smemWc[];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = t / LPN, sub = t % LPN;
    const bool valid = i < a.nNodes;
    const double dtStep = a.dtPtr ? *a.dtPtr : a.dt;
    WcRec* recs = reinterpret_cast<WcRec*>(smemWc) + (size_t)(threadIdx.x / LPN) * a.nbcap;
    stageRecords<LPN>(a, i, sub, valid, X4, V4, recs);
    double m = 0, F0 = 0;
    int si = 0;
    if (valid) {
        si = a.diagSlot[i];
        const int eb = a.n2ePtr[i], ne = a.n2ePtr[i + 1] - eb;
        for (int k = sub; k < ne; k += LPN) {
            const unsigned packed = a.n2eSlots[eb + k];
            double P[NPE], vel[NPE][DIM], rho[NPE];
            ElemGeo<DIM> G;
            elemFromRecs<DIM>(recs, packed, P, vel, rho, G);
            const int li = findSlotByte(packed, si);
            double sumP = 0, divv = 0;
#pragma unroll
            for (int q = 0; q < NPE; ++q) {
                sumP += P[q];
#pragma unroll
                for (int c = 0; c < DIM; ++c) divv += G.g[c][q] * vel[q][c];
            }
            const double pi = pick<NPE>(P, li);
            const double sI = a.K0 / NPE + a.K0p * PHI * (pi + sumP);
            const double stab = a.meduri ? G.V * PHI * (pi + sumP) : (G.V / NPE) * pi;
            F0 += -dtStep * G.V * sI * divv + stab;
            m += G.V / NPE;
        }
    }
    m = groupSum<LPN>(m);
    F0 = groupSum<LPN>(F0);
    if (valid && sub == 0) {
        const bool isFree = a.flags[i] & PFEM_NODE_FREE;
        double inv = 1.0 / m;
        if (isFree) {
            F0 = 0.0;
            inv = 1.0;
        }
        const double p = inv * F0;
        const double rho = pow((a.K0p / a.K0) * p + 1.0, 1.0 / a.K0p) * a.rhoStar;
        const WcRec& R = recs[si];
        st4(X4n + (size_t)i * 4, R.x[0], R.x[1], R.x[2], p);
        st4(V4n + (size_t)i * 4, R.v[0], R.v[1], R.v[2], rho);
    }
}

template <int DIM, int LPN, int MINB>
__global__ void __launch_bounds__(256, MINB) k_wc_mom_s(const WcArgsS a, const double* __restrict__ X4,
                                                        const double* __restrict__ V4, double* __restrict__ V4out,
                                                        double* __restrict__ A4out, double* __restrict__ X4out) {
    constexpr int NPE = DIM + 1;
    constexpr double PHI = 1.0 / ((DIM + 1) * (DIM + 2));
    extern __shared__ __align__(16) unsigned char smemWc[];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = t / LPN, sub = t % LPN;
    const bool valid = i < a.nNodes;
    const double dtStep = a.dtPtr ? *a.dtPtr : a.dt;
    WcRec* recs = reinterpret_cast<WcRec*>(smemWc) + (size_t)(threadIdx.x / LPN) * a.nbcap;
    stageRecords<LPN>(a, i, sub, valid, X4, V4, recs);
    double M = 0, F[DIM];
#pragma unroll
    for (int c = 0; c < DIM; ++c) F[c] = 0;
    int si = 0;
    if (valid) {
        si = a.diagSlot[i];
        const int eb = a.n2ePtr[i], ne = a.n2ePtr[i + 1] - eb;
        for (int k = sub; k < ne; k += LPN) {
            const unsigned packed = a.n2eSlots[eb + k];
            double P[NPE], vel[NPE][DIM], rho[NPE];
            ElemGeo<DIM> G;
            elemFromRecs<DIM>(recs, packed, P, vel, rho, G);
            const int li = findSlotByte(packed, si);
            double sumP = 0, sumR = 0;
            double Gm[DIM][DIM];
#pragma unroll
            for (int aa = 0; aa < DIM; ++aa)
#pragma unroll
                for (int c = 0; c < DIM; ++c) Gm[aa][c] = 0;
#pragma unroll
            for (int q = 0; q < NPE; ++q) {
                sumP += P[q];
                sumR += rho[q];
#pragma unroll
                for (int aa = 0; aa < DIM; ++aa)
#pragma unroll
                    for (int c = 0; c < DIM; ++c) Gm[aa][c] += vel[q][aa] * G.g[c][q];
            }
            double tr = 0;
#pragma unroll
            for (int aa = 0; aa < DIM; ++aa) tr += Gm[aa][aa];
            double gi[DIM];
#pragma unroll
            for (int c = 0; c < DIM; ++c) gi[c] = pick<NPE>(G.g[c], li);
            const double pbar = sumP / NPE;
            const double li_mass = G.V * PHI * (pick<NPE>(rho, li) + sumR);
#pragma unroll
            for (int aa = 0; aa < DIM; ++aa) {
                double sg = 0;
#pragma unroll
                for (int c = 0; c < DIM; ++c) {
                    double sig = Gm[aa][c] + Gm[c][aa];
                    if (c == aa) sig -= (2.0 / 3.0) * tr;
                    sg += a.mu * sig * gi[c];
                }
                F[aa] += -G.V * sg + G.V * pbar * gi[aa] + a.body[aa] * li_mass;
            }
            M += li_mass;
        }
    }
    M = groupSum<LPN>(M);
#pragma unroll
    for (int c = 0; c < DIM; ++c) F[c] = groupSum<LPN>(F[c]);
    if (valid && sub == 0) {
        const uint8_t fl = a.flags[i];
        const bool isFree = fl & PFEM_NODE_FREE, isBound = fl & PFEM_NODE_BOUND;
        const WcRec& R = recs[si];
        double inv = 1.0 / M;
        double acc[3] = {0, 0, 0}, vn[3] = {0, 0, 0};
#pragma unroll
        for (int c = 0; c < DIM; ++c) {
            double f = F[c], iv = inv;
            if (a.fst4) f += a.fst4[(size_t)i * 4 + c];  // facet loop of m_applyBC (MomEquation.inl:312-336)
            if (isFree && !isBound) {
                f = a.body[c];
                iv = 1.0;
            } else if (isBound && a.dirMask[i]) {
                f = a.dirVal4[(size_t)i * 4 + c];  // reference hazard 10
                iv = 1.0;
            }
            acc[c] = iv * f;
            vn[c] = R.v[c] + 0.5 * dtStep * acc[c];
        }
        st4(V4out + (size_t)i * 4, vn[0], vn[1], vn[2], R.v[3]);
        st4(A4out + (size_t)i * 4, acc[0], acc[1], acc[2], 0.0);
        st4(X4out + (size_t)i * 4, R.x[0], R.x[1], R.x[2], R.x[3]);
    }
}

// CFL (Solver.cpp:192-234) with Element::getRin (Element.cpp:226-294): one thread per element, block min -> partial
template <int DIM>
__global__ void __launch_bounds__(256) k_wc_dt(int nElems, const int* __restrict__ conn, const double* __restrict__ X4,
                                               const double* __restrict__ V4, double mu, double K0, double K0p, double sc2,
                                               double* __restrict__ partial, double thKoverCv = 0.0) {
    constexpr int NPE = DIM + 1;
    double best = 1.7976931348623157e308;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nElems; e += gridDim.x * blockDim.x) {
        int nd[NPE];
#pragma unroll
        for (int m = 0; m < NPE; ++m) nd[m] = conn[(size_t)e * NPE + m];
        double px[NPE][3], mx = 0, alphaMax = 0;
#pragma unroll
        for (int m = 0; m < NPE; ++m) {
            const double* xp = X4 + (size_t)nd[m] * 4;
            const double* vp = V4 + (size_t)nd[m] * 4;
            const double2 x01 = ld2(xp), x23 = ld2(xp + 2), v01 = ld2(vp), v23 = ld2(vp + 2);
            px[m][0] = x01.x, px[m][1] = x01.y, px[m][2] = x23.x;
            double u2 = v01.x * v01.x + v01.y * v01.y;
            if (DIM == 3) u2 += v23.x * v23.x;
            const double c2 = (K0 + K0p * x23.y) / v23.y;
            double alpha = mu / v23.y;
            if (thKoverCv > 0.0) alpha = nanMax(thKoverCv / v23.y, alpha);  // thermal diffusivity k/(cv rho), Solver.cpp:214-216
            mx = nanMax(nanMax(u2, c2), mx);
            alphaMax = nanMax(alpha * alpha, alphaMax);
        }
        const double he = elemHe<DIM>(px);
        // max over nodes of max(u2, c2, 4 alpha^2/he^2): he is per element, so the alpha term can be taken outside
        mx = nanMax(4 * alphaMax / (he * he), mx);
        const double cand = sc2 * he * he / mx;
        best = (cand < best || cand != cand) ? cand : best;  // NaN propagates (Solver.cpp:231-232)
    }
    __shared__ double sh[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double other = __shfl_xor_sync(0xffffffffu, best, o);
        best = (other < best || other != other) ? other : best;
    }
    if (lane == 0) sh[w] = best;
    __syncthreads();
    if (w == 0) {
        double b2 = lane < (blockDim.x >> 5) ? sh[lane] : 1.7976931348623157e308;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double other = __shfl_xor_sync(0xffffffffu, b2, o);
            b2 = (other < b2 || other != other) ? other : b2;
        }
        if (lane == 0) partial[blockIdx.x] = b2;
    }
}
// CFL after a two-pass / tile step, NODAL form.  The reference takes, per element, he^2 over the largest nodal
// max(u^2, c^2, 4 alpha^2/he^2) and then the minimum over the elements (Solver.cpp:200-226).  Dividing by a maximum is the
// minimum of the quotients, so the same set of candidates is   min over (element e, node n of e) of
//     ((s^2 he) he) / w_n      and      ((s^2 he) he) / (4 alpha_n^2 / (he he)),
// and every floating-point operation in them is monotone in he: for a fixed node the smallest candidate is the one of its
// smallest incident he.  So the pass runs over NODES with hmin_n = min_{e containing n} he_e (stored by the continuity node
// pass) and the nodal w_n, alpha_n^2 (stored by the momentum epilogue): same candidates, same minimum, bit for bit, without
// touching the connectivity.  On a partitioned mesh every rank covers its owned nodes (all their elements are local).
__global__ void __launch_bounds__(256) k_wc_dt_nodal(int nRows, const double* __restrict__ hmin, const double* __restrict__ cfl2,
                                                     double sc2, double* __restrict__ partial) {
    double best = 1.7976931348623157e308;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nRows; i += gridDim.x * blockDim.x) {
        const double he = hmin[i];
        if (he == 1.7976931348623157e308) continue;  // node without elements: no candidate
        const double2 c = ld2(cfl2 + (size_t)i * 2);
        const double x = sc2 * he * he;
        const double c1 = x / c.x, c2 = x / (4 * c.y / (he * he));
        const double cand = nanMin(c1, c2);
        best = nanMin(cand, best);
    }
    __shared__ double sh[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = nanMin(__shfl_xor_sync(0xffffffffu, best, o), best);
    if (lane == 0) sh[w] = best;
    __syncthreads();
    if (w == 0) {
        double b2 = lane < (blockDim.x >> 5) ? sh[lane] : 1.7976931348623157e308;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) b2 = nanMin(__shfl_xor_sync(0xffffffffu, b2, o), b2);
        if (lane == 0) partial[blockIdx.x] = b2;
    }
}
__global__ void k_min_final(const double* __restrict__ partial, int n, double* __restrict__ out) {
    double best = 1.7976931348623157e308;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        const double o = partial[k];
        best = (o < best || o != o) ? o : best;
    }
    __shared__ double sh[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double other = __shfl_xor_sync(0xffffffffu, best, o);
        best = (other < best || other != other) ? other : best;
    }
    if (lane == 0) sh[w] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k) {
            const double o = sh[k];
            best = (o < best || o != o) ? o : best;
        }
        out[0] = best;
    }
}

// chained steps: dtDev[0] = current dt, dtDev[1] = elapsed simulated time, dtDev[2] = NaN flag.  Runs after k_wc_dt of
// step n: elapsed += dt_n ; dt_{n+1} = min(sqrt(min_e ...), maxDT)   (Solver.cpp:228, Problem::updateTime)
__global__ void k_dt_chain(const double* __restrict__ partial, int n, double maxDT, double* __restrict__ dtDev) {
    double best = 1.7976931348623157e308;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        const double o = partial[k];
        best = (o < best || o != o) ? o : best;
    }
    __shared__ double sh[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double other = __shfl_xor_sync(0xffffffffu, best, o);
        best = (other < best || other != other) ? other : best;
    }
    if (lane == 0) sh[w] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k) {
            const double o = sh[k];
            best = (o < best || o != o) ? o : best;
        }
        // fmin would return maxDT for a NaN minimum: flag it first, like the reference's "NaN time step!" (Solver.cpp:228-232)
        const double root = sqrt(best);
        const double dtNew = (root != root) ? root : fmin(root, maxDT);
        dtDev[1] += dtDev[0];
        if (dtNew != dtNew) dtDev[2] = 1.0;
        dtDev[0] = dtNew;
    }
}

#include "wc_tile.cuh"

// Node order of the explicit step for the current topology / partition (per remesh): owned nodes sorted by (interface
// node first, cell of a uniform grid over their coordinates) with a counting sort.  Nothing the ABI sees is renumbered: the
// order only decides which nodes a launch covers -- the interface nodes first, so that their halo exchange overlaps the
// rest -- and, for the tile kernels, which nodes share a CTA.
void buildNodeOrder(pfem_ctx* c) {
    PhaseScope ph(c, "Build node order");
    const int n = c->nRows, dim = c->dim;
    // bounding box of the owned nodes -> uniform grid with ~64 nodes per cell (locality only: any grouping is correct)
    c->scal.reserve(SC_COUNT);
    double* boxDev = c->scal.p;
    k_tile_bbox<<<1, 1024, 0, c->stream>>>(c->X4.p, n, dim, boxDev);
    LAUNCH_CHECK(c);
    double hb[6];
    CUDA_CHECK(cudaMemcpyAsync(hb, boxDev, 6 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    double ext[3] = {0, 0, 0}, vol = 1.0, maxExt = 0.0;
    for (int d = 0; d < dim; ++d) {
        ext[d] = std::max(hb[3 + d] - hb[d], 0.0);
        maxExt = std::max(maxExt, ext[d]);
    }
    if (!(maxExt > 0.0)) maxExt = 1.0;
    for (int d = 0; d < dim; ++d) vol *= std::max(ext[d], 1e-3 * maxExt);
    double Hc = std::pow(vol / std::max(n, 1) * 64.0, 1.0 / dim);
    int nx, ny, nz;
    for (;;) {
        nx = (int)std::floor(ext[0] / Hc) + 1;
        ny = (int)std::floor(ext[1] / Hc) + 1;
        nz = dim == 3 ? (int)std::floor(ext[2] / Hc) + 1 : 1;
        if ((double)nx * ny * nz <= 8.0e6) break;
        Hc *= 1.5;
    }
    const int nCells = nx * ny * nz, nBins = 2 * nCells;
    c->tileIface.reserve((size_t)c->nNodes + 4);
    CUDA_CHECK(cudaMemsetAsync(c->tileIface.p, 0, (size_t)c->nNodes, c->stream));
    if (c->nRanks > 1 && c->plan.nSendTotal > 0) {
        k_tile_mark<<<divUp(c->plan.nSendTotal, 256), 256, 0, c->stream>>>(c->plan.sendIdx.p, c->plan.nSendTotal, c->tileIface.p);
        LAUNCH_CHECK(c);
    }
    c->tileKey.reserve((size_t)n + (size_t)2 * nBins + 16);
    int* key = c->tileKey.p;
    int* binPtr = key + n;               // nBins + 1
    int* cursor = binPtr + nBins + 2;    // nBins
    CUDA_CHECK(cudaMemsetAsync(binPtr, 0, ((size_t)2 * nBins + 8) * sizeof(int), c->stream));
    c->tilePerm.reserve((size_t)n + 4);
    if (n > 0) {
        k_tile_key<<<divUp(n, 256), 256, 0, c->stream>>>(c->X4.p, n, dim, hb[0], hb[1], hb[2], 1.0 / Hc, nx, ny, nz, c->tileIface.p, key, binPtr);
        LAUNCH_CHECK(c);
    }
    exclusiveScanInt(c, binPtr, nBins + 1, nullptr);
    if (n > 0) {
        k_tile_fill<<<divUp(n, 256), 256, 0, c->stream>>>(n, key, binPtr, cursor, c->tilePerm.p);
        LAUNCH_CHECK(c);
        k_tile_sort_bins<<<divUp(nBins, 128), 128, 0, c->stream>>>(nBins, binPtr, c->tilePerm.p);
        LAUNCH_CHECK(c);
    }
    int nIface = 0;
    CUDA_CHECK(cudaMemcpyAsync(&nIface, binPtr + nCells, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->nIfaceNodes = c->nRanks > 1 ? nIface : 0;
    if (c->nRanks > 1 && !c->commStream) {
        CUDA_CHECK(cudaStreamCreateWithFlags(&c->commStream, cudaStreamNonBlocking));
        CUDA_CHECK(cudaEventCreateWithFlags(&c->evTile, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&c->evHalo, cudaEventDisableTiming));
    }
    c->orderValid = true;
}

// (re)build the node tiles of the fused explicit step for the current topology / partition (per remesh)
void buildTiles(pfem_ctx* c) {
    if (!c->orderValid) buildNodeOrder(c);
    PhaseScope ph(c, "Build tiles");
    const int n = c->nRows, dim = c->dim, npe = dim + 1;
    (void)dim;
    static const int envT = getenv("PFEM_WC_TILE") ? atoi(getenv("PFEM_WC_TILE")) : 64;
    int T = std::max(8, std::min(64, envT));
    const int maxE = std::max(c->maxE, 1);
    while (T > 8 && (int64_t)T * maxE > 8192) T >>= 1;
    PFEM_REQUIRE((int64_t)T * maxE <= 16384, PFEM_ERR_INVALID, "wc tiles: node valence too large");
    c->tileNePrefix.reserve((size_t)n + 4);
    if (n > 0) {
        k_tile_valence<<<divUp(n, 256), 256, 0, c->stream>>>(n, c->tilePerm.p, c->n2ePtr.p, c->tileNePrefix.p);
        LAUNCH_CHECK(c);
    }
    CUDA_CHECK(cudaMemsetAsync(c->tileNePrefix.p + n, 0, sizeof(int), c->stream));
    c->scratchI.reserve(64);
    int* misc = c->scratchI.p;  // [0] total incidences, [1] max elements per tile
    CUDA_CHECK(cudaMemsetAsync(misc, 0, 4 * sizeof(int), c->stream));
    exclusiveScanInt(c, c->tileNePrefix.p, n + 1, misc);
    int h[2] = {0, 0};
    const int nIface = c->nIfaceNodes;
    CUDA_CHECK(cudaMemcpyAsync(h, misc, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    const int64_t nConn = (int64_t)c->nElems * npe;
    c->tileElems.reserve((size_t)h[0] + 8);
    c->tileLconn.reserve((size_t)h[0] * 4 + 8);
    c->tileDst16.reserve((size_t)h[0] * 4 + 8);
    int nodeListCap = 0;
    for (;;) {
        const int nTiles = divUp(std::max(n, 1), T);
        int CAP = 256;
        while (CAP < T * maxE) CAP <<= 1;
        c->tileCnt.reserve((size_t)3 * nTiles + 8);  // element count, node-list start, node-list count per tile
        int* nodeStart = c->tileCnt.p + nTiles + 2;
        int* nodeCnt = nodeStart + nTiles + 2;
        if (nodeListCap == 0) nodeListCap = nTiles * 8 * T;
        c->tileNodes.reserve((size_t)nodeListCap + 8);
        CUDA_CHECK(cudaMemsetAsync(misc + 1, 0, 6 * sizeof(int), c->stream));
        const size_t smem = (size_t)6 * CAP * sizeof(int);
        CUDA_CHECK(cudaFuncSetAttribute(k_tile_build, cudaFuncAttributeMaxDynamicSharedMemorySize, PFEM_SMEM_OPTIN));
        if (n > 0) {
            k_tile_build<<<nTiles, 256, smem, c->stream>>>(T, n, npe, CAP, c->tilePerm.p, c->tileNePrefix.p, c->n2ePtr.p, c->n2e.p, c->conn.p,
                                                          c->tileElems.p, c->tileCnt.p, c->tileDst16.p, misc + 1, nodeStart, nodeCnt,
                                                          c->tileNodes.p, nodeListCap, c->tileLconn.p);
            LAUNCH_CHECK(c);
        }
        int hm[5] = {0, 0, 0, 0, 0};  // max elements / listed nodes per tile, total node-list entries, overflow, max record slots
        CUDA_CHECK(cudaMemcpyAsync(hm, misc + 1, 5 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        if (hm[2] > nodeListCap) {  // the node lists did not fit the guess: the cursor holds the exact total
            nodeListCap = hm[2];
            continue;
        }
        c->tileT = T, c->nTiles = n > 0 ? nTiles : 0, c->tileCap = (hm[4] + 1) & ~1, c->tileNodeCap = (hm[1] + 1) & ~1, c->tileMaxElems = hm[0];
        c->tileNodeStart = nodeStart, c->tileNodeCnt = nodeCnt;
        // staged node records (80 B per listed node) + one 32-byte record per (tile node, incident element) incidence
        const size_t need = (size_t)c->tileNodeCap * 80 + (size_t)c->tileCap * 32;
        if ((need <= 200 * 1024 && hm[3] == 0) || T <= 8) {
            PFEM_REQUIRE(hm[3] == 0, PFEM_ERR_INVALID, "wc tiles: node list of a tile exceeds its build buffer");
            break;
        }
        T >>= 1;
        nodeListCap = 0;
    }
    PFEM_REQUIRE((size_t)c->tileNodeCap * 80 + (size_t)c->tileCap * 32 <= 225 * 1024, PFEM_ERR_INVALID,
                 "wc tiles: tile lists too large for shared memory");
    c->nIfaceTiles = c->nRanks > 1 ? std::min(c->nTiles, divUp(nIface, T)) : 0;
    c->tilesValid = true;
    static const bool verbose = getenv("PFEM_WC_VERBOSE") != nullptr;
    if (verbose)
        fprintf(stderr, "[pfem wc tiles] rank %d: %d owned nodes, T=%d, %d tiles (%d interface), max %d elements / %d listed nodes / "
                        "%d record slots, element-list entries %d for %d elements\n",
                c->rank, n, c->tileT, c->nTiles, c->nIfaceTiles, c->tileMaxElems, c->tileNodeCap, c->tileCap, h[0], c->nElems);
}

// PFEM_WC_CFG: 10 (default) chooses by size -- two-pass continuity + momentum + CFL-from-stored-values on meshes of >= 200 k
// elements (below that the step is launch-bound and the 5-launch gather path wins) and, on a partitioned mesh, >= 4 M
// local elements (measured at C5: 2 GPUs 2.20 -> 1.57 ms/step, but 8 GPUs 0.89 -> 1.00 ms: with 2.5 M elements per rank
// the step is exchange-latency bound and the extra launches cost more than the kernels save) | 11: two-pass always |
// 12: two-pass continuity, gather momentum | 6: direct gathers, 4 lanes per node | 0: 8 lanes per node | 7: staged records
// resident 256-thread blocks per SM the element kernels are compiled for: 3 (80 registers, default) | 2 (100) | 4 (64, spills)
int wcElemBlocks() {
    static const int b = getenv("PFEM_WC_EB") ? atoi(getenv("PFEM_WC_EB")) : 3;
    return b;
}
int wcCfgRaw() {
    static const int cfg = getenv("PFEM_WC_CFG") ? atoi(getenv("PFEM_WC_CFG")) : 10;
    return cfg;
}
bool wcTwoPass(const pfem_ctx* c) {
    const int raw = c->wcVariant ? c->wcVariant : wcCfgRaw();  // pfem_wc_set_variant overrides the environment
    // BoussinesqWC: the two-pass kernels carry the buoyancy factor too (round 2); the mixed variant 12 (gather momentum)
    // and the tiles do not
    if (c->thermalOn && (raw == 12 || raw == 13)) return raw == 12;
    if (raw == 11 || raw == 12 || raw == 13) return true;
    if (raw != 10) return false;
    return c->nElems >= 200000;
}
// tiles (element records in shared memory) instead of element records in HBM: the default at size, CDS_dpdt only
bool wcTiles(const pfem_ctx* c, const pfem_wc_params& p) {
    const int raw = c->wcVariant ? c->wcVariant : wcCfgRaw();
    if (p.eqType != PFEM_WC_CDS_DPDT || c->maxE > 255 || c->thermalOn) return false;
    return raw == 13;  // opt-in: measured slower than the two-pass kernels on B200
}
// Overlap of the halo exchanges with the interior node pass (interface nodes first, second stream): opt-in with
// PFEM_WC_OVERLAP=1.  A/B on 4 B200s at C5 (tools/wc_overlap_ab.py, 20 steps, twice each): 0.780 / 0.780 ms per step
// serialised against 0.887 / 0.827 ms overlapped -- the interface-first node order costs the node passes 0.05 ms of
// coalescing and the split passes two more launches each, which is what the two hidden exchanges (0.055 ms each) are worth.
bool wcOverlap() {
    static const bool v = getenv("PFEM_WC_OVERLAP") && atoi(getenv("PFEM_WC_OVERLAP")) == 1;
    return v;
}
bool wcTwoPassMom(const pfem_ctx* c) { return (c->wcVariant ? c->wcVariant : wcCfgRaw()) != 12; }
int wcCfg(const pfem_ctx* c) {
    const int raw = c->wcVariant ? c->wcVariant : wcCfgRaw();
    if (c->thermalOn && !wcTwoPass(c)) return 6;  // BoussinesqWC below the two-pass size: the gather kernels with the buoyancy factor
    return wcTwoPass(c) ? 10 : (raw >= 10 ? 6 : raw);
}

// kick + continuity + momentum of one step on the context's stream; dtPtr != null: dt is read from the device
void launchStep(pfem_ctx* c, const pfem_wc_params& p, double dt, const double* dtPtr) {
    const int cfg = wcCfg(c);
    c->cflMu = p.mu, c->cflK0 = p.K0, c->cflK0p = p.K0p;
    WcArgs a;
    a.conn = c->conn.p, a.n2ePtr = c->n2ePtr.p, a.n2e = c->n2e.p, a.flags = c->flags.p;
    a.dirMask = c->dirMask.p, a.dirVal4 = c->dirVal4.p, a.nNodes = c->nRows;  // rows = owned nodes
    a.dt = dt, a.dtPtr = dtPtr, a.mu = p.mu, a.K0 = p.K0, a.K0p = p.K0p, a.rhoStar = p.rhoStar, a.meduri = p.meduri;
    WcArgsS as;
    as.n2ePtr = c->n2ePtr.p, as.n2eSlots = c->n2eSlots.p, as.nbrPtr = c->nbrPtr.p, as.nbr = c->nbr.p, as.diagSlot = c->diagSlot.p;
    as.flags = c->flags.p, as.dirMask = c->dirMask.p, as.dirVal4 = c->dirVal4.p, as.nNodes = c->nRows;
    as.nbcap = std::max(c->maxNb, 1);
    as.dt = dt, as.dtPtr = dtPtr, as.mu = p.mu, as.K0 = p.K0, as.K0p = p.K0p, as.rhoStar = p.rhoStar, as.meduri = p.meduri;
    for (int d = 0; d < 3; ++d) a.body[d] = as.body[d] = p.bodyForce[d];
#define PFEM_WC_LAUNCH_RHO(MODE_, ...)                                                                    \
    do {                                                                                                  \
        const int grid_ = divUp((int64_t)c->nRows * 4, 256);                                              \
        if (c->dim == 2) k_wc_cont_rho<2, 4, 2, MODE_><<<grid_, 256, 0, c->stream>>>(a, __VA_ARGS__);       \
        else k_wc_cont_rho<3, 4, 2, MODE_><<<grid_, 256, 0, c->stream>>>(a, __VA_ARGS__);                   \
    } while (0)
    if (p.eqType == PFEM_WC_CDS_RHO) {  // ContEqWCompNewton::preCompute, before the kick and the move (Solver.cpp:244-246)
        PhaseScope ph(c, "Solving continuity eq");
        PFEM_WC_LAUNCH_RHO(3, c->X4.p, c->V4.p, nullptr, nullptr, c->wcF0.p);
        LAUNCH_CHECK(c);
    }
    {
        PhaseScope ph(c, "Update solutions");
        k_wc_kick_move<<<divUp(c->nNodes, 256), 256, 0, c->stream>>>(c->nNodes, c->dim, dt, dtPtr, c->flags.p, c->X4.p, c->V4.p,
                                                                     c->A4.p);
        LAUNCH_CHECK(c);
    }
    a.fst4 = as.fst4 = facetsForces(c, c->X4.p, true);  // on the moved mesh; Facet::isOnFreeSurface = all nodes (Facet.cpp:249-255)
    if (c->thermalOn) {  // m_solveBoussinesqWC: heat -> continuity -> momentum (Solver.cpp:278-320)
        thermalWcHeat(c, dt, dtPtr);
        a.T = c->Tn.p, a.thAlpha = c->thAlpha, a.thTr = c->thTr;
    }
#define PFEM_WC_LAUNCH_S(KERNEL, LPN_, ...)                                                                         \
    do {                                                                                                            \
        const int grid_ = divUp((int64_t)c->nRows * LPN_, 256);                                                     \
        const size_t smem_ = (size_t)(256 / LPN_) * as.nbcap * sizeof(WcRec);                                       \
        PFEM_REQUIRE(smem_ <= 200 * 1024, PFEM_ERR_INVALID, "wc_step: node valence too large for shared memory");   \
        if (c->dim == 2) {                                                                                          \
            if (smem_ > 40 * 1024)                                                                                  \
                CUDA_CHECK(cudaFuncSetAttribute(KERNEL<2, LPN_, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, PFEM_SMEM_OPTIN)); \
            KERNEL<2, LPN_, 2><<<grid_, 256, smem_, c->stream>>>(as, __VA_ARGS__);                                   \
        } else {                                                                                                    \
            if (smem_ > 40 * 1024)                                                                                  \
                CUDA_CHECK(cudaFuncSetAttribute(KERNEL<3, LPN_, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, PFEM_SMEM_OPTIN)); \
            KERNEL<3, LPN_, 2><<<grid_, 256, smem_, c->stream>>>(as, __VA_ARGS__);                                   \
        }                                                                                                           \
    } while (0)
#define PFEM_WC_LAUNCH(KERNEL, LPN_, ...)                                                     \
    do {                                                                                      \
        const int grid_ = divUp((int64_t)c->nRows * LPN_, 256);                               \
        if (c->dim == 2) KERNEL<2, LPN_, 2><<<grid_, 256, 0, c->stream>>>(a, __VA_ARGS__);      \
        else KERNEL<3, LPN_, 2><<<grid_, 256, 0, c->stream>>>(a, __VA_ARGS__);                  \
    } while (0)
    const bool tiles = wcTiles(c, p);
    // two-pass node passes of a partitioned mesh run over the node order (interface nodes first) when it exists
    const bool useOrder = c->nRanks > 1 && c->orderValid && !c->local && wcOverlap();
    bool twoPassContDone = false;
    TileArgs ta;
    ta.perm = c->tilePerm.p, ta.nePrefix = c->tileNePrefix.p, ta.tileCnt = c->tileCnt.p, ta.tileElems = c->tileElems.p;
    ta.dst16 = c->tileDst16.p, ta.T = c->tileT, ta.nRows = c->nRows, ta.tile0 = 0;
    ta.nodeStart = c->tileNodeStart, ta.nodeCnt = c->tileNodeCnt, ta.tileNodes = c->tileNodes.p, ta.lconn = c->tileLconn.p;
    ta.nodeCap = std::max(c->tileNodeCap, 2), ta.slotCap = std::max(c->tileCap, 2);
    // tile passes of a partitioned mesh: the tiles holding interface nodes run first; the exchange of their results
    // (commStream) overlaps the interior tiles; the main stream joins before the next pass reads the ghosts
    auto splitPass = [&](int nFirst, int nTotal, auto launchRange, auto exchange) {
        // (with phase timers on, the pass runs serialised so that every phase is timed on one stream)
        const bool split = c->nRanks > 1 && c->commStream && !c->local && wcOverlap() && !c->profiling;
        const int nI = split ? nFirst : 0;
        if (nI > 0) launchRange(0, nI);
        if (split) CUDA_CHECK(cudaEventRecord(c->evTile, c->stream));
        if (nTotal > nI) launchRange(nI, nTotal - nI);
        if (c->nRanks > 1) {
            if (split) {
                cudaStream_t mainStream = c->stream;
                CUDA_CHECK(cudaStreamWaitEvent(c->commStream, c->evTile, 0));
                c->stream = c->commStream;
                try {
                    exchange();
                } catch (...) {
                    c->stream = mainStream;
                    throw;
                }
                c->stream = mainStream;
                CUDA_CHECK(cudaEventRecord(c->evHalo, c->commStream));
                CUDA_CHECK(cudaStreamWaitEvent(c->stream, c->evHalo, 0));
            } else
                exchange();
        }
    };
    {
        PhaseScope ph(c, "Solving continuity eq");
        if (tiles) {
            auto range = [&](int t0, int nt) {
                TileArgs tb = ta;
                tb.tile0 = t0;
                const size_t smem = (size_t)ta.nodeCap * 80 + (size_t)std::max(c->tileCap, 2) * 32;
                if (c->dim == 2) {
                    CUDA_CHECK(cudaFuncSetAttribute(k_wc_cont_tile<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, PFEM_SMEM_OPTIN));
                    k_wc_cont_tile<2><<<nt, 256, smem, c->stream>>>(tb, a, c->X4.p, c->V4.p, c->X4b.p, c->V4b.p, c->wcHmin.p);
                } else {
                    CUDA_CHECK(cudaFuncSetAttribute(k_wc_cont_tile<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, PFEM_SMEM_OPTIN));
                    k_wc_cont_tile<3><<<nt, 256, smem, c->stream>>>(tb, a, c->X4.p, c->V4.p, c->X4b.p, c->V4b.p, c->wcHmin.p);
                }
                LAUNCH_CHECK(c);
            };
            splitPass(c->nIfaceTiles, c->nTiles, range, [&]() { commHalo(c, c->X4b.p, c->V4b.p, 4); });
        } else
        if (p.eqType == PFEM_WC_CDS_DRHODT) PFEM_WC_LAUNCH_RHO(1, c->X4.p, c->V4.p, c->X4b.p, c->V4b.p, nullptr);
        else if (p.eqType == PFEM_WC_CDS_RHO) PFEM_WC_LAUNCH_RHO(2, c->X4.p, c->V4.p, c->X4b.p, c->V4b.p, c->wcF0.p);
        else if (cfg == 10) {  // two-pass: element records, then the nodal gather
            const int ge = divUp(c->nElems, 256);
            if (c->dim == 2) k_wc_cont_elem<2, 3><<<ge, 256, 0, c->stream>>>(c->nElems, c->conn.p, c->X4.p, c->V4.p, dt, dtPtr, p.K0, p.K0p, p.meduri, c->wcContRec.p);
            else if (wcElemBlocks() == 4) k_wc_cont_elem<3, 4><<<ge, 256, 0, c->stream>>>(c->nElems, c->conn.p, c->X4.p, c->V4.p, dt, dtPtr, p.K0, p.K0p, p.meduri, c->wcContRec.p);
            else if (wcElemBlocks() == 2) k_wc_cont_elem<3, 2><<<ge, 256, 0, c->stream>>>(c->nElems, c->conn.p, c->X4.p, c->V4.p, dt, dtPtr, p.K0, p.K0p, p.meduri, c->wcContRec.p);
            else k_wc_cont_elem<3, 3><<<ge, 256, 0, c->stream>>>(c->nElems, c->conn.p, c->X4.p, c->V4.p, dt, dtPtr, p.K0, p.K0p, p.meduri, c->wcContRec.p);
            LAUNCH_CHECK(c);
            auto range = [&](int k0, int cnt) {
                WcArgs b = a;
                b.nNodes = cnt, b.k0 = k0, b.order = useOrder ? c->tilePerm.p : nullptr;
                const int gn = divUp((int64_t)cnt * 4, 256);
                if (c->dim == 2) k_wc_cont_node<2, 4><<<gn, 256, 0, c->stream>>>(b, c->wcContRec.p, c->X4.p, c->V4.p, c->X4b.p, c->V4b.p, c->wcHmin.p);
                else k_wc_cont_node<3, 4><<<gn, 256, 0, c->stream>>>(b, c->wcContRec.p, c->X4.p, c->V4.p, c->X4b.p, c->V4b.p, c->wcHmin.p);
                LAUNCH_CHECK(c);
            };
            splitPass(useOrder ? c->nIfaceNodes : 0, c->nRows, range, [&]() { commHalo(c, c->X4b.p, c->V4b.p, 4); });
            twoPassContDone = true;
        }
        else if (cfg == 7) PFEM_WC_LAUNCH_S(k_wc_cont_s, 8, c->X4.p, c->V4.p, c->X4b.p, c->V4b.p);
        else if (cfg == 0) PFEM_WC_LAUNCH(k_wc_cont, 8, c->X4.p, c->V4.p, c->X4b.p, c->V4b.p);
        else PFEM_WC_LAUNCH(k_wc_cont, 4, c->X4.p, c->V4.p, c->X4b.p, c->V4b.p);
        if (!tiles && !twoPassContDone) {
            LAUNCH_CHECK(c);
            if (c->nRanks > 1) commHalo(c, c->X4b.p, c->V4b.p, 4);  // (x, p_new) and (v_half, rho_new) of interface nodes
        }
    }
    if (tiles) {
        PhaseScope ph(c, "Solving momentum eq");
        auto range = [&](int t0, int nt) {
            TileArgs tb = ta;
            tb.tile0 = t0;
            const size_t smem = (size_t)ta.nodeCap * 80 + (size_t)ta.slotCap * 32;
            if (c->dim == 2) {
                CUDA_CHECK(cudaFuncSetAttribute(k_wc_mom_tile<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, PFEM_SMEM_OPTIN));
                k_wc_mom_tile<2><<<nt, 256, smem, c->stream>>>(tb, a, c->X4b.p, c->V4b.p, c->V4.p, c->A4.p, c->X4.p, c->wcCfl2.p);
            } else {
                CUDA_CHECK(cudaFuncSetAttribute(k_wc_mom_tile<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, PFEM_SMEM_OPTIN));
                k_wc_mom_tile<3><<<nt, 256, smem, c->stream>>>(tb, a, c->X4b.p, c->V4b.p, c->V4.p, c->A4.p, c->X4.p, c->wcCfl2.p);
            }
            LAUNCH_CHECK(c);
        };
        splitPass(c->nIfaceTiles, c->nTiles, range, [&]() {
            commHalo(c, c->V4.p, c->A4.p, 4);  // (v, rho) and acceleration of interface nodes
            const size_t g0 = (size_t)c->nRows * 4, gn = (size_t)(c->nNodes - c->nRows) * 4;  // ghosts: (x, p_new) from X4b
            if (gn) CUDA_CHECK(cudaMemcpyAsync(c->X4.p + g0, c->X4b.p + g0, gn * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        });
    } else {
        PhaseScope ph(c, "Solving momentum eq");
        if (cfg == 10 && !wcTwoPassMom(c)) PFEM_WC_LAUNCH(k_wc_mom, 4, c->X4b.p, c->V4b.p, c->V4.p, c->A4.p, c->X4.p, c->wcCfl2.p);
        else if (cfg == 10) {
            const int ge = divUp(c->nElems, 256);
            if (c->thermalOn) {  // BoussinesqWC: buoyancy-weighted body mass
                if (c->dim == 2) k_wc_mom_elem<2, 3, true><<<ge, 256, 0, c->stream>>>(c->nElems, c->conn.p, c->X4b.p, c->V4b.p, p.mu, p.bodyForce[0], p.bodyForce[1], p.bodyForce[2], c->wcElemRec.p, c->Tn.p, c->thAlpha, c->thTr);
                else k_wc_mom_elem<3, 3, true><<<ge, 256, 0, c->stream>>>(c->nElems, c->conn.p, c->X4b.p, c->V4b.p, p.mu, p.bodyForce[0], p.bodyForce[1], p.bodyForce[2], c->wcElemRec.p, c->Tn.p, c->thAlpha, c->thTr);
            } else if (c->dim == 2) k_wc_mom_elem<2, 3><<<ge, 256, 0, c->stream>>>(c->nElems, c->conn.p, c->X4b.p, c->V4b.p, p.mu, p.bodyForce[0], p.bodyForce[1], p.bodyForce[2], c->wcElemRec.p);
            else if (wcElemBlocks() == 4) k_wc_mom_elem<3, 4><<<ge, 256, 0, c->stream>>>(c->nElems, c->conn.p, c->X4b.p, c->V4b.p, p.mu, p.bodyForce[0], p.bodyForce[1], p.bodyForce[2], c->wcElemRec.p);
            else if (wcElemBlocks() == 2) k_wc_mom_elem<3, 2><<<ge, 256, 0, c->stream>>>(c->nElems, c->conn.p, c->X4b.p, c->V4b.p, p.mu, p.bodyForce[0], p.bodyForce[1], p.bodyForce[2], c->wcElemRec.p);
            else k_wc_mom_elem<3, 3><<<ge, 256, 0, c->stream>>>(c->nElems, c->conn.p, c->X4b.p, c->V4b.p, p.mu, p.bodyForce[0], p.bodyForce[1], p.bodyForce[2], c->wcElemRec.p);
            LAUNCH_CHECK(c);
            auto range = [&](int k0, int cnt) {
                WcArgs b = a;
                b.nNodes = cnt, b.k0 = k0, b.order = useOrder ? c->tilePerm.p : nullptr;
                const int gn = divUp((int64_t)cnt * 4, 256);
                if (c->dim == 2) k_wc_mom_node<2, 4><<<gn, 256, 0, c->stream>>>(b, c->wcElemRec.p, (size_t)c->nElems, c->n2eSlots.p, c->diagSlot.p, c->X4b.p, c->V4b.p, c->V4.p, c->A4.p, c->X4.p, c->wcCfl2.p);
                else k_wc_mom_node<3, 4><<<gn, 256, 0, c->stream>>>(b, c->wcElemRec.p, (size_t)c->nElems, c->n2eSlots.p, c->diagSlot.p, c->X4b.p, c->V4b.p, c->V4.p, c->A4.p, c->X4.p, c->wcCfl2.p);
                LAUNCH_CHECK(c);
            };
            splitPass(useOrder ? c->nIfaceNodes : 0, c->nRows, range, [&]() {
                commHalo(c, c->V4.p, c->A4.p, 4);  // (v, rho) and acceleration of interface nodes
                const size_t g0 = (size_t)c->nRows * 4, gn = (size_t)(c->nNodes - c->nRows) * 4;  // ghosts: (x, p_new) from X4b
                if (gn) CUDA_CHECK(cudaMemcpyAsync(c->X4.p + g0, c->X4b.p + g0, gn * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
            });
            return;
        }
        else if (cfg == 7) PFEM_WC_LAUNCH_S(k_wc_mom_s, 8, c->X4b.p, c->V4b.p, c->V4.p, c->A4.p, c->X4.p);
        else if (cfg == 0) PFEM_WC_LAUNCH(k_wc_mom, 8, c->X4b.p, c->V4b.p, c->V4.p, c->A4.p, c->X4.p);
        else PFEM_WC_LAUNCH(k_wc_mom, 4, c->X4b.p, c->V4b.p, c->V4.p, c->A4.p, c->X4.p);
        LAUNCH_CHECK(c);
        if (c->nRanks > 1) {
            commHalo(c, c->V4.p, c->A4.p, 4);  // (v, rho) and acceleration of interface nodes
            const size_t g0 = (size_t)c->nRows * 4, gn = (size_t)(c->nNodes - c->nRows) * 4;  // ghosts: (x, p_new) from X4b
            if (gn) CUDA_CHECK(cudaMemcpyAsync(c->X4.p + g0, c->X4b.p + g0, gn * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        }
    }
}

void launchDt(pfem_ctx* c, const pfem_wc_params& p, double securityCoeff, int grid, bool afterTwoPassStep = false) {
    const double sc2 = securityCoeff * securityCoeff;
    // the stored he / nodal CFL values belong to the two-pass step that has just run with the same material constants
    if (afterTwoPassStep && !c->thermalOn && p.eqType == PFEM_WC_CDS_DPDT && p.mu == c->cflMu && p.K0 == c->cflK0 && p.K0p == c->cflK0p) {
        PhaseScope ph(c, "CFL nodal pass");
        k_wc_dt_nodal<<<grid, 256, 0, c->stream>>>(c->nRows, c->wcHmin.p, c->wcCfl2.p, sc2, c->dtPartial.p);
        LAUNCH_CHECK(c);
        return;
    }
    PhaseScope ph(c, "CFL element pass");
    const double thKoverCv = c->thermalOn ? c->thK / c->thCv : 0.0;
    if (c->dim == 2)
        k_wc_dt<2><<<grid, 256, 0, c->stream>>>(c->nElems, c->conn.p, c->X4.p, c->V4.p, p.mu, p.K0, p.K0p, sc2, c->dtPartial.p, thKoverCv);
    else
        k_wc_dt<3><<<grid, 256, 0, c->stream>>>(c->nElems, c->conn.p, c->X4.p, c->V4.p, p.mu, p.K0, p.K0p, sc2, c->dtPartial.p, thKoverCv);
    LAUNCH_CHECK(c);
}

void checkStepArgs(pfem_ctx* c, const pfem_wc_params& p, double dt) {
    PFEM_REQUIRE(c->haveTopology && c->havePositions, PFEM_ERR_STATE, "wc_step: topology/positions missing");
    PFEM_REQUIRE(dt > 0 && p.K0 > 0 && p.K0p != 0, PFEM_ERR_INVALID, "wc_step: dt, K0 must be positive and K0p non-zero");
    PFEM_REQUIRE(p.eqType >= PFEM_WC_CDS_DPDT && p.eqType <= PFEM_WC_CDS_RHO, PFEM_ERR_INVALID, "wc_step: unknown eqType");
    PFEM_REQUIRE(p.eqType == PFEM_WC_CDS_DPDT || p.rhoStar > 0, PFEM_ERR_INVALID, "wc_step: rhoStar must be positive");
    const size_t n4 = (size_t)c->nNodes * 4;
    c->X4b.reserve(n4);
    c->V4b.reserve(n4);
    if (p.eqType == PFEM_WC_CDS_RHO) c->wcF0.reserve((size_t)c->nNodes);
    if (c->thermalOn) thermalPrepare(c);
    if (wcTiles(c, p) && !c->tilesValid) buildTiles(c);
    if (c->nRanks > 1 && !c->local && wcCfg(c) == 10 && !c->orderValid && wcOverlap()) buildNodeOrder(c);
    if (wcCfg(c) == 10) {  // before any graph capture
        c->wcElemRec.reserve((size_t)std::max(c->nElems, 1) * (c->dim + 1) * 4);
        c->wcContRec.reserve((size_t)std::max(c->nElems, 1) * 4);
        c->wcCfl2.reserve((size_t)c->nNodes * 2);
        c->wcHmin.reserve((size_t)c->nNodes + 4);
    }
}

}  // namespace

void wcStep(pfem_ctx* c, const pfem_wc_params& p, double dt) {
    checkStepArgs(c, p, dt);
    launchStep(c, p, dt, nullptr);
    c->cflFresh = (wcCfg(c) == 10);  // cleared by the next API call that is not pfem_wc_next_dt (capi.cu)
}

int wcNextDt(pfem_ctx* c, const pfem_wc_params& p, double securityCoeff, double maxDT, double* dtOut) {
    PFEM_REQUIRE(c->haveTopology && c->havePositions, PFEM_ERR_STATE, "wc_next_dt: topology/positions missing");
    PFEM_REQUIRE(dtOut, PFEM_ERR_INVALID, "wc_next_dt: null");
    PhaseScope ph(c, "Compute next dt");
    const int grid = std::max(1, std::min(c->smCount * 8, divUp(c->nElems, 256)));
    c->dtPartial.reserve(grid + 8);
    c->scal.reserve(SC_COUNT);
    if (!c->hScal) CUDA_CHECK(cudaMallocHost(&c->hScal, SC_COUNT * sizeof(double)));
    launchDt(c, p, securityCoeff, grid, c->cflFresh);
    k_min_final<<<1, 256, 0, c->stream>>>(c->dtPartial.p, grid, c->scal.p + SC_COUNT - 1);
    LAUNCH_CHECK(c);
    if (c->nRanks > 1) commAllReduceMin(c, c->scal.p + SC_COUNT - 1);
    CUDA_CHECK(cudaMemcpyAsync(c->hScal, c->scal.p + SC_COUNT - 1, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    const double ts = c->hScal[0];
    const double root = sqrt(ts);
    const double dt = (root != root) ? root : fmin(root, maxDT);  // Solver.cpp:228; a NaN minimum stays NaN
    *dtOut = dt;
    return (dt != dt || ts != ts) ? PFEM_NAN : PFEM_OK;
}

// nSteps explicit steps with the CFL time step recomputed ON THE DEVICE after every step (computeNextDT) and no host round
// trip in between: the 5 launches of a step (kick, continuity, momentum, CFL, dt chain) are captured once into a CUDA
// graph and replayed -- the small reference configurations (1-14 k elements, ~4e5 steps) are launch-latency bound.
int wcRun(pfem_ctx* c, const pfem_wc_params& p, int nSteps, double securityCoeff, double maxDT, double* dtInOut, double* elapsed) {
    PFEM_REQUIRE(dtInOut && nSteps >= 0, PFEM_ERR_INVALID, "wc_run: bad arguments");
    checkStepArgs(c, p, *dtInOut);
    const int grid = std::max(1, std::min(c->smCount * 8, divUp(c->nElems, 256)));
    c->dtPartial.reserve(grid + 8);
    c->scal.reserve(SC_COUNT);
    if (!c->hScal) CUDA_CHECK(cudaMallocHost(&c->hScal, SC_COUNT * sizeof(double)));
    double* dtDev = c->scal.p + SC_COUNT - 4;  // [dt, elapsed, nan]
    c->hScal[0] = *dtInOut, c->hScal[1] = 0.0, c->hScal[2] = 0.0;
    CUDA_CHECK(cudaMemcpyAsync(dtDev, c->hScal, 3 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (c->nRanks > 1) {
        // partitioned mesh: the same chain with the CFL minimum all-reduced on the device; nothing returns to the host between
        // steps (no CUDA graph: the exchanges are NCCL calls / local-transport copies enqueued per step)
        double* gmin = c->scal.p + SC_COUNT - 1;
        for (int s = 0; s < nSteps; ++s) {
            launchStep(c, p, *dtInOut, dtDev);
            launchDt(c, p, securityCoeff, grid, wcCfg(c) == 10);
            k_min_final<<<1, 256, 0, c->stream>>>(c->dtPartial.p, grid, gmin);
            LAUNCH_CHECK(c);
            commAllReduceMin(c, gmin);
            k_dt_chain<<<1, 256, 0, c->stream>>>(gmin, 1, maxDT, dtDev);
            LAUNCH_CHECK(c);
        }
        CUDA_CHECK(cudaMemcpyAsync(c->hScal, dtDev, 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        c->cflFresh = false;
        *dtInOut = c->hScal[0];
        if (elapsed) *elapsed = c->hScal[1];
        return (c->hScal[2] != 0.0 || c->hScal[0] != c->hScal[0]) ? PFEM_NAN : PFEM_OK;
    }
    const bool wasProfiling = c->profiling;
    c->profiling = false;  // no event records inside a capture
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    try {
        const long long launches0 = c->launches;
        CUDA_CHECK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        launchStep(c, p, *dtInOut, dtDev);
        launchDt(c, p, securityCoeff, grid, wcCfg(c) == 10);
        k_dt_chain<<<1, 256, 0, c->stream>>>(c->dtPartial.p, grid, maxDT, dtDev);
        LAUNCH_CHECK(c);
        CUDA_CHECK(cudaStreamEndCapture(c->stream, &graph));
        CUDA_CHECK(cudaGraphInstantiate(&exec, graph, 0));
        const long long perStep = c->launches - launches0;  // 5 kernels (6 with the CDS_rho pre-pass)
        c->launches = launches0;                            // the capture itself launched nothing
        for (int s = 0; s < nSteps; ++s) {
            CUDA_CHECK(cudaGraphLaunch(exec, c->stream));
            c->launches += perStep;
        }
        CUDA_CHECK(cudaMemcpyAsync(c->hScal, dtDev, 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
    } catch (...) {
        cudaStreamCaptureStatus st;
        if (cudaStreamIsCapturing(c->stream, &st) == cudaSuccess && st != cudaStreamCaptureStatusNone) {
            cudaGraph_t g2 = nullptr;
            cudaStreamEndCapture(c->stream, &g2);
            if (g2) cudaGraphDestroy(g2);
        }
        if (exec) cudaGraphExecDestroy(exec);
        if (graph) cudaGraphDestroy(graph);
        c->profiling = wasProfiling;
        throw;
    }
    cudaGraphExecDestroy(exec);
    cudaGraphDestroy(graph);
    c->profiling = wasProfiling;
    *dtInOut = c->hScal[0];
    if (elapsed) *elapsed = c->hScal[1];
    return (c->hScal[2] != 0.0 || c->hScal[0] != c->hScal[0]) ? PFEM_NAN : PFEM_OK;
}
