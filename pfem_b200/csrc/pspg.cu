// pspg.cu -- assembly of the monolithic PSPG system  A q = b  of MomContEqIncompNewton<dim>
// (MomContEquationPSPG.inl:7-146 m_buildAbPSPG, :149-235 m_applyBCPSPG, :238-259 m_computeTauPSPG) on the device.
//
// Design (B200: HBM-bound, no tensor cores -- 4x4 blocks are far too small):
//   * node-row GATHER instead of element scatter: one persistent warp owns one node i == one (dim+1)-row block row of A,
//     software-pipelined (header and per-lane words of the next node are fetched while this one is computed).
//     Phase 0: lanes = neighbours; each neighbour's nodal record (x, v_prev, |v_cur|) is staged once in shared memory.
//     Phase 1: lanes = incident elements of i; each lane rebuilds that element's geometry from the staged records
//              (Element.cpp:15-135, nothing stored per element), tau, and writes a 208-byte record (grad N, V, tau V, the
//              element's contribution to the RHS rows of i, slot bytes) to shared memory.
//     Phase 2: lane = one (dim+1)^2 block: 28 lanes own the off-diagonal blocks and walk, in ASCENDING ELEMENT ORDER (the
//              order in which setFromTriplets sums duplicates), the bit mask of the elements around edge (i,j), two
//              elements per trip; 4 lanes share the diagonal block and the RHS rows, reduced through shared memory.
//   * A is written exactly once, fully coalesced (1 KB per warp store), no atomics, bit-reproducible run to run.
//   * m_applyBCPSPG is fused into the epilogue: row masks / identity rows (PSPG.inl:68,81,108-128), Dirichlet column
//     elimination (PSPG.inl:216-228) as a per-row operation, free-node RHS (PSPG.inl:193-204), and 1/diag for the
//     Jacobi preconditioner.
// Storage: node-block CSR ("BSR") with BS=(dim+1): Aval[(nbrPtr[i]+slot)*BS*BS + r*BS + c] = A(i + r*N, nbr[slot] + c*N).
// Masked rows are held as explicit zero rows + unit diagonal internally; pfem_pspg_export_csc emits the reference's
// exact CSC pattern.
#include <cstdlib>

#include "common.cuh"

namespace {

// per-element record staged in shared memory by phase 1 (one per incident element of the node)
template <int DIM> struct alignas(16) ElemS {
    static constexpr int GP = (DIM == 3) ? 4 : 2;  // doubles per node gradient (16-byte aligned for LDS.128)
    double g[DIM + 1][GP];  // grad N_m = g[m][0..DIM)   (MatricesBuilder.inl:93-127)
    double V;               // element size detJ*ref; all five block coefficients are multiples of V and tau*V:
    double tauV;            //   M/dt: rho phi/dt V | K: mu V | D: V/npe | (tau/dt) C: tauV/(npe dt) | tau L: tauV/rho
    double be[4];           // this element's contribution to the RHS rows of node i (summed by the diagonal lanes)
    unsigned slots;         // slot byte of each local node in the neighbour list of i
    int li;                 // local index of node i in this element
    double pad_[(DIM == 3) ? 3 : 1];  // record stride = odd multiple of 16 B (208 | 112): LDS.128 of different records
                                      // spread over the banks
};
struct BlockCoef {
    double kmass, kvisc, kdiv, kpc, kL;  // rho phi/dt, mu, 1/npe, 1/(npe dt), 1/rho
};

struct AsmArgs {
    const int* conn;
    const int* n2ePtr;
    const int* n2e;
    const unsigned* n2eSlots;
    const unsigned* blkMask;
    const int* nbrPtr;
    const int* nbr;
    const int* diagSlot;
    const uint8_t* flags;
    const unsigned* rowDir;   // per node: dirWords words, bit s = neighbour slot s is a Dirichlet node
    const int4* nodeHdr;      // packed per-node header (k_row_dir)
    const uint8_t* dirMask;
    const double* dirVal4;
    const double* X4;
    const double* VP4;
    const double* fst4;       // nodal surface-tension force (facets.cu) or null
    double* Aval;
    double* b;
    double* dinv;
    int nNodes, ecap, nbcap, CH, dirWords;
    double rho, mu, dt, body[3];
};

__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }

// |v_cur| per node: the sqrt(nodeU) of PSPG.inl:248-253, computed once per node instead of once per element-node
__global__ void k_vnorm(const double* __restrict__ V4, double* __restrict__ VP4, int nNodes, int dim) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nNodes) return;
    double s = 0;
    for (int d = 0; d < dim; ++d) {
        const double v = V4[(size_t)n * 4 + d];
        s += v * v;
    }
    VP4[(size_t)n * 4 + 3] = sqrt(s);
}

// rowDir bit s of node i: neighbour slot s is a bound node whose tag carries a velocity BC (PSPG.inl:206-208)
__global__ void k_row_dir(int nNodes, const int* __restrict__ nbrPtr, const int* __restrict__ nbr,
                          const uint8_t* __restrict__ flags, const uint8_t* __restrict__ dirMask, int W,
                          unsigned* __restrict__ rowDir, const int* __restrict__ n2ePtr, const int* __restrict__ diagSlot,
                          int4* __restrict__ nodeHdr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nNodes) return;
    const int b0 = nbrPtr[i], nb = nbrPtr[i + 1] - b0;
    unsigned word0 = 0;
    for (int w = 0; w < W; ++w) {
        unsigned bits = 0;
        for (int s = w * 32; s < min(nb, w * 32 + 32); ++s) {
            const int nd = nbr[b0 + s];
            if (dirMask[nd] && (flags[nd] & PFEM_NODE_BOUND)) bits |= 1u << (s & 31);
        }
        rowDir[(size_t)i * W + w] = bits;
        if (w == 0) word0 = bits;
    }
    // packed per-node header of k_pspg_assemble2 (valences above 255 are routed to the first formulation)
    const int eb = n2ePtr[i], ne = n2ePtr[i + 1] - eb;
    const unsigned pk = (unsigned)min(ne, 255) | ((unsigned)min(nb, 255) << 8) | ((unsigned)(diagSlot[i] & 0xff) << 16) |
                        ((unsigned)flags[i] << 24);
    nodeHdr[i] = make_int4(eb, b0, (int)pk, (int)word0);
}

__device__ __forceinline__ int findByte(unsigned packed, int v) {
    const unsigned eq = __vcmpeq4(packed, (unsigned)v * 0x01010101u);
    return (__ffs(eq) - 1) >> 3;
}

// full (dim+1)x(dim+1) contribution of one element to block (i,j); acc is 4x4 row-major for both dims
template <int DIM>
__device__ __forceinline__ void accumBlock(const ElemS<DIM>& E, const BlockCoef& K, int li, int lj, double (&acc)[16]) {
    double gi[DIM], gj[DIM];
    if constexpr (DIM == 3) {
        const double2 a = ld2(&E.g[li][0]), a2 = ld2(&E.g[li][2]), c = ld2(&E.g[lj][0]), c2 = ld2(&E.g[lj][2]);
        gi[0] = a.x, gi[1] = a.y, gi[2] = a2.x;
        gj[0] = c.x, gj[1] = c.y, gj[2] = c2.x;
    } else {
        const double2 a = ld2(&E.g[li][0]), c = ld2(&E.g[lj][0]);
        gi[0] = a.x, gi[1] = a.y;
        gj[0] = c.x, gj[1] = c.y;
    }
    const double2 vt = ld2(&E.V);
    const double cmass = K.kmass * vt.x, cvisc = K.kvisc * vt.x, cdiv = K.kdiv * vt.x, cpc = K.kpc * vt.y, cL = K.kL * vt.y;
    double dot = 0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) dot += gi[d] * gj[d];
    const double diag = cmass * (li == lj ? 2.0 : 1.0) + cvisc * dot;
#pragma unroll
    for (int a = 0; a < DIM; ++a) {
        // velocity row a :  mu V g[c][i] g[a][j] + delta_ac (rho V phi (1+delta_ij)/dt + mu V g_i.g_j) ; -(V/npe) g[a][i]
        const double t = cvisc * gj[a];
#pragma unroll
        for (int c = 0; c < DIM; ++c) acc[a * 4 + c] += t * gi[c];
        acc[a * 4 + a] += diag;
        acc[a * 4 + DIM] -= cdiv * gi[a];
    }
    // pressure row :  (tau/dt)(V/npe) g[c][i] + (V/npe) g[c][j] ; tau (V/rho) g_i.g_j
#pragma unroll
    for (int c = 0; c < DIM; ++c) acc[DIM * 4 + c] += cpc * gi[c] + cdiv * gj[c];
    acc[DIM * 4 + DIM] += cL * dot;
}

// per-node header and the first-chunk per-lane data; prefetched one node ahead by the persistent warp
struct NodeHdr {
    int eb, ne, nb0, nb, si;
    unsigned fl, dir0;
};
struct LaneData {
    int nbrNode;      // neighbour node of slot `lane` (lane < nb)
    unsigned packed;  // slot bytes of incident element `lane` (lane < ne)
    unsigned mask;    // lanes 0..15: element mask of off-diagonal block `lane` (first 32 incident elements)
};
// per-neighbour nodal record staged once per node in shared memory: (x, y, z, -, u_prev, v_prev, w_prev, |v_cur|)
struct alignas(16) NodeRec {
    double x[4];
    double v[4];
    double pad_[2];  // 80-byte stride (odd multiple of 16 B)
};

template <int DIM, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_pspg_assemble(const AsmArgs a) {
    constexpr int NPE = DIM + 1, BS = DIM + 1;
    constexpr double REF = (DIM == 2) ? 0.5 : 0.16666666666666666666666666666667;
    constexpr double PHI = 1.0 / ((DIM + 1) * (DIM + 2));
    extern __shared__ __align__(16) unsigned char smemRaw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int gw = blockIdx.x * wpb + wib, nw = gridDim.x * wpb;
    const int CH = a.CH;
    unsigned char* wbase = smemRaw + (size_t)wib * ((size_t)a.ecap * sizeof(ElemS<DIM>) + (size_t)a.nbcap * sizeof(NodeRec));
    ElemS<DIM>* es = reinterpret_cast<ElemS<DIM>*>(wbase);
    NodeRec* nrec = reinterpret_cast<NodeRec*>(wbase + (size_t)a.ecap * sizeof(ElemS<DIM>));
    const double invdt = 1.0 / a.dt;
    BlockCoef KC;
    KC.kmass = a.rho * PHI * invdt, KC.kvisc = a.mu, KC.kdiv = 1.0 / NPE, KC.kpc = invdt / NPE, KC.kL = 1.0 / a.rho;
    // lane roles in phase 2: lanes 16..19 own the diagonal block, the other 28 lanes one off-diagonal block each
    const bool diagLane = (lane & 28) == 16;
    const int dq = lane & 3;                       // diagonal lane: handles incident elements k = dq (mod 4)
    const int offIdx = lane < 16 ? lane : lane - 4;  // off-diagonal lane: block index within a round of 28

    auto loadHdr = [&](int i, NodeHdr& h) {
        h.eb = __ldg(a.n2ePtr + i);
        h.ne = __ldg(a.n2ePtr + i + 1) - h.eb;
        h.nb0 = __ldg(a.nbrPtr + i);
        h.nb = __ldg(a.nbrPtr + i + 1) - h.nb0;
        h.si = __ldg(a.diagSlot + i);
        h.fl = __ldg(a.flags + i);
        h.dir0 = __ldg(a.rowDir + (size_t)i * a.dirWords);
    };
    auto loadLane = [&](const NodeHdr& h, LaneData& d) {
        d.nbrNode = 0, d.packed = 0, d.mask = 0;
        if (lane < h.ne) d.packed = __ldg(a.n2eSlots + h.eb + lane);
        if (lane < h.nb) d.nbrNode = __ldg(a.nbr + h.nb0 + lane);
        if (!diagLane && offIdx < h.nb - 1) {
            const int jb = offIdx + (offIdx >= h.si ? 1 : 0);
            d.mask = __ldg(a.blkMask + (size_t)(h.nb0 + jb) * CH);
        }
    };
    auto stageNode = [&](int slot, int node) {
        const double* xp = a.X4 + (size_t)node * 4;
        const double* vq = a.VP4 + (size_t)node * 4;
        const double2 x01 = ld2(xp), x23 = ld2(xp + 2), v01 = ld2(vq), v23 = ld2(vq + 2);
        NodeRec& R = nrec[slot];
        *reinterpret_cast<double2*>(&R.x[0]) = x01;
        *reinterpret_cast<double2*>(&R.x[2]) = x23;
        *reinterpret_cast<double2*>(&R.v[0]) = v01;
        *reinterpret_cast<double2*>(&R.v[2]) = v23;
    };

    // phase-1 body for one (node, incident element) pair; returns the element's contribution to the RHS rows of node i
    auto elementPhase = [&](int k, unsigned packed, int si) {
        const int li = findByte(packed, si);
        // previous velocities: only their element sum and the value at node i are needed
        double sv[DIM], vpi[DIM], usum = 0;
#pragma unroll
        for (int c = 0; c < DIM; ++c) sv[c] = 0.0, vpi[c] = 0.0;
        double px[NPE][DIM];
#pragma unroll
        for (int m = 0; m < NPE; ++m) {
            const NodeRec& R = nrec[(packed >> (8 * m)) & 0xffu];
            const double* vq = R.v;
            const double* xp = R.x;
            const double2 v01 = ld2(vq), v23 = ld2(vq + 2), x01 = ld2(xp);
            sv[0] += v01.x, sv[1] += v01.y;
            vpi[0] = (li == m) ? v01.x : vpi[0];
            vpi[1] = (li == m) ? v01.y : vpi[1];
            px[m][0] = x01.x, px[m][1] = x01.y;
            if constexpr (DIM == 3) {
                sv[2] += v23.x;
                vpi[2] = (li == m) ? v23.x : vpi[2];
                px[m][2] = xp[2];
            }
            usum += v23.y;  // |v_cur| of the node
        }
        // J, detJ, inverse (Element.cpp:15-135)
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
        double g[NPE][DIM];
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
            double s = -inv[0][d];
#pragma unroll
            for (int m = 1; m < DIM; ++m) s -= inv[m][d];
            g[0][d] = s;
#pragma unroll
            for (int m = 0; m < DIM; ++m) g[m + 1][d] = inv[m][d];
        }
        // tau (PSPG.inl:238-259): h = sqrt(ref detJ / pi) also in 3-D (reference hazard 7, reproduced)
        const double V = det * REF;
        const double h2 = REF * det / 3.14159265358979323846;
        const double U = usum / NPE;
        const double t1 = 2.0 * invdt, t3 = 4.0 * a.mu / (h2 * a.rho);
        const double tau = 1.0 / sqrt(t1 * t1 + 4.0 * U * U / h2 + 9.0 * t3 * t3);
        const double cmass = KC.kmass * V, cdiv = KC.kdiv * V, cpc = KC.kpc * (tau * V);
        ElemS<DIM>& E = es[k];
#pragma unroll
        for (int m = 0; m < NPE; ++m) {
            *reinterpret_cast<double2*>(&E.g[m][0]) = make_double2(g[m][0], g[m][1]);
            if constexpr (DIM == 3) *reinterpret_cast<double2*>(&E.g[m][2]) = make_double2(g[m][2], 0.0);
        }
        *reinterpret_cast<double2*>(&E.V) = make_double2(V, tau * V);
        *reinterpret_cast<uint2*>(&E.slots) = make_uint2(packed, (unsigned)li);
        // RHS rows of node i: be = [F + (M/dt) vPrev ; tau H + (tau/dt) C vPrev]   (PSPG.inl:53)
        double gb = 0, gs = 0, bloc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int c = 0; c < DIM; ++c) {
            double gci = g[0][c];
#pragma unroll
            for (int m = 1; m < NPE; ++m) gci = (li == m) ? g[m][c] : gci;
            gb += gci * a.body[c];
            gs += gci * sv[c];
            bloc[c] = a.rho * cdiv * a.body[c] + cmass * (vpi[c] + sv[c]);
        }
        bloc[DIM] = tau * V * gb + cpc * gs;
        *reinterpret_cast<double2*>(&E.be[0]) = make_double2(bloc[0], bloc[1]);
        *reinterpret_cast<double2*>(&E.be[2]) = make_double2(bloc[2], bloc[3]);
    };

    // ---- software pipeline prologue: header, lane data and connectivity of the first node ---------------------------
    int i = gw;
    if (i >= a.nNodes) return;
    NodeHdr H, Hn;
    LaneData L, Ln;
    loadHdr(i, H);
    loadLane(H, L);
    Hn = H, Ln = L;

    for (; i < a.nNodes; i += nw) {
        const int inext = i + nw;
        const bool haveNext = inext < a.nNodes;
        if (haveNext) loadHdr(inext, Hn);  // level-1 loads of the next node fly during phase 1

        const int eb = H.eb, ne = H.ne, nb0 = H.nb0, nb = H.nb, si = H.si;
        const bool isBound = H.fl & PFEM_NODE_BOUND, isFree = H.fl & PFEM_NODE_FREE;
        const bool maskV = isBound || isFree, maskP = isFree;
        bool anyDir = H.dir0 != 0;
        for (int w = 1; w < a.dirWords; ++w) anyDir |= a.rowDir[(size_t)i * a.dirWords + w] != 0;

        // -------------------------------------------------------------- phase 1: element geometry + RHS rows of node i
        // stage the nodal records of all neighbours once (instead of once per incident element)
        if (lane < nb) stageNode(lane, L.nbrNode);
        for (int s2 = lane + 32; s2 < nb; s2 += 32) stageNode(s2, a.nbr[nb0 + s2]);
        __syncwarp();
        if (lane < ne) elementPhase(lane, L.packed, si);
        for (int k = lane + 32; k < ne; k += 32)  // nodes with more than 32 incident elements (rare)
            elementPhase(k, a.n2eSlots[eb + k], si);
        if (haveNext) loadLane(Hn, Ln);  // level-2 loads of the next node fly during phase 2
        __syncwarp();

        // -------------------------------------------------------------- phase 2: one lane = one (dim+1)^2 block
        // lanes 0..15 : off-diagonal blocks, 16 per round, elements of the edge (i,j) in ascending order (blkMask)
        // lanes 16..31: the diagonal block, incident elements dealt round-robin, then a transposing butterfly reduction
        double bsub[BS], beTot = 0.0;
#pragma unroll
        for (int r = 0; r < BS; ++r) bsub[r] = 0.0;
        double* Arow = a.Aval + (size_t)nb0 * BS * BS;

        for (int m0 = 0; m0 == 0 || m0 < nb - 1; m0 += 28) {
            const int m = m0 + offIdx;
            const bool offAct = !diagLane && m < nb - 1;
            const bool dgAct = diagLane && m0 == 0;
            const int jb = offAct ? (m + (m >= si ? 1 : 0)) : si;
            double acc[16], beAcc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int t = 0; t < 16; ++t) acc[t] = 0.0;
            if (offAct || dgAct) {
                for (int ch = 0; ch < CH; ++ch) {
                    unsigned mk;
                    if (offAct)
                        mk = (m0 == 0 && ch == 0) ? L.mask : a.blkMask[(size_t)(nb0 + jb) * CH + ch];
                    else {
                        const int rem = ne - ch * 32;
                        const unsigned valid = rem >= 32 ? 0xffffffffu : (rem > 0 ? (1u << rem) - 1u : 0u);
                        mk = (0x11111111u << dq) & valid;
                    }
                    while (mk) {  // two elements per trip: the second one's shared-memory loads overlap the first one's math
                        const int k0 = ch * 32 + __ffs(mk) - 1;
                        mk &= mk - 1;
                        const bool two = mk != 0;
                        const int k1 = two ? ch * 32 + __ffs(mk) - 1 : k0;
                        mk &= mk - 1;
                        const ElemS<DIM>& E0 = es[k0];
                        const ElemS<DIM>& E1 = es[k1];
                        const int li0 = E0.li, li1 = E1.li;
                        const int lj0 = offAct ? findByte(E0.slots, jb) : li0;
                        const int lj1 = offAct ? findByte(E1.slots, jb) : li1;
                        accumBlock<DIM>(E0, KC, li0, lj0, acc);
                        if (two) accumBlock<DIM>(E1, KC, li1, lj1, acc);
                        if (!offAct) {  // diagonal lanes also sum the element RHS rows of node i
                            const double2 b01 = ld2(&E0.be[0]), b23 = ld2(&E0.be[2]);
                            beAcc[0] += b01.x, beAcc[1] += b01.y, beAcc[2] += b23.x, beAcc[3] += b23.y;
                            if (two) {
                                const double2 c01 = ld2(&E1.be[0]), c23 = ld2(&E1.be[2]);
                                beAcc[0] += c01.x, beAcc[1] += c01.y, beAcc[2] += c23.x, beAcc[3] += c23.y;
                            }
                        }
                    }
                }
            }
            if (m0 == 0) {
                // diagonal block + RHS rows: the 4 diagonal lanes park their partial sums (16 + 4 doubles each) in shared
                // memory (the neighbour-record area is free after phase 1); lanes 0..15 then own one diagonal entry each,
                // lanes 16..19 one RHS row each
                double* scr = reinterpret_cast<double*>(nrec);
                if (diagLane) {
#pragma unroll
                    for (int t = 0; t < 8; ++t)
                        *reinterpret_cast<double2*>(scr + dq * 20 + 2 * t) = make_double2(acc[2 * t], acc[2 * t + 1]);
                    *reinterpret_cast<double2*>(scr + dq * 20 + 16) = make_double2(beAcc[0], beAcc[1]);
                    *reinterpret_cast<double2*>(scr + dq * 20 + 18) = make_double2(beAcc[2], beAcc[3]);
                }
                __syncwarp();
                if (lane < 20) {
                    const double tot = (scr[lane] + scr[20 + lane]) + (scr[40 + lane] + scr[60 + lane]);
                    if (lane < 16) {
                        const int r = lane >> 2, cc = lane & 3;
                        if (r < BS && cc < BS) {
                            const bool rowMasked = (r < DIM) ? maskV : maskP;
                            const bool selfDir = (si < 32) ? ((H.dir0 >> si) & 1u)
                                                           : ((a.rowDir[(size_t)i * a.dirWords + (si >> 5)] >> (si & 31)) & 1u);
                            double val = tot;
                            if (rowMasked) val = (r == cc) ? 1.0 : 0.0;
                            else if (selfDir && cc < DIM && cc != r) {  // node i itself is a Dirichlet node (PSPG.inl:219-228)
                                const double sub = val * a.dirVal4[(size_t)i * 4 + cc];
#pragma unroll
                                for (int rr = 0; rr < BS; ++rr) bsub[rr] += (rr == r) ? sub : 0.0;
                                val = 0.0;
                            }
                            Arow[(size_t)si * BS * BS + r * BS + cc] = val;
                            if (r == cc) a.dinv[(size_t)i * BS + r] = (val != 0.0) ? rsqrt(fabs(val)) : 1.0;  // 1/sqrt|a_ii|
                        }
                    } else
                        beTot = tot;  // lane 16 + r holds the assembled RHS of row r
                }
            }
            if (offAct) {
                // row masks (PSPG.inl:68, 81): masked rows hold no off-diagonal entries
#pragma unroll
                for (int r = 0; r < BS; ++r) {
                    const bool rowMasked = (r < DIM) ? maskV : maskP;
                    if (rowMasked) {
#pragma unroll
                        for (int cc = 0; cc < BS; ++cc) acc[r * 4 + cc] = 0.0;
                    }
                }
                const bool colDir = anyDir && ((jb < 32) ? ((H.dir0 >> jb) & 1u)
                                                         : ((a.rowDir[(size_t)i * a.dirWords + (jb >> 5)] >> (jb & 31)) & 1u));
                if (colDir) {  // Dirichlet column elimination (PSPG.inl:216-228)
                    const double* gd = a.dirVal4 + (size_t)a.nbr[nb0 + jb] * 4;
#pragma unroll
                    for (int cc = 0; cc < DIM; ++cc) {
                        const double gv = gd[cc];
#pragma unroll
                        for (int r = 0; r < BS; ++r) {
                            bsub[r] += acc[r * 4 + cc] * gv;
                            acc[r * 4 + cc] = 0.0;
                        }
                    }
                }
                double* dst = Arow + (size_t)jb * BS * BS;
                if constexpr (BS == 4) {
#pragma unroll
                    for (int t = 0; t < 8; ++t)
                        __stcs(reinterpret_cast<double2*>(dst + 2 * t), make_double2(acc[2 * t], acc[2 * t + 1]));
                } else {
#pragma unroll
                    for (int r = 0; r < BS; ++r)
#pragma unroll
                        for (int cc = 0; cc < BS; ++cc) dst[r * BS + cc] = acc[r * 4 + cc];
                }
            }
        }
        // -------------------------------------------------------------- RHS (PSPG.inl:140-144 then :190-232)
        if (anyDir) {
#pragma unroll
            for (int r = 0; r < BS; ++r)
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) bsub[r] += __shfl_xor_sync(0xffffffffu, bsub[r], o);
        }
        if (lane >= 16 && lane < 16 + BS) {
            const int r = lane - 16;
            double bv = beTot, bs = bsub[0];
#pragma unroll
            for (int c = 1; c < BS; ++c) bs = (c == r) ? bsub[c] : bs;
            if (a.fst4 && r < DIM) bv += a.fst4[(size_t)i * 4 + r];  // facet loop of m_applyBCPSPG (PSPG.inl:155-187)
            bv -= bs;
            if (isFree) {
                if (r == DIM) bv = 0.0;
                else if (!isBound) bv = a.VP4[(size_t)i * 4 + r] + a.dt * a.body[r];
            }
            if (isBound && a.dirMask[i] && r < DIM) bv = a.dirVal4[(size_t)i * 4 + r];
            a.b[(size_t)i * BS + r] = bv;
        }
        __syncwarp();  // all lanes are done with es[] before the next node's phase 1 overwrites it
        H = Hn, L = Ln;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// k_pspg_assemble2 -- second formulation of the same gather (default since round 2; PFEM_ASM_V=1 selects the one above).
// What changed, and why (ncu of the kernel above at C4: 458 M warp instructions, fp64 pipe 29 %, issue-active 48 %; only
// ~18 of 32 lanes own a block in phase 2 and the diagonal block walks all ~23 incident elements):
//   * TWO nodes per warp: half-warp h owns node 2w+h, lane t of the half owns off-diagonal block t -- a tetrahedral mesh
//     row has ~14 off-diagonal blocks, so 28 of 32 lanes work; phase 1 runs over the concatenated incident-element list
//     of both nodes;
//   * the diagonal block is NOT summed over the incident elements: the shape-function gradients of an element sum to
//     zero, so every term of block (i,i) is a linear combination of the same terms summed over the off-diagonal blocks
//     of the row (sum_j raw_ij = 0, the i-only terms appear dim times).  20 row sums are reduced across the half-warp
//     through shared memory instead of 23 more (block, element) pairs per node;
//   * per (block, element) pair the lane accumulates 20 RAW sums from V-prescaled gradients G_m = V grad N_m
//     (sum G_j[a] g_i[c], sum V, sum V g_i, sum tau V g_i, sum G_j, sum tau G_j.g_i: 24 fp64 operations, was 38) and the
//     material constants (mu, rho phi/dt, 1/npe, ...) are applied ONCE per block after the loop;
//   * phase 1 needs one division (1/detJ) and one reciprocal square root (tau); the cofactors give G directly.
// Summation order per block is still ascending element index; the result differs from the first formulation by rounding
// only (parity bar: 1e-12 per block type against the reference's Eigen assembly).
template <int DIM> struct ElemR;
template <> struct alignas(16) ElemR<3> {
    double gi[3];      // grad N of the row node i in this element
    double V[1];       // element size -> gi|V fill 32 B
    double G[4][4];    // (V grad N_m, tau) per local node m: slot 3 holds tau in every row
    double be[4];      // this element's contribution to the RHS rows of node i
    unsigned slots;    // slot byte of each local node in the neighbour list of i
    unsigned pad_[3];  // record stride = odd multiple of 16 B (208): LDS.128 of different records spread over the banks
};
template <> struct alignas(16) ElemR<2> {
    double gi[2];
    double V[2];       // (V, -)
    double G[3][4];    // (V grad N_m [2], -, tau)
    double be[4];
    unsigned slots;
    unsigned pad_[3];  // 176 B
};
static_assert(sizeof(ElemR<3>) == 208 && sizeof(ElemR<2>) == 176, "ElemR layout");

template <int DIM, int THREADS, int MINB, bool DIRECT, bool PREFETCH>
__global__ void __launch_bounds__(THREADS, MINB) k_pspg_assemble2(const AsmArgs a) {
    constexpr int NPE = DIM + 1, BS = DIM + 1;
    constexpr int NACC = DIM * DIM + 3 * DIM + 2;  // raw[DIM][DIM], sV, sGi[DIM], sTGi[DIM], sGj[DIM], sTd
    constexpr int I_SV = DIM * DIM, I_GI = I_SV + 1, I_TGI = I_GI + DIM, I_GJ = I_TGI + DIM, I_TD = I_GJ + DIM;
    constexpr int SCR = NACC + 1;                  // scratch row stride (odd: conflict-free column reads)
    constexpr double REF = (DIM == 2) ? 0.5 : 0.16666666666666666666666666666667;
    constexpr double PHI = 1.0 / ((DIM + 1) * (DIM + 2));
    extern __shared__ __align__(16) unsigned char smemRaw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int h = lane >> 4, t = lane & 15;
    const unsigned hmask = 0xffffu << (h * 16);
    const int gw = blockIdx.x * wpb + wib, nw = gridDim.x * wpb;
    const int CH = a.CH;
    const size_t perWarp = (size_t)2 * a.ecap * sizeof(ElemR<DIM>) + (size_t)2 * a.nbcap * sizeof(NodeRec);
    unsigned char* wbase = smemRaw + (size_t)wib * perWarp;
    ElemR<DIM>* es = reinterpret_cast<ElemR<DIM>*>(wbase);          // [2][ecap]
    NodeRec* nrec = reinterpret_cast<NodeRec*>(wbase + (size_t)2 * a.ecap * sizeof(ElemR<DIM>));  // [2][nbcap]
    double* scr = reinterpret_cast<double*>(wbase);                   // row-sum scratch, aliases es[] after phase 2
    const double invdt = 1.0 / a.dt;
    const double kmass = a.rho * PHI * invdt, kvisc = a.mu, kdiv = 1.0 / NPE, kpc = invdt / NPE, kL = 1.0 / a.rho;
    const int nPairs = (a.nNodes + 1) >> 1;

    struct Hdr {
        int eb, ne, nb0, nb, si;
        unsigned fl, dir0;
    };
    auto loadRaw = [&](int i) { return i < a.nNodes ? __ldg(a.nodeHdr + i) : make_int4(0, 0, 0, 0); };
    auto unpack = [](const int4& q, Hdr& H) {
        const unsigned pk = (unsigned)q.z;
        H.eb = q.x, H.nb0 = q.y, H.ne = pk & 0xffu, H.nb = (pk >> 8) & 0xffu, H.si = (pk >> 16) & 0xffu, H.fl = pk >> 24;
        H.dir0 = (unsigned)q.w;
    };
    // per-lane words of a pair, fetched one pair ahead: the neighbour this lane stages (slot t of its half's node), the slot
    // bytes of its first phase-1 element, the element mask of its off-diagonal block
    struct LaneData {
        int nbrNode;
        unsigned packed, packed2, mask;  // packed2: slot bytes of this lane's second phase-1 element (p = lane + 32)
    };
    auto loadLane = [&](const Hdr& A0, const Hdr& A1, LaneData& L) {
        const Hdr& A = h ? A1 : A0;
        L.nbrNode = (t < A.nb) ? __ldg(a.nbr + A.nb0 + t) : -1;
        L.packed = L.packed2 = 0, L.mask = 0;
        if (lane < A0.ne) L.packed = __ldg(a.n2eSlots + A0.eb + lane);
        else if (lane - A0.ne < A1.ne) L.packed = __ldg(a.n2eSlots + A1.eb + lane - A0.ne);
        if (lane + 32 < A0.ne) L.packed2 = __ldg(a.n2eSlots + A0.eb + lane + 32);
        else if (lane + 32 - A0.ne < A1.ne) L.packed2 = __ldg(a.n2eSlots + A1.eb + lane + 32 - A0.ne);
        if (t < A.nb - 1) L.mask = __ldg(a.blkMask + (size_t)(A.nb0 + t + (t >= A.si ? 1 : 0)) * CH);
    };
    int4 q0 = loadRaw(2 * gw), q1 = loadRaw(2 * gw + 1);
    Hdr H0, H1;
    LaneData L;
    if (gw < nPairs) {
        unpack(q0, H0);
        unpack(q1, H1);
        loadLane(H0, H1, L);
    }

    for (int pr = gw; pr < nPairs; pr += nw) {
        // ---- headers of the next pair fly during phases 0/1 (every lane keeps both nodes' headers) -------------------------
        const bool haveNext = PREFETCH && pr + nw < nPairs;
        if (haveNext) {
            q0 = loadRaw(2 * (pr + nw));
            q1 = loadRaw(2 * (pr + nw) + 1);
        }
        const Hdr& H = h ? H1 : H0;
        const int i = 2 * pr + h;
        const bool live = i < a.nNodes;
        // ---- phase 0: each half stages its node's neighbour records once ----------------------------------------------------
        {
            NodeRec* my = nrec + (size_t)h * a.nbcap;
            for (int s2 = t; s2 < H.nb; s2 += 16) {
                const int node = (s2 == t) ? L.nbrNode : __ldg(a.nbr + H.nb0 + s2);
                const double* xp = a.X4 + (size_t)node * 4;
                const double* vq = a.VP4 + (size_t)node * 4;
                const double2 x01 = ld2(xp), x23 = ld2(xp + 2), v01 = ld2(vq), v23 = ld2(vq + 2);
                NodeRec& R = my[s2];
                *reinterpret_cast<double2*>(&R.x[0]) = x01;
                *reinterpret_cast<double2*>(&R.x[2]) = x23;
                *reinterpret_cast<double2*>(&R.v[0]) = v01;
                *reinterpret_cast<double2*>(&R.v[2]) = v23;
            }
        }
        __syncwarp();
        // ---- phase 1: lanes = (node, incident element) pairs of both nodes ---------------------------------------------------
        const int neTot = H0.ne + H1.ne;
        for (int p = lane; p < neTot; p += 32) {
            const int sel = p >= H0.ne ? 1 : 0;
            const int k = sel ? p - H0.ne : p;
            const Hdr& Hs = sel ? H1 : H0;
            const unsigned packed = (p == lane) ? L.packed : ((p == lane + 32) ? L.packed2 : __ldg(a.n2eSlots + Hs.eb + k));
            const NodeRec* recs = nrec + (size_t)sel * a.nbcap;
            const int li = findByte(packed, Hs.si);
            double sv[DIM], vpi[DIM], usum = 0;
#pragma unroll
            for (int c = 0; c < DIM; ++c) sv[c] = 0.0;
            double px[NPE][DIM];
#pragma unroll
            for (int m = 0; m < NPE; ++m) {
                const NodeRec& R = recs[(packed >> (8 * m)) & 0xffu];
                const double2 v01 = ld2(R.v), v23 = ld2(R.v + 2), x01 = ld2(R.x);
                sv[0] += v01.x, sv[1] += v01.y;
                px[m][0] = x01.x, px[m][1] = x01.y;
                if constexpr (DIM == 3) {
                    sv[2] += v23.x;
                    px[m][2] = R.x[2];
                }
                usum += v23.y;  // |v_cur| of the node
            }
            {  // previous velocity of the row node itself: its record sits at slot si
                const NodeRec& Ri = recs[Hs.si];
                const double2 v01 = ld2(Ri.v);
                vpi[0] = v01.x, vpi[1] = v01.y;
                if constexpr (DIM == 3) vpi[2] = Ri.v[2];
            }
            // J, cofactors, detJ (Element.cpp:15-135): G_m = V grad N_m = REF * cof_m needs no division
            double J[DIM][DIM];
#pragma unroll
            for (int d = 0; d < DIM; ++d)
#pragma unroll
                for (int m = 0; m < DIM; ++m) J[d][m] = px[m + 1][d] - px[0][d];
            double det, cof[DIM][DIM];  // cof[m][d] = detJ * invJ[m][d]
            if constexpr (DIM == 2) {
                cof[0][0] = J[1][1], cof[0][1] = -J[0][1], cof[1][0] = -J[1][0], cof[1][1] = J[0][0];
                det = J[0][0] * J[1][1] - J[1][0] * J[0][1];
            } else {
                cof[0][0] = J[1][1] * J[2][2] - J[1][2] * J[2][1];
                cof[0][1] = J[2][1] * J[0][2] - J[2][2] * J[0][1];
                cof[0][2] = J[0][1] * J[1][2] - J[0][2] * J[1][1];
                cof[1][0] = J[2][0] * J[1][2] - J[1][0] * J[2][2];
                cof[1][1] = J[0][0] * J[2][2] - J[2][0] * J[0][2];
                cof[1][2] = J[1][0] * J[0][2] - J[0][0] * J[1][2];
                cof[2][0] = J[1][0] * J[2][1] - J[2][0] * J[1][1];
                cof[2][1] = J[2][0] * J[0][1] - J[0][0] * J[2][1];
                cof[2][2] = J[0][0] * J[1][1] - J[1][0] * J[0][1];
                det = J[0][0] * J[1][1] * J[2][2] + J[0][1] * J[1][2] * J[2][0] + J[0][2] * J[1][0] * J[2][1] -
                      J[2][0] * J[1][1] * J[0][2] - J[2][1] * J[1][2] * J[0][0] - J[2][2] * J[1][0] * J[0][1];
            }
            const double rd = 1.0 / det;
            const double V = det * REF;
            double G[NPE][DIM];
#pragma unroll
            for (int d = 0; d < DIM; ++d) {
                double s0 = 0;
#pragma unroll
                for (int m = 0; m < DIM; ++m) {
                    G[m + 1][d] = REF * cof[m][d];
                    s0 -= G[m + 1][d];
                }
                G[0][d] = s0;
            }
            // tau (PSPG.inl:238-259): h^2 = ref detJ / pi also in 3-D (reference hazard 7, reproduced)
            const double invh2 = 3.14159265358979323846 * rd * (1.0 / REF);
            const double U = usum / NPE;
            const double t1 = 2.0 * invdt, t3 = 4.0 * a.mu / a.rho * invh2;
            const double tau = rsqrt(t1 * t1 + 4.0 * U * U * invh2 + 9.0 * t3 * t3);
            const double invV = rd * (1.0 / REF);
            ElemR<DIM>& E = es[(size_t)sel * a.ecap + k];
#pragma unroll
            for (int m = 0; m < NPE; ++m) {
                if constexpr (DIM == 3) {
                    *reinterpret_cast<double2*>(&E.G[m][0]) = make_double2(G[m][0], G[m][1]);
                    *reinterpret_cast<double2*>(&E.G[m][2]) = make_double2(G[m][2], tau);
                } else {
                    *reinterpret_cast<double2*>(&E.G[m][0]) = make_double2(G[m][0], G[m][1]);
                    *reinterpret_cast<double2*>(&E.G[m][2]) = make_double2(0.0, tau);
                }
            }
            double gi[DIM];  // grad N_li = G_li / V: read back by its (dynamic) row instead of a chain of selects
            {
                const double2 g01 = ld2(&E.G[li][0]);
                gi[0] = g01.x * invV, gi[1] = g01.y * invV;
                if constexpr (DIM == 3) gi[2] = E.G[li][2] * invV;
            }
            if constexpr (DIM == 3) {
                *reinterpret_cast<double2*>(&E.gi[0]) = make_double2(gi[0], gi[1]);
                *reinterpret_cast<double2*>(&E.gi[2]) = make_double2(gi[2], V);
            } else {
                *reinterpret_cast<double2*>(&E.gi[0]) = make_double2(gi[0], gi[1]);
                *reinterpret_cast<double2*>(&E.V[0]) = make_double2(V, 0.0);
            }
            // RHS rows of node i: be = [F + (M/dt) vPrev ; tau H + (tau/dt) C vPrev]   (PSPG.inl:53)
            const double cmass = kmass * V, cdiv = kdiv * V, tauV = tau * V;
            double gb = 0, gs = 0, bloc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int c = 0; c < DIM; ++c) {
                gb += gi[c] * a.body[c];
                gs += gi[c] * sv[c];
                bloc[c] = a.rho * cdiv * a.body[c] + cmass * (vpi[c] + sv[c]);
            }
            bloc[DIM] = tauV * gb + kpc * tauV * gs;
            *reinterpret_cast<double2*>(&E.be[0]) = make_double2(bloc[0], bloc[1]);
            *reinterpret_cast<double2*>(&E.be[2]) = make_double2(bloc[2], bloc[3]);
            E.slots = packed;
        }
        // per-lane words of the next pair fly during phase 2
        LaneData Ln;
        Ln.nbrNode = -1, Ln.packed = Ln.packed2 = 0, Ln.mask = 0;
        if (haveNext) {
            Hdr N0, N1;  // scoped: only the raw words (q0, q1) stay live across phase 2
            unpack(q0, N0);
            unpack(q1, N1);
            loadLane(N0, N1, Ln);
        }
        __syncwarp();

        // ---- phase 2: lane t of half h = off-diagonal block t of node i ----------------------------------------------------
        const int eb = H.eb, ne = H.ne, nb0 = H.nb0, nb = H.nb, si = H.si;
        const bool isBound = H.fl & PFEM_NODE_BOUND, isFree = H.fl & PFEM_NODE_FREE;
        const bool maskV = isBound || isFree, maskP = isFree;
        bool anyDir = H.dir0 != 0;
        for (int w = 1; w < a.dirWords; ++w) anyDir |= live && a.rowDir[(size_t)i * a.dirWords + w] != 0;
        const ElemR<DIM>* myEs = es + (size_t)h * a.ecap;
        double* Arow = a.Aval + (size_t)nb0 * BS * BS;
        double bsub[BS], dsum[DIRECT ? 1 : NACC];
#pragma unroll
        for (int r = 0; r < BS; ++r) bsub[r] = 0.0;
#pragma unroll
        for (int q = 0; q < (DIRECT ? 1 : NACC); ++q) dsum[q] = 0.0;
        int nOffMax = max(H0.nb, H1.nb) - 1;  // both halves run the same number of rounds (warp-level syncs inside)
        for (int m0 = 0; m0 < nOffMax; m0 += 16) {
            const int m = m0 + t;
            const bool act = m < nb - 1;
            const int jb = act ? (m + (m >= si ? 1 : 0)) : si;
            double acc[NACC];
#pragma unroll
            for (int q = 0; q < NACC; ++q) acc[q] = 0.0;
            if (act) {
                for (int ch = 0; ch < CH; ++ch) {
                    unsigned mk = (m0 == 0 && ch == 0) ? L.mask : __ldg(a.blkMask + (size_t)(nb0 + jb) * CH + ch);
                    while (mk) {
                        const int k = ch * 32 + __ffs(mk) - 1;
                        mk &= mk - 1;
                        const ElemR<DIM>& E = myEs[k];
                        const int lj = findByte(E.slots, jb);
                        double gi[DIM], Gj[DIM], V, tau;
                        if constexpr (DIM == 3) {
                            const double2 p0 = ld2(&E.gi[0]), p1 = ld2(&E.gi[2]), q0 = ld2(&E.G[lj][0]), q1 = ld2(&E.G[lj][2]);
                            gi[0] = p0.x, gi[1] = p0.y, gi[2] = p1.x, V = p1.y;
                            Gj[0] = q0.x, Gj[1] = q0.y, Gj[2] = q1.x, tau = q1.y;
                        } else {
                            const double2 p0 = ld2(&E.gi[0]), q0 = ld2(&E.G[lj][0]);
                            gi[0] = p0.x, gi[1] = p0.y, V = E.V[0];
                            Gj[0] = q0.x, Gj[1] = q0.y, tau = E.G[lj][3];
                        }
                        const double tauV = tau * V;
                        double d = 0;
#pragma unroll
                        for (int aa = 0; aa < DIM; ++aa) {
#pragma unroll
                            for (int cc = 0; cc < DIM; ++cc) acc[aa * DIM + cc] += Gj[aa] * gi[cc];
                            d += Gj[aa] * gi[aa];
                            acc[I_GI + aa] += V * gi[aa];
                            acc[I_TGI + aa] += tauV * gi[aa];
                            acc[I_GJ + aa] += Gj[aa];
                        }
                        acc[I_SV] += V;
                        acc[I_TD] += tau * d;
                    }
                }
                if constexpr (!DIRECT) {
#pragma unroll
                    for (int q = 0; q < NACC; ++q) dsum[q] += acc[q];
                }
                // block values from the raw sums
                double blk[16];
                double tr = 0;
#pragma unroll
                for (int aa = 0; aa < DIM; ++aa) tr += acc[aa * DIM + aa];
                const double dg = kmass * acc[I_SV] + kvisc * tr;
#pragma unroll
                for (int aa = 0; aa < DIM; ++aa) {
#pragma unroll
                    for (int cc = 0; cc < DIM; ++cc) blk[aa * 4 + cc] = kvisc * acc[aa * DIM + cc] + (aa == cc ? dg : 0.0);
                    blk[aa * 4 + DIM] = -kdiv * acc[I_GI + aa];
                    blk[DIM * 4 + aa] = kpc * acc[I_TGI + aa] + kdiv * acc[I_GJ + aa];
                }
                blk[DIM * 4 + DIM] = kL * acc[I_TD];
                // row masks (PSPG.inl:68, 81): masked rows hold no off-diagonal entries
#pragma unroll
                for (int r = 0; r < BS; ++r) {
                    const bool rowMasked = (r < DIM) ? maskV : maskP;
                    if (rowMasked) {
#pragma unroll
                        for (int cc = 0; cc < BS; ++cc) blk[r * 4 + cc] = 0.0;
                    }
                }
                const bool colDir = anyDir && ((jb < 32) ? ((H.dir0 >> jb) & 1u)
                                                         : ((a.rowDir[(size_t)i * a.dirWords + (jb >> 5)] >> (jb & 31)) & 1u));
                if (colDir) {  // Dirichlet column elimination (PSPG.inl:216-228)
                    const double* gd = a.dirVal4 + (size_t)a.nbr[nb0 + jb] * 4;
#pragma unroll
                    for (int cc = 0; cc < DIM; ++cc) {
                        const double gv = gd[cc];
#pragma unroll
                        for (int r = 0; r < BS; ++r) {
                            bsub[r] += blk[r * 4 + cc] * gv;
                            blk[r * 4 + cc] = 0.0;
                        }
                    }
                }
                double* dst = Arow + (size_t)jb * BS * BS;
                if constexpr (BS == 4) {
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        __stcs(reinterpret_cast<double2*>(dst + 2 * q), make_double2(blk[2 * q], blk[2 * q + 1]));
                } else {
#pragma unroll
                    for (int r = 0; r < BS; ++r)
#pragma unroll
                        for (int cc = 0; cc < BS; ++cc) dst[r * BS + cc] = blk[r * 4 + cc];
                }
            }
        }
        // the neighbour records of the next pair: warm L1 now, phase 0 of the next trip then finds them there
        if (Ln.nbrNode >= 0) {
            asm volatile("prefetch.global.L1 [%0];" ::"l"(a.X4 + (size_t)Ln.nbrNode * 4));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(a.VP4 + (size_t)Ln.nbrNode * 4));
        }
        // ---- RHS rows and the diagonal block ------------------------------------------------------------------------------
        double beTot = 0.0;
        if constexpr (DIRECT) {
            // ---- diagonal block + RHS rows, summed directly over the incident elements: lane (r = t&3, part = t>>2) owns row r
            // of the diagonal block and every 4th element.  Uniform code for the velocity and the pressure rows:
            //   x = G_i[r] | tau ,  y = grad N_i | G_i   ->   drow[c] += x y[c]   (mu V g_i[c] g_i[r]  |  tau V g_i[c])
            const int r = t & 3, part = t >> 2;
            const bool vrow = r < DIM;
            double drow[DIM], sG[DIM], sD = 0.0, sV = 0.0;
#pragma unroll
            for (int cc = 0; cc < DIM; ++cc) drow[cc] = 0.0, sG[cc] = 0.0;
            for (int k = part; k < ne; k += 4) {
                const ElemR<DIM>& E = myEs[k];
                const int li = findByte(E.slots, si);
                double gi[DIM], Gi[DIM], V, tau;
                if constexpr (DIM == 3) {
                    const double2 p0 = ld2(&E.gi[0]), p1 = ld2(&E.gi[2]), q0 = ld2(&E.G[li][0]), q1 = ld2(&E.G[li][2]);
                    gi[0] = p0.x, gi[1] = p0.y, gi[2] = p1.x, V = p1.y;
                    Gi[0] = q0.x, Gi[1] = q0.y, Gi[2] = q1.x, tau = q1.y;
                } else {
                    const double2 p0 = ld2(&E.gi[0]), q0 = ld2(&E.G[li][0]);
                    gi[0] = p0.x, gi[1] = p0.y, V = E.V[0];
                    Gi[0] = q0.x, Gi[1] = q0.y, tau = E.G[li][3];
                }
                beTot += E.be[r];
                double x = tau, dot = 0.0;
#pragma unroll
                for (int cc = 0; cc < DIM; ++cc) {
                    x = (r == cc) ? Gi[cc] : x;
                    dot += Gi[cc] * gi[cc];
                    sG[cc] += Gi[cc];
                }
#pragma unroll
                for (int cc = 0; cc < DIM; ++cc) drow[cc] += x * (vrow ? gi[cc] : Gi[cc]);
                sD += (vrow ? 1.0 : tau) * dot;
                sV += V;
            }
            auto red4 = [](double v) {
                v += __shfl_xor_sync(0xffffffffu, v, 4);
                v += __shfl_xor_sync(0xffffffffu, v, 8);
                return v;
            };
            beTot = red4(beTot), sD = red4(sD), sV = red4(sV);
#pragma unroll
            for (int cc = 0; cc < DIM; ++cc) drow[cc] = red4(drow[cc]), sG[cc] = red4(sG[cc]);
            // every lane of the half now holds the sums of its row r: lane (r, part) writes entry (r, cc = part)
            const int cc = part;
            if (live && r < BS && cc < BS) {
                double dsel = drow[0], gsel = sG[0], grow = sG[0];
#pragma unroll
                for (int q = 1; q < DIM; ++q) {
                    dsel = (cc == q) ? drow[q] : dsel;
                    gsel = (cc == q) ? sG[q] : gsel;
                    grow = (r == q) ? sG[q] : grow;
                }
                double val;
                if (vrow && cc < DIM) val = kvisc * dsel + (r == cc ? 2.0 * kmass * sV + kvisc * sD : 0.0);
                else if (vrow) val = -kdiv * grow;
                else if (cc < DIM) val = kpc * dsel + kdiv * gsel;
                else val = kL * sD;
                const bool rowMasked = vrow ? maskV : maskP;
                const bool selfDir = (si < 32) ? ((H.dir0 >> si) & 1u)
                                               : ((a.rowDir[(size_t)i * a.dirWords + (si >> 5)] >> (si & 31)) & 1u);
                if (rowMasked) val = (r == cc) ? 1.0 : 0.0;
                else if (selfDir && cc < DIM && cc != r) {  // node i itself is a Dirichlet node (PSPG.inl:219-228)
                    const double sub = val * a.dirVal4[(size_t)i * 4 + cc];
#pragma unroll
                    for (int rr = 0; rr < BS; ++rr) bsub[rr] += (rr == r) ? sub : 0.0;
                    val = 0.0;
                }
                Arow[(size_t)si * BS * BS + r * BS + cc] = val;
                if (r == cc) a.dinv[(size_t)i * BS + r] = (val != 0.0) ? rsqrt(fabs(val)) : 1.0;  // 1/sqrt|a_ii|
            }
        } else {
        // ---- RHS rows: 16 lanes sum the element contributions (4 interleaved partial sums per row, then 2 shuffles) ---------
        {
            const int r = t & 3, part = t >> 2;
            for (int k = part; k < ne; k += 4) beTot += myEs[k].be[r];
            beTot += __shfl_xor_sync(0xffffffffu, beTot, 4);
            beTot += __shfl_xor_sync(0xffffffffu, beTot, 8);  // lanes t < 4 hold the assembled RHS of row t
        }
        __syncwarp();  // every lane is done reading es[]: its storage becomes the row-sum scratch
        // ---- diagonal block from the row sums --------------------------------------------------------------------------------
        {
            double* mine = scr + (size_t)(h * 16 + t) * SCR;
#pragma unroll
            for (int q = 0; q < NACC; ++q) mine[q] = dsum[q];
        }
        __syncwarp();
        double* red = scr + (size_t)32 * SCR + (size_t)h * NACC;  // reduced sums of this half
        for (int q = t; q < NACC; q += 16) {
            double s0 = 0;
#pragma unroll
            for (int l = 0; l < 16; ++l) s0 += scr[(size_t)(h * 16 + l) * SCR + q];
            red[q] = s0;
        }
        __syncwarp();
        if (live && t < 16) {
            const int r = t >> 2, cc = t & 3;
            if (r < BS && cc < BS) {
                constexpr double INVD = 1.0 / DIM;
                double val;
                if (r < DIM && cc < DIM) {
                    val = -kvisc * red[r * DIM + cc];
                    if (r == cc) {
                        double tr = 0;
#pragma unroll
                        for (int aa = 0; aa < DIM; ++aa) tr += red[aa * DIM + aa];
                        val += kmass * (2.0 * INVD) * red[I_SV] - kvisc * tr;
                    }
                } else if (r < DIM) {
                    val = -kdiv * INVD * red[I_GI + r];
                } else if (cc < DIM) {
                    val = kpc * INVD * red[I_TGI + cc] - kdiv * red[I_GJ + cc];
                } else {
                    val = -kL * red[I_TD];
                }
                const bool rowMasked = (r < DIM) ? maskV : maskP;
                const bool selfDir = (si < 32) ? ((H.dir0 >> si) & 1u)
                                               : ((a.rowDir[(size_t)i * a.dirWords + (si >> 5)] >> (si & 31)) & 1u);
                if (rowMasked) val = (r == cc) ? 1.0 : 0.0;
                else if (selfDir && cc < DIM && cc != r) {  // node i itself is a Dirichlet node (PSPG.inl:219-228)
                    const double sub = val * a.dirVal4[(size_t)i * 4 + cc];
#pragma unroll
                    for (int rr = 0; rr < BS; ++rr) bsub[rr] += (rr == r) ? sub : 0.0;
                    val = 0.0;
                }
                Arow[(size_t)si * BS * BS + r * BS + cc] = val;
                if (r == cc) a.dinv[(size_t)i * BS + r] = (val != 0.0) ? rsqrt(fabs(val)) : 1.0;  // 1/sqrt|a_ii|
            }
        }
        }
        // ---- RHS (PSPG.inl:140-144 then :190-232) ----------------------------------------------------------------------------
        if (__any_sync(0xffffffffu, anyDir)) {
#pragma unroll
            for (int r = 0; r < BS; ++r)
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) bsub[r] += __shfl_xor_sync(0xffffffffu, bsub[r], o);
        }
        if (live && t < BS) {
            const int r = t;
            double bv = beTot, bs = bsub[0];
#pragma unroll
            for (int cix = 1; cix < BS; ++cix) bs = (cix == r) ? bsub[cix] : bs;
            if (a.fst4 && r < DIM) bv += a.fst4[(size_t)i * 4 + r];  // facet loop of m_applyBCPSPG (PSPG.inl:155-187)
            bv -= bs;
            if (isFree) {
                if (r == DIM) bv = 0.0;
                else if (!isBound) bv = a.VP4[(size_t)i * 4 + r] + a.dt * a.body[r];
            }
            if (isBound && a.dirMask[i] && r < DIM) bv = a.dirVal4[(size_t)i * 4 + r];
            a.b[(size_t)i * BS + r] = bv;
        }
        __syncwarp();  // scratch / es[] are free for the next pair
        if constexpr (PREFETCH) {
            unpack(q0, H0);
            unpack(q1, H1);
            L = Ln;
        } else if (pr + nw < nPairs) {
            q0 = loadRaw(2 * (pr + nw));
            q1 = loadRaw(2 * (pr + nw) + 1);
            unpack(q0, H0);
            unpack(q1, H1);
            loadLane(H0, H1, L);
        }
    }
}

// states <- q (setNodesStatesfromQ, PSPG.inl:293) ; x = x_saved + dt*v unless fixed (updateNodesPositionFromSave,
// PSPG.inl:294-295, Mesh.cpp:1246-1256)
__global__ void k_picard_update(const double* __restrict__ q, int nNodes, int dim, double dt,
                                const uint8_t* __restrict__ flags, const double* __restrict__ Xsave4,
                                double* __restrict__ X4, double* __restrict__ V4) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nNodes) return;
    const int BS = dim + 1;
    const bool fixed = flags[n] & PFEM_NODE_FIXED;
    for (int d = 0; d < dim; ++d) {
        const double v = q[(size_t)n * BS + d];
        V4[(size_t)n * 4 + d] = v;
        if (!fixed) X4[(size_t)n * 4 + d] = Xsave4[(size_t)n * 4 + d] + v * dt;
    }
    X4[(size_t)n * 4 + 3] = q[(size_t)n * BS + dim];
}

// ---- export in the reference's format: column-major m_A with the reference's pattern --------------------------------
__device__ __forceinline__ bool rowUnmasked(uint8_t f, int d1, int dim) {
    return d1 < dim ? !(f & (PFEM_NODE_BOUND | PFEM_NODE_FREE)) : !(f & PFEM_NODE_FREE);
}
template <bool FILL>
__global__ void k_export_csc(int nNodes, int dim, const int* __restrict__ nbrPtr, const int* __restrict__ nbr,
                             const int* __restrict__ diagSlot, const uint8_t* __restrict__ flags,
                             const double* __restrict__ Aval, int* __restrict__ colPtr, int* __restrict__ rowIdx,
                             double* __restrict__ val) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int BS = dim + 1;
    if (t >= (int64_t)nNodes * BS) return;
    const int d2 = (int)(t / nNodes), j = (int)(t % nNodes);  // reference column index = j + d2*nNodes = t
    const int b0 = nbrPtr[j], nb = nbrPtr[j + 1] - b0;
    int cnt = 0;
    int o = FILL ? colPtr[t] : 0;
    for (int d1 = 0; d1 < BS; ++d1) {
        for (int s = 0; s < nb; ++s) {
            const int i = nbr[b0 + s];
            const bool un = rowUnmasked(flags[i], d1, dim);
            if (!(un || (i == j && d1 == d2))) continue;
            if (FILL) {
                // locate block (i,j) in row i
                const int ib = nbrPtr[i];
                int lo = 0, hi = nbrPtr[i + 1] - ib - 1;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (nbr[ib + mid] < j) lo = mid + 1;
                    else hi = mid;
                }
                rowIdx[o] = i + d1 * nNodes;
                val[o] = Aval[((size_t)ib + lo) * BS * BS + d1 * BS + d2];
                ++o;
            } else
                ++cnt;
        }
    }
    if (!FILL) colPtr[t] = cnt;
}
__global__ void k_b_out(const double* __restrict__ b, double* __restrict__ dst, int nNodes, int BS) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= (int64_t)nNodes * BS) return;
    const int d = (int)(t / nNodes), n = (int)(t % nNodes);
    dst[t] = b[(size_t)n * BS + d];
}

template <int DIM> size_t asmSmemPerWarp(int ecap, int nbcap) {
    return (size_t)ecap * sizeof(ElemS<DIM>) + (size_t)nbcap * sizeof(NodeRec);
}

}  // namespace

void pspgAssemble(pfem_ctx* c, const pfem_pspg_params& p) {
    PFEM_REQUIRE(c->haveTopology && c->havePositions, PFEM_ERR_STATE, "pspg_assemble: topology/positions missing");
    PFEM_REQUIRE(c->haveQprev, PFEM_ERR_STATE, "pspg_assemble: qPrev missing (pfem_pspg_set_qprev)");
    PFEM_REQUIRE(p.dt > 0 && p.rho > 0, PFEM_ERR_INVALID, "pspg_assemble: dt and rho must be positive");
    const int BS = c->dim + 1;
    c->Aval.reserve((size_t)c->nBlocks * BS * BS);
    c->bvec.reserve((size_t)c->nNodes * BS);
    c->dinv.reserve((size_t)c->nNodes * BS);
    const int dirWords = (std::max(c->maxNb, 1) + 31) / 32;
    {
        PhaseScope ph(c, "Prepare matrix assembly");
        k_vnorm<<<divUp(c->nNodes, 256), 256, 0, c->stream>>>(c->V4.p, c->VP4.p, c->nNodes, c->dim);
        LAUNCH_CHECK(c);
        if (c->rowDirDirty) {
            c->rowDir.reserve((size_t)c->nNodes * dirWords + 4);
            c->nodeHdr.reserve((size_t)c->nNodes * 4 + 8);
            k_row_dir<<<divUp(c->nNodes, 128), 128, 0, c->stream>>>(c->nNodes, c->nbrPtr.p, c->nbr.p, c->flags.p, c->dirMask.p,
                                                                  dirWords, c->rowDir.p, c->n2ePtr.p, c->diagSlot.p,
                                                                  reinterpret_cast<int4*>(c->nodeHdr.p));
            LAUNCH_CHECK(c);
            c->rowDirDirty = false;
        }
    }
    AsmArgs a;
    a.fst4 = facetsForces(c, c->X4.p, false);  // "any node on the free surface" rule of PSPG.inl:164-173
    a.conn = c->conn.p, a.n2ePtr = c->n2ePtr.p, a.n2e = c->n2e.p, a.nbrPtr = c->nbrPtr.p, a.nbr = c->nbr.p;
    a.n2eSlots = c->n2eSlots.p, a.blkMask = c->blkMask.p, a.rowDir = c->rowDir.p;
    a.nodeHdr = reinterpret_cast<const int4*>(c->nodeHdr.p);
    a.diagSlot = c->diagSlot.p, a.flags = c->flags.p, a.dirMask = c->dirMask.p, a.dirVal4 = c->dirVal4.p;
    a.X4 = c->X4.p, a.VP4 = c->VP4.p, a.Aval = c->Aval.p, a.b = c->bvec.p, a.dinv = c->dinv.p;
    a.nNodes = c->nRows;  // rows assembled by this rank (owned nodes of a partitioned mesh)
    a.CH = c->maskWords;
    a.dirWords = dirWords;
    a.ecap = std::max(c->maxE, 1);
    a.nbcap = std::max(c->maxNb, 8);  // >= 640 B per warp: the area doubles as the diagonal-reduction scratch
    a.rho = p.rho, a.mu = p.mu, a.dt = p.dt;
    for (int d = 0; d < 3; ++d) a.body[d] = p.bodyForce[d];
    if (pspgNeedsGeneralPath(c)) {  // per-element factors (Bingham viscosity, Boussinesq buoyancy): general kernels, thermal.cu
        pspgAssembleGeneral(c, p, a.fst4);
        c->haveSystem = true;
        mgInvalidate(c, false);
        c->asmStamp = p.dt;
        c->haveSolution = false;
        return;
    }
    {
        PhaseScope ph(c, "Assemble system");  // = Compute triplets + Push back + Assemble matrix/vector + Apply BC
        static const int cfg = getenv("PFEM_ASM_CFG") ? atoi(getenv("PFEM_ASM_CFG")) : 0;
        static const int ver = getenv("PFEM_ASM_V") ? atoi(getenv("PFEM_ASM_V")) : 5;
        if (ver >= 2 && c->maxE <= 255) {
            // two nodes per warp; warps per block chosen so that two blocks fit in an SM's shared memory when possible
            const size_t recB = c->dim == 2 ? sizeof(ElemR<2>) : sizeof(ElemR<3>);
            const int nacc = c->dim * c->dim + 3 * c->dim + 2;
            size_t per = (size_t)2 * a.ecap * recB + (size_t)2 * a.nbcap * sizeof(NodeRec);
            const size_t scrNeed = ((size_t)32 * (nacc + 1) + 2 * nacc) * sizeof(double);  // row-sum scratch aliases the records
            if ((size_t)2 * a.ecap * recB < scrNeed) {  // tiny valences (2-D, boundary-only meshes): pad the record area
                a.ecap = (int)((scrNeed + 2 * recB - 1) / (2 * recB));
                per = (size_t)2 * a.ecap * recB + (size_t)2 * a.nbcap * sizeof(NodeRec);
            }
            PFEM_REQUIRE((size_t)2 * a.ecap * recB >= scrNeed, PFEM_ERR_INVALID, "pspg_assemble: scratch sizing");
            int wpb = 8;
            while (wpb > 1 && per * wpb > 110 * 1024) wpb >>= 1;
            if (per * wpb > 110 * 1024) {
                wpb = 8;
                while (wpb > 1 && per * wpb > 225 * 1024) wpb >>= 1;
            }
            const size_t smem = per * wpb;
            PFEM_REQUIRE(smem <= 225 * 1024, PFEM_ERR_INVALID, "pspg_assemble: node valence too large for shared memory");
            const int blocksPerSm = smem <= 110 * 1024 ? (cfg == 1 ? 3 : 2) : 1;
            const int nPairs = (c->nRows + 1) / 2;
            const int grid = std::max(1, std::min(divUp(nPairs, wpb), c->smCount * blocksPerSm * (8 / wpb)));
#define PFEM_LAUNCH_ASM2(DIM_, M_, D_, P_)                                                                                 \
    do {                                                                                                                  \
        if (smem > 40 * 1024)                                                                                             \
            CUDA_CHECK(cudaFuncSetAttribute(k_pspg_assemble2<DIM_, 256, M_, D_, P_>,                                      \
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, PFEM_SMEM_OPTIN));                     \
        k_pspg_assemble2<DIM_, 256, M_, D_, P_><<<grid, wpb * 32, smem, c->stream>>>(a);                                  \
    } while (0)
            // ver: 2 = row-sum diagonal | 3 = row-sum + next-pair prefetch | 4 = direct diagonal | 5 = direct + prefetch
            if (c->dim == 2) PFEM_LAUNCH_ASM2(2, 2, true, false);
            else if (ver == 2) PFEM_LAUNCH_ASM2(3, 2, false, false);
            else if (ver == 3) PFEM_LAUNCH_ASM2(3, 2, false, true);
            else if (ver == 5) PFEM_LAUNCH_ASM2(3, 2, true, true);
            else PFEM_LAUNCH_ASM2(3, 2, true, false);
            LAUNCH_CHECK(c);
        } else {
        const int wpb = (cfg >= 2) ? 4 : 8;
        const int blocksPerSm = (cfg == 0) ? 2 : (cfg == 1 ? 3 : (cfg == 2 ? 6 : 8));
        const size_t per = c->dim == 2 ? asmSmemPerWarp<2>(a.ecap, a.nbcap) : asmSmemPerWarp<3>(a.ecap, a.nbcap);
        const size_t smem = per * wpb;
        PFEM_REQUIRE(smem <= 225 * 1024, PFEM_ERR_INVALID, "pspg_assemble: node valence too large for shared memory");
        const int grid = std::max(1, std::min(divUp(c->nRows, wpb), c->smCount * blocksPerSm));
#define PFEM_LAUNCH_ASM(DIM_, T_, M_)                                                                                     \
    do {                                                                                                                  \
        if (smem > 40 * 1024)                                                                                             \
            CUDA_CHECK(cudaFuncSetAttribute(k_pspg_assemble<DIM_, T_, M_>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                            PFEM_SMEM_OPTIN));                                                                  \
        k_pspg_assemble<DIM_, T_, M_><<<grid, T_, smem, c->stream>>>(a);                                                  \
    } while (0)
        if (c->dim == 2) PFEM_LAUNCH_ASM(2, 256, 2);
        else if (cfg == 1) PFEM_LAUNCH_ASM(3, 256, 3);
        else if (cfg == 2) PFEM_LAUNCH_ASM(3, 128, 6);
        else if (cfg == 3) PFEM_LAUNCH_ASM(3, 128, 8);
        else PFEM_LAUNCH_ASM(3, 256, 2);
        LAUNCH_CHECK(c);
        }
    }
    c->haveSystem = true;
    mgInvalidate(c, false);  // the coarse matrices follow A
    c->asmStamp = p.dt;
    c->haveSolution = false;
}

void pspgPicardUpdate(pfem_ctx* c, double dt) {
    PFEM_REQUIRE(c->haveSolution && c->haveSnapshot, PFEM_ERR_STATE, "picard: need a solution and a position snapshot");
    PhaseScope ph(c, "Update solutions");
    if (c->nRanks > 1) commHalo(c, c->kx.p, nullptr, c->dim + 1);  // ghost nodes move with their owners' velocities
    k_picard_update<<<divUp(c->nNodes, 256), 256, 0, c->stream>>>(c->kx.p, c->nNodes, c->dim, dt, c->flags.p, c->Xsave4.p,
                                                                  c->X4.p, c->V4.p);
    LAUNCH_CHECK(c);
}

void pspgExportCsc(pfem_ctx* c, int64_t* nnz, int32_t* colPtr, int32_t* rowIdx, double* val, double* b) {
    PFEM_REQUIRE(c->haveSystem, PFEM_ERR_STATE, "export_csc: no assembled system");
    PFEM_REQUIRE(c->nRows == c->nNodes, PFEM_ERR_STATE, "export_csc: not available on a partitioned mesh");
    PFEM_REQUIRE(nnz, PFEM_ERR_INVALID, "export_csc: nnz is null");
    const int BS = c->dim + 1;
    const int64_t nDof = (int64_t)c->nNodes * BS;
    c->cscPtr.reserve(nDof + 2);
    int* total = c->scratchI.p;  // scratch word
    k_export_csc<false><<<divUp(nDof, 128), 128, 0, c->stream>>>(c->nNodes, c->dim, c->nbrPtr.p, c->nbr.p, c->diagSlot.p,
                                                                 c->flags.p, c->Aval.p, c->cscPtr.p, nullptr, nullptr);
    LAUNCH_CHECK(c);
    CUDA_CHECK(cudaMemsetAsync(c->cscPtr.p + nDof, 0, sizeof(int), c->stream));
    exclusiveScanInt(c, c->cscPtr.p, (int)nDof + 1, total);
    int h = 0;
    CUDA_CHECK(cudaMemcpyAsync(&h, total, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    *nnz = h;
    c->nnzReference = h;
    if (!colPtr) return;
    PFEM_REQUIRE(rowIdx && val && b, PFEM_ERR_INVALID, "export_csc: null output array");
    c->cscRow.reserve(h + 1);
    c->cscVal.reserve(h + 1);
    k_export_csc<true><<<divUp(nDof, 128), 128, 0, c->stream>>>(c->nNodes, c->dim, c->nbrPtr.p, c->nbr.p, c->diagSlot.p,
                                                                c->flags.p, c->Aval.p, c->cscPtr.p, c->cscRow.p, c->cscVal.p);
    LAUNCH_CHECK(c);
    c->stageD.reserve(nDof);
    k_b_out<<<divUp(nDof, 256), 256, 0, c->stream>>>(c->bvec.p, c->stageD.p, c->nNodes, BS);
    LAUNCH_CHECK(c);
    CUDA_CHECK(cudaMemcpyAsync(colPtr, c->cscPtr.p, (nDof + 1) * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(rowIdx, c->cscRow.p, (size_t)h * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(val, c->cscVal.p, (size_t)h * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(b, c->stageD.p, nDof * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
}
