// pspg.cu -- assembly of the monolithic PSPG system  A q = b  of MomContEqIncompNewton<dim>
// (MomContEquationPSPG.inl:7-146 m_buildAbPSPG, :149-235 m_applyBCPSPG, :238-259 m_computeTauPSPG) on the device.
//
// Design (B200: HBM-bound, no tensor cores -- 4x4 blocks are far too small):
//   * node-row GATHER instead of element scatter: one warp owns one node i == one (dim+1)-row block row of A.
//     Phase 1: lanes = incident elements of i; each lane recomputes that element's geometry from coordinates
//              (Element.cpp:15-135, nothing stored per element), tau, and the five scalar coefficients of the closed
//              form and stages them in shared memory.
//     Phase 2: lanes = (neighbour block, row) pairs; each lane sums, in ASCENDING ELEMENT ORDER (the order in which
//              setFromTriplets sums duplicates), the contributions of the elements that contain the edge (i,j).
//   * A is written exactly once, fully coalesced (1 KB per warp store), no atomics, bit-reproducible run to run.
//   * m_applyBCPSPG is fused into the epilogue: row masks / identity rows (PSPG.inl:68,81,108-128), Dirichlet column
//     elimination (PSPG.inl:216-228) as a per-row operation, free-node RHS (PSPG.inl:193-204), and 1/diag for the
//     Jacobi preconditioner.
// Storage: node-block CSR ("BSR") with BS=(dim+1): Aval[(nbrPtr[i]+slot)*BS*BS + r*BS + c] = A(i + r*N, nbr[slot] + c*N).
// Masked rows are held as explicit zero rows + unit diagonal internally; pfem_pspg_export_csc emits the reference's
// exact CSC pattern.
#include "common.cuh"

namespace {

template <int DIM> struct ElemS {
    double g[DIM * (DIM + 1)];  // grad N, g[d*NPE + k]   (MatricesBuilder.inl:93-127)
    double cmass;               // rho V phi / dt          M/dt   (phi = 1/((dim+1)(dim+2)))
    double cvisc;               // mu V                    K
    double cdiv;                // V / npe                 D
    double cpc;                 // (tau/dt) V / npe        (tau/dt) C
    double cL;                  // tau V / rho             tau L
};

struct AsmArgs {
    const int* conn;
    const int* n2ePtr;
    const int* n2e;
    const int* nbrPtr;
    const int* nbr;
    const int* diagSlot;
    const uint8_t* flags;
    const uint8_t* dirMask;
    const double* dirVal4;
    const double* X4;
    const double* VP4;
    double* Aval;
    double* b;
    double* dinv;
    int nNodes, ecap, nbcap;
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

template <int DIM> __device__ __forceinline__ int findByte(unsigned packed, int v) {
    const unsigned eq = __vcmpeq4(packed, (unsigned)v * 0x01010101u);
    return (__ffs(eq) - 1) >> 3;
}

// contribution of one element to row r of block (i,j)
template <int DIM>
__device__ __forceinline__ void accumRow(const ElemS<DIM>& E, int li, int lj, int r, double (&out)[DIM + 1]) {
    constexpr int NPE = DIM + 1;
    double gi[DIM], gj[DIM], dot = 0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
        gi[d] = E.g[d * NPE + li];
        gj[d] = E.g[d * NPE + lj];
        dot += gi[d] * gj[d];
    }
    const bool vrow = r < DIM;
    const int rr = vrow ? r : 0;
    const double gjr = E.g[rr * NPE + lj], gir = E.g[rr * NPE + li];
    // velocity row a=r :  mu V g[c][i] g[a][j] + delta_ac (rho V phi (1+delta_ij)/dt + mu V g_i.g_j) ; -(V/npe) g[a][i]
    // pressure row     :  (tau/dt)(V/npe) g[c][i] + (V/npe) g[c][j]                                 ; tau (V/rho) g_i.g_j
    const double ca = vrow ? E.cvisc * gjr : E.cpc;
    const double cb = vrow ? 0.0 : E.cdiv;
    const double diag = E.cmass * (li == lj ? 2.0 : 1.0) + E.cvisc * dot;
#pragma unroll
    for (int c = 0; c < DIM; ++c) {
        double v = ca * gi[c] + cb * gj[c];
        if (vrow && c == r) v += diag;
        out[c] += v;
    }
    out[DIM] += vrow ? -E.cdiv * gir : E.cL * dot;
}

template <int DIM>
__global__ void __launch_bounds__(256) k_pspg_assemble(const AsmArgs a) {
    constexpr int NPE = DIM + 1, BS = DIM + 1;
    constexpr double REF = (DIM == 2) ? 0.5 : 0.16666666666666666666666666666667;
    constexpr double PHI = 1.0 / ((DIM + 1) * (DIM + 2));
    extern __shared__ __align__(16) unsigned char smemRaw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int i = blockIdx.x * wpb + wib;
    const int CH = a.ecap >> 5;
    // per-warp shared memory carve-up
    const size_t perWarp = (size_t)a.ecap * (sizeof(ElemS<DIM>) + 4) + (size_t)a.nbcap * 4 * (1 + CH) + (size_t)a.nbcap;
    const size_t perWarpAl = (perWarp + 15) & ~(size_t)15;
    unsigned char* base = smemRaw + perWarpAl * wib;
    ElemS<DIM>* es = reinterpret_cast<ElemS<DIM>*>(base);
    unsigned* eslots = reinterpret_cast<unsigned*>(base + (size_t)a.ecap * sizeof(ElemS<DIM>));
    int* nbrS = reinterpret_cast<int*>(eslots + a.ecap);
    unsigned* emask = reinterpret_cast<unsigned*>(nbrS + a.nbcap);
    unsigned char* dirS = reinterpret_cast<unsigned char*>(emask + (size_t)a.nbcap * CH);
    if (i >= a.nNodes) return;

    const int eb = a.n2ePtr[i], ne = a.n2ePtr[i + 1] - eb;
    const int nb0 = a.nbrPtr[i], nb = a.nbrPtr[i + 1] - nb0;
    const int si = a.diagSlot[i];
    const uint8_t fl = a.flags[i];
    const bool isBound = fl & PFEM_NODE_BOUND, isFree = fl & PFEM_NODE_FREE;
    const bool maskV = isBound || isFree, maskP = isFree;
    const double invdt = 1.0 / a.dt;

    for (int s = lane; s < nb; s += 32) {
        const int nd = a.nbr[nb0 + s];
        nbrS[s] = nd;
        dirS[s] = a.dirMask[nd];
    }
    for (int s = lane; s < nb * CH; s += 32) emask[s] = 0u;
    __syncwarp();

    // ------------------------------------------------------------------ phase 1: element geometry + RHS rows of node i
    double be[BS];
#pragma unroll
    for (int r = 0; r < BS; ++r) be[r] = 0.0;
    for (int ch = 0; ch < CH; ++ch) {
        const int k = ch * 32 + lane;
        if (k < ne) {
            const int e = a.n2e[eb + k];
            int nd[NPE];
            if constexpr (DIM == 3) {
                const int4 q = *reinterpret_cast<const int4*>(a.conn + (size_t)e * 4);
                nd[0] = q.x, nd[1] = q.y, nd[2] = q.z, nd[3] = q.w;
            } else {
#pragma unroll
                for (int m = 0; m < NPE; ++m) nd[m] = a.conn[(size_t)e * NPE + m];
            }
            double px[NPE][DIM], vp[NPE][DIM], usum = 0;
#pragma unroll
            for (int m = 0; m < NPE; ++m) {
                const double* xp = a.X4 + (size_t)nd[m] * 4;
                const double* vq = a.VP4 + (size_t)nd[m] * 4;
                const double2 x01 = ld2(xp), v01 = ld2(vq), v23 = ld2(vq + 2);
                px[m][0] = x01.x, px[m][1] = x01.y;
                vp[m][0] = v01.x, vp[m][1] = v01.y;
                if constexpr (DIM == 3) {
                    px[m][2] = xp[2];
                    vp[m][2] = v23.x;
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
            ElemS<DIM> E;
#pragma unroll
            for (int d = 0; d < DIM; ++d) {
                double s = -inv[0][d];
#pragma unroll
                for (int m = 1; m < DIM; ++m) s -= inv[m][d];
                E.g[d * NPE] = s;
#pragma unroll
                for (int m = 0; m < DIM; ++m) E.g[d * NPE + m + 1] = inv[m][d];
            }
            // tau (PSPG.inl:238-259): h = sqrt(ref detJ / pi) also in 3-D (reference hazard 7, reproduced)
            const double V = det * REF;
            const double h2 = REF * det / 3.14159265358979323846;
            const double U = usum / NPE;
            const double t1 = 2.0 * invdt, t3 = 4.0 * a.mu / (h2 * a.rho);
            const double tau = 1.0 / sqrt(t1 * t1 + 4.0 * U * U / h2 + 9.0 * t3 * t3);
            E.cmass = a.rho * V * PHI * invdt;
            E.cvisc = a.mu * V;
            E.cdiv = V / NPE;
            E.cpc = tau * invdt * E.cdiv;
            E.cL = tau * V / a.rho;
            // slots of the element's nodes in the neighbour list of i (sorted -> binary search)
            unsigned packed = 0;
            int li = 0;
#pragma unroll
            for (int m = 0; m < NPE; ++m) {
                int lo = 0, hi = nb - 1;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (nbrS[mid] < nd[m]) lo = mid + 1;
                    else hi = mid;
                }
                packed |= (unsigned)lo << (8 * m);
                if (nd[m] == i) li = m;
                atomicOr(&emask[lo * CH + ch], 1u << lane);
            }
            if constexpr (DIM == 2) packed |= 0xff000000u;  // unused byte never matches a slot (< 255)
            es[k] = E;
            eslots[k] = packed;
            // RHS rows of node i: be = [F + (M/dt) vPrev ; tau H + (tau/dt) C vPrev]   (PSPG.inl:53)
            double sumvp[DIM], gb = 0, gs = 0;
#pragma unroll
            for (int c = 0; c < DIM; ++c) {
                double s = 0;
#pragma unroll
                for (int m = 0; m < NPE; ++m) s += vp[m][c];
                sumvp[c] = s;
                const double gci = E.g[c * NPE + li];
                gb += gci * a.body[c];
                gs += gci * s;
            }
            double vpi[DIM];
#pragma unroll
            for (int c = 0; c < DIM; ++c) {
                double t = vp[0][c];
#pragma unroll
                for (int m = 1; m < NPE; ++m) t = (li == m) ? vp[m][c] : t;
                vpi[c] = t;
            }
#pragma unroll
            for (int c = 0; c < DIM; ++c) be[c] += a.rho * E.cdiv * a.body[c] + E.cmass * (vpi[c] + sumvp[c]);
            be[DIM] += tau * V * gb + E.cpc * gs;
        }
    }
#pragma unroll
    for (int r = 0; r < BS; ++r)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) be[r] += __shfl_xor_sync(0xffffffffu, be[r], o);
    __syncwarp();

    // ------------------------------------------------------------------ phase 2: block rows
    const int grp = lane >> 2, r = lane & 3;
    const bool rowActive = r < BS;
    const bool rowMasked = (r < DIM) ? maskV : maskP;
    double bsub = 0.0;
    double* Arow = a.Aval + (size_t)nb0 * BS * BS;

    auto finishBlock = [&](int jb, double (&out)[BS], bool writer) {
        // row masks + identity (PSPG.inl:68, 81, 108-128), then Dirichlet column elimination (PSPG.inl:216-228)
        if (rowMasked) {
#pragma unroll
            for (int c = 0; c < BS; ++c) out[c] = (jb == si && c == r) ? 1.0 : 0.0;
        } else if (dirS[jb]) {
            const double* gd = a.dirVal4 + (size_t)nbrS[jb] * 4;
#pragma unroll
            for (int c = 0; c < DIM; ++c) {
                if (!(jb == si && c == r)) {
                    bsub += out[c] * gd[c];
                    out[c] = 0.0;
                }
            }
        }
        if (writer && rowActive) {
            double* dst = Arow + (size_t)jb * BS * BS + r * BS;
            if constexpr (BS == 4) {
                *reinterpret_cast<double2*>(dst) = make_double2(out[0], out[1]);
                *reinterpret_cast<double2*>(dst + 2) = make_double2(out[2], out[3]);
            } else {
#pragma unroll
                for (int c = 0; c < BS; ++c) dst[c] = out[c];
            }
        }
    };

    // off-diagonal blocks: 8 blocks per round, ascending element order inside each
    for (int m0 = 0; m0 < nb - 1; m0 += 8) {
        const int m = m0 + grp;
        const bool act = m < nb - 1;
        const int jb = act ? (m + (m >= si ? 1 : 0)) : 0;
        double out[BS];
#pragma unroll
        for (int c = 0; c < BS; ++c) out[c] = 0.0;
        if (act && rowActive && !rowMasked) {
            for (int ch = 0; ch < CH; ++ch) {
                unsigned mk = emask[jb * CH + ch];
                while (mk) {
                    const int k = ch * 32 + __ffs(mk) - 1;
                    mk &= mk - 1;
                    const unsigned sl = eslots[k];
                    accumRow<DIM>(es[k], findByte<DIM>(sl, si), findByte<DIM>(sl, jb), r, out);
                }
            }
        }
        if (act) finishBlock(jb, out, true);
    }
    // diagonal block: every incident element contributes; split them over the 8 lane groups, then reduce
    {
        double out[BS];
#pragma unroll
        for (int c = 0; c < BS; ++c) out[c] = 0.0;
        if (rowActive && !rowMasked) {
            for (int k = grp; k < ne; k += 8) {
                const int li = findByte<DIM>(eslots[k], si);
                accumRow<DIM>(es[k], li, li, r, out);
            }
        }
#pragma unroll
        for (int c = 0; c < BS; ++c) {
            out[c] += __shfl_xor_sync(0xffffffffu, out[c], 4);
            out[c] += __shfl_xor_sync(0xffffffffu, out[c], 8);
            out[c] += __shfl_xor_sync(0xffffffffu, out[c], 16);
        }
        if (grp == 0) {
            finishBlock(si, out, true);
            if (rowActive) {
                double d = out[0];
#pragma unroll
                for (int c = 1; c < BS; ++c) d = (c == r) ? out[c] : d;
                a.dinv[(size_t)i * BS + r] = (d != 0.0) ? 1.0 / d : 1.0;
            }
        }
    }
    bsub += __shfl_xor_sync(0xffffffffu, bsub, 4);
    bsub += __shfl_xor_sync(0xffffffffu, bsub, 8);
    bsub += __shfl_xor_sync(0xffffffffu, bsub, 16);

    // ------------------------------------------------------------------ RHS (PSPG.inl:140-144 then :190-232)
    if (grp == 0 && rowActive) {
        double bv = be[0];
#pragma unroll
        for (int c = 1; c < BS; ++c) bv = (c == r) ? be[c] : bv;
        bv -= bsub;
        if (isFree) {
            if (r == DIM) bv = 0.0;
            else if (!isBound) bv = a.VP4[(size_t)i * 4 + r] + a.dt * a.body[r];
        }
        if (isBound && a.dirMask[i] && r < DIM) bv = a.dirVal4[(size_t)i * 4 + r];
        a.b[(size_t)i * BS + r] = bv;
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
    const int CH = ecap >> 5;
    const size_t perWarp = (size_t)ecap * (sizeof(ElemS<DIM>) + 4) + (size_t)nbcap * 4 * (1 + CH) + (size_t)nbcap;
    return (perWarp + 15) & ~(size_t)15;
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
    {
        PhaseScope ph(c, "Prepare matrix assembly");
        k_vnorm<<<divUp(c->nNodes, 256), 256, 0, c->stream>>>(c->V4.p, c->VP4.p, c->nNodes, c->dim);
        LAUNCH_CHECK(c);
    }
    AsmArgs a;
    a.conn = c->conn.p, a.n2ePtr = c->n2ePtr.p, a.n2e = c->n2e.p, a.nbrPtr = c->nbrPtr.p, a.nbr = c->nbr.p;
    a.diagSlot = c->diagSlot.p, a.flags = c->flags.p, a.dirMask = c->dirMask.p, a.dirVal4 = c->dirVal4.p;
    a.X4 = c->X4.p, a.VP4 = c->VP4.p, a.Aval = c->Aval.p, a.b = c->bvec.p, a.dinv = c->dinv.p;
    a.nNodes = c->nNodes;
    a.ecap = ((std::max(c->maxE, 1) + 31) / 32) * 32;
    a.nbcap = ((std::max(c->maxNb, 1) + 7) / 8) * 8;
    a.rho = p.rho, a.mu = p.mu, a.dt = p.dt;
    for (int d = 0; d < 3; ++d) a.body[d] = p.bodyForce[d];
    {
        PhaseScope ph(c, "Assemble system");  // = Compute triplets + Push back + Assemble matrix/vector + Apply BC
        int wpb = 8;
        size_t per = c->dim == 2 ? asmSmemPerWarp<2>(a.ecap, a.nbcap) : asmSmemPerWarp<3>(a.ecap, a.nbcap);
        while (wpb > 1 && per * wpb > 200 * 1024) wpb >>= 1;
        const size_t smem = per * wpb;
        PFEM_REQUIRE(smem <= 227 * 1024, PFEM_ERR_INVALID, "pspg_assemble: node valence too large for shared memory");
        if (c->dim == 2) {
            if (smem > 48 * 1024)
                CUDA_CHECK(cudaFuncSetAttribute(k_pspg_assemble<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_pspg_assemble<2><<<divUp(c->nNodes, wpb), wpb * 32, smem, c->stream>>>(a);
        } else {
            if (smem > 48 * 1024)
                CUDA_CHECK(cudaFuncSetAttribute(k_pspg_assemble<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_pspg_assemble<3><<<divUp(c->nNodes, wpb), wpb * 32, smem, c->stream>>>(a);
        }
        LAUNCH_CHECK(c);
    }
    c->haveSystem = true;
    c->haveSolution = false;
}

void pspgPicardUpdate(pfem_ctx* c, double dt) {
    PFEM_REQUIRE(c->haveSolution && c->haveSnapshot, PFEM_ERR_STATE, "picard: need a solution and a position snapshot");
    PhaseScope ph(c, "Update solutions");
    k_picard_update<<<divUp(c->nNodes, 256), 256, 0, c->stream>>>(c->kx.p, c->nNodes, c->dim, dt, c->flags.p, c->Xsave4.p,
                                                                  c->X4.p, c->V4.p);
    LAUNCH_CHECK(c);
}

void pspgExportCsc(pfem_ctx* c, int64_t* nnz, int32_t* colPtr, int32_t* rowIdx, double* val, double* b) {
    PFEM_REQUIRE(c->haveSystem, PFEM_ERR_STATE, "export_csc: no assembled system");
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
