// wc_tile.cuh -- fused ("tile") formulation of the explicit weakly-compressible passes; included by wc.cu inside its
// anonymous namespace (it reuses loadElem / elemHe / WcArgs and the nodal epilogues' arithmetic).
//
// Why: the two-pass kernels compute every element once but hand the results to the node pass through HBM -- 128 B of
// momentum records per element written and read back (5.2 GB of the 10.9 GB a C5 step moves, profiles/r1_ncu_wc_twopass.md);
// the gather kernels move nothing extra but recompute every element once per node (4x in 3-D).  Here a CTA owns a TILE of
// ~64 spatially close nodes: it computes every element incident to the tile ONCE, keeps the per-(element, local node)
// records in SHARED memory, and the tile's nodes then sum their records in ascending element index (the reference's serial
// scatter order, ContEquation.inl:398-408, MomEquation.inl:279-298; same 4-lane order as the two-pass kernels, so the
// results are bit-identical to them) and apply the nodal epilogue.  Elements that touch two tiles are computed twice (about
// 1.5x on a spatially sorted tile, against 4x for the gather kernels); nothing per element ever reaches HBM.
// Per remesh (buildTiles): nodes are sorted by (interface node first, cell of a uniform grid over their coordinates) with a
// counting sort -- no renumbering of anything the ABI sees -- and cut into tiles of T consecutive nodes; one CTA per tile
// sorts and uniquifies the incident element ids of its nodes in shared memory (bitonic sort) and stores, per (node, incident
// element), the element's index in the tile list and the node's local index in it (16 bits).
// On a partitioned mesh the tiles that contain interface nodes come first: they are computed by a first launch, after
// which the halo exchange of their results overlaps the launch over the interior tiles (wc.cu: launchStep).
#pragma once

// staged node record: (x, y, z, p | u, v, w, rho) padded to 80 B -- an odd multiple of 16 B, so that LDS.128 of different
// records spread over all banks (a 64-byte stride would put every record's first quarter in 2 of the 8 bank groups)
constexpr int TILE_NSTRIDE = 10;

struct TileArgs {
    const int* perm;        // owned nodes in tile order
    const int* nePrefix;    // prefix sum of the valences in tile order: tile t's element list starts at nePrefix[t*T]
    const int* tileCnt;     // per tile: number of elements
    const int* tileElems;
    const unsigned short* dst16;  // per tile element (same indexing as tileElems) and local node: the SLOT of that (node, element)
                                  // incidence in the tile's record array -- the tile nodes' incidence lists laid end to end --
                                  // or 0xffff when the node belongs to another tile
    const int* nodeStart;   // per tile: offset of its node list (tile nodes + the other nodes of its elements) in tileNodes
    const int* nodeCnt;
    const int* tileNodes;
    const unsigned short* lconn;  // per tile element (same indexing as tileElems): 4 indices into the tile's node list
    int T, nRows, tile0, nodeCap, slotCap;
};

// ---- build ----------------------------------------------------------------------------------------------------------------
__global__ void k_tile_bbox(const double* __restrict__ X, int n, int dim, double* __restrict__ out) {
    __shared__ double slo[3][32], shi[3][32];
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        for (int d = 0; d < dim; ++d) {
            const double v = X[(size_t)i * 4 + d];
            lo[d] = fmin(lo[d], v);
            hi[d] = fmax(hi[d], v);
        }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int d = 0; d < 3; ++d) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[d] = fmin(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
            hi[d] = fmax(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
        }
        if (lane == 0) slo[d][w] = lo[d], shi[d][w] = hi[d];
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        const int d = threadIdx.x;
        double a = 1e300, b = -1e300;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) a = fmin(a, slo[d][k]), b = fmax(b, shi[d][k]);
        if (d >= dim) a = b = 0.0;
        out[d] = a;
        out[3 + d] = b;
    }
}
__global__ void k_tile_mark(const int* __restrict__ sendIdx, int nSend, unsigned char* __restrict__ iface) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nSend) iface[sendIdx[t]] = 1;
}
// key = cell of the node (interface nodes in the first nCells bins, the others in the next nCells); bin counts
__global__ void k_tile_key(const double* __restrict__ X, int n, int dim, double lox, double loy, double loz, double invH, int nx, int ny,
                           int nz, const unsigned char* __restrict__ iface, int* __restrict__ key, int* __restrict__ count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* x = X + (size_t)i * 4;
    const int ix = min(nx - 1, max(0, (int)floor((x[0] - lox) * invH)));
    const int iy = min(ny - 1, max(0, (int)floor((x[1] - loy) * invH)));
    const int iz = dim == 3 ? min(nz - 1, max(0, (int)floor((x[2] - loz) * invH))) : 0;
    const int k = (iz * ny + iy) * nx + ix + (iface[i] ? 0 : nx * ny * nz);
    key[i] = k;
    atomicAdd(&count[k], 1);
}
__global__ void k_tile_fill(int n, const int* __restrict__ key, const int* __restrict__ binPtr, int* __restrict__ cursor,
                            int* __restrict__ perm) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int k = key[i];
    perm[binPtr[k] + atomicAdd(&cursor[k], 1)] = i;
}
// members of a bin in ascending node id: the tile composition (hence the shared-memory layout) is the same run to run
__global__ void k_tile_sort_bins(int nBins, const int* __restrict__ binPtr, int* __restrict__ perm) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nBins) return;
    const int s = binPtr[b], len = binPtr[b + 1] - s;
    int* v = perm + s;
    for (int i = 1; i < len; ++i) {
        const int k = v[i];
        int j = i - 1;
        while (j >= 0 && v[j] > k) {
            v[j + 1] = v[j];
            --j;
        }
        v[j + 1] = k;
    }
}
__global__ void k_tile_valence(int n, const int* __restrict__ perm, const int* __restrict__ n2ePtr, int* __restrict__ ne) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int i = perm[k];
    ne[k] = n2ePtr[i + 1] - n2ePtr[i];
}
// One CTA per tile: the sorted set of elements incident to its nodes, and per (node, incident element) the index in it.
__global__ void __launch_bounds__(256) k_tile_build(int T, int nRows, int npe, int CAP, const int* __restrict__ perm,
                                                    const int* __restrict__ nePrefix, const int* __restrict__ n2ePtr,
                                                    const int* __restrict__ n2e, const int* __restrict__ conn, int* __restrict__ tileElems,
                                                    int* __restrict__ tileCnt, unsigned short* __restrict__ dst16, int* __restrict__ maxCnt,
                                                    int* __restrict__ nodeStart, int* __restrict__ nodeCnt, int* __restrict__ tileNodes,
                                                    int nodeListCap, unsigned short* __restrict__ lconn) {
    // maxCnt: [0] max elements per tile, [1] max nodes per tile list, [2] bump cursor of tileNodes, [3] overflow flag,
    //         [4] max incidences (record slots) per tile
    extern __shared__ int sm[];
    int* keys = sm;          // CAP
    int* uniq = sm + CAP;    // CAP
    int* cand = sm + 2 * CAP;  // npe*CAP (<= 4 CAP): nodes of the tile's elements
    __shared__ int warpTot[8], total, nodeOff;
    const int t = blockIdx.x, tid = threadIdx.x;
    const int k0 = t * T, k1 = min(nRows, k0 + T);
    const int base = nePrefix[k0], nInc = nePrefix[k1] - base;
    for (int j = tid; j < CAP; j += 256) keys[j] = 0x7fffffff;
    __syncthreads();
    for (int s = tid >> 2; s < k1 - k0; s += 64) {  // 4 lanes per node
        const int i = perm[k0 + s];
        const int eb = n2ePtr[i], ne = n2ePtr[i + 1] - eb, off = nePrefix[k0 + s] - base;
        for (int k = tid & 3; k < ne; k += 4) keys[off + k] = n2e[eb + k];
    }
    __syncthreads();
    // bitonic sort of CAP keys
    for (int size = 2; size <= CAP; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int j = tid; j < CAP / 2; j += 256) {
                const int lo = 2 * j - (j & (stride - 1));
                const int hi = lo + stride;
                const bool up = (lo & size) == 0;
                const int a = keys[lo], b = keys[hi];
                if ((a > b) == up) {
                    keys[lo] = b;
                    keys[hi] = a;
                }
            }
            __syncthreads();
        }
    // unique: every thread owns CAP/256 consecutive entries
    const int per = CAP / 256;
    int cnt = 0;
    for (int q = 0; q < per; ++q) {
        const int j = tid * per + q;
        if (j < nInc && (j == 0 || keys[j] != keys[j - 1])) ++cnt;
    }
    int inc = cnt;
    const int lane = tid & 31, w = tid >> 5;
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    if (lane == 31) warpTot[w] = inc;
    __syncthreads();
    if (tid == 0) {
        int s = 0;
        for (int q = 0; q < 8; ++q) {
            const int v = warpTot[q];
            warpTot[q] = s;
            s += v;
        }
        total = s;
    }
    __syncthreads();
    int pos = inc - cnt + warpTot[w];
    for (int q = 0; q < per; ++q) {
        const int j = tid * per + q;
        if (j < nInc && (j == 0 || keys[j] != keys[j - 1])) {
            uniq[pos] = keys[j];
            tileElems[base + pos] = keys[j];
            ++pos;
        }
    }
    __syncthreads();
    const int nU = total;
    if (tid == 0) {
        tileCnt[t] = nU;
        atomicMax(maxCnt, nU);
    }
    // slot of every (tile node, incident element) incidence = its position in the tile nodes' incidence lists laid end to end;
    // stored per (tile element, local node) so that the element phase scatters its records straight to where the node phase
    // reads them contiguously, in ascending element order
    for (int j = tid; j < nU * 4; j += 256) dst16[(size_t)base * 4 + j] = 0xffffu;
    __syncthreads();
    for (int s = tid >> 2; s < k1 - k0; s += 64) {
        const int i = perm[k0 + s];
        const int eb = n2ePtr[i], ne = n2ePtr[i + 1] - eb, off = nePrefix[k0 + s] - base;
        for (int k = tid & 3; k < ne; k += 4) {
            const int e = n2e[eb + k];
            int lo = 0, hi = nU - 1;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (uniq[mid] < e) lo = mid + 1;
                else hi = mid;
            }
            int li = 0;
            for (int q = 1; q < npe; ++q) li = (conn[(size_t)e * npe + q] == i) ? q : li;
            dst16[((size_t)base + lo) * 4 + li] = (unsigned short)(off + k);
        }
    }
    if (tid == 0) atomicMax(maxCnt + 4, nInc);
    // ---- the tile's node list = sorted distinct nodes of its elements; local connectivity of every element ----------------
    const int CAP2 = 4 * CAP, nCand = nU * npe;
    for (int j = tid; j < CAP2; j += 256) cand[j] = j < nCand ? conn[(size_t)uniq[j / npe] * npe + (j % npe)] : 0x7fffffff;
    __syncthreads();
    for (int size = 2; size <= CAP2; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int j = tid; j < CAP2 / 2; j += 256) {
                const int lo = 2 * j - (j & (stride - 1));
                const int hi = lo + stride;
                const bool up = (lo & size) == 0;
                const int a = cand[lo], b = cand[hi];
                if ((a > b) == up) {
                    cand[lo] = b;
                    cand[hi] = a;
                }
            }
            __syncthreads();
        }
    const int per2 = CAP2 / 256;
    int cnt2 = 0;
    for (int q = 0; q < per2; ++q) {
        const int j = tid * per2 + q;
        if (j < nCand && (j == 0 || cand[j] != cand[j - 1])) ++cnt2;
    }
    int inc2 = cnt2;
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc2, o);
        if (lane >= o) inc2 += v;
    }
    __syncthreads();
    if (lane == 31) warpTot[w] = inc2;
    __syncthreads();
    if (tid == 0) {
        int s2 = 0;
        for (int q = 0; q < 8; ++q) {
            const int v = warpTot[q];
            warpTot[q] = s2;
            s2 += v;
        }
        total = s2;
        nodeOff = atomicAdd(maxCnt + 2, s2);
        nodeStart[t] = nodeOff;
        nodeCnt[t] = s2;
        atomicMax(maxCnt + 1, s2);
        if (s2 > CAP) atomicOr(maxCnt + 3, 1);
    }
    __syncthreads();
    const int nLoc = total;
    int pos2 = inc2 - cnt2 + warpTot[w];
    for (int q = 0; q < per2; ++q) {  // distinct nodes -> keys[] (free since the element sort) and the global list
        const int j = tid * per2 + q;
        if (j < nCand && (j == 0 || cand[j] != cand[j - 1])) {
            if (pos2 < CAP) keys[pos2] = cand[j];
            if (nodeOff + pos2 < nodeListCap) tileNodes[nodeOff + pos2] = cand[j];  // (too small: the host re-runs with the total)
            ++pos2;
        }
    }
    __syncthreads();
    if (nLoc <= CAP)
        for (int j = tid; j < nCand; j += 256) {
            const int nd = conn[(size_t)uniq[j / npe] * npe + (j % npe)];
            int lo = 0, hi = nLoc - 1;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (keys[mid] < nd) lo = mid + 1;
                else hi = mid;
            }
            lconn[((size_t)base + j / npe) * 4 + (j % npe)] = (unsigned short)lo;
        }
}

// ---- run time ---------------------------------------------------------------------------------------------------------------
// Stage the records (x, y, z, p | u, v, w, rho) of the tile's node list in shared memory: one independent pair of 32-byte loads
// per thread, all in flight at once -- the element phase then never waits on global memory.
__device__ __forceinline__ void stageTileNodes(const TileArgs& ta, int t, const double* __restrict__ X4, const double* __restrict__ V4,
                                               double* __restrict__ nodeS) {
    const int off = ta.nodeStart[t], n = ta.nodeCnt[t];
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        const int nd = __ldg(ta.tileNodes + off + k);
        const D4 xr = ld4(X4 + (size_t)nd * 4), vr = ld4(V4 + (size_t)nd * 4);
        double* r = nodeS + (size_t)k * TILE_NSTRIDE;
        *reinterpret_cast<double2*>(r) = make_double2(xr.x, xr.y);
        *reinterpret_cast<double2*>(r + 2) = make_double2(xr.z, xr.w);
        *reinterpret_cast<double2*>(r + 4) = make_double2(vr.x, vr.y);
        *reinterpret_cast<double2*>(r + 6) = make_double2(vr.z, vr.w);
    }
}
// element j of the tile from the staged records: same values, same arithmetic as loadElem
template <int DIM>
__device__ __forceinline__ void loadElemStaged(const double* __restrict__ nodeS, const unsigned short* __restrict__ lc,
                                               double (&px3)[DIM + 1][3], double (&xw)[DIM + 1], double (&vel)[DIM + 1][DIM],
                                               double (&vw)[DIM + 1], ElemGeo<DIM>& G) {
    constexpr int NPE = DIM + 1;
    const uint2 w = *reinterpret_cast<const uint2*>(lc);  // 4 x 16-bit local node indices
    const unsigned li[4] = {w.x & 0xffffu, w.x >> 16, w.y & 0xffffu, w.y >> 16};
    double px[NPE][DIM];
#pragma unroll
    for (int m = 0; m < NPE; ++m) {
        const double* r = nodeS + (size_t)li[m] * TILE_NSTRIDE;
        const double2 x01 = *reinterpret_cast<const double2*>(r), x23 = *reinterpret_cast<const double2*>(r + 2);
        const double2 v01 = *reinterpret_cast<const double2*>(r + 4), v23 = *reinterpret_cast<const double2*>(r + 6);
        px[m][0] = x01.x, px[m][1] = x01.y;
        px3[m][0] = x01.x, px3[m][1] = x01.y, px3[m][2] = x23.x;
        vel[m][0] = v01.x, vel[m][1] = v01.y;
        if constexpr (DIM == 3) {
            px[m][2] = x23.x;
            vel[m][2] = v23.x;
        }
        xw[m] = x23.y;
        vw[m] = v23.y;
    }
    buildGeo<DIM>(px, G);
}

// ---- continuity, CDS_dpdt: element phase into shared memory, then the ordered nodal sum + epilogue -----------------------
template <int DIM>
__global__ void __launch_bounds__(256, 2) k_wc_cont_tile(const TileArgs ta, const WcArgs a, const double* __restrict__ X4,
                                                         const double* __restrict__ V4, double* __restrict__ X4n, double* __restrict__ V4n,
                                                         double* __restrict__ hminOut) {
    constexpr int NPE = DIM + 1;
    constexpr double PHI = 1.0 / ((DIM + 1) * (DIM + 2));
    extern __shared__ __align__(32) double smemT[];
    double* nodeS = smemT;                                        // nodeCap x TILE_NSTRIDE
    double* recS = smemT + (size_t)ta.nodeCap * TILE_NSTRIDE;     // slotCap x (F0 contribution, V/NPE, he, -)
    const int t = ta.tile0 + blockIdx.x, tid = threadIdx.x;
    const int k0 = t * ta.T;
    const int start = ta.nePrefix[k0], cnt = ta.tileCnt[t];
    const double dtStep = a.dtPtr ? *a.dtPtr : a.dt;
    stageTileNodes(ta, t, X4, V4, nodeS);
    __syncthreads();
    for (int j = tid; j < cnt; j += 256) {
        double P[NPE], vel[NPE][DIM], rho[NPE], px[NPE][3];
        ElemGeo<DIM> G;
        loadElemStaged<DIM>(nodeS, ta.lconn + ((size_t)start + j) * 4, px, P, vel, rho, G);
        double sumP = 0, divv = 0;
#pragma unroll
        for (int q = 0; q < NPE; ++q) {
            sumP += P[q];
#pragma unroll
            for (int c = 0; c < DIM; ++c) divv += G.g[c][q] * vel[q][c];
        }
        // F0_i = -dt V divv (K0/NPE + K0' PHI (p_i + sumP)) + (meduri ? V PHI (p_i + sumP) : (V/NPE) p_i)  (as k_wc_cont_elem)
        const double adv = -dtStep * G.V * divv;
        const double alpha = adv * (a.K0 / NPE + a.K0p * PHI * sumP) + (a.meduri ? G.V * PHI * sumP : 0.0);
        const double beta = adv * (a.K0p * PHI) + (a.meduri ? G.V * PHI : G.V / NPE);
        const double he = elemHe<DIM>(px), mq = G.V / NPE;
        const uint2 dw = *reinterpret_cast<const uint2*>(ta.dst16 + ((size_t)start + j) * 4);
        const unsigned dst[4] = {dw.x & 0xffffu, dw.x >> 16, dw.y & 0xffffu, dw.y >> 16};
#pragma unroll
        for (int q = 0; q < NPE; ++q)
            if (dst[q] != 0xffffu) {
                double* r = recS + (size_t)dst[q] * 4;
                *reinterpret_cast<double2*>(r) = make_double2(alpha + beta * P[q], mq);
                r[2] = he;
            }
    }
    __syncthreads();
    const int s = tid >> 2, sub = tid & 3;
    const int k = k0 + s;
    const bool valid = s < ta.T && k < ta.nRows;
    const int i = valid ? __ldg(ta.perm + k) : 0;
    double m = 0, F0 = 0, hmin = 1.7976931348623157e308;
    if (valid) {
        const int off = ta.nePrefix[k] - start, ne = ta.nePrefix[k + 1] - ta.nePrefix[k];
        for (int kk = sub; kk < ne; kk += 4) {
            const double* r = recS + (size_t)(off + kk) * 4;
            const double2 r01 = *reinterpret_cast<const double2*>(r);
            F0 += r01.x;
            m += r01.y;
            hmin = nanMin(r[2], hmin);
        }
    }
    m = groupSum<4>(m);
    F0 = groupSum<4>(F0);
    hmin = groupMin<4>(hmin);
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

// ---- momentum: per-(element, local node) records (F, lumped rho-mass) in shared memory, plane per local node ------------------
template <int DIM>
__global__ void __launch_bounds__(256, 2) k_wc_mom_tile(const TileArgs ta, const WcArgs a, const double* __restrict__ X4,
                                                        const double* __restrict__ V4, double* __restrict__ V4out,
                                                        double* __restrict__ A4out, double* __restrict__ X4out, double* __restrict__ cfl2) {
    constexpr int NPE = DIM + 1;
    constexpr double PHI = 1.0 / ((DIM + 1) * (DIM + 2));
    extern __shared__ __align__(32) double smemT[];
    double* nodeS = smemT;                                        // nodeCap x TILE_NSTRIDE
    double* recS = smemT + (size_t)ta.nodeCap * TILE_NSTRIDE;     // slotCap x (Fx, Fy, Fz, lumped mass)
    const int t = ta.tile0 + blockIdx.x, tid = threadIdx.x;
    const int k0 = t * ta.T;
    const int start = ta.nePrefix[k0], cnt = ta.tileCnt[t];
    const double dtStep = a.dtPtr ? *a.dtPtr : a.dt;
    stageTileNodes(ta, t, X4, V4, nodeS);
    __syncthreads();
    for (int j = tid; j < cnt; j += 256) {
        double P[NPE], vel[NPE][DIM], rho[NPE], px[NPE][3];
        ElemGeo<DIM> G;
        loadElemStaged<DIM>(nodeS, ta.lconn + ((size_t)start + j) * 4, px, P, vel, rho, G);
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
                sig[aa][c] = a.mu * sv;
            }
        const uint2 dw = *reinterpret_cast<const uint2*>(ta.dst16 + ((size_t)start + j) * 4);
        const unsigned dst[4] = {dw.x & 0xffffu, dw.x >> 16, dw.y & 0xffffu, dw.y >> 16};
#pragma unroll
        for (int q = 0; q < NPE; ++q) {
            if (dst[q] == 0xffffu) continue;  // node q belongs to another tile (which computes this element too)
            const double li_mass = G.V * PHI * (rho[q] + sumR);  // lumped rho-mass == sum_g w (N.rho) N_q
            double F[3] = {0.0, 0.0, 0.0};
#pragma unroll
            for (int aa = 0; aa < DIM; ++aa) {
                double sg = 0;
#pragma unroll
                for (int c = 0; c < DIM; ++c) sg += sig[aa][c] * G.g[c][q];
                F[aa] = -G.V * sg + G.V * pbar * G.g[aa][q] + a.body[aa] * li_mass;
            }
            double* r = recS + (size_t)dst[q] * 4;
            *reinterpret_cast<double2*>(r) = make_double2(F[0], F[1]);
            *reinterpret_cast<double2*>(r + 2) = make_double2(F[2], li_mass);
        }
    }
    __syncthreads();
    const int s = tid >> 2, sub = tid & 3;
    const int k = k0 + s;
    const bool valid = s < ta.T && k < ta.nRows;
    const int i = valid ? __ldg(ta.perm + k) : 0;
    double M = 0, F[DIM];
#pragma unroll
    for (int c = 0; c < DIM; ++c) F[c] = 0;
    if (valid) {
        const int off = ta.nePrefix[k] - start, ne = ta.nePrefix[k + 1] - ta.nePrefix[k];
        for (int kk = sub; kk < ne; kk += 4) {
            const double* r = recS + (size_t)(off + kk) * 4;
            const double2 r01 = *reinterpret_cast<const double2*>(r), r23 = *reinterpret_cast<const double2*>(r + 2);
            F[0] += r01.x;
            F[1] += r01.y;
            if constexpr (DIM == 3) F[2] += r23.x;
            M += r23.y;
        }
    }
    M = groupSum<4>(M);
#pragma unroll
    for (int c = 0; c < DIM; ++c) F[c] = groupSum<4>(F[c]);
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
