// mg.cu -- aggregation multigrid preconditioner for the BiCGSTAB solve of the monolithic PSPG system.
//
// The reference factorises m_A with Eigen::SparseLU every Picard iterate (MomContEquationPSPG.inl:281-290).  The Krylov
// replacement with a node-block Jacobi preconditioner needs O(1/h) iterations (1622 at 2 M tets); the Schur complement of
// the (v,p) system is a pressure Laplacian, so a multilevel correction is what removes the mesh dependence.  Design:
//   * monolithic: every level keeps the (dim+1)x(dim+1) node blocks, so all levels reuse the node-block SpMV kernel
//     (spmv.cuh) -- smoothing sweep and residual are SpMV epilogues, nothing else touches A;
//   * aggregates = nodes that fall in the same cell of a uniform grid laid over the level's node coordinates (PFEM keeps
//     the node spacing near h_char everywhere, Mesh.cpp:27-147), cell size chosen so that an aggregate holds ~2^dim nodes;
//     piecewise-constant prolongation in the physical variables (v, p), Galerkin coarse matrices P^T A P;
//   * node-block Jacobi smoothing with an l1-type LOCAL damping theta / r_i, r_i = sum_j ||A_ii^-1 A_ij||_inf taken in the
//     equilibrated variables: the (v,p) coupling gives D^-1 A complex eigenvalues and a uniform damping that is stable on
//     one mesh diverges on the next (measured: 0.5 fine at 1.3 M tets, divergent at 2 M); V(nu,nu) cycle (nu+1 sweeps on the cheap coarse levels), over-correction
//     of the coarse update (standard for unsmoothed aggregation), dense inverse on the coarsest level (<= 32 nodes);
//   * symbolic part (aggregates, coarse patterns, fine-block -> coarse-slot map) once per topology; numeric part
//     (Galerkin sums, block inverses, coarsest inverse) once per assembly.  Everything is gather-style and ordered, no
//     floating-point atomics: the preconditioner, hence the whole solve, is bit-reproducible run to run.
// On a partitioned mesh the hierarchy is GLOBAL: aggregates are rank-local, but the Galerkin operators keep the couplings
// across ranks (ghost aggregates + a halo plan per distributed level); once a level has fewer than ~50 k nodes in total it
// is replicated on every rank (all-gather of its rows) and the rest of the hierarchy, down to the dense coarsest solve, is
// computed redundantly -- no exchange on the small levels, one all-gather of the restricted residual per cycle.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#include "common.cuh"
#include "spmv.cuh"

struct MgLevel {
    int n = 0;        // rows (nodes) of this level
    int nVec = 0;     // vector extent in nodes (level 0 of a partitioned mesh: owned + ghost)
    int maxNb = 0;
    int64_t nBlocks = 0;
    const int* nbrPtr = nullptr;  // level 0 aliases the context's pattern and values
    const int* nbr = nullptr;
    const int* diagSlot = nullptr;
    const double* Aval = nullptr;
    const double* X = nullptr;    // 4 doubles per node
    DevBuf<int> nbrPtrB, nbrB, diagSlotB;
    DevBuf<double> AvalB, XB;
    DevBuf<float> Af;             // fp32 copy of Aval streamed by the cycle (PFEM_MG_FP32=0: not used)
    // transfer to the next level
    int nc = 0;
    DevBuf<int> agg, aggPtr, aggNodes, cslot;
    // numeric
    DevBuf<double> Dw;            // omega * A_ii^-1
    DevBuf<float> DwF;            // fp32 copy for the fp32-vector cycle
    DevBuf<double> sc;            // 1/sqrt|a_dd| per dof (l1 damping works in the equilibrated variables)
    double omega = 0.5;           // damping of this level's smoother (tuned or fixed)
    DevBuf<double> b, xa, xb, t, xo;
    DevBuf<double> w, xo2;        // W-cycle: right-hand side and result of the second coarse visit
    // partitioned mesh: a DISTRIBUTED level holds this rank's rows; its vectors carry a ghost tail that
    // `plan` refreshes from the owners before every sweep.  Below a size threshold the next level is REPLICATED: every rank
    // holds all of its rows (this rank contributed rows [rowOff, rowOff + ncLocal)) and works on it redundantly, so the
    // small levels need no exchange at all -- one all-gather of the restricted residual per cycle.
    bool distributed = false;
    HaloPlan* plan = nullptr;            // level 0: the context's plan
    std::unique_ptr<HaloPlan> ownPlan;   // deeper distributed levels
    bool nextReplicated = false;
    int ncLocal = 0, rowOff = 0;
    std::vector<int64_t> rowCounts, rowDispls;  // per rank, coarse rows
    std::vector<int64_t> blkCounts, blkDispls;  // per rank, coarse blocks
    void resetLogical() {  // a recycled level object keeps its buffers, nothing else
        n = nVec = maxNb = nc = ncLocal = rowOff = 0;
        nBlocks = 0;
        nbrPtr = nbr = diagSlot = nullptr;
        Aval = X = nullptr;
        distributed = nextReplicated = false;
        plan = nullptr;
        omega = 0.5;
    }
};
struct MgHierarchy {
    // The level objects (and every device buffer they own) persist across symbolic rebuilds: an incompressible run remeshes
    // every time step, and freeing / re-allocating ~100 device buffers per step costs tens of milliseconds of cudaFree /
    // cudaMalloc (measured: 60-360 ms per rebuild, growing).  `lev` lists the levels of the current hierarchy.
    std::vector<std::unique_ptr<MgLevel>> pool;
    std::vector<MgLevel*> lev;
    DevBuf<double> sBox, sXc, sTmp, sPart;  // scratch of the symbolic phase (grow-only)
    DevBuf<int> sKey, sCell;
    MgLevel& levelObject(size_t l) {
        while (pool.size() <= l) pool.push_back(std::make_unique<MgLevel>());
        return *pool[l];
    }
    DevBuf<double> dense;  // coarsest level: A^-1, nD x nD
    DevBuf<int> flag;
    int nD = 0;
    bool denseOk = false;
    bool symbolicValid = false, numericValid = false, symbolicFailed = false;
    int nu = 2, nuCoarse = 2;
    int preFine = -1, postFine = -1, preCoarse = -1, postCoarse = -1;  // -1: nu / nuCoarse (PFEM_MG_PRE|POST|PREC|POSTC)
    bool f32v = false;         // the cycle runs on fp32 vectors (flexible GMRES outside); level buffers are reinterpreted
    DevBuf<float> inF, outF;   // fp32 copies of the cycle's right-hand side and result
    bool rhsIsF32 = false;     // the caller writes inF itself (mgRhsF): no conversion pass at the head of the cycle
    int wFrom = -1;  // W-cycle: the coarse correction of every level >= wFrom is computed twice (-1: V-cycle)
    double fixedOmega = 0.0;  // > 0: the caller's damping on every level; 0: tuned per level (tuneDamping)
    double over = 1.5;
    bool tuned = false;
    double stamp = 0.0;       // dt of the system the dampings were tuned for
    // one captured CUDA graph per output vector: a cycle is ~40 launches, most of them on tiny coarse levels
    struct GraphSlot {
        double* out = nullptr;
        cudaGraphExec_t exec = nullptr;
        unsigned long long sig = 0;
        int launches = 0;
    };
    std::vector<GraphSlot> graphs;
    bool graphBroken = false;
    void dropGraphs() {
        for (auto& g : graphs)
            if (g.exec) cudaGraphExecDestroy(g.exec);
        graphs.clear();
    }
    ~MgHierarchy() { dropGraphs(); }
};

namespace {

inline long long __double_as_longlong_host(double v) {
    long long r;
    memcpy(&r, &v, sizeof r);
    return r;
}

constexpr int COARSEST_NODES = 32;
constexpr int NBR_CAP = 512;  // distinct coarse neighbours a warp can collect

// ---- symbolic ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_bbox(const double* __restrict__ X, int n, int dim, double* __restrict__ out) {
    __shared__ double slo[3][32], shi[3][32];
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
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
        out[(size_t)blockIdx.x * 6 + d] = a;
        out[(size_t)blockIdx.x * 6 + 3 + d] = b;
    }
}
// box of the per-block boxes (minimum / maximum: order-independent, so the result is the same bits as one block's)
__global__ void k_bbox_final(const double* __restrict__ part, int nPart, double* __restrict__ out) {
    const int d = threadIdx.x;
    if (d >= 3) return;
    double a = 1e300, b = -1e300;
    for (int k = 0; k < nPart; ++k) a = fmin(a, part[(size_t)k * 6 + d]), b = fmax(b, part[(size_t)k * 6 + 3 + d]);
    out[d] = a;
    out[3 + d] = b;
}

// mean element size -> node spacing h0 of level 0 (two-stage ordered sum: deterministic)
__global__ void __launch_bounds__(256) k_vol_partial(const int* __restrict__ conn, int nElems, int dim,
                                                     const double* __restrict__ X4, double* __restrict__ partial) {
    __shared__ double sh[8];
    const int npe = dim + 1;
    double s = 0;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nElems; e += gridDim.x * blockDim.x) {
        const int* en = conn + (size_t)e * npe;
        const double* p0 = X4 + (size_t)en[0] * 4;
        double J[3][3];
        for (int m = 0; m < dim; ++m) {
            const double* pm = X4 + (size_t)en[m + 1] * 4;
            for (int d = 0; d < dim; ++d) J[d][m] = pm[d] - p0[d];
        }
        double det;
        if (dim == 2) det = J[0][0] * J[1][1] - J[1][0] * J[0][1];
        else
            det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                  J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
        s += fabs(det);  // = dim! * element size: the volume of the cube whose Kuhn split has this element size
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int k = 0; k < 8; ++k) t += sh[k];
        partial[blockIdx.x] = t;
    }
}
__global__ void k_vol_final(const double* __restrict__ partial, int nb, double* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double t = 0;
        for (int k = 0; k < nb; ++k) t += partial[k];
        *out = t;
    }
}

__global__ void k_cell_key(const double* __restrict__ X, int n, int dim, double lox, double loy, double loz, double invH, int nx,
                           int ny, int nz, int* __restrict__ key, int* __restrict__ flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* x = X + (size_t)i * 4;
    const int ix = min(nx - 1, max(0, (int)floor((x[0] - lox) * invH + 1e-6)));
    const int iy = min(ny - 1, max(0, (int)floor((x[1] - loy) * invH + 1e-6)));
    const int iz = dim == 3 ? min(nz - 1, max(0, (int)floor((x[2] - loz) * invH + 1e-6))) : 0;
    const int k = (iz * ny + iy) * nx + ix;
    key[i] = k;
    flag[k] = 1;
}
__global__ void k_assign_agg(int n, const int* __restrict__ key, const int* __restrict__ cellId, int* __restrict__ agg,
                             int* __restrict__ count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int a = cellId[key[i]];
    agg[i] = a;
    atomicAdd(&count[a], 1);
}
__global__ void k_fill_members(int n, const int* __restrict__ agg, const int* __restrict__ aggPtr, int* __restrict__ cursor,
                               int* __restrict__ aggNodes) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int a = agg[i];
    aggNodes[aggPtr[a] + atomicAdd(&cursor[a], 1)] = i;
}
// ascending member order (fixes the summation order of every later gather) + coarse coordinates = member mean
__global__ void k_sort_members(int nc, int dim, const int* __restrict__ aggPtr, int* __restrict__ aggNodes,
                               const double* __restrict__ X, double* __restrict__ Xc) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= nc) return;
    const int b = aggPtr[a], len = aggPtr[a + 1] - b;
    int* v = aggNodes + b;
    for (int i = 1; i < len; ++i) {
        const int k = v[i];
        int j = i - 1;
        while (j >= 0 && v[j] > k) {
            v[j + 1] = v[j];
            --j;
        }
        v[j + 1] = k;
    }
    double s[3] = {0, 0, 0};
    for (int i = 0; i < len; ++i)
        for (int d = 0; d < dim; ++d) s[d] += X[(size_t)v[i] * 4 + d];
    for (int d = 0; d < 3; ++d) Xc[(size_t)a * 4 + d] = len > 0 ? s[d] / len : 0.0;
    Xc[(size_t)a * 4 + 3] = 0.0;
}

// coarse neighbour list of aggregate I = sorted set { agg[j] : i in I, j in nbr(i), j owned }.  One warp per aggregate.
template <bool FILL>
__global__ void k_coarse_nbr(int nc, const int* __restrict__ aggPtr, const int* __restrict__ aggNodes, const int* __restrict__ agg,
                             int nFine, const int* __restrict__ nbrPtr, const int* __restrict__ nbr, int* __restrict__ cPtr,
                             int* __restrict__ cNbr, int* __restrict__ cDiag, int* __restrict__ misc, int rowOff) {
    __shared__ int lists[8][NBR_CAP];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int I = blockIdx.x * 8 + w;
    if (I >= nc) return;
    int* list = lists[w];
    int cnt = 0;
    for (int m = aggPtr[I]; m < aggPtr[I + 1]; ++m) {
        const int i = aggNodes[m];
        const int b0 = nbrPtr[i], nb = nbrPtr[i + 1] - b0;
        for (int s0 = 0; s0 < nb; s0 += 32) {
            const int s = s0 + lane;
            int J = -1;
            if (s < nb) {
                const int j = nbr[b0 + s];
                if (j < nFine) J = agg[j];
            }
            bool isNew = J >= 0;
            const int lim = min(cnt, NBR_CAP);
            for (int k = 0; k < lim; ++k)
                if (list[k] == J) isNew = false;
            const unsigned same = __match_any_sync(0xffffffffu, isNew ? J : -1 - lane);
            const bool leader = isNew && (__ffs(same) - 1 == lane);
            const unsigned lead = __ballot_sync(0xffffffffu, leader);
            const int pos = cnt + __popc(lead & ((1u << lane) - 1u));
            if (leader && pos < NBR_CAP) list[pos] = J;
            cnt += __popc(lead);
            __syncwarp();
        }
    }
    if (cnt > NBR_CAP) {
        if (lane == 0) atomicOr(misc + 1, 1);  // overflow: the caller gives up on the hierarchy
        cnt = NBR_CAP;
    }
    if (!FILL) {
        if (lane == 0) {
            cPtr[I] = cnt;
            atomicMax(misc, cnt);
        }
    } else {
        const int base = cPtr[I];
        for (int k = lane; k < cnt; k += 32) {
            const int v = list[k];
            int rank = 0;
            for (int q = 0; q < cnt; ++q) rank += list[q] < v;
            cNbr[base + rank] = v;
            if (v == I + rowOff) cDiag[I] = rank;  // rowOff: this rank's first row of a replicated level
        }
    }
}
// slot of every fine block in the coarse row of its aggregate (-1: ghost column, dropped)
__global__ void k_cslot(int nFine, const int* __restrict__ nbrPtr, const int* __restrict__ nbr, const int* __restrict__ agg,
                        const int* __restrict__ cPtr, const int* __restrict__ cNbr, int* __restrict__ cslot, int nCols = -1) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nFine) return;
    if (nCols < 0) nCols = nFine;  // columns with an aggregate: the rows, plus the ghost tail of a distributed level
    const int I = agg[i];
    const int c0 = cPtr[I], cn = cPtr[I + 1] - c0;
    for (int k = nbrPtr[i]; k < nbrPtr[i + 1]; ++k) {
        const int j = nbr[k];
        int res = -1;
        if (j < nCols) {
            const int J = agg[j];
            int lo = 0, hi = cn - 1;
            while (lo <= hi) {
                const int mid = (lo + hi) >> 1;
                const int v = cNbr[c0 + mid];
                if (v == J) {
                    res = mid;
                    break;
                }
                if (v < J) lo = mid + 1;
                else hi = mid - 1;
            }
        }
        cslot[k] = res;
    }
}

__global__ void k_agg_shift_to_double(int n, int* __restrict__ agg, int shift, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int a = agg[i] + shift;
    agg[i] = a;
    out[i] = (double)a;
}
__global__ void k_double_to_int(int n, const double* __restrict__ in, int* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int)in[i];
}

// ---- numeric ----------------------------------------------------------------------------------------------------------
// Galerkin sum A_c(I,J) = sum_{i in I, j in J} A(i,j), block-wise.  One half-warp per coarse row, lane = block entry: every
// lane only ever touches "its" entry of each accumulator block, members and fine blocks are walked in ascending order.
template <int BS>
__global__ void k_galerkin(int nc, const int* __restrict__ aggPtr, const int* __restrict__ aggNodes,
                                                  const int* __restrict__ fPtr, const double* __restrict__ fA,
                                                  const int* __restrict__ cslot, const int* __restrict__ cPtr,
                                                  double* __restrict__ cA, int nbcap) {
    constexpr int BB = BS * BS;
    extern __shared__ double accAll[];
    const int l = threadIdx.x & 15;
    const int I = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    double* acc = accAll + (size_t)(threadIdx.x >> 4) * nbcap * BB;
    if (I >= nc || l >= BB) return;
    const int c0 = cPtr[I], cn = cPtr[I + 1] - c0;
    for (int k = 0; k < cn; ++k) acc[k * BB + l] = 0.0;
    // four fine blocks per step with their loads issued together (the walk is latency-bound: one dependent global load per
    // block otherwise); the sums keep the ascending order
    for (int m = aggPtr[I]; m < aggPtr[I + 1]; ++m) {
        const int i = aggNodes[m];
        const int k1 = fPtr[i + 1];
        for (int k = fPtr[i]; k < k1; k += 4) {
            int cs[4];
            double a[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool in = k + u < k1;
                cs[u] = in ? cslot[k + u] : -1;
                a[u] = in ? fA[(size_t)(k + u) * BB + l] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (cs[u] >= 0) acc[cs[u] * BB + l] += a[u];
        }
    }
    for (int k = 0; k < cn; ++k) cA[(size_t)(c0 + k) * BB + l] = acc[k * BB + l];
}

// Dw_i = omega * A_ii^-1 (Gauss-Jordan with partial pivoting); singular block -> omega * diag^-1
template <int BS>
__global__ void k_mg_block_inv(int n, const int* __restrict__ nbrPtr, const int* __restrict__ diagSlot,
                               const double* __restrict__ Aval, double omega, double* __restrict__ Dw) {
    const int nd = blockIdx.x * blockDim.x + threadIdx.x;
    if (nd >= n) return;
    const double* Ad = Aval + ((size_t)nbrPtr[nd] + diagSlot[nd]) * BS * BS;
    double M[BS][2 * BS];
#pragma unroll
    for (int r = 0; r < BS; ++r)
#pragma unroll
        for (int c = 0; c < BS; ++c) {
            M[r][c] = Ad[r * BS + c];
            M[r][BS + c] = (r == c) ? 1.0 : 0.0;
        }
    bool singular = false;
#pragma unroll
    for (int k = 0; k < BS; ++k) {
        int piv = k;
        double best = fabs(M[k][k]);
#pragma unroll
        for (int r = k + 1; r < BS; ++r)
            if (fabs(M[r][k]) > best) best = fabs(M[r][k]), piv = r;
        if (!(best > 0.0)) singular = true;
#pragma unroll
        for (int r = k + 1; r < BS; ++r)
            if (r == piv) {
#pragma unroll
                for (int c = 0; c < 2 * BS; ++c) {
                    const double t = M[k][c];
                    M[k][c] = M[r][c];
                    M[r][c] = t;
                }
            }
        const double inv = 1.0 / M[k][k];
#pragma unroll
        for (int c = 0; c < 2 * BS; ++c) M[k][c] *= inv;
#pragma unroll
        for (int r = 0; r < BS; ++r)
            if (r != k) {
                const double f = M[r][k];
#pragma unroll
                for (int c = 0; c < 2 * BS; ++c) M[r][c] -= f * M[k][c];
            }
    }
#pragma unroll
    for (int r = 0; r < BS; ++r)
#pragma unroll
        for (int c = 0; c < BS; ++c) {
            double v = omega * M[r][BS + c];
            if (singular) {
                const double d = Ad[r * BS + r];
                v = (r == c) ? (d != 0.0 ? omega / d : omega) : 0.0;
            }
            Dw[(size_t)nd * BS * BS + r * BS + c] = v;
        }
}

// l1-type local damping: Dw_i *= theta / r_i with r_i = sum_j ||A_ii^-1 A_ij||_inf (>= 1), the inf-norm of block row i of
// D^-1 A.  Rows of a consistent mass matrix get theta/2.5 (3-D), rows next to slivers or with strong (v,p) coupling get
// less, without one bad row dictating the damping of the whole level.  Half-warp per row, lane = entry of the 4x4 product.
// sc = 1/sqrt|a_dd| per dof: the block norms below are taken in the equilibrated variables (v and p differ by 1e6 in scale)
template <int BS>
__global__ void k_mg_diag_scale(int n, const int* __restrict__ nbrPtr, const int* __restrict__ diagSlot,
                                const double* __restrict__ Aval, double* __restrict__ sc) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * BS) return;
    const int i = t / BS, d = t % BS;
    const double a = fabs(Aval[((size_t)nbrPtr[i] + diagSlot[i]) * BS * BS + d * BS + d]);
    sc[t] = a > 0.0 ? rsqrt(a) : 1.0;
}
template <int BS>
__global__ void k_mg_l1_scale(int n, const int* __restrict__ nbrPtr, const int* __restrict__ nbr, const double* __restrict__ Aval,
                              const double* __restrict__ sc, double theta, double wcap, double* __restrict__ Dw, double* __restrict__ rmax,
                              float* __restrict__ Af) {
    constexpr int BB = BS * BS;
    const int l = threadIdx.x & 15;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const bool act = i < n;
    const int r = l / BS, cc = l % BS;  // lanes >= BB idle (2-D)
    const bool ent = act && l < BB;
    double dinvRow[BS];
#pragma unroll
    for (int k = 0; k < BS; ++k) dinvRow[k] = ent ? Dw[(size_t)i * BB + r * BS + k] : 0.0;
    double ri = 0;
    const int b0 = act ? nbrPtr[i] : 0, b1 = act ? nbrPtr[i + 1] : 0;
    const unsigned hm = 0xffffu << (threadIdx.x & 16);  // this half-warp
    const double scMine = ent ? sc[(size_t)i * BS + r] : 1.0;
    // the block entry, the neighbour index and its scale are fetched one block ahead (two for the index): the walk is
    // latency-bound otherwise (round 2: 603 -> ~300 us at C4)
    int nb1 = (ent && b0 < b1) ? nbr[b0] : 0;
    int nb2 = (ent && b0 + 1 < b1) ? nbr[b0 + 1] : 0;
    double a1 = (ent && b0 < b1) ? Aval[(size_t)b0 * BB + l] : 0.0;
    double s1 = (ent && b0 < b1) ? sc[(size_t)nb1 * BS + cc] : 0.0;
    for (int k = b0; k < b1; ++k) {
        const bool more = ent && k + 1 < b1;
        const double a2 = more ? Aval[(size_t)(k + 1) * BB + l] : 0.0;
        const double s2 = more ? sc[(size_t)nb2 * BS + cc] : 0.0;
        const int nb3 = (ent && k + 2 < b1) ? nbr[k + 2] : 0;
        // every lane holds its own entry of the block (one coalesced 128-byte row per half-warp); the 4x4 product takes the
        // column entries from the owning lanes
        const double aMine = a1;
        if (Af && ent) Af[(size_t)k * BB + l] = (float)aMine;  // the cycle's fp32 copy of A, written in the same pass
        double p = 0;
#pragma unroll
        for (int q = 0; q < BS; ++q) {
            const double aq = __shfl_sync(hm, aMine, (threadIdx.x & 16) + min(q * BS + cc, 15));
            p += dinvRow[q] * aq;
        }
        if (ent) p *= s1 / scMine;  // S_i^-1 (D^-1 A)_ij S_j
        p = ent ? fabs(p) : 0.0;
        // row sums over cc (lanes of the same r), then max over r
        double rs = 0;
#pragma unroll
        for (int q = 0; q < BS; ++q) {
            const double v = __shfl_sync(hm, p, (threadIdx.x & 16) + min(r * BS + q, 15));
            rs += (l < BB) ? v : 0.0;
        }
        double mx = rs;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(hm, mx, o));
        ri += mx;
        a1 = a2, s1 = s2, nb2 = nb3;
    }
    ri = fmax(ri, 1.0);
    if (ent) Dw[(size_t)i * BB + l] *= fmin(theta / ri, wcap);
    if (act && l == 0 && rmax) atomicMax(reinterpret_cast<unsigned long long*>(rmax), (unsigned long long)__double_as_longlong(ri));
}

// coarsest level (<= 32 nodes, <= 128 dofs): dense matrix gathered into shared memory, inverted in place by Gauss-Jordan
// with partial pivoting (row swaps, columns unscrambled at the end), written back as a dense nD x nD inverse.  One CTA.
__global__ void __launch_bounds__(1024) k_dense_invert(int n, int BS, const int* __restrict__ nbrPtr, const int* __restrict__ nbr,
                                                       const double* __restrict__ Aval, double* __restrict__ Dinv,
                                                       int* __restrict__ flag) {
    extern __shared__ double sm[];
    const int nD = n * BS;
    double* a = sm;                   // nD x nD
    double* colk = sm + nD * nD;      // nD
    int* piv = reinterpret_cast<int*>(colk + nD);
    __shared__ double sval[32];
    __shared__ int sidx[32];
    __shared__ int pivRow;
    __shared__ double pivVal;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, w = tid >> 5, nw = nt >> 5;
    for (int t = tid; t < nD * nD; t += nt) a[t] = 0.0;
    __syncthreads();
    for (int i = tid; i < n; i += nt)
        for (int k = nbrPtr[i]; k < nbrPtr[i + 1]; ++k) {
            const int j = nbr[k];
            for (int r = 0; r < BS; ++r)
                for (int c = 0; c < BS; ++c) a[(i * BS + r) * nD + j * BS + c] = Aval[(size_t)k * BS * BS + r * BS + c];
        }
    __syncthreads();
    for (int k = 0; k < nD; ++k) {
        double best = -1.0;
        int bi = k;
        for (int r = k + tid; r < nD; r += nt) {
            const double v = fabs(a[r * nD + k]);
            if (v > best) best = v, bi = r;
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) best = ov, bi = oi;
        }
        if (lane == 0) sval[w] = best, sidx[w] = bi;
        __syncthreads();
        if (tid == 0) {
            double b = sval[0];
            int id = sidx[0];
            for (int q = 1; q < nw; ++q)
                if (sval[q] > b || (sval[q] == b && sidx[q] < id)) b = sval[q], id = sidx[q];
            pivRow = id;
            piv[k] = id;
            pivVal = b > 1e-300 ? a[id * nD + k] : 0.0;
            if (!(b > 1e-300)) *flag = 1;
        }
        __syncthreads();
        const int p = pivRow;
        const double d = pivVal;
        if (d == 0.0) return;
        // swap rows k and p; pivot row: a[k][k] <- 1, then divide by the pivot
        for (int c = tid; c < nD; c += nt) {
            const double up = a[p * nD + c], uk = a[k * nD + c];
            a[p * nD + c] = uk;
            a[k * nD + c] = ((c == k) ? 1.0 : up) / d;
        }
        __syncthreads();
        for (int r = tid; r < nD; r += nt) colk[r] = a[r * nD + k];
        __syncthreads();
        for (int t = tid; t < nD * nD; t += nt) {
            const int r = t / nD, c = t - r * nD;
            if (r == k) continue;
            const double cur = (c == k) ? 0.0 : a[t];
            a[t] = cur - colk[r] * a[k * nD + c];
        }
        __syncthreads();
    }
    for (int k = nD - 1; k >= 0; --k) {  // undo the row swaps as column swaps, in reverse order
        const int p = piv[k];
        if (p != k)
            for (int r = tid; r < nD; r += nt) {
                const double u = a[r * nD + k];
                a[r * nD + k] = a[r * nD + p];
                a[r * nD + p] = u;
            }
        __syncthreads();
    }
    for (int t = tid; t < nD * nD; t += nt) Dinv[t] = a[t];
}
// x = A^-1 b: warp per row
template <typename VT>
__global__ void __launch_bounds__(1024) k_dense_apply(int nD, const double* __restrict__ Dinv, const VT* __restrict__ b,
                                                      VT* __restrict__ x) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int r = w; r < nD; r += nw) {
        double s = 0;
        for (int c = lane; c < nD; c += 32) s += Dinv[(size_t)r * nD + c] * (double)b[c];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) x[r] = (VT)s;
    }
}

// ---- cycle ------------------------------------------------------------------------------------------------------------
// first sweep from a zero guess: x = Dw b
template <int BS, typename VT>
__global__ void k_mg_jacobi0(int nDof, const VT* __restrict__ Dw, const VT* __restrict__ b, VT* __restrict__ x) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nDof) return;
    const int base = (i / BS) * BS;
    VT a = 0;
#pragma unroll
    for (int c = 0; c < BS; ++c) a += Dw[(size_t)i * BS + c] * b[base + c];
    x[i] = a;
}
template <int BS, typename VT>
__global__ void k_mg_restrict(int nc, const int* __restrict__ aggPtr, const int* __restrict__ aggNodes,
                              const VT* __restrict__ r, VT* __restrict__ bc) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nc * BS) return;
    const int I = t / BS, d = t % BS;
    VT s = 0;
    for (int m = aggPtr[I]; m < aggPtr[I + 1]; ++m) s += r[(size_t)aggNodes[m] * BS + d];
    bc[t] = s;
}
template <int BS, typename VT>
__global__ void k_mg_prolong(int nDof, const int* __restrict__ agg, const VT* __restrict__ xc, double over,
                             VT* __restrict__ x, bool set) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nDof) return;
    const VT e = (VT)over * xc[(size_t)agg[i / BS] * BS + (i % BS)];
    x[i] = set ? e : x[i] + e;
}

// deterministic start vector of the power iteration (all frequencies present) and ordered 2-norm, single CTA
__global__ void k_mg_noise(int nDof, double* __restrict__ x) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nDof) return;
    unsigned h = (unsigned)i * 2654435761u;
    h ^= h >> 15;
    h *= 2246822519u;
    h ^= h >> 13;
    x[i] = (double)(h & 0xffffu) / 32768.0 - 1.0;
}
__global__ void __launch_bounds__(1024) k_mg_norm2(int nDof, const double* __restrict__ x, double* __restrict__ out) {
    __shared__ double sh[32];
    double s = 0;
    for (int i = threadIdx.x; i < nDof; i += blockDim.x) s += x[i] * x[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int k = 0; k < 32; ++k) t += sh[k];
        *out = t;
    }
}

template <typename VT> __global__ void k_mg_add(int n, const VT* __restrict__ a, VT* __restrict__ x) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] += a[i];
}
__global__ void k_to_double(size_t n, const float* __restrict__ f, double* __restrict__ a) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) a[i] = (double)f[i];
}
__global__ void k_to_float(size_t n, const double* __restrict__ a, float* __restrict__ f) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) f[i] = (float)a[i];
}
bool mgFp32() {
    static const bool v = !(getenv("PFEM_MG_FP32") && atoi(getenv("PFEM_MG_FP32")) == 0);
    return v;
}

template <typename VT> inline const VT* levelDw(const MgLevel& L);
template <> inline const double* levelDw<double>(const MgLevel& L) { return L.Dw.p; }
template <> inline const float* levelDw<float>(const MgLevel& L) { return L.DwF.p; }
template <typename VT> inline VT* vecOf(DevBuf<double>& b) { return reinterpret_cast<VT*>(b.p); }

template <int EPI, typename VT>
void launchSpmv(pfem_ctx* c, const MgLevel& L, int BS, const VT* x, VT* y, const VT* b) {
    // partitioned mesh: the sweeps of a distributed level act on the GLOBAL matrix of that level (ghost entries of the
    // iterate follow their owners); replicated levels need no exchange
    if (L.distributed && c->nRanks > 1) commHaloPlanT<VT>(c, *L.plan, const_cast<VT*>(x), nullptr, BS);
    SpmvEpiT<VT> e;
    e.b = b;
    e.Dw = levelDw<VT>(L);
    const int grid = std::max(1, std::min(c->smCount * 8, divUp(L.n, 8)));
    std::unique_ptr<PhaseScope> ph;
    if (c->profileDetail && L.Aval == c->Aval.p) ph.reset(new PhaseScope(c, EPI == EPI_SMOOTH ? "MG smooth L0" : "MG resid L0"));
    if constexpr (sizeof(VT) == 4) {  // fp32 vectors: always with the fp32 matrix copy
        PFEM_REQUIRE(L.Af.p && L.DwF.p, PFEM_ERR_STATE, "multigrid: fp32 level data missing");
        if (BS == 4 && L.maxNb > 16)
            k_spmv<4, 3, EPI, float, 4, float><<<grid, 256, 0, c->stream>>>(L.n, L.nbrPtr, L.nbr, L.Af.p, x, y, nullptr, nullptr, 0, -1,
                                                                          -1, nullptr, nullptr, e);
        else if (BS == 4)
            k_spmv<4, 4, EPI, float, 2, float><<<grid, 256, 0, c->stream>>>(L.n, L.nbrPtr, L.nbr, L.Af.p, x, y, nullptr, nullptr, 0, -1,
                                                                          -1, nullptr, nullptr, e);
        else
            k_spmv<3, 3, EPI, float, 2, float><<<grid, 256, 0, c->stream>>>(L.n, L.nbrPtr, L.nbr, L.Af.p, x, y, nullptr, nullptr, 0, -1,
                                                                          -1, nullptr, nullptr, e);
        LAUNCH_CHECK(c);
        return;
    } else {
    static const bool minb4 = !(getenv("PFEM_MG_MINB4") && atoi(getenv("PFEM_MG_MINB4")) == 0);  // 64 registers, 32 warps/SM
    // cp.async ring variant: measured equal to the 64-register kernel (126 vs 118 us smoothing sweep at 2 M tets): the sweep is
    // bound by instruction issue and gather latency (ncu: 47 % issue active, 67 M instructions, DRAM 3.1 TB/s), not by the
    // bytes of A in flight.  Opt-in (PFEM_MG_RING=3|4|6).
    static const int ringDepth = getenv("PFEM_MG_RING") ? atoi(getenv("PFEM_MG_RING")) : 0;
    if (mgFp32() && L.Af.p && BS == 4 && L.maxNb <= 16 && ringDepth > 0) {
        // fine level of a tetrahedral mesh: cp.async ring, DEPTH-1 rows of A in flight per warp
        const int g = std::max(1, std::min(c->smCount * 4, divUp(L.n, 8)));
        auto launch = [&](auto kern, int depth) {
            const size_t smem = (size_t)8 * depth * 16 * 16 * sizeof(float);
            if (smem > 40 * 1024) CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PFEM_SMEM_OPTIN));
            kern<<<g, 256, smem, c->stream>>>(L.n, L.nbrPtr, L.nbr, L.Af.p, x, y, e);
        };
        if (ringDepth == 3) launch(k_spmv_ring<EPI, float, 3, 4>, 3);
        else if (ringDepth == 6) launch(k_spmv_ring<EPI, float, 6, 4>, 6);
        else launch(k_spmv_ring<EPI, float, 4, 4>, 4);
        LAUNCH_CHECK(c);
        return;
    }
    if (mgFp32() && L.Af.p) {
        if (BS == 4 && L.maxNb > 16)  // aggregated levels: ~27 blocks per row
            k_spmv<4, 2, EPI, float, 4><<<grid, 256, 0, c->stream>>>(L.n, L.nbrPtr, L.nbr, L.Af.p, x, y, nullptr, nullptr, 0, -1, -1,
                                                                   nullptr, nullptr, e);
        else if (BS == 4 && minb4)
            k_spmv<4, 4, EPI, float><<<std::max(1, std::min(c->smCount * 8, divUp(L.n, 8))), 256, 0, c->stream>>>(
                L.n, L.nbrPtr, L.nbr, L.Af.p, x, y, nullptr, nullptr, 0, -1, -1, nullptr, nullptr, e);
        else if (BS == 4)
            k_spmv<4, 3, EPI, float><<<grid, 256, 0, c->stream>>>(L.n, L.nbrPtr, L.nbr, L.Af.p, x, y, nullptr, nullptr, 0, -1, -1,
                                                                nullptr, nullptr, e);
        else
            k_spmv<3, 3, EPI, float><<<grid, 256, 0, c->stream>>>(L.n, L.nbrPtr, L.nbr, L.Af.p, x, y, nullptr, nullptr, 0, -1, -1,
                                                                nullptr, nullptr, e);
        LAUNCH_CHECK(c);
        return;
    }
    if (BS == 4)
        k_spmv<4, 3, EPI><<<grid, 256, 0, c->stream>>>(L.n, L.nbrPtr, L.nbr, L.Aval, x, y, nullptr, nullptr, 0, -1, -1, nullptr,
                                                     nullptr, e);
    else
        k_spmv<3, 3, EPI><<<grid, 256, 0, c->stream>>>(L.n, L.nbrPtr, L.nbr, L.Aval, x, y, nullptr, nullptr, 0, -1, -1, nullptr,
                                                     nullptr, e);
    LAUNCH_CHECK(c);
    }
}

template <typename VT>
void cycle(pfem_ctx* c, MgHierarchy& H, int l, const VT* b, VT* out) {
    const int BS = c->dim + 1;
    MgLevel& L = *H.lev[l];
    const int nDof = L.n * BS;
    const bool last = l + 1 == (int)H.lev.size();
    if (last && H.denseOk) {
        k_dense_apply<VT><<<1, 1024, 0, c->stream>>>(H.nD, H.dense.p, b, out);
        LAUNCH_CHECK(c);
        return;
    }
    VT* cur = vecOf<VT>(L.xa);
    VT* oth = vecOf<VT>(L.xb);
    auto jac0 = [&](VT* dst) {
        if (BS == 4) k_mg_jacobi0<4, VT><<<divUp(nDof, 256), 256, 0, c->stream>>>(nDof, levelDw<VT>(L), b, dst);
        else k_mg_jacobi0<3, VT><<<divUp(nDof, 256), 256, 0, c->stream>>>(nDof, levelDw<VT>(L), b, dst);
        LAUNCH_CHECK(c);
    };
    const int nu = l == 0 ? H.nu : H.nuCoarse;  // coarse sweeps are cheap: the fine level may run fewer than the rest
    int nPre = last ? 4 * nu : nu, nPost = last ? 0 : nu;
    if (!last) {  // asymmetric cycles: without pre-smoothing the residual is b itself (no matrix pass before the restriction)
        const int pre = l == 0 ? H.preFine : H.preCoarse, post = l == 0 ? H.postFine : H.postCoarse;
        if (pre >= 0) nPre = pre;
        if (post >= 0) nPost = post;
        if (nPre + nPost == 0) nPost = 1;
    }
    if (nPre > 0) jac0(cur);
    for (int k = 1; k < nPre; ++k) {
        VT* dst = (last && k == nPre - 1) ? out : oth;
        launchSpmv<EPI_SMOOTH, VT>(c, L, BS, cur, dst, b);
        std::swap(cur, oth);
        if (dst == out) cur = out;
    }
    if (last) {  // coarsest level without a dense inverse: smoothing only
        if (cur != out) CUDA_CHECK(cudaMemcpyAsync(out, cur, (size_t)nDof * sizeof(VT), cudaMemcpyDeviceToDevice, c->stream));
        return;
    }
    MgLevel& C = *H.lev[l + 1];
    const VT* resid = b;
    if (nPre > 0) {
        launchSpmv<EPI_RESID, VT>(c, L, BS, cur, vecOf<VT>(L.t), b);
        resid = vecOf<VT>(L.t);
    }
    {
        // my aggregates: all coarse rows, or rows [rowOff, rowOff + ncLocal) of a replicated next level (then all-gathered)
        const int ncL = L.ncLocal;
        VT* Cb = vecOf<VT>(C.b);
        VT* dstB = Cb + (size_t)L.rowOff * BS;
        if (ncL > 0) {
            if (BS == 4) k_mg_restrict<4, VT><<<divUp(ncL * BS, 256), 256, 0, c->stream>>>(ncL, L.aggPtr.p, L.aggNodes.p, resid, dstB);
            else k_mg_restrict<3, VT><<<divUp(ncL * BS, 256), 256, 0, c->stream>>>(ncL, L.aggPtr.p, L.aggNodes.p, resid, dstB);
            LAUNCH_CHECK(c);
        }
        if (L.nextReplicated) {
            std::vector<int64_t> cnt(L.rowCounts), dsp(L.rowDispls);
            for (auto& v : cnt) v *= BS * (int64_t)sizeof(VT);
            for (auto& v : dsp) v *= BS * (int64_t)sizeof(VT);
            commAllGatherBytes(c, dstB, Cb, cnt, dsp);
        }
    }
    cycle<VT>(c, H, l + 1, vecOf<VT>(C.b), vecOf<VT>(C.xo));
    // W-cycle: a second visit of the coarse level on the residual of the first (unsmoothed aggregation loses its mesh
    // independence with V-cycles; the coarse levels cost a few percent of the cycle).  Not when that level is solved exactly.
    const bool coarseExact = (l + 2 == (int)H.lev.size()) && H.denseOk;
    if (H.wFrom >= 0 && l >= H.wFrom && !coarseExact) {
        launchSpmv<EPI_RESID, VT>(c, C, BS, vecOf<VT>(C.xo), vecOf<VT>(C.w), vecOf<VT>(C.b));
        cycle<VT>(c, H, l + 1, vecOf<VT>(C.w), vecOf<VT>(C.xo2));
        const int nc = C.n * BS;
        k_mg_add<VT><<<divUp(nc, 256), 256, 0, c->stream>>>(nc, vecOf<VT>(C.xo2), vecOf<VT>(C.xo));
        LAUNCH_CHECK(c);
    }
    VT* corrected = nPost == 0 ? out : cur;  // no post-smoothing: the corrected iterate is the result
    if (nPost == 0 && nPre > 0) CUDA_CHECK(cudaMemcpyAsync(out, cur, (size_t)nDof * sizeof(VT), cudaMemcpyDeviceToDevice, c->stream));
    if (BS == 4) k_mg_prolong<4, VT><<<divUp(nDof, 256), 256, 0, c->stream>>>(nDof, L.agg.p, vecOf<VT>(C.xo), H.over, corrected, nPre == 0);
    else k_mg_prolong<3, VT><<<divUp(nDof, 256), 256, 0, c->stream>>>(nDof, L.agg.p, vecOf<VT>(C.xo), H.over, corrected, nPre == 0);
    LAUNCH_CHECK(c);
    for (int k = 1; k <= nPost; ++k) {
        VT* dst = (k == nPost) ? out : oth;
        launchSpmv<EPI_SMOOTH, VT>(c, L, BS, cur, dst, b);
        std::swap(cur, oth);
    }
}

// one cycle from the fp64 right-hand side in lev[0]->b.  fp64 vectors: the result is written to `out`.  fp32 vectors: the
// result stays in H.outF and mgApply converts it into the caller's vector OUTSIDE the captured graph, so that one graph serves
// every output vector (FGMRES hands a different z_j per iteration).
void runCycle(pfem_ctx* c, MgHierarchy& H, double* out) {
    MgLevel& L0 = *H.lev[0];
    if (!H.f32v) {
        cycle<double>(c, H, 0, L0.b.p, out);
        return;
    }
    const size_t nDof = (size_t)L0.n * (c->dim + 1);
    const int grid = std::max(1, std::min(c->smCount * 8, divUp((int64_t)nDof, 256)));
    if (!H.rhsIsF32) {
        k_to_float<<<grid, 256, 0, c->stream>>>(nDof, L0.b.p, H.inF.p);
        LAUNCH_CHECK(c);
    }
    cycle<float>(c, H, 0, H.inF.p, H.outF.p);
}
void convertResult(pfem_ctx* c, MgHierarchy& H, double* out) {
    if (!H.f32v) return;
    const size_t nDof = (size_t)H.lev[0]->n * (c->dim + 1);
    const int grid = std::max(1, std::min(c->smCount * 8, divUp((int64_t)nDof, 256)));
    k_to_double<<<grid, 256, 0, c->stream>>>(nDof, H.outF.p, out);
    LAUNCH_CHECK(c);
}

// total size below which the next level is replicated on every rank (read at every symbolic build: tests change it)
int mgReplicateNodes() { return getenv("PFEM_MG_REPL_NODES") ? atoi(getenv("PFEM_MG_REPL_NODES")) : 50000; }

// Rank-local part of a coarsening step: aggregates of the OWNED nodes of L (grid cells over their coordinates), member
// lists in ascending order, coarse coordinates written to Xc (4 doubles per aggregate).  Returns false when this rank's
// share cannot be coarsened; nc = local aggregate count.
bool localAggregate(pfem_ctx* c, MgHierarchy& H, MgLevel& L, double& cellSize, int& ncOut, DevBuf<double>& XcLocal) {
    DevBuf<double>& box = H.sBox;
    DevBuf<int>& key = H.sKey;
    DevBuf<int>& cell = H.sCell;
    const int dim = c->dim, n = L.n;
    ncOut = 0;
    if (n < 1) return false;
    const double target = dim == 3 ? 8.0 : 4.0;
    c->scratchI.reserve((size_t)std::max(L.nVec, c->nNodes) + 64);
    H.flag.reserve(16);
    box.reserve(8);
    {
        const int nb = std::max(1, std::min(64, n / 4096));
        box.reserve(8 + (size_t)nb * 6);
        if (nb == 1)
            k_bbox<<<1, 1024, 0, c->stream>>>(L.X, n, dim, box.p);
        else {
            k_bbox<<<nb, 1024, 0, c->stream>>>(L.X, n, dim, box.p + 8);
            LAUNCH_CHECK(c);
            k_bbox_final<<<1, 32, 0, c->stream>>>(box.p + 8, nb, box.p);
        }
    }
    LAUNCH_CHECK(c);
    double hb[6];
    CUDA_CHECK(cudaMemcpyAsync(hb, box.p, 6 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    double ext[3], vol = 1.0, maxExt = 0.0;
    for (int d = 0; d < dim; ++d) {
        ext[d] = std::max(hb[3 + d] - hb[d], 0.0);
        maxExt = std::max(maxExt, ext[d]);
    }
    if (!(maxExt > 0.0)) return false;
    for (int d = 0; d < dim; ++d) vol *= std::max(ext[d], 1e-3 * maxExt);
    // cell edge: the caller's suggestion (twice the node spacing of this level), else twice the spacing of a filled box
    double Hc = cellSize > 0.0 ? cellSize : 2.0 * std::pow(vol / n, 1.0 / dim);
    const double slack = cellSize > 0.0 ? 3.0 : 1.7;  // a suggested size is only overridden when it is far off
    L.agg.reserve((size_t)L.nVec + 4);
    key.reserve(n);
    int nc = 0;
    for (int attempt = 0; attempt < 6; ++attempt) {
        int nx, ny, nz;
        for (;;) {
            nx = (int)std::floor(ext[0] / Hc) + 1;
            ny = (int)std::floor(ext[1] / Hc) + 1;
            nz = dim == 3 ? (int)std::floor(ext[2] / Hc) + 1 : 1;
            if ((double)nx * ny * nz <= 64.0e6) break;
            Hc *= 1.5;
        }
        const int ncell = nx * ny * nz;
        cell.reserve((size_t)ncell + 2);
        CUDA_CHECK(cudaMemsetAsync(cell.p, 0, ((size_t)ncell + 2) * sizeof(int), c->stream));
        k_cell_key<<<divUp(n, 256), 256, 0, c->stream>>>(L.X, n, dim, hb[0], hb[1], hb[2], 1.0 / Hc, nx, ny, nz, key.p, cell.p);
        LAUNCH_CHECK(c);
        exclusiveScanInt(c, cell.p, ncell + 1, H.flag.p + 4);
        CUDA_CHECK(cudaMemcpyAsync(&nc, H.flag.p + 4, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        const double ratio = (double)n / std::max(nc, 1);
        if (ratio > slack * target || ratio < target / (slack + 0.8)) {
            if (attempt < 5) {
                Hc *= std::pow(target / ratio, 1.0 / dim);
                continue;
            }
        }
        break;
    }
    if (nc < 1 || nc > 0.8 * n) return false;
    cellSize = Hc;
    L.aggPtr.reserve((size_t)nc + 2);
    L.aggNodes.reserve(n);
    CUDA_CHECK(cudaMemsetAsync(L.aggPtr.p, 0, ((size_t)nc + 2) * sizeof(int), c->stream));
    k_assign_agg<<<divUp(n, 256), 256, 0, c->stream>>>(n, key.p, cell.p, L.agg.p, L.aggPtr.p);
    LAUNCH_CHECK(c);
    exclusiveScanInt(c, L.aggPtr.p, nc + 1, nullptr);
    CUDA_CHECK(cudaMemsetAsync(c->scratchI.p, 0, (size_t)nc * sizeof(int), c->stream));
    k_fill_members<<<divUp(n, 256), 256, 0, c->stream>>>(n, L.agg.p, L.aggPtr.p, c->scratchI.p, L.aggNodes.p);
    LAUNCH_CHECK(c);
    XcLocal.reserve((size_t)nc * 4 + 4);
    k_sort_members<<<divUp(nc, 128), 128, 0, c->stream>>>(nc, dim, L.aggPtr.p, L.aggNodes.p, L.X, XcLocal.p);
    LAUNCH_CHECK(c);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));  // key/cell/box are released on return
    ncOut = nc;
    return true;
}

// one coarsening step; returns false when the level cannot be coarsened further.  On a partitioned mesh every decision
// (coarsen or stop, keep the next level distributed or replicate it) is taken from all-gathered numbers, so that all ranks
// build the same number of levels of the same kind and later issue the same sequence of collectives.
bool coarsen(pfem_ctx* c, MgHierarchy& H, MgLevel& L, MgLevel& C, double& cellSize) {
    const int dim = c->dim, n = L.n, BS = dim + 1;
    const bool multi = L.distributed && c->nRanks > 1;
    int nc = 0;
    DevBuf<double>& XcLocal = H.sXc;
    bool ok = localAggregate(c, H, L, cellSize, nc, XcLocal);
    const int R = c->nRanks, me = c->rank;
    std::vector<double> all(2 * (size_t)R, 0.0);
    bool replicate = false;
    int64_t ncGlobal = nc;
    if (multi) {
        const double mine[2] = {ok ? 1.0 : 0.0, (double)nc};
        commAllGatherHost(c, mine, 2, all.data());
        ncGlobal = 0;
        for (int r = 0; r < R; ++r) {
            ok = ok && all[2 * r] != 0.0;
            ncGlobal += (int64_t)all[2 * r + 1];
        }
        replicate = ncGlobal <= mgReplicateNodes();
    }
    if (!ok) return false;
    L.nextReplicated = false;
    L.ncLocal = nc;
    L.rowOff = 0;
    H.flag.reserve(16);
    int* misc = H.flag.p;
    DevBuf<double>& tmp = H.sTmp;  // aggregate ids of the fine vector entries as doubles (halo exchange payload)

    if (!multi) {
        // ---- single rank, or a replicated level: the whole level is here ------------------------------------------------
        L.nc = nc;
        C.n = C.nVec = nc;
        C.distributed = false;
        C.plan = nullptr;
        C.XB.reserve((size_t)nc * 4 + 4);
        CUDA_CHECK(cudaMemcpyAsync(C.XB.p, XcLocal.p, (size_t)nc * 4 * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        C.nbrPtrB.reserve((size_t)nc + 2);
        C.diagSlotB.reserve(nc);
        CUDA_CHECK(cudaMemsetAsync(C.nbrPtrB.p, 0, ((size_t)nc + 2) * sizeof(int), c->stream));
        CUDA_CHECK(cudaMemsetAsync(misc, 0, 4 * sizeof(int), c->stream));
        k_coarse_nbr<false><<<divUp(nc, 8), 256, 0, c->stream>>>(nc, L.aggPtr.p, L.aggNodes.p, L.agg.p, n, L.nbrPtr, L.nbr, C.nbrPtrB.p,
                                                                nullptr, nullptr, misc, 0);
        LAUNCH_CHECK(c);
        exclusiveScanInt(c, C.nbrPtrB.p, nc + 1, misc + 2);
        int hf[4];
        CUDA_CHECK(cudaMemcpyAsync(hf, misc, 4 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        if (hf[1] != 0) return false;  // a coarse row with more than NBR_CAP neighbours
        C.maxNb = hf[0];
        C.nBlocks = hf[2];
        C.nbrB.reserve((size_t)C.nBlocks + 4);
        k_coarse_nbr<true><<<divUp(nc, 8), 256, 0, c->stream>>>(nc, L.aggPtr.p, L.aggNodes.p, L.agg.p, n, L.nbrPtr, L.nbr, C.nbrPtrB.p,
                                                               C.nbrB.p, C.diagSlotB.p, misc, 0);
        LAUNCH_CHECK(c);
        L.cslot.reserve((size_t)L.nBlocks + 4);
        k_cslot<<<divUp(n, 128), 128, 0, c->stream>>>(n, L.nbrPtr, L.nbr, L.agg.p, C.nbrPtrB.p, C.nbrB.p, L.cslot.p);
        LAUNCH_CHECK(c);
    } else if (replicate) {
        // ---- distributed level -> replicated next level ---------------------------------------------------------------------
        L.rowCounts.assign(R, 0), L.rowDispls.assign(R, 0);
        for (int r = 0; r < R; ++r) L.rowCounts[r] = (int64_t)all[2 * r + 1];
        for (int r = 1; r < R; ++r) L.rowDispls[r] = L.rowDispls[r - 1] + L.rowCounts[r - 1];
        const int ncG = (int)ncGlobal, off = (int)L.rowDispls[me];
        L.nextReplicated = true;
        L.rowOff = off;
        L.nc = ncG;
        // global aggregate ids: owned entries shifted by this rank's row offset, ghost entries from their owners
        tmp.reserve((size_t)L.nVec + 4);
        k_agg_shift_to_double<<<divUp(n, 256), 256, 0, c->stream>>>(n, L.agg.p, off, tmp.p);
        LAUNCH_CHECK(c);
        commHaloPlan(c, *L.plan, tmp.p, nullptr, 1);
        if (L.nVec > n) {
            k_double_to_int<<<divUp(L.nVec - n, 256), 256, 0, c->stream>>>(L.nVec - n, tmp.p + n, L.agg.p + n);
            LAUNCH_CHECK(c);
        }
        C.n = C.nVec = ncG;
        C.distributed = false;
        C.plan = nullptr;
        auto bytes = [&](const std::vector<int64_t>& v, int64_t unit) {
            std::vector<int64_t> o(v);
            for (auto& x : o) x *= unit;
            return o;
        };
        // coarse coordinates
        C.XB.reserve((size_t)ncG * 4 + 4);
        commAllGatherV(c, XcLocal.p, C.XB.p, bytes(L.rowCounts, 4), bytes(L.rowDispls, 4));
        // row lengths of my rows -> all ranks -> global row pointers
        C.nbrPtrB.reserve((size_t)ncG + 2);
        C.diagSlotB.reserve((size_t)ncG + 2);
        CUDA_CHECK(cudaMemsetAsync(C.nbrPtrB.p, 0, ((size_t)ncG + 2) * sizeof(int), c->stream));
        CUDA_CHECK(cudaMemsetAsync(misc, 0, 4 * sizeof(int), c->stream));
        k_coarse_nbr<false><<<divUp(nc, 8), 256, 0, c->stream>>>(nc, L.aggPtr.p, L.aggNodes.p, L.agg.p, L.nVec, L.nbrPtr, L.nbr,
                                                                C.nbrPtrB.p + off, nullptr, nullptr, misc, off);
        LAUNCH_CHECK(c);
        commAllGatherBytes(c, C.nbrPtrB.p + off, C.nbrPtrB.p, bytes(L.rowCounts, sizeof(int)), bytes(L.rowDispls, sizeof(int)));
        exclusiveScanInt(c, C.nbrPtrB.p, ncG + 1, misc + 2);
        int hf[4];
        CUDA_CHECK(cudaMemcpyAsync(hf, misc, 4 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        std::vector<int> ptrHost((size_t)ncG + 1);
        CUDA_CHECK(cudaMemcpyAsync(ptrHost.data(), C.nbrPtrB.p, ((size_t)ncG + 1) * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        const double mine2[2] = {(double)hf[0], (double)hf[1]};
        std::vector<double> all2(2 * (size_t)R);
        commAllGatherHost(c, mine2, 2, all2.data());
        int maxNb = 0;
        bool overflow = false;
        for (int r = 0; r < R; ++r) {
            maxNb = std::max(maxNb, (int)all2[2 * r]);
            overflow = overflow || all2[2 * r + 1] != 0.0;
        }
        if (overflow) return false;
        C.maxNb = maxNb;
        C.nBlocks = hf[2];
        L.blkCounts.assign(R, 0), L.blkDispls.assign(R, 0);
        for (int r = 0; r < R; ++r) {
            L.blkDispls[r] = ptrHost[L.rowDispls[r]];
            L.blkCounts[r] = ptrHost[L.rowDispls[r] + L.rowCounts[r]] - ptrHost[L.rowDispls[r]];
        }
        C.nbrB.reserve((size_t)C.nBlocks + 4);
        k_coarse_nbr<true><<<divUp(nc, 8), 256, 0, c->stream>>>(nc, L.aggPtr.p, L.aggNodes.p, L.agg.p, L.nVec, L.nbrPtr, L.nbr,
                                                               C.nbrPtrB.p + off, C.nbrB.p, C.diagSlotB.p + off, misc, off);
        LAUNCH_CHECK(c);
        commAllGatherBytes(c, C.nbrB.p + L.blkDispls[me], C.nbrB.p, bytes(L.blkCounts, sizeof(int)), bytes(L.blkDispls, sizeof(int)));
        commAllGatherBytes(c, C.diagSlotB.p + off, C.diagSlotB.p, bytes(L.rowCounts, sizeof(int)), bytes(L.rowDispls, sizeof(int)));
        L.cslot.reserve((size_t)L.nBlocks + 4);
        k_cslot<<<divUp(n, 128), 128, 0, c->stream>>>(n, L.nbrPtr, L.nbr, L.agg.p, C.nbrPtrB.p, C.nbrB.p, L.cslot.p, L.nVec);
        LAUNCH_CHECK(c);
    } else {
        // ---- distributed level -> distributed next level: ghost aggregates + the next level's halo plan -----------------------
        L.nc = nc;
        tmp.reserve((size_t)L.nVec + 4);
        k_agg_shift_to_double<<<divUp(n, 256), 256, 0, c->stream>>>(n, L.agg.p, 0, tmp.p);
        LAUNCH_CHECK(c);
        commHaloPlan(c, *L.plan, tmp.p, nullptr, 1);
        std::vector<double> aggAll((size_t)L.nVec);
        CUDA_CHECK(cudaMemcpyAsync(aggAll.data(), tmp.p, (size_t)L.nVec * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        if (!C.ownPlan) {
            C.ownPlan.reset(new HaloPlan());
            C.ownPlan->sendIdx.accounting = C.ownPlan->sendBuf.accounting = &c->deviceBytes;
        }
        HaloPlan& CP = *C.ownPlan;
        CP.clear();
        std::vector<int> ghostAgg((size_t)(L.nVec - n), 0);
        int nGhostC = 0;
        for (const auto& p : L.plan->peers) {
            // what I receive: the distinct aggregates (owner-local ids) of the fine ghosts of this peer, ascending
            std::vector<int> u;
            u.reserve(p.recvCount);
            for (int k = 0; k < p.recvCount; ++k) u.push_back((int)aggAll[(size_t)p.recvStart + k]);
            std::sort(u.begin(), u.end());
            u.erase(std::unique(u.begin(), u.end()), u.end());
            for (int k = 0; k < p.recvCount; ++k) {
                const int a = (int)aggAll[(size_t)p.recvStart + k];
                ghostAgg[(size_t)p.recvStart - n + k] = nc + nGhostC + (int)(std::lower_bound(u.begin(), u.end(), a) - u.begin());
            }
            // what I send: the distinct aggregates of the fine nodes this peer holds as ghosts, ascending -- the same list
            // the peer derives from the ids it received, so both sides agree on the order without another exchange
            std::vector<int> sset;
            sset.reserve(p.sendCount);
            for (int k = 0; k < p.sendCount; ++k) sset.push_back((int)aggAll[(size_t)L.plan->sendIdxHost[(size_t)p.sendOff + k]]);
            std::sort(sset.begin(), sset.end());
            sset.erase(std::unique(sset.begin(), sset.end()), sset.end());
            CP.peers.push_back({p.rank, (int)CP.sendIdxHost.size(), (int)sset.size(), nc + nGhostC, (int)u.size()});
            CP.sendIdxHost.insert(CP.sendIdxHost.end(), sset.begin(), sset.end());
            nGhostC += (int)u.size();
        }
        commFinishPlan(c, CP);
        if (L.nVec > n)
            CUDA_CHECK(cudaMemcpyAsync(L.agg.p + n, ghostAgg.data(), (size_t)(L.nVec - n) * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        C.n = nc;
        C.nVec = nc + nGhostC;
        C.distributed = true;
        C.plan = C.ownPlan.get();
        C.XB.reserve((size_t)C.nVec * 4 + 4);
        CUDA_CHECK(cudaMemsetAsync(C.XB.p, 0, ((size_t)C.nVec * 4) * sizeof(double), c->stream));
        CUDA_CHECK(cudaMemcpyAsync(C.XB.p, XcLocal.p, (size_t)nc * 4 * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        C.nbrPtrB.reserve((size_t)nc + 2);
        C.diagSlotB.reserve(nc);
        CUDA_CHECK(cudaMemsetAsync(C.nbrPtrB.p, 0, ((size_t)nc + 2) * sizeof(int), c->stream));
        CUDA_CHECK(cudaMemsetAsync(misc, 0, 4 * sizeof(int), c->stream));
        k_coarse_nbr<false><<<divUp(nc, 8), 256, 0, c->stream>>>(nc, L.aggPtr.p, L.aggNodes.p, L.agg.p, L.nVec, L.nbrPtr, L.nbr,
                                                                C.nbrPtrB.p, nullptr, nullptr, misc, 0);
        LAUNCH_CHECK(c);
        exclusiveScanInt(c, C.nbrPtrB.p, nc + 1, misc + 2);
        int hf[4];
        CUDA_CHECK(cudaMemcpyAsync(hf, misc, 4 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        const double mine2[2] = {(double)hf[0], (double)hf[1]};
        std::vector<double> all2(2 * (size_t)R);
        commAllGatherHost(c, mine2, 2, all2.data());
        bool overflow = false;
        for (int r = 0; r < R; ++r) overflow = overflow || all2[2 * r + 1] != 0.0;
        if (overflow) return false;
        C.maxNb = hf[0];
        C.nBlocks = hf[2];
        C.nbrB.reserve((size_t)C.nBlocks + 4);
        k_coarse_nbr<true><<<divUp(nc, 8), 256, 0, c->stream>>>(nc, L.aggPtr.p, L.aggNodes.p, L.agg.p, L.nVec, L.nbrPtr, L.nbr,
                                                               C.nbrPtrB.p, C.nbrB.p, C.diagSlotB.p, misc, 0);
        LAUNCH_CHECK(c);
        L.cslot.reserve((size_t)L.nBlocks + 4);
        k_cslot<<<divUp(n, 128), 128, 0, c->stream>>>(n, L.nbrPtr, L.nbr, L.agg.p, C.nbrPtrB.p, C.nbrB.p, L.cslot.p, L.nVec);
        LAUNCH_CHECK(c);
    }
    C.nbrPtr = C.nbrPtrB.p, C.nbr = C.nbrB.p, C.diagSlot = C.diagSlotB.p, C.X = C.XB.p;
    C.AvalB.reserve((size_t)C.nBlocks * BS * BS + 8);
    C.Aval = C.AvalB.p;
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return true;
}

void allocVectors(pfem_ctx* c, MgLevel& L, int BS) {
    const size_t nv = (size_t)L.nVec * BS + 8;
    for (auto* v : {&L.b, &L.xa, &L.xb, &L.t, &L.xo, &L.w, &L.xo2}) {
        const bool fresh = nv > v->cap;
        v->reserve(nv);
        if (fresh) CUDA_CHECK(cudaMemsetAsync(v->p, 0, v->cap * sizeof(double), c->stream));  // ghost entries stay zero
    }
    L.Dw.reserve((size_t)L.n * BS * BS + 8);
}

void buildSymbolic(pfem_ctx* c, MgHierarchy& H) {
    PhaseScope ph(c, "Preconditioner pattern");
    const int BS = c->dim + 1;
    H.dropGraphs();
    H.lev.clear();
    H.symbolicFailed = false;
    MgLevel* L0 = &H.levelObject(0);
    L0->resetLogical();
    L0->n = c->nRows, L0->nVec = c->nNodes, L0->maxNb = c->maxNb, L0->nBlocks = c->nBlocks;
    L0->distributed = c->nRanks > 1;
    L0->plan = &c->plan;
    H.lev.push_back(L0);
    // node spacing of level 0 from the mean element size: h0 = (mean |detJ|)^(1/dim)
    double cellSize = 0.0;
    if (c->nElems > 0) {
        const int nb = 256;
        DevBuf<double>& part = H.sPart;
        part.reserve(nb + 8);
        k_vol_partial<<<nb, 256, 0, c->stream>>>(c->conn.p, c->nElems, c->dim, c->X4.p, part.p);
        LAUNCH_CHECK(c);
        k_vol_final<<<1, 32, 0, c->stream>>>(part.p, nb, part.p + nb);
        LAUNCH_CHECK(c);
        double tot = 0;
        CUDA_CHECK(cudaMemcpyAsync(&tot, part.p + nb, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        if (tot > 0.0) cellSize = 2.0 * std::pow(tot / c->nElems, 1.0 / c->dim);
    }
    for (int l = 0; l < 16; ++l) {
        MgLevel& L = *H.lev[l];
        if (l == 0) {  // aliases may have been reallocated since the last build
            L.nbrPtr = c->nbrPtr.p, L.nbr = c->nbr.p, L.diagSlot = c->diagSlot.p, L.Aval = c->Aval.p, L.X = c->X4.p;
        }
        allocVectors(c, L, BS);
        if (!L.distributed && L.n <= COARSEST_NODES) break;  // (a distributed level's size is a per-rank number: coarsen() decides)
        MgLevel* C = &H.levelObject((size_t)l + 1);
        C->resetLogical();
        if (!coarsen(c, H, L, *C, cellSize)) break;
        cellSize *= 2.0;
        H.lev.push_back(C);
    }
    H.symbolicValid = true;
    H.numericValid = false;
}

// Dw = damping * A_ii^-1: uniform damping L.omega, or (l1 = true) theta / r_i per node
void blockInverse(pfem_ctx* c, MgLevel& L, int BS, bool l1 = false, double theta = 1.0, float* Af = nullptr) {
    const double w = l1 ? 1.0 : L.omega;
    if (BS == 4) k_mg_block_inv<4><<<divUp(L.n, 128), 128, 0, c->stream>>>(L.n, L.nbrPtr, L.diagSlot, L.Aval, w, L.Dw.p);
    else k_mg_block_inv<3><<<divUp(L.n, 128), 128, 0, c->stream>>>(L.n, L.nbrPtr, L.diagSlot, L.Aval, w, L.Dw.p);
    LAUNCH_CHECK(c);
    if (!l1) return;
    static const double wcap = getenv("PFEM_MG_WCAP") ? atof(getenv("PFEM_MG_WCAP")) : 1.0;
    // distributed level: the scales of the ghost columns come from their owners (halo exchange below)
    if (L.sc.cap < (size_t)L.nVec * BS + 8) {
        L.sc.reserve((size_t)L.nVec * BS + 8);
        std::vector<double> ones(L.sc.cap, 1.0);
        CUDA_CHECK(cudaMemcpyAsync(L.sc.p, ones.data(), L.sc.cap * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
    }
    if (BS == 4) {
        k_mg_diag_scale<4><<<divUp(L.n * BS, 256), 256, 0, c->stream>>>(L.n, L.nbrPtr, L.diagSlot, L.Aval, L.sc.p);
        LAUNCH_CHECK(c);
        if (L.distributed && c->nRanks > 1) commHaloPlan(c, *L.plan, L.sc.p, nullptr, BS);  // scales of the ghost columns
        k_mg_l1_scale<4><<<divUp((int64_t)L.n * 16, 256), 256, 0, c->stream>>>(L.n, L.nbrPtr, L.nbr, L.Aval, L.sc.p, theta, wcap, L.Dw.p, nullptr, Af);
    } else {
        k_mg_diag_scale<3><<<divUp(L.n * BS, 256), 256, 0, c->stream>>>(L.n, L.nbrPtr, L.diagSlot, L.Aval, L.sc.p);
        LAUNCH_CHECK(c);
        if (L.distributed && c->nRanks > 1) commHaloPlan(c, *L.plan, L.sc.p, nullptr, BS);
        k_mg_l1_scale<3><<<divUp((int64_t)L.n * 16, 256), 256, 0, c->stream>>>(L.n, L.nbrPtr, L.nbr, L.Aval, L.sc.p, theta, wcap, L.Dw.p, nullptr, Af);
    }
    LAUNCH_CHECK(c);
}

// Damped node-block Jacobi is only a smoother if it does not amplify anything: the (v,p) coupling gives D^-1 A complex
// eigenvalues, so the admissible damping depends on dt, the material and the mesh.  Largest damping of a short ladder
// whose error propagation I - omega D^-1 A has a spectral radius (power iteration, deterministic start) below 0.97.
void tuneDamping(pfem_ctx* c, MgHierarchy& H, MgLevel& L, int BS) {
    static const double ladder[] = {0.7, 0.55, 0.45, 0.36, 0.28, 0.2, 0.12};
    const int nDof = L.n * BS, nIt = 14, nTail = 6;
    H.flag.reserve(16);
    double* dNorm = reinterpret_cast<double*>(H.flag.p + 10);  // 8-byte aligned: flag.p is cudaMalloc'ed, offset 40 B
    CUDA_CHECK(cudaMemsetAsync(L.t.p, 0, (size_t)nDof * sizeof(double), c->stream));
    for (double w : ladder) {
        L.omega = w;
        blockInverse(c, L, BS);
        k_mg_noise<<<divUp(nDof, 256), 256, 0, c->stream>>>(nDof, L.xa.p);
        LAUNCH_CHECK(c);
        double* cur = L.xa.p;
        double* oth = L.xb.p;
        double nrm[32];
        for (int it = 0; it <= nIt; ++it) {
            if (it > 0) {
                launchSpmv<EPI_SMOOTH, double>(c, L, BS, cur, oth, L.t.p);
                std::swap(cur, oth);
            }
            k_mg_norm2<<<1, 1024, 0, c->stream>>>(nDof, cur, dNorm);
            LAUNCH_CHECK(c);
            CUDA_CHECK(cudaMemcpyAsync(&nrm[it], dNorm, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        }
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        const double a = nrm[nIt - nTail], b = nrm[nIt];
        const double rho = (a > 0.0 && b == b) ? std::pow(b / a, 0.5 / nTail) : 0.0;  // norms are squared
        if (b == b && rho <= 0.97) break;
    }
    // the work vectors go back to their resting state (ghost entries were never touched)
    CUDA_CHECK(cudaMemsetAsync(L.xa.p, 0, (size_t)nDof * sizeof(double), c->stream));
    CUDA_CHECK(cudaMemsetAsync(L.xb.p, 0, (size_t)nDof * sizeof(double), c->stream));
}

void buildNumeric(pfem_ctx* c, MgHierarchy& H) {
    PhaseScope ph(c, "Preconditioner setup");
    const int BS = c->dim + 1, BB = BS * BS;
    MgLevel& L0 = *H.lev[0];
    L0.nbrPtr = c->nbrPtr.p, L0.nbr = c->nbr.p, L0.diagSlot = c->diagSlot.p, L0.Aval = c->Aval.p, L0.X = c->X4.p;
    for (size_t l = 0; l < H.lev.size(); ++l) {
        MgLevel& L = *H.lev[l];
        static const int dampMode = getenv("PFEM_MG_DAMP") ? atoi(getenv("PFEM_MG_DAMP")) : 0;  // 0 l1 | 1 uniform | 2 tuned
        std::unique_ptr<PhaseScope> sub(new PhaseScope(c, l == 0 ? "MG smoother setup L0" : "MG smoother setup coarse"));
        const bool l1Path = dampMode == 0 || c->nRanks > 1;  // (the tuned ladder takes per-rank decisions: not on a partitioned mesh)
        if (mgFp32()) {
            const size_t nv = (size_t)L.nBlocks * BB;
            L.Af.reserve(nv + 8);
            if (!l1Path) {  // the l1 pass reads every entry of A anyway and writes the fp32 copy itself
                k_to_float<<<std::max(1, std::min(c->smCount * 8, divUp((int64_t)nv, 1024))), 256, 0, c->stream>>>(nv, L.Aval, L.Af.p);
                LAUNCH_CHECK(c);
            }
        }
        if (l1Path) {
            L.omega = H.fixedOmega > 0.0 ? H.fixedOmega : 2.0;  // theta: 2.0 measured best-robust (2.5 turns unstable in 2-D)
            blockInverse(c, L, BS, true, L.omega, mgFp32() ? L.Af.p : nullptr);
        } else if (H.fixedOmega > 0.0 || dampMode == 1) {
            L.omega = H.fixedOmega > 0.0 ? H.fixedOmega : 0.45;
            blockInverse(c, L, BS);
        } else if (!H.tuned)
            tuneDamping(c, H, L, BS);  // leaves Dw for the accepted damping
        else
            blockInverse(c, L, BS);
        if (mgFp32()) {  // fp32 copy of the damped block inverses for the fp32-vector cycle
            const size_t nd = (size_t)L.n * BB;
            L.DwF.reserve(nd + 8);
            k_to_float<<<std::max(1, std::min(c->smCount * 8, divUp((int64_t)nd, 1024))), 256, 0, c->stream>>>(nd, L.Dw.p, L.DwF.p);
            LAUNCH_CHECK(c);
        }
        sub.reset();
        if (l + 1 == H.lev.size()) break;
        PhaseScope sub2(c, l == 0 ? "MG Galerkin L0" : "MG Galerkin coarse");
        MgLevel& C = *H.lev[l + 1];
        const int nbcap = std::max(C.maxNb, 1);
        const int hwPerBlock = std::max(1, std::min(8, (int)((96 * 1024) / ((size_t)nbcap * BB * sizeof(double)))));
        const size_t smem = (size_t)hwPerBlock * nbcap * BB * sizeof(double);
        const int ncL = L.ncLocal;  // my aggregates = the coarse rows I sum (all of them unless the next level is replicated)
        const int* cPtrMine = C.nbrPtr + L.rowOff;
        if (ncL > 0) {
            if (BS == 4) {
                if (smem > 40 * 1024) CUDA_CHECK(cudaFuncSetAttribute(k_galerkin<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, PFEM_SMEM_OPTIN));
                k_galerkin<4><<<divUp(ncL, hwPerBlock), hwPerBlock * 16, smem, c->stream>>>(ncL, L.aggPtr.p, L.aggNodes.p, L.nbrPtr, L.Aval,
                                                                                          L.cslot.p, cPtrMine, C.AvalB.p, nbcap);
            } else {
                if (smem > 40 * 1024) CUDA_CHECK(cudaFuncSetAttribute(k_galerkin<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, PFEM_SMEM_OPTIN));
                k_galerkin<3><<<divUp(ncL, hwPerBlock), hwPerBlock * 16, smem, c->stream>>>(ncL, L.aggPtr.p, L.aggNodes.p, L.nbrPtr, L.Aval,
                                                                                          L.cslot.p, cPtrMine, C.AvalB.p, nbcap);
            }
            LAUNCH_CHECK(c);
        }
        if (L.nextReplicated) {  // every rank gets every row of the replicated level
            std::vector<int64_t> cnt(L.blkCounts), dsp(L.blkDispls);
            for (auto& v : cnt) v *= BB;
            for (auto& v : dsp) v *= BB;
            commAllGatherV(c, C.AvalB.p + (size_t)L.blkDispls[c->rank] * BB, C.AvalB.p, cnt, dsp);
        }
    }
    // coarsest level: dense inverse when it is small enough
    MgLevel& Lc = *H.lev.back();
    H.denseOk = false;
    PhaseScope sub3(c, "MG coarsest inverse");
    if (Lc.n <= COARSEST_NODES && Lc.nVec == Lc.n) {
        H.nD = Lc.n * BS;
        H.dense.reserve((size_t)H.nD * H.nD + 8);
        H.flag.reserve(16);
        CUDA_CHECK(cudaMemsetAsync(H.flag.p + 8, 0, sizeof(int), c->stream));
        const size_t smem = ((size_t)H.nD * H.nD + H.nD) * sizeof(double) + (size_t)H.nD * sizeof(int);
        CUDA_CHECK(cudaFuncSetAttribute(k_dense_invert, cudaFuncAttributeMaxDynamicSharedMemorySize, PFEM_SMEM_OPTIN));
        k_dense_invert<<<1, 1024, smem, c->stream>>>(Lc.n, BS, Lc.nbrPtr, Lc.nbr, Lc.Aval, H.dense.p, H.flag.p + 8);
        LAUNCH_CHECK(c);
        int bad = 0;
        CUDA_CHECK(cudaMemcpyAsync(&bad, H.flag.p + 8, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        H.denseOk = bad == 0;
    }
    static const bool verbose = getenv("PFEM_MG_VERBOSE") != nullptr;
    if (verbose && !H.tuned) {
        fprintf(stderr, "[pfem mg] levels:");
        for (auto& L : H.lev) fprintf(stderr, " %d(nb<=%d, omega %.2f)", L->n, L->maxNb, L->omega);
        fprintf(stderr, "  coarsest dense: %s\n", H.denseOk ? "yes" : "no");
    }
    H.tuned = true;
    H.numericValid = true;
}

}  // namespace

void mgInvalidate(pfem_ctx* c, bool symbolic) {
    if (!c->mg) return;
    c->mg->numericValid = false;
    if (symbolic) c->mg->symbolicValid = false;
}
void mgDestroy(pfem_ctx* c) {
    delete c->mg;
    c->mg = nullptr;
}
// (re)build what is stale; false: no usable hierarchy (single level) -> the caller keeps node-block Jacobi
bool mgSetup(pfem_ctx* c) {
    if (!c->mg) c->mg = new MgHierarchy();
    MgHierarchy& H = *c->mg;
    static const int envNu = getenv("PFEM_MG_NU") ? atoi(getenv("PFEM_MG_NU")) : 0;
    static const double envOmega = getenv("PFEM_MG_OMEGA") ? atof(getenv("PFEM_MG_OMEGA")) : 0.0;
    static const double envOver = getenv("PFEM_MG_OVER") ? atof(getenv("PFEM_MG_OVER")) : 0.0;
    // default sweeps: V(3,3) in 3-D, V(2,2) in 2-D, 3 on the coarse levels (measured under FGMRES at C4 / Delaunay cloud / 2-D
    // n=700: 3-D 33 -> 23 cycles and 33.9 -> 27.9 ms with the third sweep, 2-D 39 -> 42 cycles and slower)
    const bool explicitNu = c->mgSweeps > 0 || envNu > 0;
    const int nu = c->mgSweeps > 0 ? c->mgSweeps : (envNu > 0 ? envNu : (c->dim == 3 ? 3 : 2));
    const double omega = c->mgDamping > 0 ? c->mgDamping : (envOmega > 0 ? envOmega : 0.0);  // 0: tuned per level
    const double over = envOver > 0 ? envOver : 1.5;
    if (omega != H.fixedOmega) H.numericValid = false, H.tuned = false;  // Dw carries the damping
    if (c->asmStamp != H.stamp) H.tuned = false;                          // another dt: the spectrum moved
    static const int envNuC = getenv("PFEM_MG_NUC") ? atoi(getenv("PFEM_MG_NUC")) : 0;
    H.nu = nu, H.nuCoarse = envNuC > 0 ? envNuC : (explicitNu ? nu + 1 : 3), H.fixedOmega = omega, H.over = over, H.stamp = c->asmStamp;
    static const int envW = getenv("PFEM_MG_W") ? atoi(getenv("PFEM_MG_W")) : -1;
    H.wFrom = envW;
    auto envInt = [](const char* name) { return getenv(name) ? atoi(getenv(name)) : -1; };
    static const int envPre = envInt("PFEM_MG_PRE"), envPost = envInt("PFEM_MG_POST"), envPreC = envInt("PFEM_MG_PREC"),
                     envPostC = envInt("PFEM_MG_POSTC");
    H.preFine = envPre, H.postFine = envPost, H.preCoarse = envPreC, H.postCoarse = envPostC;
    // fp32 vectors inside the cycle: only under the flexible GMRES (the caller says so) and with the fp32 matrix copies
    const bool envF32V = !(getenv("PFEM_MG_FP32V") && atoi(getenv("PFEM_MG_FP32V")) == 0);  // read per solve: tests switch it
    H.f32v = c->mgFlexible && envF32V && mgFp32();
    H.rhsIsF32 = false;  // until the Krylov method asks for the fp32 right-hand side (mgRhsF)
    if (!H.symbolicValid) {
        buildSymbolic(c, H);
        H.tuned = false;
    }
    if (H.lev.size() < 2) return false;
    if (!H.numericValid) buildNumeric(c, H);
    return true;
}
double* mgRhs(pfem_ctx* c) { return c->mg->lev[0]->b.p; }
float* mgRhsF(pfem_ctx* c) {
    MgHierarchy& H = *c->mg;
    if (!H.f32v) return nullptr;
    const size_t nAll = (size_t)H.lev[0]->nVec * (c->dim + 1);
    H.inF.reserve(nAll + 8);
    H.outF.reserve(nAll + 8);
    H.rhsIsF32 = true;
    return H.inF.p;
}
int mgLevelCount(pfem_ctx* c) { return c->mg ? (int)c->mg->lev.size() : 0; }
// out = V-cycle(rhs held in mgRhs()): an approximation of A^-1 rhs in the physical variables
void mgApply(pfem_ctx* c, double* out) {
    PhaseScope ph(c, "Preconditioner apply");
    MgHierarchy& H = *c->mg;
    static const bool noGraph = getenv("PFEM_MG_NOGRAPH") != nullptr;
    // partitioned mesh: the cycle runs un-graphed.  Capturing the NCCL point-to-point exchanges of the cycle into the graph
    // (PFEM_MG_GRAPH_NCCL=1) was tried on 2 and 4 B200s and DEADLOCKED in the first replay (round 2, profiles/README.md): the
    // ranks capture two graphs each (one per output vector) whose grouped send/recv lists differ per rank; kept opt-in only.
    static const bool graphNccl = getenv("PFEM_MG_GRAPH_NCCL") && atoi(getenv("PFEM_MG_GRAPH_NCCL")) == 1;
    const bool multiNoGraph = c->nRanks > 1 && (c->local || !graphNccl);
    if (H.f32v) {  // reserved here, not inside a capture
        const size_t nAll = (size_t)H.lev[0]->nVec * (c->dim + 1);
        H.inF.reserve(nAll + 8);
        H.outF.reserve(nAll + 8);
    }
    double* const userOut = out;
    if (H.f32v) out = nullptr;  // the captured part ends in H.outF: one graph for every output vector
    if (noGraph || H.graphBroken || c->profileDetail || multiNoGraph) {
        runCycle(c, H, out);
        convertResult(c, H, userOut);
        return;
    }
    // everything a captured cycle bakes in: buffers that can be reallocated, and the cycle parameters
    unsigned long long sig = 1469598103934665603ull;
    auto mix = [&](unsigned long long v) { sig = (sig ^ v) * 1099511628211ull; };
    mix((unsigned long long)(uintptr_t)H.lev[0]->Aval), mix((unsigned long long)(uintptr_t)H.lev[0]->nbr);
    mix((unsigned long long)(uintptr_t)H.lev[0]->b.p), mix((unsigned long long)H.lev.size()), mix((unsigned long long)H.nu), mix((unsigned long long)H.nuCoarse);
    mix((unsigned long long)(H.preFine + 1)), mix((unsigned long long)(H.postFine + 1)), mix((unsigned long long)(H.preCoarse + 1)), mix((unsigned long long)(H.postCoarse + 1));
    mix((unsigned long long)(H.wFrom + 7)), mix(H.f32v ? (H.rhsIsF32 ? 3ull : 2ull) : 1ull);
    if (H.f32v) mix((unsigned long long)(uintptr_t)H.inF.p), mix((unsigned long long)(uintptr_t)H.outF.p);
    mix((unsigned long long)__double_as_longlong_host(H.over)), mix(H.denseOk ? 1ull : 0ull), mix((unsigned long long)H.nD);
    for (auto& L : H.lev) {
        mix((unsigned long long)(uintptr_t)L->Dw.p), mix((unsigned long long)(uintptr_t)L->Af.p), mix((unsigned long long)L->n);
        mix((unsigned long long)(uintptr_t)L->DwF.p);
        if (L->distributed && L->plan) {  // a captured exchange bakes in the pack buffer and the send list
            L->plan->sendBuf.reserve((size_t)L->plan->nSendTotal * (c->dim + 1) + 8);
            mix((unsigned long long)(uintptr_t)L->plan->sendBuf.p), mix((unsigned long long)(uintptr_t)L->plan->sendIdx.p);
        }
    }
    for (auto& g : H.graphs)
        if (g.out == out && g.sig == sig) {
            CUDA_CHECK(cudaGraphLaunch(g.exec, c->stream));
            c->launches += g.launches;
            convertResult(c, H, userOut);
            return;
        }
    for (size_t k = 0; k < H.graphs.size();)  // stale capture for this output vector
        if (H.graphs[k].out == out) {
            cudaGraphExecDestroy(H.graphs[k].exec);
            H.graphs.erase(H.graphs.begin() + k);
        } else
            ++k;
    const bool wasProfiling = c->profiling;
    const int64_t launches0 = c->launches;
    c->profiling = false;  // no event records inside a capture
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    bool ok = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (ok) {
        try {
            runCycle(c, H, out);
        } catch (...) {
            ok = false;
        }
        if (cudaStreamEndCapture(c->stream, &graph) != cudaSuccess || !graph) ok = false;
        if (ok && cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) ok = false;
        if (graph) cudaGraphDestroy(graph);
    }
    c->profiling = wasProfiling;
    const int per = (int)(c->launches - launches0);
    c->launches = launches0;
    if (!ok) {
        cudaGetLastError();
        H.graphBroken = true;
        runCycle(c, H, out);
        convertResult(c, H, userOut);
        return;
    }
    MgHierarchy::GraphSlot g;
    g.out = out, g.exec = exec, g.sig = sig, g.launches = per;
    H.graphs.push_back(g);
    CUDA_CHECK(cudaGraphLaunch(exec, c->stream));
    c->launches += per;
    convertResult(c, H, userOut);
}
