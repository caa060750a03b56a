// topology.cu -- per-remesh pattern build on the device.
//
// Replaces the pattern half of Eigen::SparseMatrix::setFromTriplets (MomContEquationPSPG.inl:136) and the
// Node::m_elements / m_neighbourNodes lists the reference keeps per node (Node.hpp:95-99): from the element
// connectivity alone it builds
//   n2ePtr/n2e : node -> incident elements, ascending element index (the order in which setFromTriplets sums
//                duplicates, so the assembly kernel can reproduce that order),
//   nbrPtr/nbr : node -> sorted neighbour nodes incl. itself == the (dim+1)x(dim+1) block-row pattern of m_A.
// All kernels are integer/HBM-bound; one warp per node for the neighbour pass.
#include "common.cuh"

namespace {

constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void k_convert_conn(const unsigned long long* __restrict__ in, int* __restrict__ conn, int64_t n, int nNodes,
                               int* __restrict__ cnt, int* __restrict__ bad) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long v = in[i];
    if (v >= (unsigned long long)nNodes) {
        atomicOr(bad, 1);
        conn[i] = 0;
        return;
    }
    conn[i] = (int)v;
    atomicAdd(&cnt[(int)v], 1);
}

// tile-local exclusive scan; tile totals to blockSums
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tile(int* __restrict__ data, int n, int* __restrict__ blockSums) {
    __shared__ int warpSums[SCAN_THREADS / 32];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < n) ? data[base + k] : 0;
        sum += v[k];
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warpSums[w] = inc;
    __syncthreads();
    if (w == 0) {
        int s = (lane < SCAN_THREADS / 32) ? warpSums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        if (lane < SCAN_THREADS / 32) warpSums[lane] = s;  // inclusive over warps
    }
    __syncthreads();
    int excl = inc - sum + (w > 0 ? warpSums[w - 1] : 0);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) data[base + k] = excl;
        excl += v[k];
    }
    if (threadIdx.x == SCAN_THREADS - 1 && blockSums) blockSums[blockIdx.x] = excl;
}

__global__ void k_scan_add(int* __restrict__ data, int n, const int* __restrict__ blockSums) {
    const int i = blockIdx.x * SCAN_TILE + threadIdx.x;
    const int add = blockSums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        int idx = i + k * SCAN_THREADS;
        if (idx < n) data[idx] += add;
    }
}

__global__ void k_fill_n2e(const int* __restrict__ conn, int64_t n, int npe, const int* __restrict__ ptr,
                           int* __restrict__ cursor, int* __restrict__ n2e) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int node = conn[i];
    const int pos = atomicAdd(&cursor[node], 1);
    n2e[ptr[node] + pos] = (int)(i / npe);
}

// ascending element index per node (lists are short: ~6 in 2-D, ~24 in 3-D); also the valence maximum
__global__ void k_sort_n2e(const int* __restrict__ ptr, int* __restrict__ n2e, int nNodes, int* __restrict__ maxE) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    int len = 0;
    if (n < nNodes) {
        const int b = ptr[n];
        len = ptr[n + 1] - b;
        int* a = n2e + b;
        for (int i = 1; i < len; ++i) {
            const int key = a[i];
            int j = i - 1;
            while (j >= 0 && a[j] > key) {
                a[j + 1] = a[j];
                --j;
            }
            a[j + 1] = key;
        }
    }
    for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
    if ((threadIdx.x & 31) == 0 && len > 0) atomicMax(maxE, len);
}

// FREE bit := node belongs to no element (Node.inl:48-51), whatever the caller passed
__global__ void k_fix_free_flag(uint8_t* __restrict__ flags, const int* __restrict__ ptr, int nNodes) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nNodes) return;
    uint8_t f = flags[n] & ~(uint8_t)PFEM_NODE_FREE;
    if (ptr[n + 1] == ptr[n]) f |= PFEM_NODE_FREE;
    flags[n] = f;
}

// One warp per node.  FILL=false: count distinct neighbours.  FILL=true: write them sorted + diagSlot, and for every
// incident element k the slots of its nodes in that sorted list (n2eSlots, one byte per local node) and for every
// block (i, slot) the bit mask of the incident elements that contain that neighbour (blkMask, CH words per block).
// Both are what the assembly kernel needs to gather without searching.
template <bool FILL>
__global__ void k_neighbours(const int* __restrict__ conn, int npe, const int* __restrict__ n2ePtr,
                             const int* __restrict__ n2e, int nNodes, int candCap, int nbCap, int CH,
                             int* __restrict__ nbrCntOrPtr, int* __restrict__ nbr, int* __restrict__ diagSlot,
                             int* __restrict__ maxNb, unsigned* __restrict__ n2eSlots, unsigned* __restrict__ blkMask) {
    extern __shared__ int smem[];
    const int warpInBlock = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int node = blockIdx.x * (blockDim.x >> 5) + warpInBlock;
    if (node >= nNodes) return;
    const int perWarp = candCap + (FILL ? nbCap * (1 + CH) : 0);
    unsigned* cand = reinterpret_cast<unsigned*>(smem) + (size_t)warpInBlock * perWarp;
    unsigned* sorted = cand + candCap;     // FILL only
    unsigned* masks = sorted + nbCap;      // FILL only, nbCap*CH
    const int eb = n2ePtr[node], ne = n2ePtr[node + 1] - eb;
    const int C = ne * npe;
    if (ne == 0) {  // isolated node: the block row is its own diagonal block
        if (lane == 0) {
            if (FILL) {
                nbr[nbrCntOrPtr[node]] = node;
                diagSlot[node] = 0;
                for (int ch = 0; ch < CH; ++ch) blkMask[(size_t)nbrCntOrPtr[node] * CH + ch] = 0u;
            } else
                nbrCntOrPtr[node] = 1;
        }
        return;
    }
    // candidates = the nodes of the incident elements; warp-wide bitonic sort in shared memory (padded with 0xffffffff to a
    // power of two), then the distinct values are the entries that differ from their left neighbour.  Round 2: replaces an
    // O(C^2) duplicate sweep + rank count (1.65 -> ~0.7 ms for the two passes at 2 M tets).
    int P = 32;
    while (P < C) P <<= 1;
    for (int c = lane; c < P; c += 32) cand[c] = c < C ? (unsigned)conn[(size_t)n2e[eb + c / npe] * npe + (c % npe)] : 0xffffffffu;
    __syncwarp();
    for (int k = 2; k <= P; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = lane; t < (P >> 1); t += 32) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int q = i | j;
                const unsigned x = cand[i], y = cand[q];
                const bool up = (i & k) == 0;
                if ((x > y) == up) cand[i] = y, cand[q] = x;
            }
            __syncwarp();
        }
    int nDistinct = 0;
    for (int c0 = 0; c0 < P; c0 += 32) {  // (uniform trip count: the ballot needs the whole warp)
        const int c = c0 + lane;
        const unsigned v = cand[c];
        const bool first = v != 0xffffffffu && (c == 0 || cand[c - 1] != v);
        const unsigned bal = __ballot_sync(0xffffffffu, first);
        if (FILL && first) {
            const int rank = nDistinct + __popc(bal & ((1u << lane) - 1u));
            nbr[nbrCntOrPtr[node] + rank] = (int)v;
            sorted[rank] = v;
            if ((int)v == node) diagSlot[node] = rank;
        }
        nDistinct += __popc(bal);
    }
    if (!FILL) {
        if (lane == 0) {
            nbrCntOrPtr[node] = nDistinct;
            atomicMax(maxNb, nDistinct);
        }
        return;
    }
    __syncwarp();
    const int nb0 = nbrCntOrPtr[node], nb = nbrCntOrPtr[node + 1] - nb0;
    for (int t = lane; t < nb * CH; t += 32) masks[t] = 0u;
    __syncwarp();
    // slot of every candidate = lower_bound in the sorted distinct list; one byte per (element, local node)
    unsigned char* slotBytes = reinterpret_cast<unsigned char*>(n2eSlots + eb);
    for (int c = lane; c < C; c += 32) {
        const unsigned v = (unsigned)conn[(size_t)n2e[eb + c / npe] * npe + (c % npe)];  // original (element, local node) order
        int lo = 0, hi = nb - 1;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (sorted[mid] < v) lo = mid + 1;
            else hi = mid;
        }
        const int k = c / npe, m = c % npe;
        slotBytes[k * 4 + m] = (unsigned char)lo;
        if (npe == 3 && m == 0) slotBytes[k * 4 + 3] = 0xff;  // unused 4th byte never matches a slot
        atomicOr(&masks[lo * CH + (k >> 5)], 1u << (k & 31));
    }
    __syncwarp();
    for (int t = lane; t < nb * CH; t += 32) blkMask[(size_t)nb0 * CH + t] = masks[t];
}

}  // namespace

// In-place exclusive scan of n ints (recursive tile scan); if totalOut != null the grand total is stored there
// (device pointer).
static void scanRec(pfem_ctx* c, int* data, int n, int* scratch, int* totalOut) {
    const int tiles = divUp(n, SCAN_TILE);
    k_scan_tile<<<tiles, SCAN_THREADS, 0, c->stream>>>(data, n, scratch);
    LAUNCH_CHECK(c);
    if (tiles == 1) {
        if (totalOut) CUDA_CHECK(cudaMemcpyAsync(totalOut, scratch, sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
        return;
    }
    scanRec(c, scratch, tiles, scratch + tiles + 1, totalOut);
    k_scan_add<<<tiles, SCAN_THREADS, 0, c->stream>>>(data, n, scratch);
    LAUNCH_CHECK(c);
}

void exclusiveScanInt(pfem_ctx* c, int* data, int n, int* totalOut) {
    size_t need = 0;
    for (int m = n;;) {
        const int tiles = divUp(m, SCAN_TILE);
        need += tiles + 1;
        if (tiles == 1) break;
        m = tiles;
    }
    c->scanScratch.reserve(need + 8);
    scanRec(c, data, n, c->scanScratch.p, totalOut);
}

void topoBuild(pfem_ctx* c, int64_t nNodes64, int64_t nElems64, const uint64_t* elemNodes, const uint8_t* flagsHost) {
    PhaseScope ph(c, "Build pattern");
    const int npe = c->dim + 1;
    PFEM_REQUIRE(nNodes64 > 0 && nElems64 >= 0, PFEM_ERR_INVALID, "set_topology: empty mesh");
    PFEM_REQUIRE(nNodes64 < (1ll << 30) && nElems64 * npe < (1ll << 31), PFEM_ERR_INVALID,
                 "set_topology: mesh too large for int32 device indices");
    const int nNodes = (int)nNodes64, nElems = (int)nElems64;
    const int64_t nConn = (int64_t)nElems * npe;
    c->nNodes = nNodes;
    c->nElems = nElems;
    c->nRows = nNodes;  // until pfem_set_partition narrows it to the owned nodes
    c->plan.clear();
    c->tilesValid = c->orderValid = false;
    c->nFacets = c->nFstNodes = 0;  // facets belong to the previous mesh (pfem_set_facets)
    c->haveTopology = false;
    c->haveSystem = c->haveSolution = c->haveQprev = c->haveSnapshot = c->havePositions = c->haveDirichlet = false;
    c->nnzReference = -1;
    c->rowDirDirty = true;
    c->haveTemperature = c->haveTemperatureBc = c->haveHeatSystem = false;
    mgInvalidate(c, true);

    c->conn.reserve(nConn + 4);
    c->flags.reserve(nNodes);
    c->n2ePtr.reserve(nNodes + 2);
    c->n2e.reserve(nConn + 4);
    c->nbrPtr.reserve(nNodes + 2);
    c->diagSlot.reserve(nNodes);
    c->stage64.reserve(nConn + 4);
    c->scratchI.reserve((size_t)nNodes + 64);
    int* cursor = c->scratchI.p;
    int* misc = c->scratchI.p + nNodes;  // [0]=bad, [1]=maxE, [2]=maxNb, [3]=total
    CUDA_CHECK(cudaMemcpyAsync(c->stage64.p, elemNodes, nConn * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(c->flags.p, flagsHost, nNodes, cudaMemcpyHostToDevice, c->stream));
    CUDA_CHECK(cudaMemsetAsync(c->n2ePtr.p, 0, (nNodes + 2) * sizeof(int), c->stream));
    CUDA_CHECK(cudaMemsetAsync(c->scratchI.p, 0, ((size_t)nNodes + 64) * sizeof(int), c->stream));
    if (nConn > 0) {
        k_convert_conn<<<divUp(nConn, 256), 256, 0, c->stream>>>(c->stage64.p, c->conn.p, nConn, nNodes, c->n2ePtr.p, misc);
        LAUNCH_CHECK(c);
    }
    exclusiveScanInt(c, c->n2ePtr.p, nNodes + 1, nullptr);
    if (nConn > 0) {
        k_fill_n2e<<<divUp(nConn, 256), 256, 0, c->stream>>>(c->conn.p, nConn, npe, c->n2ePtr.p, cursor, c->n2e.p);
        LAUNCH_CHECK(c);
    }
    k_sort_n2e<<<divUp(nNodes, 128), 128, 0, c->stream>>>(c->n2ePtr.p, c->n2e.p, nNodes, misc + 1);
    LAUNCH_CHECK(c);
    k_fix_free_flag<<<divUp(nNodes, 256), 256, 0, c->stream>>>(c->flags.p, c->n2ePtr.p, nNodes);
    LAUNCH_CHECK(c);
    int h[4];
    CUDA_CHECK(cudaMemcpyAsync(h, misc, 4 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    PFEM_REQUIRE(h[0] == 0, PFEM_ERR_INVALID, "set_topology: element node index out of range");
    c->maxE = h[1];
    PFEM_REQUIRE(c->maxE <= 4096, PFEM_ERR_INVALID, "set_topology: node valence above 4096 elements");

    // neighbour lists
    const int warps = 8;
    int candCap = 32;  // power of two >= the largest candidate count (bitonic sort in k_neighbours)
    while (candCap < max(c->maxE, 1) * npe) candCap <<= 1;
    const int CH = (max(c->maxE, 1) + 31) / 32;
    c->maskWords = CH;
    size_t smem = (size_t)warps * candCap * sizeof(int);
    if (smem > 40 * 1024)
        CUDA_CHECK(cudaFuncSetAttribute(k_neighbours<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PFEM_SMEM_OPTIN));
    CUDA_CHECK(cudaMemsetAsync(c->nbrPtr.p, 0, (nNodes + 2) * sizeof(int), c->stream));
    k_neighbours<false><<<divUp(nNodes, warps), warps * 32, smem, c->stream>>>(
        c->conn.p, npe, c->n2ePtr.p, c->n2e.p, nNodes, candCap, 0, CH, c->nbrPtr.p, nullptr, nullptr, misc + 2, nullptr, nullptr);
    LAUNCH_CHECK(c);
    exclusiveScanInt(c, c->nbrPtr.p, nNodes + 1, misc + 3);
    CUDA_CHECK(cudaMemcpyAsync(h, misc, 4 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->maxNb = h[2];
    c->nBlocks = h[3];
    PFEM_REQUIRE(c->maxNb <= 254, PFEM_ERR_INVALID, "set_topology: node with more than 254 neighbours");
    c->nbr.reserve(c->nBlocks + 4);
    c->n2eSlots.reserve(nConn + 4);
    c->blkMask.reserve((size_t)c->nBlocks * CH + 4);
    const int nbCap = c->maxNb;
    smem = (size_t)warps * (candCap + nbCap * (1 + CH)) * sizeof(int);
    if (smem > 40 * 1024)
        CUDA_CHECK(cudaFuncSetAttribute(k_neighbours<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PFEM_SMEM_OPTIN));
    k_neighbours<true><<<divUp(nNodes, warps), warps * 32, smem, c->stream>>>(
        c->conn.p, npe, c->n2ePtr.p, c->n2e.p, nNodes, candCap, nbCap, CH, c->nbrPtr.p, c->nbr.p, c->diagSlot.p, nullptr,
        c->n2eSlots.p, c->blkMask.p);
    LAUNCH_CHECK(c);

    // nodal fields sized for the new mesh
    const size_t n4 = (size_t)nNodes * 4;
    c->X4.reserve(n4);
    c->Xsave4.reserve(n4);
    c->V4.reserve(n4);
    c->A4.reserve(n4);
    c->VP4.reserve(n4);
    c->dirMask.reserve(nNodes);
    c->dirVal4.reserve(n4);
    CUDA_CHECK(cudaMemsetAsync(c->X4.p, 0, n4 * sizeof(double), c->stream));
    CUDA_CHECK(cudaMemsetAsync(c->V4.p, 0, n4 * sizeof(double), c->stream));
    CUDA_CHECK(cudaMemsetAsync(c->A4.p, 0, n4 * sizeof(double), c->stream));
    CUDA_CHECK(cudaMemsetAsync(c->VP4.p, 0, n4 * sizeof(double), c->stream));
    CUDA_CHECK(cudaMemsetAsync(c->dirMask.p, 0, nNodes, c->stream));
    CUDA_CHECK(cudaMemsetAsync(c->dirVal4.p, 0, n4 * sizeof(double), c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->haveTopology = true;
}
