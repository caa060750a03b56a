// partition.cu -- spatial (RCB) partition of a mesh over the GPUs of one box, with one ghost-element layer
// (SURVEY.md section 8e).  Host code (C++/OpenMP), once per remesh; the reference has no counterpart (it is a single
// OpenMP process) -- the call site is where it hands the new mesh over after Mesh::remesh (Mesh.cpp:919-926).
//
// Scheme -- "owner computes" (same as the numpy statement in pfem_b200/partition.py, which the tests compare it with):
//   * nodes are split by recursive coordinate bisection into nRanks parts (nRanks need not be a power of two: cuts are
//     proportional); the order along the widest axis is the total order (coordinate, node index), so the split does not
//     depend on how the selection algorithm treats ties;
//   * a rank keeps every element incident to at least one of its nodes, in ascending GLOBAL element index, so that the
//     node-gather kernels sum element contributions in exactly the order a single GPU would;
//   * local node numbering = owned nodes (ascending global id) followed by ghost nodes grouped by owner rank (ascending
//     global id inside a group): every halo receive lands in one contiguous range of the nodal arrays;
//   * per peer: send list = local ids of owned nodes that are ghosts on that peer, in the peer's ghost order.
#include <algorithm>
#include <cstring>
#include <memory>
#include <numeric>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "common.cuh"

struct PartLocal {
    bool built = false;
    int64_t nOwned = 0;
    std::vector<int64_t> l2gNodes, l2gElems;
    std::vector<uint64_t> conn;  // local connectivity, row-major
    std::vector<int32_t> peerRank, sendIdx;
    std::vector<int64_t> sendOffsets, recvStart, recvCount;
};
struct pfem_partition {
    int dim = 0, nRanks = 1;
    int64_t nNodes = 0, nElems = 0;
    std::vector<int32_t> owner;
    const uint64_t* conn = nullptr;  // the caller's connectivity: must stay valid until pfem_partition_destroy
    std::vector<PartLocal> local;
    std::string err;
};

namespace {

void rcbSplit(const double* x, int64_t nNodes, int dim, int32_t* idx, int64_t count, int r0, int nr, int32_t* owner) {
    if (nr == 1 || count == 0) {
        for (int64_t k = 0; k < count; ++k) owner[idx[k]] = r0;
        return;
    }
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int64_t k = 0; k < count; ++k)
        for (int d = 0; d < dim; ++d) {
            const double v = x[(size_t)d * nNodes + idx[k]];
            lo[d] = v < lo[d] ? v : lo[d];
            hi[d] = v > hi[d] ? v : hi[d];
        }
    int axis = 0;
    for (int d = 1; d < dim; ++d)
        if (hi[d] - lo[d] > hi[axis] - lo[axis]) axis = d;
    const double* xa = x + (size_t)axis * nNodes;
    const int nl = nr / 2;
    const int64_t cut = (count * nl) / nr;
    auto less = [xa](int32_t a, int32_t b) { return xa[a] < xa[b] || (xa[a] == xa[b] && a < b); };
    if (cut > 0 && cut < count) std::nth_element(idx, idx + cut, idx + count, less);
#pragma omp task default(shared) if (count > 200000)
    rcbSplit(x, nNodes, dim, idx, cut, r0, nl, owner);
#pragma omp task default(shared) if (count > 200000)
    rcbSplit(x, nNodes, dim, idx + cut, count - cut, r0 + nl, nr - nl, owner);
#pragma omp taskwait
}

void buildLocal(pfem_partition& P, int rank) {
    PartLocal& L = P.local[rank];
    if (L.built) return;
    const int npe = P.dim + 1;
    const int64_t nn = P.nNodes, ne = P.nElems;
    const int32_t* owner = P.owner.data();
    const uint64_t* conn = P.conn;
    // my elements: any node owned here; ascending global index (chunked so that the concatenation keeps the order)
    int nChunks = 1;
#ifdef _OPENMP
    nChunks = std::max(1, omp_get_max_threads());
#endif
    std::vector<std::vector<int64_t>> chunkE(nChunks);
#pragma omp parallel for schedule(static, 1)
    for (int ch = 0; ch < nChunks; ++ch) {
        const int64_t e0 = ne * ch / nChunks, e1 = ne * (ch + 1) / nChunks;
        auto& out = chunkE[ch];
        for (int64_t e = e0; e < e1; ++e) {
            bool mine = false;
            for (int k = 0; k < npe; ++k) mine = mine || owner[conn[(size_t)e * npe + k]] == rank;
            if (mine) out.push_back(e);
        }
    }
    L.l2gElems.clear();
    for (auto& v : chunkE) L.l2gElems.insert(L.l2gElems.end(), v.begin(), v.end());
    const int64_t nle = (int64_t)L.l2gElems.size();
    // owned nodes, then ghosts by (owner, global id)
    L.l2gNodes.clear();
    for (int64_t n = 0; n < nn; ++n)
        if (owner[n] == rank) L.l2gNodes.push_back(n);
    L.nOwned = (int64_t)L.l2gNodes.size();
    std::vector<uint8_t> seen((size_t)nn, 0);
    std::vector<int64_t> ghosts;
    for (int64_t le = 0; le < nle; ++le)
        for (int k = 0; k < npe; ++k) {
            const int64_t n = (int64_t)conn[(size_t)L.l2gElems[le] * npe + k];
            if (owner[n] != rank && !seen[n]) {
                seen[n] = 1;
                ghosts.push_back(n);
            }
        }
    std::sort(ghosts.begin(), ghosts.end(), [owner](int64_t a, int64_t b) { return owner[a] < owner[b] || (owner[a] == owner[b] && a < b); });
    L.l2gNodes.insert(L.l2gNodes.end(), ghosts.begin(), ghosts.end());
    std::vector<int32_t> g2l((size_t)nn, -1);
    for (size_t l = 0; l < L.l2gNodes.size(); ++l) g2l[L.l2gNodes[l]] = (int32_t)l;
    L.conn.resize((size_t)nle * npe);
#pragma omp parallel for schedule(static)
    for (int64_t le = 0; le < nle; ++le)
        for (int k = 0; k < npe; ++k) L.conn[(size_t)le * npe + k] = (uint64_t)g2l[conn[(size_t)L.l2gElems[le] * npe + k]];
    // send side: my owned nodes that share an element with a node owned by rank q  ->  (q, node), unique, ascending
    std::vector<uint64_t> keys;
    for (int64_t le = 0; le < nle; ++le) {
        const uint64_t* en = conn + (size_t)L.l2gElems[le] * npe;
        bool mixed = false;
        for (int k = 1; k < npe; ++k) mixed = mixed || owner[en[k]] != owner[en[0]];
        if (!mixed) continue;
        for (int a = 0; a < npe; ++a) {
            if (owner[en[a]] != rank) continue;
            for (int b = 0; b < npe; ++b)
                if (owner[en[b]] != rank) keys.push_back((uint64_t)owner[en[b]] * (uint64_t)nn + en[a]);
        }
    }
    std::sort(keys.begin(), keys.end());
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    std::vector<int32_t> peerSet;
    for (int64_t g : ghosts) peerSet.push_back(owner[g]);
    for (uint64_t k : keys) peerSet.push_back((int32_t)(k / (uint64_t)nn));
    std::sort(peerSet.begin(), peerSet.end());
    peerSet.erase(std::unique(peerSet.begin(), peerSet.end()), peerSet.end());
    L.peerRank = peerSet;
    L.sendIdx.clear();
    L.sendOffsets.assign(1, 0);
    L.recvStart.clear();
    L.recvCount.clear();
    size_t kpos = 0, gpos = 0;
    for (int32_t p : peerSet) {
        while (kpos < keys.size() && (int32_t)(keys[kpos] / (uint64_t)nn) == p) {
            L.sendIdx.push_back(g2l[keys[kpos] % (uint64_t)nn]);
            ++kpos;
        }
        L.sendOffsets.push_back((int64_t)L.sendIdx.size());
        const size_t g0 = gpos;
        while (gpos < ghosts.size() && owner[ghosts[gpos]] == p) ++gpos;
        L.recvStart.push_back(L.nOwned + (int64_t)(gpos > g0 ? g0 : 0));
        L.recvCount.push_back((int64_t)(gpos - g0));
    }
    L.built = true;
}

}  // namespace

extern "C" {

int pfem_partition_create(pfem_partition** out, int dim, int64_t nNodes, int64_t nElems, const uint64_t* elemNodes, const double* x,
                          int nRanks) {
    if (!out) return PFEM_ERR_INVALID;
    *out = nullptr;
    if ((dim != 2 && dim != 3) || nNodes <= 0 || nElems < 0 || nRanks < 1 || !x || (nElems > 0 && !elemNodes) || nNodes >= (1ll << 31))
        return PFEM_ERR_INVALID;
    try {
        std::unique_ptr<pfem_partition> P(new pfem_partition);
        P->dim = dim, P->nRanks = nRanks, P->nNodes = nNodes, P->nElems = nElems;
        const int npe = dim + 1;
        P->conn = elemNodes;
        int bad = 0;
        const int64_t nConn = nElems * npe;
#pragma omp parallel for reduction(| : bad) schedule(static)
        for (int64_t k = 0; k < nConn; ++k) bad |= elemNodes[k] >= (uint64_t)nNodes ? 1 : 0;
        if (bad) return PFEM_ERR_INVALID;
        P->owner.assign((size_t)nNodes, 0);
        std::vector<int32_t> idx((size_t)nNodes);
        std::iota(idx.begin(), idx.end(), 0);
#pragma omp parallel
#pragma omp single
        rcbSplit(x, nNodes, dim, idx.data(), nNodes, 0, nRanks, P->owner.data());
        P->local.resize(nRanks);
        *out = P.release();
    } catch (...) {
        return PFEM_ERR_INVALID;
    }
    return PFEM_OK;
}
int pfem_partition_destroy(pfem_partition* P) {
    delete P;
    return PFEM_OK;
}
int pfem_partition_owner(const pfem_partition* P, int32_t* owner) {
    if (!P || !owner) return PFEM_ERR_INVALID;
    memcpy(owner, P->owner.data(), (size_t)P->nNodes * sizeof(int32_t));
    return PFEM_OK;
}
int pfem_partition_local_sizes(pfem_partition* P, int rank, int64_t* nLocalNodes, int64_t* nOwned, int64_t* nLocalElems, int32_t* nPeers,
                               int64_t* nSend) {
    if (!P || rank < 0 || rank >= P->nRanks) return PFEM_ERR_INVALID;
    try {
        buildLocal(*P, rank);
    } catch (...) {
        return PFEM_ERR_INVALID;
    }
    const PartLocal& L = P->local[rank];
    if (nLocalNodes) *nLocalNodes = (int64_t)L.l2gNodes.size();
    if (nOwned) *nOwned = L.nOwned;
    if (nLocalElems) *nLocalElems = (int64_t)L.l2gElems.size();
    if (nPeers) *nPeers = (int32_t)L.peerRank.size();
    if (nSend) *nSend = (int64_t)L.sendIdx.size();
    return PFEM_OK;
}
int pfem_partition_local_get(pfem_partition* P, int rank, int64_t* l2gNodes, int64_t* l2gElems, uint64_t* localConn, int32_t* peerRank,
                             int64_t* sendOffsets, int32_t* sendIdx, int64_t* recvStart, int64_t* recvCount) {
    if (!P || rank < 0 || rank >= P->nRanks) return PFEM_ERR_INVALID;
    try {
        buildLocal(*P, rank);
    } catch (...) {
        return PFEM_ERR_INVALID;
    }
    const PartLocal& L = P->local[rank];
    auto cp = [](auto* dst, const auto& v) {
        if (dst && !v.empty()) memcpy(dst, v.data(), v.size() * sizeof(v[0]));
    };
    cp(l2gNodes, L.l2gNodes);
    cp(l2gElems, L.l2gElems);
    cp(localConn, L.conn);
    cp(peerRank, L.peerRank);
    cp(sendOffsets, L.sendOffsets);
    cp(sendIdx, L.sendIdx);
    cp(recvStart, L.recvStart);
    cp(recvCount, L.recvCount);
    return PFEM_OK;
}

}  // extern "C"
