// comm.cu -- multi-GPU plumbing (one process per GPU): NCCL communicator, halo exchange of nodal records, all-reduces.
//
// Sharding scheme (pfem_b200/partition.py): RCB over node coordinates; a rank owns its nodes'
// rows and holds one ghost-element layer, so the node-gather kernels need NO reduction of partial sums -- after each
// element pass the owners send the updated nodal records (32-byte X4/V4/A4 records or (dim+1)-double Krylov entries)
// of their interface nodes; ghosts of one owner are contiguous, so receives land directly in the nodal arrays.
// Messages are small (tens of KB .. ~1 MB): latency-bound, one grouped ncclSend/ncclRecv per exchange over NVLink.
// NCCL is dlopen'ed so that the library loads (and exports its symbols) on a box without NCCL/GPU; a missing NCCL is
// a loud PFEM_ERR_COMM, never a fallback.
// Second transport, "local": the ranks are contexts of ONE process, each driven by its own host thread (the reference is
// a single process: this is how its one main thread's workers reach several GPUs, and how the parity tests run a
// partitioned mesh on a box with a single GPU).  Exchanges are device-to-device copies between the ranks' buffers, ordered
// by CUDA events across the ranks' streams; the host threads only meet at barriers.  Same call sequence as the NCCL path.
#include <dlfcn.h>

#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.cuh"

// minimal NCCL ABI (nccl.h 2.x; stable since 2.0)
typedef struct ncclComm* ncclComm_t;
typedef struct {
    char internal[128];
} ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclInt8 = 0, ncclUint8 = 1, ncclInt32 = 2, ncclUint32 = 3, ncclInt64 = 4, ncclUint64 = 5, ncclFloat16 = 6, ncclFloat32 = 7, ncclFloat64 = 8 } ncclDataType_t;
typedef enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 } ncclRedOp_t;

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi* loadNccl() {
    static NcclApi published;
    static std::mutex mtx;
    std::lock_guard<std::mutex> lock(mtx);
    if (published.handle) return &published;
    NcclApi api;  // filled locally; published only when every symbol has resolved
    const char* env = getenv("PFEM_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        if (!n) continue;
        api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) pfemThrow(PFEM_ERR_COMM, std::string("cannot dlopen NCCL (set PFEM_NCCL_LIB): ") + dlerror());
    auto sym = [&](const char* s) {
        void* p = dlsym(api.handle, s);
        if (!p) pfemThrow(PFEM_ERR_COMM, std::string("NCCL symbol missing: ") + s);
        return p;
    };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
    api.Send = (decltype(api.Send))sym("ncclSend");
    api.Recv = (decltype(api.Recv))sym("ncclRecv");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    published = api;
    return &published;
}
#define NCCL_CHECK(api, expr)                                                                             \
    do {                                                                                                  \
        ncclResult_t _r = (expr);                                                                         \
        if (_r != ncclSuccess) pfemThrow(PFEM_ERR_COMM, std::string(#expr) + " -> " + (api)->GetErrorString(_r)); \
    } while (0)

namespace {
// buf[(e*width + c)] = arr[idx[e]*width + c]
template <typename T>
__global__ void k_pack(const int* __restrict__ idx, int nSend, int width, const T* __restrict__ arr, T* __restrict__ buf) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nSend * width) return;
    const int e = t / width, cidx = t % width;
    buf[t] = arr[(size_t)idx[e] * width + cidx];
}
// local transport all-reduce: every rank stores its values in its row of a pinned, device-mapped host bank, then every
// rank combines the rows in RANK ORDER -> the same bits on every rank
__global__ void k_slots_store(const double* __restrict__ src, int n, double* __restrict__ row) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) row[t] = src[t];
}
__global__ void k_slots_reduce(const double* __restrict__ bank, int rowStride, int nRanks, int n, int isMin, double* __restrict__ dst) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    double a = bank[t];
    for (int r = 1; r < nRanks; ++r) {
        const double b = bank[(size_t)r * rowStride + t];
        if (isMin) a = (b < a || b != b) ? b : a;  // NaN propagates, like the CFL reductions
        else a += b;
    }
    dst[t] = a;
}
}  // namespace

// ---- local transport ------------------------------------------------------------------------------------------------
struct LocalGroup {
    static constexpr int SLOT_CAP = 64;
    int n = 0;
    std::mutex m;
    std::condition_variable cv;
    int arrived = 0;
    unsigned long long gen = 0;
    bool aborted = false;
    struct Pub {
        pfem_ctx* ctx = nullptr;
        const double* src = nullptr;     // packed send buffer / all-gather source of the current exchange
        const HaloPlan* plan = nullptr;
    };
    std::vector<Pub> pub;
    double* bank = nullptr;  // pinned + mapped: n rows of SLOT_CAP doubles
    void barrier() {
        std::unique_lock<std::mutex> lk(m);
        if (aborted) pfemThrow(PFEM_ERR_COMM, "local communicator aborted by another rank");
        const unsigned long long g = gen;
        if (++arrived == n) {
            arrived = 0;
            ++gen;
            cv.notify_all();
            return;
        }
        const bool ok = cv.wait_for(lk, std::chrono::seconds(300), [&] { return gen != g || aborted; });
        if (!ok) {
            aborted = true;
            cv.notify_all();
            pfemThrow(PFEM_ERR_COMM, "local communicator: barrier timed out (ranks issued different call sequences?)");
        }
        if (gen == g && aborted) pfemThrow(PFEM_ERR_COMM, "local communicator aborted by another rank");
    }
};

LocalGroup* commLocalCreate(int nRanks) {
    PFEM_REQUIRE(nRanks >= 1 && nRanks <= 64, PFEM_ERR_INVALID, "comm_local_create: 1 <= nRanks <= 64");
    LocalGroup* g = new LocalGroup;
    g->n = nRanks;
    g->pub.resize(nRanks);
    if (cudaHostAlloc(&g->bank, (size_t)nRanks * LocalGroup::SLOT_CAP * sizeof(double), cudaHostAllocPortable | cudaHostAllocMapped) !=
        cudaSuccess) {
        cudaGetLastError();
        delete g;
        pfemThrow(PFEM_ERR_CUDA, "comm_local_create: cannot allocate the pinned reduction bank");
    }
    return g;
}
void commLocalDestroy(LocalGroup* g) {
    if (!g) return;
    if (g->bank) cudaFreeHost(g->bank);
    delete g;
}
void commInitLocal(pfem_ctx* c, LocalGroup* g, int rank) {
    PFEM_REQUIRE(g && rank >= 0 && rank < g->n, PFEM_ERR_INVALID, "comm_init_local: bad arguments");
    PFEM_REQUIRE(!c->comm && !c->local, PFEM_ERR_STATE, "comm_init_local: communicator already initialised");
    CUDA_CHECK(cudaEventCreateWithFlags(&c->evPacked, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&c->evCopied, cudaEventDisableTiming));
    {
        std::lock_guard<std::mutex> lk(g->m);
        PFEM_REQUIRE(!g->pub[rank].ctx, PFEM_ERR_STATE, "comm_init_local: rank already taken");
        g->pub[rank].ctx = c;
    }
    c->local = g;
    c->nRanks = g->n;
    c->rank = rank;
}
void commAbort(pfem_ctx* c) {
    if (!c || !c->local) return;
    std::lock_guard<std::mutex> lk(c->local->m);
    c->local->aborted = true;
    c->local->cv.notify_all();
}

void commUniqueId(void* id128) {
    PFEM_REQUIRE(id128, PFEM_ERR_INVALID, "comm_unique_id: null");
    NcclApi* api = loadNccl();
    ncclUniqueId id;
    NCCL_CHECK(api, api->GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
}

void commInit(pfem_ctx* c, int nRanks, int rank, const void* id128) {
    PFEM_REQUIRE(nRanks >= 1 && rank >= 0 && rank < nRanks && id128, PFEM_ERR_INVALID, "comm_init: bad arguments");
    PFEM_REQUIRE(!c->comm && !c->local, PFEM_ERR_STATE, "comm_init: communicator already initialised");
    NcclApi* api = loadNccl();
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t comm = nullptr;
    NCCL_CHECK(api, api->CommInitRank(&comm, nRanks, id, rank));
    c->nccl = api;
    c->comm = comm;
    c->nRanks = nRanks;
    c->rank = rank;
}

void commDestroy(pfem_ctx* c) {
    if (c->comm && c->nccl) c->nccl->CommDestroy((ncclComm_t)c->comm);
    c->comm = nullptr;
    if (c->local) {
        std::lock_guard<std::mutex> lk(c->local->m);
        c->local->pub[c->rank].ctx = nullptr;
    }
    c->local = nullptr;
    if (c->evPacked) cudaEventDestroy(c->evPacked);
    if (c->evCopied) cudaEventDestroy(c->evCopied);
    c->evPacked = c->evCopied = nullptr;
}

// Partition description of the LOCAL mesh given to pfem_set_topology: rows [0, nOwned) are computed here, nodes
// [nOwned, nNodes) are ghosts; per peer a send list (local owned ids) and a contiguous ghost range to receive into.
void commSetPartition(pfem_ctx* c, int64_t nOwned, int nPeers, const int32_t* peerRank, const int64_t* sendOffsets,
                      const int32_t* sendIdx, const int64_t* recvStart, const int64_t* recvCount) {
    PFEM_REQUIRE(c->haveTopology, PFEM_ERR_STATE, "set_partition: call pfem_set_topology (local mesh) first");
    PFEM_REQUIRE(nOwned >= 0 && nOwned <= c->nNodes && nPeers >= 0, PFEM_ERR_INVALID, "set_partition: bad sizes");
    PFEM_REQUIRE(nPeers == 0 || (peerRank && sendOffsets && recvStart && recvCount), PFEM_ERR_INVALID, "set_partition: null");
    c->nRows = (int)nOwned;
    mgInvalidate(c, true);
    HaloPlan& P = c->plan;
    P.clear();
    int64_t total = nPeers ? sendOffsets[nPeers] : 0;
    for (int p = 0; p < nPeers; ++p) {
        PFEM_REQUIRE(peerRank[p] >= 0 && peerRank[p] < c->nRanks && peerRank[p] != c->rank, PFEM_ERR_INVALID,
                     "set_partition: bad peer rank");
        PFEM_REQUIRE(recvStart[p] >= nOwned && recvStart[p] + recvCount[p] <= c->nNodes, PFEM_ERR_INVALID,
                     "set_partition: receive range outside the ghost nodes");
        P.peers.push_back({peerRank[p], (int)sendOffsets[p], (int)(sendOffsets[p + 1] - sendOffsets[p]), (int)recvStart[p],
                           (int)recvCount[p]});
    }
    for (int64_t k = 0; k < total; ++k)
        PFEM_REQUIRE(sendIdx[k] >= 0 && sendIdx[k] < nOwned, PFEM_ERR_INVALID, "set_partition: send index is not an owned node");
    P.sendIdxHost.assign(sendIdx, sendIdx + total);
    commFinishPlan(c, P);
    c->haveSystem = c->haveSolution = false;
    c->tilesValid = c->orderValid = false;
}
void commFinishPlan(pfem_ctx* c, HaloPlan& P) {
    const size_t total = P.sendIdxHost.size();
    P.nSendTotal = (int)total;
    P.sendIdx.reserve(total + 4);
    if (total > 0)
        CUDA_CHECK(cudaMemcpyAsync(P.sendIdx.p, P.sendIdxHost.data(), total * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    P.sendBuf.reserve(total * 8 + 8);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

// owners -> ghosts for up to two nodal arrays of `width` values of type T per node
template <typename T> void commHaloPlanT(pfem_ctx* c, HaloPlan& P, T* arr0, T* arr1, int width) {
    if (c->nRanks <= 1) return;
    PFEM_REQUIRE(c->comm || c->local, PFEM_ERR_COMM, "halo exchange without a communicator");
    if (!c->local && P.peers.empty()) return;  // the local transport meets at barriers: every rank takes part
    PhaseScope ph(c, "Halo exchange");
    const int nArr = arr1 ? 2 : 1;
    T* arrs[2] = {arr0, arr1};
    P.sendBuf.reserve((size_t)P.nSendTotal * width * nArr + 8);  // sized in doubles: enough for either type
    T* sendBuf = reinterpret_cast<T*>(P.sendBuf.p);
    const ncclDataType_t ncclT = sizeof(T) == 8 ? ncclFloat64 : ncclFloat32;
    for (int k = 0; k < nArr; ++k) {
        if (P.nSendTotal > 0) {
            k_pack<T><<<divUp((int64_t)P.nSendTotal * width, 256), 256, 0, c->stream>>>(
                P.sendIdx.p, P.nSendTotal, width, arrs[k], sendBuf + (size_t)k * P.nSendTotal * width);
            LAUNCH_CHECK(c);
        }
    }
    if (c->local) {
        LocalGroup* g = c->local;
        CUDA_CHECK(cudaEventRecord(c->evPacked, c->stream));
        g->pub[c->rank].src = P.sendBuf.p;
        g->pub[c->rank].plan = &P;
        g->barrier();
        for (const auto& p : P.peers) {
            if (p.recvCount <= 0) continue;
            const LocalGroup::Pub& o = g->pub[p.rank];
            PFEM_REQUIRE(o.ctx && o.plan, PFEM_ERR_COMM, "local halo: peer has no plan");
            const HaloPlan::Peer* q = nullptr;
            for (const auto& e : o.plan->peers)
                if (e.rank == c->rank) q = &e;
            PFEM_REQUIRE(q && q->sendCount == p.recvCount, PFEM_ERR_COMM, "local halo: send/receive counts of a pair differ");
            CUDA_CHECK(cudaStreamWaitEvent(c->stream, o.ctx->evPacked, 0));
            for (int k = 0; k < nArr; ++k)
                CUDA_CHECK(cudaMemcpyAsync(arrs[k] + (size_t)p.recvStart * width,
                                           reinterpret_cast<const T*>(o.src) + ((size_t)k * o.plan->nSendTotal + q->sendOff) * width,
                                           (size_t)p.recvCount * width * sizeof(T), cudaMemcpyDefault, c->stream));
        }
        CUDA_CHECK(cudaEventRecord(c->evCopied, c->stream));
        g->barrier();
        for (const auto& p : P.peers)  // my next pack must not overwrite the buffer a peer is still copying from
            if (p.sendCount > 0) CUDA_CHECK(cudaStreamWaitEvent(c->stream, g->pub[p.rank].ctx->evCopied, 0));
        return;
    }
    NcclApi* api = c->nccl;
    NCCL_CHECK(api, api->GroupStart());
    for (int k = 0; k < nArr; ++k) {
        for (const auto& p : P.peers) {
            if (p.sendCount > 0)
                NCCL_CHECK(api, api->Send(sendBuf + ((size_t)k * P.nSendTotal + p.sendOff) * width,
                                          (size_t)p.sendCount * width, ncclT, p.rank, (ncclComm_t)c->comm, c->stream));
            if (p.recvCount > 0)
                NCCL_CHECK(api, api->Recv(arrs[k] + (size_t)p.recvStart * width, (size_t)p.recvCount * width, ncclT,
                                          p.rank, (ncclComm_t)c->comm, c->stream));
        }
    }
    NCCL_CHECK(api, api->GroupEnd());
}
template void commHaloPlanT<double>(pfem_ctx*, HaloPlan&, double*, double*, int);
template void commHaloPlanT<float>(pfem_ctx*, HaloPlan&, float*, float*, int);
void commHaloPlan(pfem_ctx* c, HaloPlan& P, double* arr0, double* arr1, int width) { commHaloPlanT<double>(c, P, arr0, arr1, width); }
void commHalo(pfem_ctx* c, double* arr0, double* arr1, int width) { commHaloPlan(c, c->plan, arr0, arr1, width); }

static void localAllReduce(pfem_ctx* c, double* buf, int count, bool isMin) {
    LocalGroup* g = c->local;
    PFEM_REQUIRE(count <= LocalGroup::SLOT_CAP, PFEM_ERR_INVALID, "local all-reduce: too many values");
    // the previous all-reduce's readers are done with the bank (their events were recorded before its second barrier)
    for (int r = 0; r < g->n; ++r)
        if (g->pub[r].ctx && g->pub[r].ctx != c) CUDA_CHECK(cudaStreamWaitEvent(c->stream, g->pub[r].ctx->evCopied, 0));
    k_slots_store<<<1, 64, 0, c->stream>>>(buf, count, g->bank + (size_t)c->rank * LocalGroup::SLOT_CAP);
    LAUNCH_CHECK(c);
    CUDA_CHECK(cudaEventRecord(c->evPacked, c->stream));
    g->barrier();
    for (int r = 0; r < g->n; ++r)
        if (g->pub[r].ctx && g->pub[r].ctx != c) CUDA_CHECK(cudaStreamWaitEvent(c->stream, g->pub[r].ctx->evPacked, 0));
    k_slots_reduce<<<1, 64, 0, c->stream>>>(g->bank, LocalGroup::SLOT_CAP, g->n, count, isMin ? 1 : 0, buf);
    LAUNCH_CHECK(c);
    CUDA_CHECK(cudaEventRecord(c->evCopied, c->stream));
    g->barrier();
}

void commAllReduceMin(pfem_ctx* c, double* devScalar) {
    if (c->nRanks <= 1) return;
    if (c->local) return localAllReduce(c, devScalar, 1, true);
    PFEM_REQUIRE(c->comm, PFEM_ERR_COMM, "no communicator");
    NCCL_CHECK(c->nccl, c->nccl->AllReduce(devScalar, devScalar, 1, ncclFloat64, ncclMin, (ncclComm_t)c->comm, c->stream));
}

void commAllReduceSum(pfem_ctx* c, double* buf, int count) {
    if (c->nRanks <= 1) return;
    if (c->local) return localAllReduce(c, buf, count, false);
    PFEM_REQUIRE(c->comm, PFEM_ERR_COMM, "no communicator");
    NCCL_CHECK(c->nccl, c->nccl->AllReduce(buf, buf, (size_t)count, ncclFloat64, ncclSum, (ncclComm_t)c->comm, c->stream));
}

void commAllGatherBytes(pfem_ctx* c, const void* srcV, void* dstV, const std::vector<int64_t>& counts, const std::vector<int64_t>& displs) {
    PFEM_REQUIRE((int)counts.size() == c->nRanks && (int)displs.size() == c->nRanks, PFEM_ERR_INVALID, "all-gather: bad counts");
    const char* src = static_cast<const char*>(srcV);
    char* dst = static_cast<char*>(dstV);
    char* mine = dst + displs[c->rank];
    if (src != mine && counts[c->rank] > 0)
        CUDA_CHECK(cudaMemcpyAsync(mine, src, (size_t)counts[c->rank], cudaMemcpyDeviceToDevice, c->stream));
    if (c->nRanks <= 1) return;
    PhaseScope ph(c, "All-gather");
    if (c->local) {
        LocalGroup* g = c->local;
        CUDA_CHECK(cudaEventRecord(c->evPacked, c->stream));
        g->pub[c->rank].src = reinterpret_cast<const double*>(mine);
        g->barrier();
        for (int r = 0; r < g->n; ++r) {
            if (r == c->rank || counts[r] <= 0) continue;
            PFEM_REQUIRE(g->pub[r].ctx, PFEM_ERR_COMM, "local all-gather: missing rank");
            CUDA_CHECK(cudaStreamWaitEvent(c->stream, g->pub[r].ctx->evPacked, 0));
            CUDA_CHECK(cudaMemcpyAsync(dst + displs[r], g->pub[r].src, (size_t)counts[r], cudaMemcpyDefault, c->stream));
        }
        CUDA_CHECK(cudaEventRecord(c->evCopied, c->stream));
        g->barrier();
        for (int r = 0; r < g->n; ++r)
            if (r != c->rank && g->pub[r].ctx) CUDA_CHECK(cudaStreamWaitEvent(c->stream, g->pub[r].ctx->evCopied, 0));
        return;
    }
    PFEM_REQUIRE(c->comm, PFEM_ERR_COMM, "no communicator");
    NcclApi* api = c->nccl;
    NCCL_CHECK(api, api->GroupStart());
    for (int r = 0; r < c->nRanks; ++r) {
        if (r == c->rank) continue;
        if (counts[c->rank] > 0)
            NCCL_CHECK(api, api->Send(mine, (size_t)counts[c->rank], ncclInt8, r, (ncclComm_t)c->comm, c->stream));
        if (counts[r] > 0)
            NCCL_CHECK(api, api->Recv(dst + displs[r], (size_t)counts[r], ncclInt8, r, (ncclComm_t)c->comm, c->stream));
    }
    NCCL_CHECK(api, api->GroupEnd());
}
void commAllGatherV(pfem_ctx* c, const double* src, double* dst, const std::vector<int64_t>& counts, const std::vector<int64_t>& displs) {
    std::vector<int64_t> cb(counts), db(displs);
    for (auto& v : cb) v *= (int64_t)sizeof(double);
    for (auto& v : db) v *= (int64_t)sizeof(double);
    commAllGatherBytes(c, src, dst, cb, db);
}

void commAllGatherHost(pfem_ctx* c, const double* mine, int n, double* all) {
    if (c->nRanks <= 1) {
        memcpy(all, mine, (size_t)n * sizeof(double));
        return;
    }
    c->commScratch.reserve((size_t)c->nRanks * n + 8);
    std::vector<int64_t> counts(c->nRanks, n), displs(c->nRanks);
    for (int r = 0; r < c->nRanks; ++r) displs[r] = (int64_t)r * n;
    CUDA_CHECK(cudaMemcpyAsync(c->commScratch.p + displs[c->rank], mine, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    commAllGatherV(c, c->commScratch.p + displs[c->rank], c->commScratch.p, counts, displs);
    CUDA_CHECK(cudaMemcpyAsync(all, c->commScratch.p, (size_t)c->nRanks * n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
}
