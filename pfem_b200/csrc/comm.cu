// comm.cu -- multi-GPU plumbing (one process per GPU): NCCL communicator, halo exchange of nodal records, all-reduces.
//
// Sharding scheme (pfem_b200/partition.py): RCB over node coordinates; a rank owns its nodes'
// rows and holds one ghost-element layer, so the node-gather kernels need NO reduction of partial sums -- after each
// element pass the owners send the updated nodal records (32-byte X4/V4/A4 records or (dim+1)-double Krylov entries)
// of their interface nodes; ghosts of one owner are contiguous, so receives land directly in the nodal arrays.
// Messages are small (tens of KB .. ~1 MB): latency-bound, one grouped ncclSend/ncclRecv per exchange over NVLink.
// NCCL is dlopen'ed so that the library loads (and exports its symbols) on a box without NCCL/GPU; a missing NCCL is
// a loud PFEM_ERR_COMM, never a fallback.
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>

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
    static NcclApi api;
    if (api.handle) return &api;
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
    return &api;
}
#define NCCL_CHECK(api, expr)                                                                             \
    do {                                                                                                  \
        ncclResult_t _r = (expr);                                                                         \
        if (_r != ncclSuccess) pfemThrow(PFEM_ERR_COMM, std::string(#expr) + " -> " + (api)->GetErrorString(_r)); \
    } while (0)

namespace {
// buf[(e*width + c)] = arr[idx[e]*width + c]
__global__ void k_pack(const int* __restrict__ idx, int nSend, int width, const double* __restrict__ arr,
                       double* __restrict__ buf) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nSend * width) return;
    const int e = t / width, cidx = t % width;
    buf[t] = arr[(size_t)idx[e] * width + cidx];
}
}  // namespace

void commUniqueId(void* id128) {
    PFEM_REQUIRE(id128, PFEM_ERR_INVALID, "comm_unique_id: null");
    NcclApi* api = loadNccl();
    ncclUniqueId id;
    NCCL_CHECK(api, api->GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
}

void commInit(pfem_ctx* c, int nRanks, int rank, const void* id128) {
    PFEM_REQUIRE(nRanks >= 1 && rank >= 0 && rank < nRanks && id128, PFEM_ERR_INVALID, "comm_init: bad arguments");
    PFEM_REQUIRE(!c->comm, PFEM_ERR_STATE, "comm_init: communicator already initialised");
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
    c->peers.clear();
    int64_t total = nPeers ? sendOffsets[nPeers] : 0;
    for (int p = 0; p < nPeers; ++p) {
        PFEM_REQUIRE(peerRank[p] >= 0 && peerRank[p] < c->nRanks && peerRank[p] != c->rank, PFEM_ERR_INVALID,
                     "set_partition: bad peer rank");
        PFEM_REQUIRE(recvStart[p] >= nOwned && recvStart[p] + recvCount[p] <= c->nNodes, PFEM_ERR_INVALID,
                     "set_partition: receive range outside the ghost nodes");
        c->peers.push_back({peerRank[p], (int)sendOffsets[p], (int)(sendOffsets[p + 1] - sendOffsets[p]), (int)recvStart[p],
                            (int)recvCount[p]});
    }
    c->nSendTotal = (int)total;
    c->sendIdx.reserve(total + 4);
    if (total > 0) {
        for (int64_t k = 0; k < total; ++k)
            PFEM_REQUIRE(sendIdx[k] >= 0 && sendIdx[k] < nOwned, PFEM_ERR_INVALID, "set_partition: send index is not an owned node");
        CUDA_CHECK(cudaMemcpyAsync(c->sendIdx.p, sendIdx, total * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    }
    c->sendBuf.reserve((size_t)total * 8 + 8);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->haveSystem = c->haveSolution = false;
}

// owners -> ghosts for up to two nodal arrays of `width` doubles per node, one NCCL group
void commHalo(pfem_ctx* c, double* arr0, double* arr1, int width) {
    if (c->nRanks <= 1 || c->peers.empty()) return;
    PFEM_REQUIRE(c->comm, PFEM_ERR_COMM, "halo exchange without a communicator");
    PhaseScope ph(c, "Halo exchange");
    NcclApi* api = c->nccl;
    const int nArr = arr1 ? 2 : 1;
    double* arrs[2] = {arr0, arr1};
    c->sendBuf.reserve((size_t)c->nSendTotal * width * nArr + 8);
    for (int k = 0; k < nArr; ++k) {
        if (c->nSendTotal > 0) {
            k_pack<<<divUp((int64_t)c->nSendTotal * width, 256), 256, 0, c->stream>>>(
                c->sendIdx.p, c->nSendTotal, width, arrs[k], c->sendBuf.p + (size_t)k * c->nSendTotal * width);
            LAUNCH_CHECK(c);
        }
    }
    NCCL_CHECK(api, api->GroupStart());
    for (int k = 0; k < nArr; ++k) {
        for (const auto& p : c->peers) {
            if (p.sendCount > 0)
                NCCL_CHECK(api, api->Send(c->sendBuf.p + ((size_t)k * c->nSendTotal + p.sendOff) * width,
                                          (size_t)p.sendCount * width, ncclFloat64, p.rank, (ncclComm_t)c->comm, c->stream));
            if (p.recvCount > 0)
                NCCL_CHECK(api, api->Recv(arrs[k] + (size_t)p.recvStart * width, (size_t)p.recvCount * width, ncclFloat64,
                                          p.rank, (ncclComm_t)c->comm, c->stream));
        }
    }
    NCCL_CHECK(api, api->GroupEnd());
}

void commAllReduceMin(pfem_ctx* c, double* devScalar) {
    PFEM_REQUIRE(c->comm, PFEM_ERR_COMM, "no communicator");
    NCCL_CHECK(c->nccl, c->nccl->AllReduce(devScalar, devScalar, 1, ncclFloat64, ncclMin, (ncclComm_t)c->comm, c->stream));
}

void commAllReduceSum(pfem_ctx* c, double* buf, int count) {
    if (c->nRanks <= 1) return;
    PFEM_REQUIRE(c->comm, PFEM_ERR_COMM, "no communicator");
    NCCL_CHECK(c->nccl, c->nccl->AllReduce(buf, buf, (size_t)count, ncclFloat64, ncclSum, (ncclComm_t)c->comm, c->stream));
}
