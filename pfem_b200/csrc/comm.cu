// comm.cu -- NCCL plumbing for the multi-GPU (one process per GPU) path.  NCCL is dlopen'ed so that the library loads
// (and exports its symbols) on a box without NCCL/GPU; a missing NCCL is a loud PFEM_ERR_COMM, never a fallback.
#include <dlfcn.h>

#include <cstdlib>

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

void commAllReduceMin(pfem_ctx* c, double* devScalar) {
    PFEM_REQUIRE(c->comm, PFEM_ERR_COMM, "no communicator");
    NCCL_CHECK(c->nccl, c->nccl->AllReduce(devScalar, devScalar, 1, ncclFloat64, ncclMin, (ncclComm_t)c->comm, c->stream));
}

void commAllReduceSumInterface(pfem_ctx* c, double* buf, int count) {
    PFEM_REQUIRE(c->comm, PFEM_ERR_COMM, "no communicator");
    NCCL_CHECK(c->nccl, c->nccl->AllReduce(buf, buf, (size_t)count, ncclFloat64, ncclSum, (ncclComm_t)c->comm, c->stream));
}
