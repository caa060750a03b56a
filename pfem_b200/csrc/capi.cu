// capi.cu -- the extern "C" surface declared in include/pfem_b200.h.  Every entry point converts failures into a
// status code + pfem_last_error() string; nothing throws across the ABI and nothing falls back to the CPU.
#include <cmath>
#include <cstring>

#include "common.cuh"

static thread_local std::string g_createError;

void pfemFlushPhases(pfem_ctx* c) {
    if (c->pending.empty()) return;
    cudaStreamSynchronize(c->stream);
    for (auto& p : c->pending) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess) {
            auto& acc = c->phases[p.name];
            acc.ms += ms;
            acc.calls += 1;
        }
        c->eventPool.push_back(p.e0);
        c->eventPool.push_back(p.e1);
    }
    c->pending.clear();
}

#define API_BEGIN(ctx)                                   \
    if (!(ctx)) return PFEM_ERR_INVALID;                 \
    try {                                                \
        cudaSetDevice((ctx)->device);                    \
        (ctx)->cflFresh = false; /* any call may change positions/states; pfem_wc_step sets it again */
#define API_END(ctx)                                     \
    }                                                    \
    catch (const PfemFail& f) {                          \
        (ctx)->err = f.msg;                              \
        cudaGetLastError();                              \
        if (f.code < 0) commAbort(ctx);                  \
        return f.code;                                   \
    }                                                    \
    catch (const std::exception& e) {                    \
        (ctx)->err = e.what();                           \
        return PFEM_ERR_INVALID;                         \
    }                                                    \
    catch (...) {                                        \
        (ctx)->err = "unknown exception";                \
        return PFEM_ERR_INVALID;                         \
    }                                                    \
    return PFEM_OK;
// the same handlers for entry points that return a status of their own: nothing may unwind through extern "C"
#define API_CATCH(ctx)                                   \
    catch (const PfemFail& f) {                          \
        (ctx)->err = f.msg;                              \
        cudaGetLastError();                              \
        if (f.code < 0) commAbort(ctx);                  \
        return f.code;                                   \
    }                                                    \
    catch (const std::exception& e) {                    \
        (ctx)->err = e.what();                           \
        return PFEM_ERR_INVALID;                         \
    }                                                    \
    catch (...) {                                        \
        (ctx)->err = "unknown exception";                \
        return PFEM_ERR_INVALID;                         \
    }

extern "C" {

int pfem_abi_version(void) { return PFEM_ABI_VERSION; }

int pfem_create(pfem_ctx** out, int dim, int device) {
    if (!out) return PFEM_ERR_INVALID;
    *out = nullptr;
    if (dim != 2 && dim != 3) return PFEM_ERR_INVALID;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0 || device < 0 || device >= count) {
        g_createError = std::string("pfem_create: no usable CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback";
        cudaGetLastError();
        return PFEM_ERR_CUDA;
    }
    pfem_ctx* c = new pfem_ctx;
    c->dim = dim;
    c->device = device;
    try {
        CUDA_CHECK(cudaSetDevice(device));
        cudaDeviceProp prop;
        CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
        c->smCount = prop.multiProcessorCount;
        CUDA_CHECK(cudaStreamCreateWithFlags(&c->ownStream, cudaStreamNonBlocking));
        c->stream = c->ownStream;
    } catch (const PfemFail& f) {
        g_createError = f.msg;
        delete c;
        return f.code;
    } catch (...) {
        g_createError = "pfem_create: unexpected exception";
        delete c;
        return PFEM_ERR_CUDA;
    }
    *out = c;
    return PFEM_OK;
}

int pfem_destroy(pfem_ctx* c) {
    if (!c) return PFEM_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    commDestroy(c);
    mgDestroy(c);
    for (auto& p : c->pending) {
        cudaEventDestroy(p.e0);
        cudaEventDestroy(p.e1);
    }
    for (auto e : c->eventPool) cudaEventDestroy(e);
    if (c->hScal) cudaFreeHost(c->hScal);
    if (c->commStream) cudaStreamDestroy(c->commStream);
    if (c->evTile) cudaEventDestroy(c->evTile);
    if (c->evHalo) cudaEventDestroy(c->evHalo);
    cudaStream_t s = c->ownStream;
    delete c;
    if (s) cudaStreamDestroy(s);
    return PFEM_OK;
}

const char* pfem_last_error(const pfem_ctx* c) { return c ? c->err.c_str() : g_createError.c_str(); }

int pfem_set_stream(pfem_ctx* c, void* s) {
    API_BEGIN(c)
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->stream = s ? (cudaStream_t)s : c->ownStream;
    API_END(c)
}

int pfem_get_info(const pfem_ctx* cc, pfem_info* info) {
    if (!cc || !info) return PFEM_ERR_INVALID;
    pfem_ctx* c = const_cast<pfem_ctx*>(cc);
    info->dim = c->dim, info->device = c->device, info->nRanks = c->nRanks, info->rank = c->rank;
    info->nNodes = c->nNodes, info->nElems = c->nElems, info->nDof = (int64_t)c->nNodes * (c->dim + 1);
    info->nnzBlocks = c->nBlocks, info->nnzReference = c->nnzReference, info->deviceBytes = (int64_t)c->deviceBytes;
    info->maxElemsPerNode = c->maxE, info->maxNeighbours = c->maxNb;
    return PFEM_OK;
}

int pfem_set_topology(pfem_ctx* c, int64_t nNodes, int64_t nElems, const uint64_t* elemNodes, const uint8_t* flags) {
    API_BEGIN(c)
    PFEM_REQUIRE(elemNodes || nElems == 0, PFEM_ERR_INVALID, "set_topology: elemNodes is null");
    PFEM_REQUIRE(flags, PFEM_ERR_INVALID, "set_topology: nodeFlags is null");
    topoBuild(c, nNodes, nElems, elemNodes, flags);
    API_END(c)
}
int pfem_set_positions(pfem_ctx* c, const double* x) {
    API_BEGIN(c)
    fieldsSetPositions(c, x);
    API_END(c)
}
int pfem_get_positions(pfem_ctx* c, double* x) {
    API_BEGIN(c)
    fieldsGetPositions(c, x);
    API_END(c)
}
int pfem_snapshot_positions(pfem_ctx* c) {
    API_BEGIN(c)
    PhaseScope ph(c, "Save/restore nodelist");
    fieldsSnapshot(c);
    API_END(c)
}
int pfem_restore_positions(pfem_ctx* c) {
    API_BEGIN(c)
    PhaseScope ph(c, "Save/restore nodelist");
    fieldsRestore(c);
    API_END(c)
}
int pfem_move_positions(pfem_ctx* c, const double* delta, int fromSnapshot) {
    API_BEGIN(c)
    fieldsMove(c, delta, fromSnapshot);
    API_END(c)
}
int pfem_set_states(pfem_ctx* c, int first, int count, const double* q) {
    API_BEGIN(c)
    fieldsSetStates(c, first, count, q);
    API_END(c)
}
int pfem_get_states(pfem_ctx* c, int first, int count, double* q) {
    API_BEGIN(c)
    fieldsGetStates(c, first, count, q);
    API_END(c)
}
int pfem_set_dirichlet(pfem_ctx* c, const uint8_t* mask, const double* values) {
    API_BEGIN(c)
    fieldsSetDirichlet(c, mask, values);
    API_END(c)
}

int pfem_set_facets(pfem_ctx* c, int64_t nFacets, const uint64_t* facetNodes, const uint64_t* outNode, const uint64_t* elemIndex) {
    API_BEGIN(c)
    facetsSet(c, nFacets, facetNodes, outNode, elemIndex);
    API_END(c)
}
int pfem_set_surface_tension(pfem_ctx* c, double gamma) {
    API_BEGIN(c)
    PFEM_REQUIRE(gamma >= 0 && gamma == gamma, PFEM_ERR_INVALID, "set_surface_tension: gamma must be >= 0");
    c->gammaST = gamma;
    API_END(c)
}

int pfem_set_thermal(pfem_ctx* c, const pfem_thermal_params* t) {
    API_BEGIN(c)
    if (!t) {
        c->thermalOn = false;
    } else {
        PFEM_REQUIRE(t->cv > 0 && t->k >= 0, PFEM_ERR_INVALID, "set_thermal: cv must be positive, k non-negative");
        c->thermalOn = true;
        c->thK = t->k, c->thCv = t->cv, c->thAlpha = t->alpha, c->thTr = t->Tr;
    }
    API_END(c)
}
int pfem_set_bingham(pfem_ctx* c, int on, double tau0, double mReg) {
    API_BEGIN(c)
    PFEM_REQUIRE(!on || (tau0 >= 0 && mReg >= 0), PFEM_ERR_INVALID, "set_bingham: tau0 and mReg must be non-negative");
    c->binghamOn = on != 0;
    c->binghamTau0 = tau0, c->binghamM = mReg;
    API_END(c)
}
int pfem_set_temperature(pfem_ctx* c, const double* T) {
    API_BEGIN(c)
    thermalSetTemperature(c, T);
    API_END(c)
}
int pfem_get_temperature(pfem_ctx* c, double* T) {
    API_BEGIN(c)
    thermalGetTemperature(c, T);
    API_END(c)
}
int pfem_set_temperature_bc(pfem_ctx* c, const uint8_t* mask, const double* values) {
    API_BEGIN(c)
    thermalSetBc(c, mask, values);
    API_END(c)
}
int pfem_heat_assemble(pfem_ctx* c, double rho, double cv, double k, double dt, const double* thetaPrev) {
    API_BEGIN(c)
    heatAssemble(c, rho, cv, k, dt, thetaPrev);
    API_END(c)
}
int pfem_heat_solve(pfem_ctx* c, double relTol, int maxIter, double* T, int* iters, double* relRes) {
    if (!c) return PFEM_ERR_INVALID;
    try {
        cudaSetDevice(c->device);
        c->cflFresh = false;
        return heatSolve(c, relTol, maxIter, T, iters, relRes);
    } API_CATCH(c)
}
int pfem_heat_export_csc(pfem_ctx* c, int64_t* nnz, int32_t* colPtr, int32_t* rowIdx, double* val, double* b) {
    API_BEGIN(c)
    heatExport(c, nnz, colPtr, rowIdx, val, b);
    API_END(c)
}

int pfem_fs_assemble_vapp(pfem_ctx* c, const pfem_pspg_params* p, double gammaFS, const double* qPrev) {
    API_BEGIN(c)
    PFEM_REQUIRE(p, PFEM_ERR_INVALID, "fs_assemble_vapp: params is null");
    fsAssembleVapp(c, *p, gammaFS, qPrev);
    API_END(c)
}
int pfem_fs_assemble_pcorr(pfem_ctx* c, double rho, double dt, double gammaFS, const double* vTilde, const double* pPrev) {
    API_BEGIN(c)
    fsAssemblePcorr(c, rho, dt, gammaFS, vTilde, pPrev);
    API_END(c)
}
int pfem_fs_assemble_vcorr(pfem_ctx* c, double rho, double dt, const double* deltaP) {
    API_BEGIN(c)
    fsAssembleVcorr(c, rho, dt, deltaP);
    API_END(c)
}
int pfem_fs_get_rhs(pfem_ctx* c, double* b) {
    API_BEGIN(c)
    fsGetRhs(c, b);
    API_END(c)
}
int pfem_fs_solve(pfem_ctx* c, double relTol, int maxIter, double* x, int* iters, double* relRes) {
    if (!c) return PFEM_ERR_INVALID;
    try {
        cudaSetDevice(c->device);
        c->cflFresh = false;
        return fsSolve(c, relTol, maxIter, x, iters, relRes);
    } API_CATCH(c)
}

int pfem_pspg_set_qprev(pfem_ctx* c, const double* qPrev) {
    API_BEGIN(c)
    fieldsSetQprev(c, qPrev);
    API_END(c)
}
int pfem_pspg_assemble_resident(pfem_ctx* c, const pfem_pspg_params* p) {
    API_BEGIN(c)
    PFEM_REQUIRE(p, PFEM_ERR_INVALID, "pspg_assemble: params is null");
    pspgAssemble(c, *p);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    API_END(c)
}
int pfem_pspg_assemble(pfem_ctx* c, const pfem_pspg_params* p, const double* qPrev) {
    API_BEGIN(c)
    PFEM_REQUIRE(p, PFEM_ERR_INVALID, "pspg_assemble: params is null");
    fieldsSetQprev(c, qPrev);
    pspgAssemble(c, *p);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    API_END(c)
}
int pfem_pspg_solve(pfem_ctx* c, double relTol, int maxIter, double* q, int* iters, double* relRes) {
    if (!c) return PFEM_ERR_INVALID;
    int status = PFEM_OK;
    try {
        cudaSetDevice(c->device);
        status = krylovSolve(c, relTol, maxIter, iters, relRes, false);
        if (q) krylovFetchSolution(c, q);
    } API_CATCH(c)
    return status;
}
int pfem_pspg_set_preconditioner(pfem_ctx* c, int kind, int sweeps, double damping) {
    API_BEGIN(c)
    PFEM_REQUIRE(kind >= PFEM_PRECOND_AUTO && kind <= PFEM_PRECOND_MG, PFEM_ERR_INVALID, "set_preconditioner: unknown kind");
    PFEM_REQUIRE(sweeps <= 16 && damping <= 8.0, PFEM_ERR_INVALID, "set_preconditioner: sweeps <= 16, damping <= 8");
    c->precondKind = kind;
    c->mgSweeps = sweeps > 0 ? sweeps : 0;
    c->mgDamping = damping > 0 ? damping : 0.0;
    API_END(c)
}
int pfem_pspg_get_preconditioner(pfem_ctx* c, int* kindUsed, int* levelsOut) {
    API_BEGIN(c)
    if (kindUsed) *kindUsed = c->lastPrecond;
    if (levelsOut) *levelsOut = c->lastPrecond == PFEM_PRECOND_MG ? mgLevelCount(c) : 1;
    API_END(c)
}
int pfem_pspg_residual(pfem_ctx* c, const double* q, double* res) {
    API_BEGIN(c)
    PFEM_REQUIRE(res, PFEM_ERR_INVALID, "pspg_residual: null");
    double* xi = c->kx.p;
    if (q) {
        c->ks.reserve((size_t)c->nNodes * (c->dim + 1));
        krylovLoadVector(c, q, c->ks.p);
        xi = c->ks.p;
    } else
        PFEM_REQUIRE(c->haveSolution, PFEM_ERR_STATE, "pspg_residual: no device solution");
    *res = krylovResidualNorm(c, xi);
    API_END(c)
}
int pfem_pspg_picard_iter(pfem_ctx* c, const pfem_pspg_params* p, const double* qPrev, double relTol, int maxIter,
                          double* q, double* resAxf, int* iters) {
    if (!c) return PFEM_ERR_INVALID;
    int status = PFEM_OK;
    try {
        cudaSetDevice(c->device);
        c->cflFresh = false;
        PFEM_REQUIRE(p, PFEM_ERR_INVALID, "picard_iter: params is null");
        PFEM_REQUIRE(c->haveSnapshot, PFEM_ERR_STATE, "picard_iter: call pfem_snapshot_positions first (m_prepare)");
        if (qPrev) fieldsSetQprev(c, qPrev);
        double rr = 0;
        status = krylovSolve(c, relTol, maxIter, iters, &rr, /*warmStart=*/true);  // m_solver (PSPG.inl:281-290)
        if (status == PFEM_OK) {
            pspgPicardUpdate(c, p->dt);  // PSPG.inl:293-295
            pspgAssemble(c, *p);         // PSPG.inl:298-300
            c->haveSolution = true;      // kx still holds q^{k+1}
            const double r = krylovResidualNorm(c, c->kx.p);  // PSPG.inl:368
            if (resAxf) *resAxf = r;
            if (r != r) status = PFEM_NAN;
        }
        if (q) krylovFetchSolution(c, q);
    } API_CATCH(c)
    return status;
}
int pfem_pspg_export_csc(pfem_ctx* c, int64_t* nnz, int32_t* colPtr, int32_t* rowIdx, double* val, double* b) {
    API_BEGIN(c)
    pspgExportCsc(c, nnz, colPtr, rowIdx, val, b);
    API_END(c)
}
int pfem_pspg_matvec(pfem_ctx* c, const double* x, double* y) {
    API_BEGIN(c)
    PFEM_REQUIRE(x && y, PFEM_ERR_INVALID, "matvec: null");
    const size_t n = (size_t)c->nNodes * (c->dim + 1);
    c->ks.reserve(n);
    c->kt.reserve(n);
    krylovLoadVector(c, x, c->ks.p);
    krylovMatvec(c, c->ks.p, c->kt.p);
    krylovStoreVector(c, c->kt.p, y);
    API_END(c)
}

int pfem_wc_set_variant(pfem_ctx* c, int variant) {
    API_BEGIN(c)
    PFEM_REQUIRE(variant == 0 || variant == 6 || variant == 7 || variant == 11 || variant == 12 || variant == 13, PFEM_ERR_INVALID,
                 "wc_set_variant: 0 (by size), 6 (gather), 7 (staged gather), 11 (two-pass), 12 (two-pass continuity + gather momentum), "
                 "13 (tiles: element records in shared memory)");
    c->wcVariant = variant;
    API_END(c)
}
int pfem_wc_step(pfem_ctx* c, const pfem_wc_params* p, double dt) {
    API_BEGIN(c)
    PFEM_REQUIRE(p, PFEM_ERR_INVALID, "wc_step: params is null");
    wcStep(c, *p, dt);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    API_END(c)
}
int pfem_wc_next_dt(pfem_ctx* c, const pfem_wc_params* p, double securityCoeff, double maxDT, double* dt) {
    if (!c) return PFEM_ERR_INVALID;
    try {
        cudaSetDevice(c->device);
        PFEM_REQUIRE(p, PFEM_ERR_INVALID, "wc_next_dt: params is null");
        return wcNextDt(c, *p, securityCoeff, maxDT, dt);
    } API_CATCH(c)
}

int pfem_wc_run(pfem_ctx* c, const pfem_wc_params* p, int nSteps, double securityCoeff, double maxDT, double* dt, double* elapsed) {
    if (!c) return PFEM_ERR_INVALID;
    try {
        cudaSetDevice(c->device);
        c->cflFresh = false;
        PFEM_REQUIRE(p, PFEM_ERR_INVALID, "wc_run: params is null");
        return wcRun(c, *p, nSteps, securityCoeff, maxDT, dt, elapsed);
    } API_CATCH(c)
}

int pfem_comm_unique_id(void* id128) {
    try {
        commUniqueId(id128);
    } catch (const PfemFail& f) {
        g_createError = f.msg;
        return f.code;
    } catch (const std::exception& e) {
        g_createError = e.what();
        return PFEM_ERR_COMM;
    } catch (...) {
        g_createError = "unknown exception";
        return PFEM_ERR_COMM;
    }
    return PFEM_OK;
}
int pfem_comm_init(pfem_ctx* c, int nRanks, int rank, const void* id128) {
    API_BEGIN(c)
    commInit(c, nRanks, rank, id128);
    API_END(c)
}
int pfem_comm_local_create(int nRanks, void** group) {
    if (!group) return PFEM_ERR_INVALID;
    *group = nullptr;
    try {
        *group = commLocalCreate(nRanks);
    } catch (const PfemFail& f) {
        g_createError = f.msg;
        return f.code;
    } catch (...) {
        g_createError = "comm_local_create: unexpected exception";
        return PFEM_ERR_COMM;
    }
    return PFEM_OK;
}
int pfem_comm_local_destroy(void* group) {
    commLocalDestroy(static_cast<LocalGroup*>(group));
    return PFEM_OK;
}
int pfem_comm_abort(pfem_ctx* c) {
    if (!c) return PFEM_ERR_INVALID;
    commAbort(c);
    return PFEM_OK;
}
int pfem_comm_init_local(pfem_ctx* c, void* group, int rank) {
    API_BEGIN(c)
    commInitLocal(c, static_cast<LocalGroup*>(group), rank);
    API_END(c)
}

int pfem_set_partition(pfem_ctx* c, int64_t nOwned, int nPeers, const int32_t* peerRank, const int64_t* sendOffsets,
                       const int32_t* sendIdx, const int64_t* recvStart, const int64_t* recvCount) {
    API_BEGIN(c)
    commSetPartition(c, nOwned, nPeers, peerRank, sendOffsets, sendIdx, recvStart, recvCount);
    API_END(c)
}

int pfem_profile_enable(pfem_ctx* c, int on) {
    API_BEGIN(c)
    pfemFlushPhases(c);
    c->profiling = on != 0;
    c->profileDetail = on == 2;
    API_END(c)
}
int pfem_profile_reset(pfem_ctx* c) {
    API_BEGIN(c)
    pfemFlushPhases(c);
    c->phases.clear();
    API_END(c)
}
int pfem_profile_get(pfem_ctx* c, const char* phase, double* ms, int64_t* calls) {
    API_BEGIN(c)
    PFEM_REQUIRE(phase, PFEM_ERR_INVALID, "profile_get: null phase");
    pfemFlushPhases(c);
    auto it = c->phases.find(phase);
    if (ms) *ms = it == c->phases.end() ? 0.0 : it->second.ms;
    if (calls) *calls = it == c->phases.end() ? 0 : it->second.calls;
    API_END(c)
}
int pfem_launch_count(const pfem_ctx* c, int64_t* launches) {
    if (!c || !launches) return PFEM_ERR_INVALID;
    *launches = c->launches;
    return PFEM_OK;
}

}  // extern "C"
