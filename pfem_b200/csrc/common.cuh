// common.cuh -- context, device buffers, launch/phase bookkeeping shared by the pfem_b200 translation units.
// B200 (sm_100a) only; fp64 throughout.  Nothing in this library runs the physics on the CPU.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "../../include/pfem_b200.h"

// dynamic shared memory opt-in: always the device maximum (227 KB) -- the attribute is per-function process state, and rank
// threads of one process (local communicator) would otherwise race on per-launch values
#define PFEM_SMEM_OPTIN (225 * 1024)  /* the 227 KB per-block limit minus room for the kernels' static shared memory */
#define PFEM_MAX_STATES 8  // 2*dim+2 for the weakly-compressible problem in 3-D

// ---- error plumbing: every API entry point is `try { ... } catch (PfemFail&)`, nothing escapes the C ABI ----------
struct PfemFail {
    int code;
    std::string msg;
};
[[noreturn]] inline void pfemThrow(int code, const std::string& m) { throw PfemFail{code, m}; }
#define CUDA_CHECK(expr)                                                                                        \
    do {                                                                                                        \
        cudaError_t _e = (expr);                                                                                \
        if (_e != cudaSuccess)                                                                                  \
            pfemThrow(PFEM_ERR_CUDA, std::string(#expr) + " -> " + cudaGetErrorString(_e) + " @" + __FILE__ + ":" + \
                                         std::to_string(__LINE__));                                             \
    } while (0)
#define PFEM_REQUIRE(cond, code, msg) \
    do {                              \
        if (!(cond)) pfemThrow((code), (msg)); \
    } while (0)

// ---- grow-only device buffer (remeshing every step must not cudaMalloc every step) --------------------------------
template <class T> struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;  // elements
    size_t* accounting = nullptr;
    void reserve(size_t n) {
        if (n <= cap) return;
        size_t want = n + n / 8 + 64;
        release();
        CUDA_CHECK(cudaMalloc(&p, want * sizeof(T)));
        cap = want;
        if (accounting) *accounting += want * sizeof(T);
    }
    void release() {
        if (p) {
            cudaFree(p);
            if (accounting) *accounting -= cap * sizeof(T);
        }
        p = nullptr;
        cap = 0;
    }
    ~DevBuf() { release(); }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
};

struct PhaseAcc {
    double ms = 0;
    int64_t calls = 0;
};
struct PendingPhase {
    std::string name;
    cudaEvent_t e0, e1;
};

struct NcclApi;      // dlopen'ed NCCL entry points (comm.cu)
struct LocalGroup;   // in-process transport: several contexts (ranks) driven by threads of one process (comm.cu)

// Halo plan of one (multigrid) level of a partitioned mesh: rows [0, nOwned) are computed here, the entries behind them
// are ghosts grouped by owner; per peer a send list (local owned ids) and one contiguous ghost range to receive into.
struct HaloPlan {
    struct Peer {
        int rank, sendOff, sendCount, recvStart, recvCount;
    };
    std::vector<Peer> peers;
    DevBuf<int> sendIdx;     // local ids of owned entries to send, concatenated per peer
    std::vector<int> sendIdxHost;
    DevBuf<double> sendBuf;  // packed send data (grown on demand)
    int nSendTotal = 0;
    void clear() {
        peers.clear();
        sendIdxHost.clear();
        nSendTotal = 0;
    }
};
struct MgHierarchy;  // multigrid preconditioner state (mg.cu)

// slots of the Krylov scalar bank kept on the device (krylov.cu)
enum { SC_RHO0 = 0, SC_RHO1, SC_ALPHA, SC_OMEGA, SC_DONE, SC_ITERS, SC_RES2, SC_BNORM2, SC_TOL2, SC_BAD, SC_TMP0, SC_TMP1, SC_COUNT = 16 };
// slots of the per-block partial-sum bank
enum { PS_RHO = 0, PS_SIGMA, PS_TS, PS_TT, PS_RR, PS_AUX, PS_COUNT = 8 };

struct pfem_ctx {
    int dim = 0, device = 0, nRanks = 1, rank = 0;
    int smCount = 148;
    cudaStream_t stream = nullptr, ownStream = nullptr;
    std::string err;
    int64_t launches = 0;
    size_t deviceBytes = 0;

    // ---- mesh (device) ----
    int nNodes = 0, nElems = 0;
    int nRows = 0;             // rows computed by this rank: nNodes, or the owned-node count of a partitioned mesh
    int64_t nBlocks = 0;       // node-block non-zeros
    int64_t nnzReference = -1; // lazily counted
    int maxE = 0, maxNb = 0;
    bool haveTopology = false, havePositions = false, haveSnapshot = false, haveDirichlet = false;
    bool haveQprev = false, haveSystem = false, haveSolution = false;
    DevBuf<int> conn;          // nElems*(dim+1)
    DevBuf<uint8_t> flags;     // PFEM_NODE_* ; FREE recomputed from the incidence (Node.inl:48-51)
    DevBuf<uint8_t> dirMask;
    DevBuf<double> dirVal4;    // 4 doubles per node
    DevBuf<int> n2ePtr, n2e;   // node -> incident elements, ascending element index
    DevBuf<int> nbrPtr, nbr;   // node -> sorted neighbour nodes (incl. itself) == block-row pattern
    DevBuf<int> diagSlot;      // position of the node in its own neighbour list
    DevBuf<unsigned> n2eSlots; // per (node, incident element): slot bytes of the element's nodes in the node's nbr list
    DevBuf<unsigned> blkMask;  // per block (i,slot): maskWords words, bit k = incident element k of i contains the neighbour
    int maskWords = 1;
    DevBuf<unsigned> rowDir;   // per node: bit s = neighbour slot s is a Dirichlet node (rebuilt when topology/BC change)
    bool rowDirDirty = true;
    DevBuf<int> nodeHdr;       // per node int4 (n2ePtr, nbrPtr, ne | nb<<8 | diagSlot<<16 | flags<<24, rowDir word 0): one load per node
    DevBuf<int> scratchI;      // counters / cursors
    DevBuf<int> scanScratch;   // tile sums of exclusiveScanInt
    DevBuf<unsigned long long> stage64;
    DevBuf<double> stageD;     // host<->device staging in ABI (SoA) layout
    DevBuf<uint8_t> stageB;

    // ---- nodal fields (device, 4 doubles per node, both dims) ----
    DevBuf<double> X4;         // (x, y, z, p)
    DevBuf<double> Xsave4;     // snapshot of X4 (saveNodesList)
    DevBuf<double> V4;         // (u, v, w, rho)
    DevBuf<double> A4;         // (ax, ay, az, -)
    DevBuf<double> VP4;        // (u_prev, v_prev, w_prev, |v_cur|)  -- PSPG assembly input
    DevBuf<double> X4b, V4b;   // ping-pong partners for the explicit step

    // ---- PSPG system ----
    DevBuf<double> Aval;       // nBlocks * BS*BS, block row-major
    DevBuf<double> bvec;       // nNodes*BS, internal dof = node*BS + d
    DevBuf<double> dinv;       // symmetric Jacobi scale 1/sqrt|a_ii|
    DevBuf<double> Wblk;       // node-block Jacobi: A_ii^-1 S_i^-1 per node
    MgHierarchy* mg = nullptr; // aggregation multigrid preconditioner (mg.cu)
    int precondKind = 0;       // PFEM_PRECOND_*: 0 auto
    int mgSweeps = 0;          // 0: default
    double mgDamping = 0.0;    // 0: default
    int lastPrecond = 0;       // what the last solve used
    double asmStamp = 0.0;     // dt of the assembled system (the multigrid dampings are re-tuned when it changes)
    DevBuf<double> kx, kr, kr0, kp, kp2, kv, ks, kt, kph, ksh;  // Krylov vectors (internal dof order)
    DevBuf<double> partial;    // PS_COUNT * reduceBlocks
    DevBuf<double> genRec;     // general assembly kernels: per-element record (grad N, V, tau, viscosity), 16 doubles
    DevBuf<double> fsVec;      // fractional-step systems: vTilde input / scalar-system solution, [d][nNodes]
    DevBuf<uint8_t> fsMask;    // row mask of the scalar systems (zero: pressure; bound nodes: velocity correction)
    const uint8_t* heatMask = nullptr;  // the Dirichlet-row mask the scalar system in hA was assembled with
    int fsWhich = -1;          // fractional-step system assembled last (0 velocity prediction, 1 pressure, 2 velocity correction)
    bool mgFlexible = false;   // the Krylov method tolerates a slightly non-linear preconditioner (FGMRES): fp32 cycle vectors
    DevBuf<double> gmV, gmZ;   // FGMRES: orthonormal basis (owned dofs) and preconditioned directions (owned + ghost dofs)
    DevBuf<double> gmBank;     // (restart + 2) * reduceBlocks partial dot products
    DevBuf<double> gmS;        // Hessenberg column, rotations, g, y (GmLayout in krylov.cu)
    DevBuf<double> scal;       // SC_COUNT
    double* hScal = nullptr;   // pinned mirror
    int reduceBlocks = 0;

    // ---- export scratch ----
    DevBuf<int> cscPtr;
    DevBuf<int> cscRow;
    DevBuf<double> cscVal;

    // ---- explicit step ----
    DevBuf<double> dtPartial;
    DevBuf<double> wcF0;  // CDS_rho: F0 = sum_e M_e rho_e on the configuration before the move
    DevBuf<double> wcElemRec;  // two-pass explicit step: per-(local node, element) momentum records
    DevBuf<double> wcContRec;  // per-element continuity record (alpha, beta, V/NPE, he)
    DevBuf<double> wcCfl2;     // per node (max(u^2, c^2), alpha^2) of the state the last two-pass step produced
    DevBuf<double> wcHmin;     // per node: smallest he = 2 r_in among its incident elements (mesh of the last two-pass step)
    int wcVariant = 0;         // pfem_wc_set_variant: 0 = by size (PFEM_WC_CFG), 6 gather, 11 two-pass, 12 mixed
    bool cflFresh = false;     // wcContRec.he / wcCfl2 describe the current positions and states
    bool tilesValid = false;   // node tiles of the fused explicit step (wc_tile.cuh) match the current topology/partition
    bool orderValid = false;   // tilePerm (interface nodes first, then by grid cell) matches the current topology/partition
    int nIfaceNodes = 0;
    int tileT = 0, nTiles = 0, nIfaceTiles = 0, tileCap = 0;  // nodes per tile, tiles, tiles holding interface nodes, max record slots per tile
    DevBuf<int> tilePerm, tileNePrefix, tileElems, tileCnt, tileKey, tileNodes;
    const int* tileNodeStart = nullptr;  // views into tileCnt
    const int* tileNodeCnt = nullptr;
    int tileNodeCap = 0;                 // max entries of a tile's node list
    DevBuf<unsigned short> tileDst16, tileLconn;
    int tileMaxElems = 0;
    DevBuf<uint8_t> tileIface;
    cudaStream_t commStream = nullptr;  // halo exchanges that overlap the interior tiles (partitioned explicit step)
    cudaEvent_t evTile = nullptr, evHalo = nullptr;
    double cflMu = 0, cflK0 = 0, cflK0p = 0;

    // ---- temperature-dependent / shear-rate-dependent factors and the heat equations (thermal.cu) ----
    bool thermalOn = false;    // Boussinesq factors (pfem_set_thermal)
    double thK = 0, thCv = 1, thAlpha = 0, thTr = 0;
    bool binghamOn = false;    // Bingham regularised viscosity (pfem_set_bingham)
    double binghamTau0 = 0, binghamM = 0;
    bool haveTemperature = false, haveTemperatureBc = false, haveHeatSystem = false;
    DevBuf<double> Tn, Tnb;    // nodal temperature (+ ping-pong partner of the explicit heat step)
    DevBuf<uint8_t> tMask;     // node carries a temperature Dirichlet condition ("<type>T" Lua function)
    DevBuf<double> tVal;
    DevBuf<double> hA, hb, hTheta;            // implicit heat system on the node pattern
    DevBuf<double> cgR, cgZ, cgP, cgAp, cgD;  // conjugate-gradient vectors

    // ---- free-surface facets / surface tension (facets.cu) ----
    int nFacets = 0, nFstNodes = 0;
    double gammaST = 0.0;      // MomContEqIncompNewton::m_gamma / MomEqWCompNewton::m_gamma
    DevBuf<int> facetRec;      // per facet: dim facet nodes, out node, element index
    DevBuf<int> fstNode, fstPtr, fstItem;  // nodes touched by a facet's element -> their facets, ascending facet index
    DevBuf<double> fst4;       // nodal surface-tension force, 4 doubles per node

    // ---- multi-GPU ----
    NcclApi* nccl = nullptr;
    void* comm = nullptr;
    LocalGroup* local = nullptr;   // in-process transport (pfem_comm_init_local) instead of NCCL
    cudaEvent_t evPacked = nullptr, evCopied = nullptr;  // local transport: stream ordering between the ranks' streams
    HaloPlan plan;                 // halo plan of the local mesh (pfem_set_partition)
    DevBuf<double> commScratch;    // staging of small host <-> device collectives

    // ---- profiling ----
    bool profiling = false;
    bool profileDetail = false;  // per-kernel phases inside the multigrid cycle (no CUDA graph then)
    std::map<std::string, PhaseAcc> phases;
    std::vector<PendingPhase> pending;
    std::vector<cudaEvent_t> eventPool;

    pfem_ctx() {
        for (auto* b : {&conn, &n2ePtr, &n2e, &nbrPtr, &nbr, &diagSlot, &scratchI, &scanScratch, &cscPtr, &cscRow, &plan.sendIdx,
                        &facetRec, &fstNode, &fstPtr, &fstItem})
            b->accounting = &deviceBytes;
        fst4.accounting = &deviceBytes;
        wcElemRec.accounting = &deviceBytes;
        wcContRec.accounting = &deviceBytes;
        wcCfl2.accounting = &deviceBytes;
        wcHmin.accounting = &deviceBytes;
        for (auto* b : {&flags, &dirMask, &stageB}) b->accounting = &deviceBytes;
        for (auto* b : {&dirVal4, &stageD, &X4, &Xsave4, &V4, &A4, &VP4, &X4b, &V4b, &Aval, &bvec, &dinv, &Wblk, &kx, &kr, &kr0,
                        &kp, &kp2, &kv, &ks, &kt, &kph, &ksh, &partial, &scal, &cscVal, &dtPartial, &plan.sendBuf, &commScratch})
            b->accounting = &deviceBytes;
        stage64.accounting = &deviceBytes;
        n2eSlots.accounting = &deviceBytes;
        blkMask.accounting = &deviceBytes;
        rowDir.accounting = &deviceBytes;
        nodeHdr.accounting = &deviceBytes;
        for (auto* b : {&Tn, &Tnb, &tVal, &hA, &hb, &hTheta, &cgR, &cgZ, &cgP, &cgAp, &cgD}) b->accounting = &deviceBytes;
        tMask.accounting = &deviceBytes;
        for (auto* b : {&tilePerm, &tileNePrefix, &tileElems, &tileCnt, &tileKey, &tileNodes}) b->accounting = &deviceBytes;
        tileDst16.accounting = &deviceBytes;
        tileLconn.accounting = &deviceBytes;
        tileIface.accounting = &deviceBytes;
    }
};

// ---- phase timing (names mirror the reference's m_accumalatedTimes keys) ------------------------------------------
struct PhaseScope {
    pfem_ctx* c;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    std::string name;
    PhaseScope(pfem_ctx* ctx, const char* n) : c(ctx), name(n) {
        if (!c->profiling) return;
        auto get = [&]() {
            cudaEvent_t e;
            if (!c->eventPool.empty()) {
                e = c->eventPool.back();
                c->eventPool.pop_back();
            } else
                cudaEventCreate(&e);
            return e;
        };
        e0 = get();
        e1 = get();
        cudaEventRecord(e0, c->stream);
    }
    ~PhaseScope() {
        if (!e0) return;
        cudaEventRecord(e1, c->stream);
        c->pending.push_back({name, e0, e1});
    }
};
void pfemFlushPhases(pfem_ctx* c);

inline void countLaunch(pfem_ctx* c, int n = 1) { c->launches += n; }
#define LAUNCH_CHECK(c)                 \
    do {                                \
        countLaunch(c);                 \
        CUDA_CHECK(cudaGetLastError()); \
    } while (0)

inline int divUp(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- per-module entry points (implemented in the .cu files) -------------------------------------------------------
// topology.cu
void topoBuild(pfem_ctx* c, int64_t nNodes, int64_t nElems, const uint64_t* elemNodes, const uint8_t* flags);
void exclusiveScanInt(pfem_ctx* c, int* data, int n, int* totalOut /*device, may be null*/);
// fields.cu
void fieldsSetPositions(pfem_ctx* c, const double* x);
void fieldsGetPositions(pfem_ctx* c, double* x);
void fieldsSetStates(pfem_ctx* c, int first, int count, const double* q);
void fieldsGetStates(pfem_ctx* c, int first, int count, double* q);
void fieldsSetDirichlet(pfem_ctx* c, const uint8_t* mask, const double* values);
void fieldsSnapshot(pfem_ctx* c);
void fieldsRestore(pfem_ctx* c);
void fieldsMove(pfem_ctx* c, const double* delta, int fromSnapshot);
void fieldsSetQprev(pfem_ctx* c, const double* qPrev);
// pspg.cu
void pspgAssemble(pfem_ctx* c, const pfem_pspg_params& p);
void pspgExportCsc(pfem_ctx* c, int64_t* nnz, int32_t* colPtr, int32_t* rowIdx, double* val, double* b);
void pspgPicardUpdate(pfem_ctx* c, double dt);  // states <- q ; X = Xsave + dt*v
// krylov.cu
int krylovSolve(pfem_ctx* c, double relTol, int maxIter, int* iters, double* relRes, bool warmStart);
void krylovFetchSolution(pfem_ctx* c, double* q);
void krylovLoadVector(pfem_ctx* c, const double* qHost, double* dst);  // ABI layout -> internal dof order
void krylovStoreVector(pfem_ctx* c, const double* src, double* qHost);
double krylovResidualNorm(pfem_ctx* c, double* xInternal);
void krylovMatvec(pfem_ctx* c, double* xInternal, double* yInternal);
// mg.cu
void mgInvalidate(pfem_ctx* c, bool symbolic);
void mgDestroy(pfem_ctx* c);
bool mgSetup(pfem_ctx* c);
double* mgRhs(pfem_ctx* c);
float* mgRhsF(pfem_ctx* c);  // non-null when the cycle runs on fp32 vectors: the caller writes the right-hand side there instead
int mgLevelCount(pfem_ctx* c);
void mgApply(pfem_ctx* c, double* out);
// wc.cu
void wcStep(pfem_ctx* c, const pfem_wc_params& p, double dt);
int wcNextDt(pfem_ctx* c, const pfem_wc_params& p, double securityCoeff, double maxDT, double* dt);
int wcRun(pfem_ctx* c, const pfem_wc_params& p, int nSteps, double securityCoeff, double maxDT, double* dtInOut, double* elapsed);
// thermal.cu
bool pspgNeedsGeneralPath(const pfem_ctx* c);
void pspgAssembleGeneral(pfem_ctx* c, const pfem_pspg_params& p, const double* fst4);
void thermalSetTemperature(pfem_ctx* c, const double* T);
void thermalGetTemperature(pfem_ctx* c, double* T);
void thermalSetBc(pfem_ctx* c, const uint8_t* mask, const double* values);
void thermalWcHeat(pfem_ctx* c, double dt, const double* dtPtr);
void thermalPrepare(pfem_ctx* c);
void heatAssemble(pfem_ctx* c, double rho, double cv, double k, double dt, const double* thetaPrevHost);
int heatSolve(pfem_ctx* c, double relTol, int maxIter, double* Tout, int* itersOut, double* relResOut);
void heatExport(pfem_ctx* c, int64_t* nnz, int32_t* colPtr, int32_t* rowIdx, double* val, double* b);
void fsAssembleVapp(pfem_ctx* c, const pfem_pspg_params& p, double gammaFS, const double* qPrevHost);
void fsAssemblePcorr(pfem_ctx* c, double rho, double dt, double gammaFS, const double* vTildeHost, const double* pPrevHost);
void fsAssembleVcorr(pfem_ctx* c, double rho, double dt, const double* deltaPHost);
void fsGetRhs(pfem_ctx* c, double* b);
int fsSolve(pfem_ctx* c, double relTol, int maxIter, double* xHost, int* itersOut, double* relResOut);
// facets.cu
void facetsSet(pfem_ctx* c, int64_t nFacets, const uint64_t* facetNodes, const uint64_t* outNode, const uint64_t* elemIndex);
const double* facetsForces(pfem_ctx* c, const double* X4, bool allNodesRule);  // null when the facet terms are off
// comm.cu
void commUniqueId(void* id128);
void commInit(pfem_ctx* c, int nRanks, int rank, const void* id128);
void commDestroy(pfem_ctx* c);
void commSetPartition(pfem_ctx* c, int64_t nOwned, int nPeers, const int32_t* peerRank, const int64_t* sendOffsets,
                      const int32_t* sendIdx, const int64_t* recvStart, const int64_t* recvCount);
void commHalo(pfem_ctx* c, double* arr0, double* arr1, int width);  // owners -> ghosts, up to two nodal arrays (level-0 plan)
void commHaloPlan(pfem_ctx* c, HaloPlan& plan, double* arr0, double* arr1, int width);
template <typename T> void commHaloPlanT(pfem_ctx* c, HaloPlan& plan, T* arr0, T* arr1, int width);  // T = double | float
void commAllReduceSum(pfem_ctx* c, double* buf, int count);
void commAllReduceMin(pfem_ctx* c, double* devScalar);
// every rank contributes counts[rank] doubles from `src`; all of them land in `dst` at displs[r] on every rank (device buffers;
// src may alias dst + displs[rank])
void commAllGatherV(pfem_ctx* c, const double* src, double* dst, const std::vector<int64_t>& counts, const std::vector<int64_t>& displs);
// the same for raw bytes (counts and displacements in bytes)
void commAllGatherBytes(pfem_ctx* c, const void* src, void* dst, const std::vector<int64_t>& counts, const std::vector<int64_t>& displs);
// install a halo plan on the device (send list upload, buffers)
void commFinishPlan(pfem_ctx* c, HaloPlan& P);
// small host-side all-gather: n doubles per rank -> nRanks*n doubles, rank-major (synchronises the stream)
void commAllGatherHost(pfem_ctx* c, const double* mine, int n, double* all);
void commInitLocal(pfem_ctx* c, LocalGroup* g, int rank);
LocalGroup* commLocalCreate(int nRanks);
void commLocalDestroy(LocalGroup* g);
void commAbort(pfem_ctx* c);  // a failing rank releases the others from their barriers (local transport)
