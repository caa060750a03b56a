// fields.cu -- nodal data movement between the ABI layout q[n + s*nNodes] (StatesFromToQ.hpp:9-34) and the
// device layout (4 doubles per node: X4=(x,y,z,p), V4=(u,v,w,rho), A4=(ax,ay,az,-)), plus the position
// snapshot / move operations of Mesh.cpp:1019-1053, 1101-1137, 1238-1277.
//
// State index -> device slot (dim = 2|3):   s <  dim      -> V4[s]          velocity
//                                            s == dim      -> X4.w           pressure
//                                            s == dim+1    -> V4.w           density      (WC only)
//                                            s >= dim+2    -> A4[s-dim-2]    acceleration (WC only)
#include "common.cuh"

namespace {

struct Slot {
    int arr;   // 0 X4, 1 V4, 2 A4
    int comp;
};
__host__ __device__ inline Slot stateSlot(int dim, int s) {
    if (s < dim) return {1, s};
    if (s == dim) return {0, 3};
    if (s == dim + 1) return {1, 3};
    return {2, s - dim - 2};
}

__global__ void k_soa_to_aos(const double* __restrict__ src, int nNodes, int count, double* __restrict__ X4,
                             double* __restrict__ V4, double* __restrict__ A4, int dim, int first) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nNodes) return;
    for (int k = 0; k < count; ++k) {
        const Slot sl = stateSlot(dim, first + k);
        double* dst = sl.arr == 0 ? X4 : (sl.arr == 1 ? V4 : A4);
        dst[(size_t)n * 4 + sl.comp] = src[(size_t)k * nNodes + n];
    }
}
__global__ void k_aos_to_soa(double* __restrict__ dst, int nNodes, int count, const double* __restrict__ X4,
                             const double* __restrict__ V4, const double* __restrict__ A4, int dim, int first) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nNodes) return;
    for (int k = 0; k < count; ++k) {
        const Slot sl = stateSlot(dim, first + k);
        const double* src = sl.arr == 0 ? X4 : (sl.arr == 1 ? V4 : A4);
        dst[(size_t)k * nNodes + n] = src[(size_t)n * 4 + sl.comp];
    }
}
__global__ void k_pos_in(const double* __restrict__ src, int nNodes, int dim, double* __restrict__ X4) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nNodes) return;
    for (int d = 0; d < 3; ++d) X4[(size_t)n * 4 + d] = d < dim ? src[(size_t)d * nNodes + n] : 0.0;
}
__global__ void k_pos_out(double* __restrict__ dst, int nNodes, int dim, const double* __restrict__ X4) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nNodes) return;
    for (int d = 0; d < dim; ++d) dst[(size_t)d * nNodes + n] = X4[(size_t)n * 4 + d];
}
__global__ void k_dir_in(const double* __restrict__ src, int nNodes, int dim, double* __restrict__ D4) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nNodes) return;
    for (int d = 0; d < 4; ++d) D4[(size_t)n * 4 + d] = d < dim ? src[(size_t)d * nNodes + n] : 0.0;
}
// Mesh::updateNodesPosition(FromSave): x = base + delta unless isFixed (Mesh.cpp:1106-1116, 1246-1256)
__global__ void k_move(const double* __restrict__ delta, int nNodes, int dim, const uint8_t* __restrict__ flags,
                       const double* __restrict__ base4, double* __restrict__ X4) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nNodes) return;
    if (flags[n] & PFEM_NODE_FIXED) return;
    for (int d = 0; d < dim; ++d) X4[(size_t)n * 4 + d] = base4[(size_t)n * 4 + d] + delta[(size_t)d * nNodes + n];
}
// keep the pressure slot (X4.w) when copying positions back and forth
__global__ void k_copy_xyz(const double* __restrict__ src4, double* __restrict__ dst4, int nNodes) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nNodes) return;
    for (int d = 0; d < 3; ++d) dst4[(size_t)n * 4 + d] = src4[(size_t)n * 4 + d];
}
// VP4 = (qPrev velocity, -) ; the |v_cur| slot is filled by the assembly prologue
__global__ void k_qprev_in(const double* __restrict__ src, int nNodes, int dim, double* __restrict__ VP4) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nNodes) return;
    for (int d = 0; d < 3; ++d) VP4[(size_t)n * 4 + d] = d < dim ? src[(size_t)d * nNodes + n] : 0.0;
}

void needTopo(pfem_ctx* c) { PFEM_REQUIRE(c->haveTopology, PFEM_ERR_STATE, "call pfem_set_topology first"); }
int grid(pfem_ctx* c) { return divUp(c->nNodes, 256); }

}  // namespace

void fieldsSetPositions(pfem_ctx* c, const double* x) {
    needTopo(c);
    PFEM_REQUIRE(x, PFEM_ERR_INVALID, "set_positions: null");
    const size_t n = (size_t)c->dim * c->nNodes;
    c->stageD.reserve(n);
    CUDA_CHECK(cudaMemcpyAsync(c->stageD.p, x, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    k_pos_in<<<grid(c), 256, 0, c->stream>>>(c->stageD.p, c->nNodes, c->dim, c->X4.p);
    LAUNCH_CHECK(c);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->havePositions = true;
}
void fieldsGetPositions(pfem_ctx* c, double* x) {
    needTopo(c);
    PFEM_REQUIRE(x, PFEM_ERR_INVALID, "get_positions: null");
    const size_t n = (size_t)c->dim * c->nNodes;
    c->stageD.reserve(n);
    k_pos_out<<<grid(c), 256, 0, c->stream>>>(c->stageD.p, c->nNodes, c->dim, c->X4.p);
    LAUNCH_CHECK(c);
    CUDA_CHECK(cudaMemcpyAsync(x, c->stageD.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
}
void fieldsSetStates(pfem_ctx* c, int first, int count, const double* q) {
    needTopo(c);
    PFEM_REQUIRE(q && first >= 0 && count > 0 && first + count <= 2 * c->dim + 2, PFEM_ERR_INVALID,
                 "set_states: state range outside [0, 2*dim+2)");
    const size_t n = (size_t)count * c->nNodes;
    c->stageD.reserve(n);
    CUDA_CHECK(cudaMemcpyAsync(c->stageD.p, q, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    k_soa_to_aos<<<grid(c), 256, 0, c->stream>>>(c->stageD.p, c->nNodes, count, c->X4.p, c->V4.p, c->A4.p, c->dim, first);
    LAUNCH_CHECK(c);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
}
void fieldsGetStates(pfem_ctx* c, int first, int count, double* q) {
    needTopo(c);
    PFEM_REQUIRE(q && first >= 0 && count > 0 && first + count <= 2 * c->dim + 2, PFEM_ERR_INVALID,
                 "get_states: state range outside [0, 2*dim+2)");
    const size_t n = (size_t)count * c->nNodes;
    c->stageD.reserve(n);
    k_aos_to_soa<<<grid(c), 256, 0, c->stream>>>(c->stageD.p, c->nNodes, count, c->X4.p, c->V4.p, c->A4.p, c->dim, first);
    LAUNCH_CHECK(c);
    CUDA_CHECK(cudaMemcpyAsync(q, c->stageD.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
}
void fieldsSetDirichlet(pfem_ctx* c, const uint8_t* mask, const double* values) {
    needTopo(c);
    PFEM_REQUIRE(mask && values, PFEM_ERR_INVALID, "set_dirichlet: null");
    const size_t n = (size_t)c->dim * c->nNodes;
    c->stageD.reserve(n);
    CUDA_CHECK(cudaMemcpyAsync(c->dirMask.p, mask, c->nNodes, cudaMemcpyHostToDevice, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(c->stageD.p, values, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    k_dir_in<<<grid(c), 256, 0, c->stream>>>(c->stageD.p, c->nNodes, c->dim, c->dirVal4.p);
    LAUNCH_CHECK(c);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->haveDirichlet = true;
    c->rowDirDirty = true;
}
void fieldsSnapshot(pfem_ctx* c) {
    needTopo(c);
    PFEM_REQUIRE(c->havePositions, PFEM_ERR_STATE, "the nodes list does not exist!");  // Mesh.cpp:1049-1050
    CUDA_CHECK(cudaMemcpyAsync(c->Xsave4.p, c->X4.p, (size_t)c->nNodes * 4 * sizeof(double), cudaMemcpyDeviceToDevice,
                               c->stream));
    c->haveSnapshot = true;
}
void fieldsRestore(pfem_ctx* c) {
    needTopo(c);
    PFEM_REQUIRE(c->haveSnapshot, PFEM_ERR_STATE, "the nodes list was not saved before or does not exist!");  // Mesh.cpp:1021-1022
    k_copy_xyz<<<grid(c), 256, 0, c->stream>>>(c->Xsave4.p, c->X4.p, c->nNodes);
    LAUNCH_CHECK(c);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
}
void fieldsMove(pfem_ctx* c, const double* delta, int fromSnapshot) {
    needTopo(c);
    PFEM_REQUIRE(delta, PFEM_ERR_INVALID, "move_positions: null");
    PFEM_REQUIRE(c->havePositions, PFEM_ERR_STATE, "move_positions: no positions");
    if (fromSnapshot) PFEM_REQUIRE(c->haveSnapshot, PFEM_ERR_STATE, "you did not save the nodes list!");  // Mesh.cpp:1240-1241
    const size_t n = (size_t)c->dim * c->nNodes;
    c->stageD.reserve(n);
    CUDA_CHECK(cudaMemcpyAsync(c->stageD.p, delta, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    k_move<<<grid(c), 256, 0, c->stream>>>(c->stageD.p, c->nNodes, c->dim, c->flags.p, fromSnapshot ? c->Xsave4.p : c->X4.p,
                                           c->X4.p);
    LAUNCH_CHECK(c);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
}
void fieldsSetQprev(pfem_ctx* c, const double* qPrev) {
    needTopo(c);
    PFEM_REQUIRE(qPrev, PFEM_ERR_INVALID, "set_qprev: null");
    const size_t n = (size_t)c->dim * c->nNodes;  // only the velocity part of qPrev is read (PSPG.inl:44, 201)
    c->stageD.reserve(n);
    CUDA_CHECK(cudaMemcpyAsync(c->stageD.p, qPrev, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    k_qprev_in<<<grid(c), 256, 0, c->stream>>>(c->stageD.p, c->nNodes, c->dim, c->VP4.p);
    LAUNCH_CHECK(c);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->haveQprev = true;
}
