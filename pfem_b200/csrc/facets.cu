// Surface tension on free-surface facets: the nodal force vector
//     FST = -gamma * detJ_f * refSize(dim-1) * sum_g w_g * Be^T * T(P) * P          (MatricesBuilder.inl:389-403)
// with P the tangential projector of the facet normal in Voigt form (getP, :167-190), T(P) (getT, :193-214), Be the strain
// matrix of the element behind the facet, Facet::computeJ/DetJ/Normal (Facet.cpp:16-77, 130-210), added to the velocity
// rows of the right-hand side by the facet loops of m_applyBCPSPG (MomContEquationPSPG.inl:155-187, a facet counts when
// ANY of its nodes is on the free surface) and MomEqWCompNewton::m_applyBC (WCompNewton/MomEquation.inl:312-336,
// Facet::isOnFreeSurface = ALL nodes, Facet.cpp:249-255).
//
// Gather, never scatter: one thread owns one node touched by a facet and sums the contributions of
// its facets in ascending facet index -- the order of the reference's serial facet loop -- so the result is
// bit-reproducible.  Only column (node, d) of Be is needed per contribution: Be^T(T P) restricted to the node's rows is
// a 3-term dot product of the node's shape-function gradient with the symmetric tensor T P.  The number of facets is
// O(nNodes^(2/3)); this kernel is noise next to the element kernels and is launched only when gamma > 0.
#include "common.cuh"

#include <algorithm>
#include <vector>

namespace {

template <int DIM>
__global__ void __launch_bounds__(128) k_fst(int nTouched, const int* __restrict__ fstNode, const int* __restrict__ fstPtr,
                                             const int* __restrict__ fstItem, const int* __restrict__ facetRec,
                                             const int* __restrict__ conn, const uint8_t* __restrict__ flags,
                                             const double* __restrict__ X4, double gamma, int allNodesRule,
                                             double* __restrict__ fst4) {
    constexpr int NPE = DIM + 1, NPF = DIM, REC = DIM + 2;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nTouched) return;
    const int node = fstNode[t];
    double acc[3] = {0.0, 0.0, 0.0};
    for (int it = fstPtr[t]; it < fstPtr[t + 1]; ++it) {
        const int* rec = facetRec + (size_t)fstItem[it] * REC;
        int fn[NPF], nOnFS = 0;
#pragma unroll
        for (int k = 0; k < NPF; ++k) {
            fn[k] = rec[k];
            nOnFS += (flags[fn[k]] & PFEM_NODE_FREE_SURFACE) ? 1 : 0;
        }
        if (allNodesRule ? (nOnFS != NPF) : (nOnFS == 0)) continue;
        const int outNode = rec[NPF], elem = rec[NPF + 1];
        double xf[NPF][3], xo[3];
#pragma unroll
        for (int k = 0; k < NPF; ++k)
#pragma unroll
            for (int d = 0; d < 3; ++d) xf[k][d] = X4[(size_t)fn[k] * 4 + d];
#pragma unroll
        for (int d = 0; d < 3; ++d) xo[d] = X4[(size_t)outNode * 4 + d];
        // facet measure detJ_f * refSize and unit normal pointing away from the out node
        double nrm[3] = {0.0, 0.0, 0.0}, measure;
        if constexpr (DIM == 2) {
            const double J00 = (xf[1][0] - xf[0][0]) / 2, J10 = (xf[1][1] - xf[0][1]) / 2;
            measure = sqrt(J00 * J00 + J10 * J10) * 2.0;
            nrm[0] = xf[1][1] - xf[0][1];
            nrm[1] = xf[0][0] - xf[1][0];
            if (nrm[0] * (xo[0] - xf[0][0]) + nrm[1] * (xo[1] - xf[0][1]) > 0) nrm[0] = -nrm[0], nrm[1] = -nrm[1];
            const double norm = sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1]);
            nrm[0] /= norm, nrm[1] /= norm;
        } else {
            const double a[3] = {xf[1][0] - xf[0][0], xf[1][1] - xf[0][1], xf[1][2] - xf[0][2]};
            const double b[3] = {xf[2][0] - xf[0][0], xf[2][1] - xf[0][1], xf[2][2] - xf[0][2]};
            nrm[0] = a[1] * b[2] - a[2] * b[1];
            nrm[1] = a[2] * b[0] - a[0] * b[2];
            nrm[2] = a[0] * b[1] - a[1] * b[0];
            const double norm = sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
            measure = norm * 0.5;
            if ((xo[0] - xf[0][0]) * nrm[0] + (xo[1] - xf[0][1]) * nrm[1] + (xo[2] - xf[0][2]) * nrm[2] > 0)
                nrm[0] = -nrm[0], nrm[1] = -nrm[1], nrm[2] = -nrm[2];
            nrm[0] /= norm, nrm[1] /= norm, nrm[2] /= norm;
        }
        // S = T(P) P as a symmetric tensor: S_Voigt = T * P_Voigt (getP / getT)
        double S[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        if constexpr (DIM == 2) {
            const double P0 = 1 - nrm[0] * nrm[0], P1 = 1 - nrm[1] * nrm[1], P2 = -nrm[0] * nrm[1];
            const double s0 = (P0 * P0) * P0 + (P2 * P2) * P1 + (2 * P0 * P2) * P2;
            const double s1 = (P2 * P2) * P0 + (P1 * P1) * P1 + (2 * P2 * P1) * P2;
            const double s2 = (P0 * P2) * P0 + (P2 * P1) * P1 + (P0 * P1 + P2 * P2) * P2;
            S[0][0] = s0, S[1][1] = s1, S[0][1] = S[1][0] = s2;
        } else {
            const double P[6] = {1 - nrm[0] * nrm[0], 1 - nrm[1] * nrm[1], 1 - nrm[2] * nrm[2],
                                 -nrm[0] * nrm[1], -nrm[0] * nrm[2], -nrm[1] * nrm[2]};
            double T[6][6];
            T[0][0] = P[0] * P[0]; T[0][1] = P[3] * P[3]; T[0][2] = P[4] * P[4]; T[0][3] = 2 * P[0] * P[3]; T[0][5] = 2 * P[4] * P[3]; T[0][4] = 2 * P[0] * P[4];
            T[1][0] = T[0][1];     T[1][1] = P[1] * P[1]; T[1][2] = P[5] * P[5]; T[1][3] = 2 * P[3] * P[1]; T[1][5] = 2 * P[5] * P[1]; T[1][4] = 2 * P[3] * P[5];
            T[2][0] = T[0][2];     T[2][1] = T[1][2];     T[2][2] = P[2] * P[2]; T[2][3] = 2 * P[5] * P[4]; T[2][5] = 2 * P[5] * P[2]; T[2][4] = 2 * P[2] * P[4];
            T[3][0] = P[0] * P[3]; T[3][1] = P[3] * P[1]; T[3][2] = P[4] * P[5]; T[3][3] = P[0] * P[1] + P[3] * P[3]; T[3][5] = P[4] * P[1] + P[5] * P[3]; T[3][4] = P[4] * P[3] + P[0] * P[5];
            T[5][0] = P[3] * P[4]; T[5][1] = P[5] * P[1]; T[5][2] = P[5] * P[2]; T[5][3] = P[4] * P[1] + P[5] * P[3]; T[5][5] = P[1] * P[2] + P[5] * P[5]; T[5][4] = P[5] * P[4] + P[3] * P[2];
            T[4][0] = P[0] * P[4]; T[4][1] = P[3] * P[5]; T[4][2] = P[2] * P[4]; T[4][3] = P[4] * P[3] + P[0] * P[5]; T[4][5] = P[5] * P[4] + P[3] * P[2]; T[4][4] = P[2] * P[0] + P[4] * P[4];
            double s[6];
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                double sum = 0;
#pragma unroll
                for (int k = 0; k < 6; ++k) sum += T[r][k] * P[k];
                s[r] = sum;
            }
            S[0][0] = s[0], S[1][1] = s[1], S[2][2] = s[2];     // Voigt order [xx, yy, zz, xy, xz, yz] (MB.inl:147-160)
            S[0][1] = S[1][0] = s[3], S[0][2] = S[2][0] = s[4], S[1][2] = S[2][1] = s[5];
        }
        // gradient of this node's shape function in the element behind the facet (Element.cpp:15-135, MB.inl:93-127)
        int en[NPE], local = 0;
        double xe[NPE][DIM];
#pragma unroll
        for (int k = 0; k < NPE; ++k) {
            en[k] = conn[(size_t)elem * NPE + k];
            local = (en[k] == node) ? k : local;
#pragma unroll
            for (int d = 0; d < DIM; ++d) xe[k][d] = X4[(size_t)en[k] * 4 + d];
        }
        double J[DIM][DIM], inv[DIM][DIM];
#pragma unroll
        for (int d = 0; d < DIM; ++d)
#pragma unroll
            for (int k = 0; k < DIM; ++k) J[d][k] = xe[k + 1][d] - xe[0][d];
        if constexpr (DIM == 2) {
            const double det = J[0][0] * J[1][1] - J[1][0] * J[0][1];
            inv[0][0] = J[1][1] / det, inv[0][1] = -J[0][1] / det, inv[1][0] = -J[1][0] / det, inv[1][1] = J[0][0] / det;
        } else {
            const double det = J[0][0] * J[1][1] * J[2][2] + J[0][1] * J[1][2] * J[2][0] + J[0][2] * J[1][0] * J[2][1] -
                               J[2][0] * J[1][1] * J[0][2] - J[2][1] * J[1][2] * J[0][0] - J[2][2] * J[1][0] * J[0][1];
            inv[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) / det;
            inv[0][1] = (J[2][1] * J[0][2] - J[2][2] * J[0][1]) / det;
            inv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
            inv[1][0] = (J[2][0] * J[1][2] - J[1][0] * J[2][2]) / det;
            inv[1][1] = (J[0][0] * J[2][2] - J[2][0] * J[0][2]) / det;
            inv[1][2] = (J[1][0] * J[0][2] - J[0][0] * J[1][2]) / det;
            inv[2][0] = (J[1][0] * J[2][1] - J[2][0] * J[1][1]) / det;
            inv[2][1] = (J[2][0] * J[0][1] - J[0][0] * J[2][1]) / det;
            inv[2][2] = (J[0][0] * J[1][1] - J[1][0] * J[0][1]) / det;
        }
        double g[DIM];
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
            double s0 = 0;
#pragma unroll
            for (int k = 0; k < DIM; ++k) s0 -= inv[k][d];
            g[d] = s0;
#pragma unroll
            for (int k = 0; k < DIM; ++k) g[d] = (local == k + 1) ? inv[k][d] : g[d];
        }
        // (Be^T S)_(node,d) = sum_c grad_c * S[c][d]; three Gauss points with constant integrand, weights sum to one
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
            double bs = 0;
#pragma unroll
            for (int c = 0; c < DIM; ++c) bs += g[c] * S[c][d];
            acc[d] += -(gamma * bs) * measure;
        }
    }
    double* o = fst4 + (size_t)node * 4;
    o[0] = acc[0], o[1] = acc[1], o[2] = acc[2];
}

}  // namespace

// Facet::m_nodesIndexes / m_outNodeIndex / m_elementIndex of Mesh::m_facetsList (Mesh3D.cpp:218-262), local numbering.
void facetsSet(pfem_ctx* c, int64_t nFacets, const uint64_t* facetNodes, const uint64_t* outNode, const uint64_t* elemIndex) {
    PFEM_REQUIRE(c->haveTopology, PFEM_ERR_STATE, "set_facets: call pfem_set_mesh first");
    PFEM_REQUIRE(c->nRanks == 1, PFEM_ERR_STATE, "set_facets: facet terms are implemented for single-GPU contexts");
    c->nFacets = 0;
    c->nFstNodes = 0;
    if (nFacets <= 0) return;
    PFEM_REQUIRE(facetNodes && outNode && elemIndex, PFEM_ERR_INVALID, "set_facets: null array");
    const int dim = c->dim, rec = dim + 2;
    std::vector<int> hRec((size_t)nFacets * rec);
    std::vector<std::pair<int, int>> inc;  // (node, facet): every node of the element behind a facet receives a force
    inc.reserve((size_t)nFacets * (dim + 1));
    for (int64_t f = 0; f < nFacets; ++f) {
        for (int k = 0; k < dim; ++k) {
            const uint64_t n = facetNodes[f * dim + k];
            PFEM_REQUIRE(n < (uint64_t)c->nNodes, PFEM_ERR_INVALID, "set_facets: node index out of range");
            hRec[f * rec + k] = (int)n;
            inc.emplace_back((int)n, (int)f);
        }
        PFEM_REQUIRE(outNode[f] < (uint64_t)c->nNodes && elemIndex[f] < (uint64_t)c->nElems, PFEM_ERR_INVALID,
                     "set_facets: out node / element index out of range");
        hRec[f * rec + dim] = (int)outNode[f];
        hRec[f * rec + dim + 1] = (int)elemIndex[f];
        inc.emplace_back((int)outNode[f], (int)f);
    }
    std::stable_sort(inc.begin(), inc.end());  // by node, then ascending facet index = the reference's summation order
    std::vector<int> hNode, hPtr, hItem(inc.size());
    for (size_t i = 0; i < inc.size(); ++i) {
        if (i == 0 || inc[i].first != inc[i - 1].first) {
            hNode.push_back(inc[i].first);
            hPtr.push_back((int)i);
        }
        hItem[i] = inc[i].second;
    }
    hPtr.push_back((int)inc.size());
    c->facetRec.reserve(hRec.size());
    c->fstNode.reserve(hNode.size());
    c->fstPtr.reserve(hPtr.size());
    c->fstItem.reserve(hItem.size());
    c->fst4.reserve((size_t)c->nNodes * 4);
    CUDA_CHECK(cudaMemcpyAsync(c->facetRec.p, hRec.data(), hRec.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(c->fstNode.p, hNode.data(), hNode.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(c->fstPtr.p, hPtr.data(), hPtr.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(c->fstItem.p, hItem.data(), hItem.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CUDA_CHECK(cudaMemsetAsync(c->fst4.p, 0, (size_t)c->nNodes * 4 * sizeof(double), c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));  // the host vectors go out of scope
    c->nFacets = (int)nFacets;
    c->nFstNodes = (int)hNode.size();
}

// Returns the nodal force array (4 doubles per node, zero for nodes no facet touches) computed on positions X4, or null
// when the facet terms are off (gamma < 1e-15: PSPG.inl:157, MomEquation.inl:314; or no facets).
const double* facetsForces(pfem_ctx* c, const double* X4, bool allNodesRule) {
    if (c->nFacets == 0 || c->nFstNodes == 0 || c->gammaST < 1e-15) return nullptr;
    PhaseScope ph(c, "Apply boundary conditions");
    const int grid = divUp(c->nFstNodes, 128);
    if (c->dim == 2)
        k_fst<2><<<grid, 128, 0, c->stream>>>(c->nFstNodes, c->fstNode.p, c->fstPtr.p, c->fstItem.p, c->facetRec.p, c->conn.p,
                                             c->flags.p, X4, c->gammaST, allNodesRule ? 1 : 0, c->fst4.p);
    else
        k_fst<3><<<grid, 128, 0, c->stream>>>(c->nFstNodes, c->fstNode.p, c->fstPtr.p, c->fstItem.p, c->facetRec.p, c->conn.p,
                                             c->flags.p, X4, c->gammaST, allNodesRule ? 1 : 0, c->fst4.p);
    LAUNCH_CHECK(c);
    return c->fst4.p;
}
