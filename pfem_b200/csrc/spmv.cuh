// spmv.cuh -- node-block SpMV kernel and the block-reduction helpers shared by krylov.cu (BiCGSTAB) and mg.cu (multigrid
// smoother / residual).  Included inside each translation unit's anonymous namespace users; everything here is static.
#pragma once
#include "common.cuh"

namespace {

constexpr int RB_THREADS = 256;

// ---- block reduction helpers ------------------------------------------------------------------------------------------
__device__ __forceinline__ double warpSum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <int NV> __device__ __forceinline__ void blockSumStore(double (&v)[NV], double* partial, int stride, const int* slots) {
    __shared__ double sh[NV][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const double s = warpSum(v[k]);
        if (lane == 0) sh[k][w] = s;
    }
    __syncthreads();
    if (w == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double s = lane < nw ? sh[k][lane] : 0.0;
            s = warpSum(s);
            if (lane == 0) partial[(size_t)slots[k] * stride + blockIdx.x] = s;
        }
    }
}
// every block reduces the bank entry `slot` identically (fixed order) -> same bits in every block
__device__ __forceinline__ double bankSum(const double* __restrict__ partial, int stride, int slot, int nPart) {
    __shared__ double sh[32];
    __shared__ double result;
    double s = 0;
    for (int k = threadIdx.x; k < nPart; k += blockDim.x) s += partial[(size_t)slot * stride + k];
    s = warpSum(s);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) sh[w] = s;
    __syncthreads();
    if (w == 0) {
        double t = lane < nw ? sh[lane] : 0.0;
        t = warpSum(t);
        if (lane == 0) result = t;
    }
    __syncthreads();
    return result;
}

// ---- SpMV y = A x with up to two fused dots  (y,w1) and (y,y) ------------------------------------------------------
// Software-pipelined over block rows: row pointers are fetched two rows ahead and column indices one row ahead, so the
// only loads a row waits for are its own A rows (streamed once, evict-first) and the x gathers (L2-resident).
template <int BS> struct RowLoad {
    double2 a01, a23, x01, x23;
};
template <int BS>
__device__ __forceinline__ void loadSlot(const double* __restrict__ Aval, const double* __restrict__ x, int blk, int col,
                                         int r, RowLoad<BS>& L) {
    const double* ap = Aval + ((size_t)blk * BS + r) * BS;
    const double* xp = x + (size_t)col * BS;
    if constexpr (BS == 4) {
        L.a01 = __ldcs(reinterpret_cast<const double2*>(ap));
        L.a23 = __ldcs(reinterpret_cast<const double2*>(ap + 2));
        L.x01 = __ldg(reinterpret_cast<const double2*>(xp));
        L.x23 = __ldg(reinterpret_cast<const double2*>(xp + 2));
    } else {
        L.a01 = make_double2(__ldcs(ap), __ldcs(ap + 1));
        L.a23 = make_double2(__ldcs(ap + 2), 0.0);
        L.x01 = make_double2(__ldg(xp), __ldg(xp + 1));
        L.x23 = make_double2(__ldg(xp + 2), 0.0);
    }
}
// fp32 copy of the matrix (multigrid levels: the preconditioner may use a rounded A, it stays a fixed linear operator):
// half the bytes per block row, products and sums still in fp64.  The row stays packed (4 registers) until it is used.
template <int BS> struct RowLoadF {
    float4 a;
    double2 x01, x23;
};
template <int BS>
__device__ __forceinline__ void loadSlot(const float* __restrict__ Aval, const double* __restrict__ x, int blk, int col,
                                         int r, RowLoadF<BS>& L) {
    const float* ap = Aval + ((size_t)blk * BS + r) * BS;
    const double* xp = x + (size_t)col * BS;
    if constexpr (BS == 4) {
        L.a = __ldcs(reinterpret_cast<const float4*>(ap));
        L.x01 = __ldg(reinterpret_cast<const double2*>(xp));
        L.x23 = __ldg(reinterpret_cast<const double2*>(xp + 2));
    } else {
        L.a = make_float4(__ldcs(ap), __ldcs(ap + 1), __ldcs(ap + 2), 0.f);
        L.x01 = make_double2(__ldg(xp), __ldg(xp + 1));
        L.x23 = make_double2(__ldg(xp + 2), 0.0);
    }
}
template <int BS> __device__ __forceinline__ double dotSlot(const RowLoadF<BS>& L) {
    return (double)L.a.x * L.x01.x + (double)L.a.y * L.x01.y + (double)L.a.z * L.x23.x + (double)L.a.w * L.x23.y;
}
// fp32 matrix AND fp32 vectors (multigrid cycle under the flexible GMRES): one 16-byte load each for the block row and for
// the node's x, four FFMA, no conversions
template <int BS> struct RowLoadFF {
    float4 a, x;
};
template <int BS>
__device__ __forceinline__ void loadSlot(const float* __restrict__ Aval, const float* __restrict__ x, int blk, int col, int r,
                                         RowLoadFF<BS>& L) {
    const float* ap = Aval + ((size_t)blk * BS + r) * BS;
    const float* xp = x + (size_t)col * BS;
    if constexpr (BS == 4) {
        L.a = __ldcs(reinterpret_cast<const float4*>(ap));
        L.x = __ldg(reinterpret_cast<const float4*>(xp));
    } else {
        L.a = make_float4(__ldcs(ap), __ldcs(ap + 1), __ldcs(ap + 2), 0.f);
        L.x = make_float4(__ldg(xp), __ldg(xp + 1), __ldg(xp + 2), 0.f);
    }
}
template <int BS> __device__ __forceinline__ float dotSlot(const RowLoadFF<BS>& L) {
    return L.a.x * L.x.x + L.a.y * L.x.y + L.a.z * L.x.z + L.a.w * L.x.w;
}
template <int BS, typename AT, typename VT = double> struct RowLoadOf {
    using type = RowLoad<BS>;
};
template <int BS> struct RowLoadOf<BS, float, double> {
    using type = RowLoadF<BS>;
};
template <int BS> struct RowLoadOf<BS, float, float> {
    using type = RowLoadFF<BS>;
};
template <int BS> __device__ __forceinline__ double dotSlot(const RowLoad<BS>& L) {
    return L.a01.x * L.x01.x + L.a01.y * L.x01.y + L.a23.x * L.x23.x + L.a23.y * L.x23.y;
}

// Epilogue of a block row.  EPI_PLAIN: y = rowScale .* (A x) with the fused dots (BiCGSTAB).  EPI_RESID: y = b - A x.
// EPI_SMOOTH: damped node-block Jacobi sweep y = x + Dw_i (b_i - (A x)_i), Dw_i = omega A_ii^-1 (multigrid smoother).
enum { EPI_PLAIN = 0, EPI_RESID = 1, EPI_SMOOTH = 2 };
template <typename VT> struct SpmvEpiT {
    const VT* b = nullptr;
    const VT* Dw = nullptr;
};
using SpmvEpi = SpmvEpiT<double>;

// SL: block slots per lane group held in registers (8 groups x SL blocks per row on the pipelined path; 2 covers the
// <= 16 blocks of a tetrahedral mesh row, 4 the ~27 of an aggregated level)
template <int BS, int MINB, int EPI = EPI_PLAIN, typename AT = double, int SL = 2, typename VT = double>
__global__ void __launch_bounds__(256, MINB) k_spmv(int nNodes, const int* __restrict__ nbrPtr, const int* __restrict__ nbr,
                                              const AT* __restrict__ Aval, const VT* __restrict__ x,
                                              VT* __restrict__ y, const double* __restrict__ w1, double* partial,
                                              int stride, int slotYW, int slotYY, const double* __restrict__ scal,
                                              const double* __restrict__ rowScale, const SpmvEpiT<VT> epi = SpmvEpiT<VT>()) {
    static_assert(EPI != EPI_PLAIN || sizeof(VT) == 8, "the Krylov SpMV (fused dots, row scale) is fp64");
    // Pipeline (round 2): the loads of row i+1 (A rows, x gathers) are issued right after the products of row i, BEFORE its
    // shuffle reduction and epilogue, into the registers row i has just released -- the reduction / epilogue latency
    // (ncu: 14 % of the stall samples on the late Dw loads, 7 dependent shuffles) hides under the next row's memory latency.
    // For that the column indices run two rows ahead and the row pointers three.
    using Row = typename RowLoadOf<BS, AT, VT>::type;
    const int lane = threadIdx.x & 31, grp = lane >> 2, r = lane & 3;
    const int warpsPerBlock = blockDim.x >> 5;
    const int gw = blockIdx.x * warpsPerBlock + (threadIdx.x >> 5), nw = gridDim.x * warpsPerBlock;
    double accYW = 0, accYY = 0;
    const bool frozen = scal && scal[SC_DONE] != 0.0;
    if (!frozen && gw < nNodes) {
        int i = gw;
        int pb0, pb1, qb0 = 0, qb1 = 0, rb0 = 0, rb1 = 0;  // block ranges of rows i, i+nw, i+2nw
        pb0 = __ldg(nbrPtr + i);
        pb1 = __ldg(nbrPtr + i + 1);
        if (i + nw < nNodes) {
            qb0 = __ldg(nbrPtr + i + nw);
            qb1 = __ldg(nbrPtr + i + nw + 1);
        }
        if (i + 2 * nw < nNodes) {
            rb0 = __ldg(nbrPtr + i + 2 * nw);
            rb1 = __ldg(nbrPtr + i + 2 * nw + 1);
        }
        Row Ls[SL];
        if (r < BS) {
#pragma unroll
            for (int k = 0; k < SL; ++k)
                if (grp + 8 * k < pb1 - pb0) loadSlot<BS>(Aval, x, pb0 + grp + 8 * k, __ldg(nbr + pb0 + grp + 8 * k), r, Ls[k]);
        }
        int cs[SL];  // column indices of row i+nw
#pragma unroll
        for (int k = 0; k < SL; ++k) cs[k] = (grp + 8 * k < qb1 - qb0) ? __ldg(nbr + qb0 + grp + 8 * k) : -1;
        for (; i < nNodes; i += nw) {
            int fb0 = 0, fb1 = 0;
            if (i + 3 * nw < nNodes) {
                fb0 = __ldg(nbrPtr + i + 3 * nw);
                fb1 = __ldg(nbrPtr + i + 3 * nw + 1);
            }
            int ds[SL];  // column indices of row i+2nw
#pragma unroll
            for (int k = 0; k < SL; ++k) ds[k] = (grp + 8 * k < rb1 - rb0) ? __ldg(nbr + rb0 + grp + 8 * k) : -1;
            const int nb = pb1 - pb0;
            // epilogue operands are fetched with the row, not after its reduction
            const bool epiLane = grp == 0 && r < BS;
            const size_t oe = (size_t)i * BS + (r < BS ? r : 0);
            VT e0 = 0, e1 = 0;
            if (epiLane) {
                if constexpr (EPI == EPI_PLAIN) {
                    if (rowScale) e0 = (VT)rowScale[oe];
                    if (slotYW >= 0) e1 = (VT)w1[oe];
                } else if constexpr (EPI == EPI_RESID) {
                    e0 = epi.b[oe];
                } else {
                    e0 = epi.b[oe];
                    e1 = x[oe];
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(epi.Dw + oe * BS));  // the row of Dw: no registers held
                }
            }
            VT acc = 0;
            if (r < BS) {
#pragma unroll
                for (int k = 0; k < SL; ++k)
                    if (grp + 8 * k < nb) acc += dotSlot<BS>(Ls[k]);
                for (int s = grp + 8 * SL; s < nb; s += 8) {  // rows with more than 8*SL blocks (rare)
                    Row L;
                    loadSlot<BS>(Aval, x, pb0 + s, __ldg(nbr + pb0 + s), r, L);
                    acc += dotSlot<BS>(L);
                }
                // next row's operands, into the registers this row has released
#pragma unroll
                for (int k = 0; k < SL; ++k)
                    if (cs[k] >= 0) loadSlot<BS>(Aval, x, qb0 + grp + 8 * k, cs[k], r, Ls[k]);
            }
            acc += __shfl_xor_sync(0xffffffffu, acc, 4);
            acc += __shfl_xor_sync(0xffffffffu, acc, 8);
            acc += __shfl_xor_sync(0xffffffffu, acc, 16);
            if constexpr (EPI == EPI_PLAIN) {
                if (epiLane) {
                    if (rowScale) acc *= e0;  // symmetric Jacobi scaling: y = S A x
                    y[oe] = acc;
                    if (slotYW >= 0) accYW += acc * e1;
                    if (slotYY >= 0) accYY += acc * acc;
                }
            } else if constexpr (EPI == EPI_RESID) {
                if (epiLane) y[oe] = e0 - acc;
            } else {
                const VT res = epiLane ? e0 - acc : (VT)0;
                VT upd = 0;
#pragma unroll
                for (int cc = 0; cc < BS; ++cc) {
                    const VT rc = __shfl_sync(0xffffffffu, res, cc);
                    if (epiLane) upd += epi.Dw[oe * BS + cc] * rc;
                }
                if (epiLane) y[oe] = e1 + upd;
            }
            pb0 = qb0, pb1 = qb1, qb0 = rb0, qb1 = rb1, rb0 = fb0, rb1 = fb1;
#pragma unroll
            for (int k = 0; k < SL; ++k) cs[k] = ds[k];
        }
    }
    if (slotYW >= 0 || slotYY >= 0) {
        double v[2] = {accYW, accYY};
        const int slots[2] = {slotYW >= 0 ? slotYW : PS_AUX, slotYY >= 0 ? slotYY : PS_AUX + 1};
        blockSumStore<2>(v, partial, stride, slots);
    }
}

// ---- SpMV with an asynchronous shared-memory ring for A (4x4 blocks) ------------------------------------------------------
// The register-pipelined kernel above keeps ONE block row (~1-2 KB) of A in flight per warp; with the 24-32 warps an SM
// holds that is below the ~45 KB per SM that HBM3e needs to stream at full rate, so it is latency-bound -- clearly so once
// A is fp32 (multigrid levels).  Here every warp owns a DEPTH-deep ring of row buffers filled with cp.async (16 B per lane,
// no registers, no copy-engine descriptors): DEPTH-1 rows per warp stay in flight while the current one is multiplied.
// Column indices and the x gathers (L2-resident) stay on the LDG path.  Rows with more than ROWCAP blocks: the tail beyond
// ROWCAP is read directly.
__device__ __forceinline__ void cpAsync16(void* smemDst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smemDst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cpAsyncWait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int EPI, typename AT, int DEPTH, int MINB, int ROWCAP = 16>
__global__ void __launch_bounds__(256, MINB) k_spmv_ring(int nNodes, const int* __restrict__ nbrPtr, const int* __restrict__ nbr,
                                                   const AT* __restrict__ Aval, const double* __restrict__ x,
                                                   double* __restrict__ y, const SpmvEpi epi) {
    constexpr int BS = 4, BB = 16;
    constexpr int CPB = BB * (int)sizeof(AT) / 16;       // 16-byte chunks per block: 4 (fp32) | 8 (fp64)
    constexpr int ROWBYTES = ROWCAP * BB * (int)sizeof(AT);
    extern __shared__ __align__(16) unsigned char ringAll[];
    const int lane = threadIdx.x & 31, grp = lane >> 2, r = lane & 3;
    const int wib = threadIdx.x >> 5, warpsPerBlock = blockDim.x >> 5;
    const int gw = blockIdx.x * warpsPerBlock + wib, nw = gridDim.x * warpsPerBlock;
    unsigned char* ring = ringAll + (size_t)wib * DEPTH * ROWBYTES;

    // block ranges of rows i, i+nw, ..., i+DEPTH*nw (the last one is only a prefetch of the pointers)
    int pb[DEPTH + 1], pe[DEPTH + 1];
#pragma unroll
    for (int d = 0; d <= DEPTH; ++d) {
        const int row = gw + d * nw;
        pb[d] = pe[d] = 0;
        if (row < nNodes) {
            pb[d] = __ldg(nbrPtr + row);
            pe[d] = __ldg(nbrPtr + row + 1);
        }
    }
    auto issue = [&](int b0, int b1, int slot) {
        const int nchunk = min(b1 - b0, ROWCAP) * CPB;
        const unsigned char* src = reinterpret_cast<const unsigned char*>(Aval + (size_t)b0 * BB);
        unsigned char* dst = ring + (size_t)slot * ROWBYTES;
        for (int e = lane; e < nchunk; e += 32) cpAsync16(dst + (size_t)e * 16, src + (size_t)e * 16);
        cpAsyncCommit();
    };
#pragma unroll
    for (int d = 0; d < DEPTH - 1; ++d) issue(pb[d], pe[d], d);
    int cs[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) cs[k] = (grp + 8 * k < pe[0] - pb[0]) ? __ldg(nbr + pb[0] + grp + 8 * k) : -1;
    int slot = 0;
    for (int i = gw; i < nNodes; i += nw) {
        // keep the ring full: row i + (DEPTH-1) nw goes into the slot that was consumed last iteration
        issue(pb[DEPTH - 1], pe[DEPTH - 1], (slot + DEPTH - 1) % DEPTH);
        int nb0 = 0, nb1 = 0;
        if (i + (DEPTH + 1) * nw < nNodes) {
            nb0 = __ldg(nbrPtr + i + (DEPTH + 1) * nw);
            nb1 = __ldg(nbrPtr + i + (DEPTH + 1) * nw + 1);
        }
        int ds[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) ds[k] = (grp + 8 * k < pe[1] - pb[1]) ? __ldg(nbr + pb[1] + grp + 8 * k) : -1;
        const int b0 = pb[0], nb = pe[0] - pb[0];
        // x gathers first (longest latency), then wait for this row's A
        double2 x01[2], x23[2];
#pragma unroll
        for (int k = 0; k < 2; ++k)
            if (cs[k] >= 0) {
                const double* xp = x + (size_t)cs[k] * BS;
                x01[k] = __ldg(reinterpret_cast<const double2*>(xp));
                x23[k] = __ldg(reinterpret_cast<const double2*>(xp + 2));
            }
        const size_t o = (size_t)i * BS + r;
        double e0 = 0, e1 = 0;
        if (grp == 0) {
            if constexpr (EPI != EPI_PLAIN) e0 = epi.b[o];
            if constexpr (EPI == EPI_SMOOTH) {
                e1 = x[o];
                asm volatile("prefetch.global.L1 [%0];" ::"l"(epi.Dw + o * BS));
            }
        }
        cpAsyncWait<DEPTH - 1>();
        __syncwarp();
        const unsigned char* buf = ring + (size_t)slot * ROWBYTES;
        double acc = 0;
#pragma unroll
        for (int k = 0; k < 2; ++k)
            if (cs[k] >= 0) {
                const AT* ap = reinterpret_cast<const AT*>(buf) + ((grp + 8 * k) * BS + r) * BS;
                if constexpr (sizeof(AT) == 4) {
                    const float4 v = *reinterpret_cast<const float4*>(ap);
                    acc += (double)v.x * x01[k].x + (double)v.y * x01[k].y + (double)v.z * x23[k].x + (double)v.w * x23[k].y;
                } else {
                    const double2 a01 = *reinterpret_cast<const double2*>(ap), a23 = *reinterpret_cast<const double2*>(ap + 2);
                    acc += a01.x * x01[k].x + a01.y * x01[k].y + a23.x * x23[k].x + a23.y * x23[k].y;
                }
            }
        for (int s = grp + 16; s < nb; s += 8) {  // rows with more than 16 blocks (rare): direct loads
            typename RowLoadOf<BS, AT>::type L;
            loadSlot<BS>(Aval, x, b0 + s, __ldg(nbr + b0 + s), r, L);
            acc += dotSlot<BS>(L);
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        acc += __shfl_xor_sync(0xffffffffu, acc, 8);
        acc += __shfl_xor_sync(0xffffffffu, acc, 16);
        if constexpr (EPI == EPI_RESID) {
            if (grp == 0) y[o] = e0 - acc;
        } else if constexpr (EPI == EPI_SMOOTH) {
            const double res = (grp == 0) ? e0 - acc : 0.0;
            double upd = 0;
#pragma unroll
            for (int cc = 0; cc < BS; ++cc) {
                const double rc = __shfl_sync(0xffffffffu, res, cc);
                if (grp == 0) upd += epi.Dw[o * BS + cc] * rc;
            }
            if (grp == 0) y[o] = e1 + upd;
        } else {
            if (grp == 0) y[o] = acc;
        }
        __syncwarp();  // every lane is done with this slot before the next iteration refills it
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) pb[d] = pb[d + 1], pe[d] = pe[d + 1];
        pb[DEPTH] = nb0, pe[DEPTH] = nb1;
        cs[0] = ds[0], cs[1] = ds[1];
        slot = (slot + 1) % DEPTH;
    }
    cpAsyncWait<0>();
}

}  // namespace
