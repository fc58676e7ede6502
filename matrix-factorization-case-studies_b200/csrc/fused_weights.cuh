// Pieces shared by the fused weights-update kernels (iterate.cu, iterate_aa.cu): the
// per-sample statistics that ride along with the batched QP, their fixed-order reduction to
// one partial per CTA, and the "last CTA sums everything" tail.
//
// Layout of a statistics vector (NST = KP * KP + 2 doubles, KP = 8 * KPL):
//   [i * KP + j]  sum over samples of z_i z_j   (Z'Z of the new weights)
//   [KP * KP]     first trace term  (caller defined, e.g. tr(B Z_old))
//   [KP * KP + 1] second trace term (e.g. tr(B Z_new))
#pragma once

#include "qp_core.cuh"

namespace cdr {

constexpr int kFusedWarps = 8;                 // warps per CTA of the fused weights kernels
constexpr int kFusedThreads = kFusedWarps * 32;
constexpr int kFusedMaxK = 16;

// fixed-order sum of n values spaced `stride` doubles apart, read around L1, 32 loads in
// flight (the tail of the last batch reads nothing and adds zeros)
__device__ __forceinline__ double strided_sum_cg32(const double* base, long stride, int n)
{
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    for (int i = 0; i < n; i += 32) {
        double v[32];
#pragma unroll
        for (int q = 0; q < 32; ++q) v[q] = (i + q < n) ? __ldcg(base + (long)(i + q) * stride) : 0.0;
#pragma unroll
        for (int q = 0; q < 32; q += 4) {
            a0 += v[q];
            a1 += v[q + 1];
            a2 += v[q + 2];
            a3 += v[q + 3];
        }
    }
    return (a0 + a1) + (a2 + a3);
}

// samples per warp and grid of a fused weights kernel
static inline void fused_grid(int T, int* spw, int* blocks)
{
    // one sample per warp while the batch is small (the kernel is latency bound and no sample
    // should wait in lock step for a slower neighbour), up to four for large batches
    int s = 4;
    if (T <= 148 * 4 * 6) s = 1;
    else if (T <= 148 * 4 * 16) s = 2;
    *spw = s;
    const int warps = (T + s - 1) / s;
    *blocks = (warps + kFusedWarps - 1) / kFusedWarps;
}

// Statistics of the samples of this CTA -> cta_part[blockIdx.x][NST]; returns true in every
// thread of the last CTA of the grid to get here (`ticket` counts arrivals; the caller's tail
// resets it to zero).  x: the solved weights of the lane's components; valid / present as in
// qp_solve; tr_a, tr_b: group-uniform trace contributions of the group's sample.
// wsum: kFusedWarps * NST doubles of shared memory.  All threads of the CTA must call this.
template <int KPL>
__device__ __forceinline__ bool fused_sample_statistics(const double (&x)[KPL],
                                                        const bool (&present)[KPL], bool valid,
                                                        int k, double tr_a, double tr_b,
                                                        double* wsum, double* cta_part,
                                                        unsigned int* ticket)
{
    constexpr int KP = 8 * KPL;
    constexpr int NST = KP * KP + 2;
    __shared__ int is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane & 7, q = lane >> 3;

    double zz[KP][KPL];
#pragma unroll
    for (int i = 0; i < KP; ++i) {
        const double zi = __shfl_sync(CDR_FULL_MASK, x[i % KPL], (lane & 24) + i / KPL);
#pragma unroll
        for (int r = 0; r < KPL; ++r) zz[i][r] = (valid && present[r] && i < k) ? zi * x[r] : 0.0;
    }
    if (!valid) {
        tr_a = 0.0;
        tr_b = 0.0;
    }
    // the (up to four) samples of the warp, in group order
#pragma unroll
    for (int i = 0; i < KP; ++i)
#pragma unroll
        for (int r = 0; r < KPL; ++r) {
            const double v = zz[i][r];
            const double v0 = __shfl_sync(CDR_FULL_MASK, v, g), v1 = __shfl_sync(CDR_FULL_MASK, v, g + 8);
            const double v2 = __shfl_sync(CDR_FULL_MASK, v, g + 16), v3 = __shfl_sync(CDR_FULL_MASK, v, g + 24);
            zz[i][r] = ((v0 + v1) + v2) + v3;
        }
    {
        const double a0 = __shfl_sync(CDR_FULL_MASK, tr_a, 0), a1 = __shfl_sync(CDR_FULL_MASK, tr_a, 8);
        const double a2 = __shfl_sync(CDR_FULL_MASK, tr_a, 16), a3 = __shfl_sync(CDR_FULL_MASK, tr_a, 24);
        const double b0 = __shfl_sync(CDR_FULL_MASK, tr_b, 0), b1 = __shfl_sync(CDR_FULL_MASK, tr_b, 8);
        const double b2 = __shfl_sync(CDR_FULL_MASK, tr_b, 16), b3 = __shfl_sync(CDR_FULL_MASK, tr_b, 24);
        tr_a = ((a0 + a1) + a2) + a3;
        tr_b = ((b0 + b1) + b2) + b3;
    }
    if (q == 0) {
#pragma unroll
        for (int i = 0; i < KP; ++i)
#pragma unroll
            for (int r = 0; r < KPL; ++r) wsum[warp * NST + i * KP + g * KPL + r] = zz[i][r];
        if (g == 0) {
            wsum[warp * NST + KP * KP] = tr_a;
            wsum[warp * NST + KP * KP + 1] = tr_b;
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < NST; e += blockDim.x) {
        double s = wsum[e];
#pragma unroll
        for (int w = 1; w < kFusedWarps; ++w) s += wsum[w * NST + e];
        __stcg(cta_part + (long)blockIdx.x * NST + e, s);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    const bool last = is_last != 0;
    if (last) __threadfence();
    return last;
}

// Last CTA: fin[0..NST) = sum over the CTAs of the grid of cta_part, fixed order.
// fin: fused_fin_doubles(NST) doubles of shared memory.  All threads of the CTA must call this.
template <int KPL>
__device__ __forceinline__ void fused_final_sum(const double* cta_part, double* fin)
{
    constexpr int KP = 8 * KPL;
    constexpr int NST = KP * KP + 2;
    constexpr int PH = (kFusedThreads / NST) > 0 ? (kFusedThreads / NST) : 1;   // 3 or 1
    double* phs = fin + NST;                                                     // [PH][NST]
    const int nblk = gridDim.x;
    if (PH > 1) {
        const int e = (int)threadIdx.x % NST, ph = (int)threadIdx.x / NST;
        if (ph < PH)
            phs[ph * NST + e] = strided_sum_cg32(cta_part + (long)ph * NST + e, (long)PH * NST,
                                                 (nblk - ph + PH - 1) / PH);
    } else {
        for (int e = threadIdx.x; e < NST; e += blockDim.x)
            phs[e] = strided_sum_cg32(cta_part + e, NST, nblk);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < NST; e += blockDim.x) {
        double s = phs[e];
#pragma unroll
        for (int ph = 1; ph < PH; ++ph) s += phs[ph * NST + e];
        fin[e] = s;
    }
    __syncthreads();
}
// doubles of shared memory behind `fin` for a statistics vector of nst entries
static inline size_t fused_fin_doubles(int nst) { return (size_t)4 * nst; }

template <auto Kern>
static int ensure_dyn_smem(size_t smem)
{
    static size_t configured = 48 * 1024;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(Kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured = smem;
    }
    return 0;
}

static inline size_t align256(size_t n) { return (n + 255) / 256 * 256; }

static inline cdr_small_gram_desc gram_desc(const double* A, long sAi, long sAn, int ka,
                                            const double* B, long sBj, long sBn, int kb, int n,
                                            double* out, int mode)
{
    cdr_small_gram_desc ds;
    ds.A = A; ds.B = B; ds.out = out;
    ds.sAi = sAi; ds.sAn = sAn; ds.sBj = sBj; ds.sBn = sBn;
    ds.ka = ka; ds.kb = kb; ds.n = n; ds.mode = mode; ds.scale = 1.0;
    return ds;
}

#define CDR_TRY(call)                 \
    do {                              \
        const int rc__ = (call);      \
        if (rc__ != 0) return rc__;   \
    } while (0)

}  // namespace cdr
