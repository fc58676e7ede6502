// Simplex-projection building blocks shared by simplex.cu, qp.cu and aa_steps.cu.
//
// The reference (simplex_projection.py:13-27) sorts the vector and scans for the
// threshold t with  x_i - t > 0  on the support and  sum(max(x - t, 0)) = 1.
// The same t is the fixed point of Michelot's iteration
//     t <- (sum_{x_i > t} x_i - 1) / #{x_i > t},
// which needs only reductions (no sort) and maps onto warps / CTAs directly;
// t is also  max_m (sum of the m largest - 1) / m,  which the 8-wide per-sample
// version below uses.
#pragma once

#include "cdr_common.cuh"

namespace cdr {

// CTA-wide threshold of the n values work[i * stride].  scratch: 64 doubles of
// shared memory.  Deterministic (fixed reduction order).  All threads return t.
__device__ __forceinline__ double block_simplex_threshold(const double* work, long stride, int n,
                                                          double* scratch)
{
    // Start from t0 = max((sum - 1) / n, max - 1): both are lower bounds of the threshold t*
    // (the first is Michelot's own first step; the largest element alone contributes
    // max - t* <= 1).  When the projection is nearly a vertex -- the usual case for
    // P(x - alpha g) in the dictionary SPG -- this saves most of the ~11 iterations the
    // plain start needs (measured on the HadISST-shaped problem: 2-4 instead of 10-12).
    double r[1] = {0.0};
    double mx = -INFINITY;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double v = work[(long)i * stride];
        r[0] += v;
        mx = fmax(mx, v);
    }
    block_sum<1>(r, scratch);
    mx = block_max(mx, scratch);
    double t = fmax((r[0] - 1.0) / (double)n, mx - 1.0);
    int count_prev = n;
    for (int it = 0; it < 4096; ++it) {
        double q[2] = {0.0, 0.0};
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const double v = work[(long)i * stride];
            if (v > t) {
                q[0] += v;
                q[1] += 1.0;
            }
        }
        block_sum<2>(q, scratch);
        const int cnt = (int)q[1];
        if (cnt == count_prev || cnt <= 0) break;
        // monotone in exact arithmetic; the fmax keeps the active set shrinking
        // under rounding so the loop terminates
        t = fmax(t, (q[0] - 1.0) / q[1]);
        count_prev = cnt;
    }
    return t;
}

// Threshold for one vector of <= 8 values spread over an aligned group of 8
// lanes (one value per lane, absent components = -inf).  Uses
// t = max_i (sum_{v_j >= v_i} v_j - 1) / #{v_j >= v_i}.  All lanes of the warp
// must call this together.
__device__ __forceinline__ double group8_simplex_threshold(double v)
{
    double s = 0.0, n = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const double vj = __shfl_sync(CDR_FULL_MASK, v, j, 8);
        if (vj >= v) {
            s += vj;
            n += 1.0;
        }
    }
    double t = (s - 1.0) / n;          // -inf on absent lanes
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) t = fmax(t, __shfl_xor_sync(CDR_FULL_MASK, t, o, 8));
    return t;
}

// Michelot threshold for a vector of 8*KPL values, KPL per lane of an aligned
// group of 8 lanes (absent components = -inf).  The four groups of a warp run in
// lock step; a finished group keeps executing the shuffles with frozen state.
template <int KPL>
__device__ __forceinline__ double group8_simplex_threshold_multi(const double (&v)[KPL])
{
    double t = -INFINITY;
    int count_prev = -1;
    bool finished = false;
    while (__any_sync(CDR_FULL_MASK, !finished)) {
        double s = 0.0;
        int c = 0;
#pragma unroll
        for (int r = 0; r < KPL; ++r) {
            if (v[r] > t) {
                s += v[r];
                c += 1;
            }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            s += __shfl_xor_sync(CDR_FULL_MASK, s, o, 8);
            c += __shfl_xor_sync(CDR_FULL_MASK, c, o, 8);
        }
        if (!finished) {
            if (c == count_prev || c <= 0) {
                finished = true;
            } else {
                t = fmax(t, (s - 1.0) / (double)c);
                count_prev = c;
            }
        }
    }
    return t;
}

template <int KPL>
__device__ __forceinline__ void group8_project(const double (&v)[KPL], double (&out)[KPL])
{
    double t;
    if constexpr (KPL == 1) {
        t = group8_simplex_threshold(v[0]);
    } else {
        t = group8_simplex_threshold_multi<KPL>(v);
    }
#pragma unroll
    for (int r = 0; r < KPL; ++r) out[r] = fmax(v[r] - t, 0.0);
}

}  // namespace cdr
