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
    block_sum_and_max(r[0], mx, scratch);
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

// {RN(1 / n), -n} for n = 0..8: operands of div_small_int
static __device__ double2 kSmallDiv[9] = {{0.0, -0.0},      {1.0, -1.0},       {0.5, -2.0},
                                           {1.0 / 3.0, -3.0}, {0.25, -4.0},      {0.2, -5.0},
                                           {1.0 / 6.0, -6.0}, {1.0 / 7.0, -7.0}, {0.125, -8.0}};

// a / n for an integer 1 <= n <= 8, correctly rounded, i.e. the same double as the IEEE
// division (which costs ~40 dependent instructions here): with y = RN(1 / n), q0 = RN(a y) is
// within one ulp of a / n, the residual a - n q0 is exact in one FMA, and RN(q0 + r y) is
// RN(a / n) (Markstein's theorem; compared with a / n on 6.4e8 random operands).
__device__ __forceinline__ double div_small_int(double a, int n)
{
    const double2 c = kSmallDiv[n];
    const double q0 = a * c.x;
    const double r = fma(c.y, q0, a);
    return fma(r, c.x, q0);
}

// Threshold for one vector of <= 8 values spread over an aligned group of 8 lanes (one value
// per lane, absent components = -inf), as simplex_projection.py:13-27 defines it: with the
// values sorted in decreasing order, t = (sum of the rho largest - 1) / rho for the largest
// rho whose smallest member is still above that quotient.  Lane i forms the candidate
// c_i = (sum of the values >= v_i  -  1) / #{values >= v_i}; the lanes with v_i > c_i are the
// support, rho is their number (one ballot) and t is the candidate of a lane of rank rho.
// Near ties rounding can make the lanes' tests inconsistent (no lane of rank rho in the
// support): then t = max_i c_i, which is the same number in exact arithmetic.
// All lanes of the warp must call this together.
__device__ __forceinline__ double group8_simplex_threshold(double v)
{
    // sum and number of the values >= v (pairwise sum: three dependent additions)
    double term[8];
    int n = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const double vj = __shfl_sync(CDR_FULL_MASK, v, j, 8);
        const bool ge = vj >= v;
        term[j] = ge ? vj : 0.0;
        n += ge ? 1 : 0;
    }
    const double s = ((term[0] + term[1]) + (term[2] + term[3])) +
                     ((term[4] + term[5]) + (term[6] + term[7]));
    const double c = div_small_int(s - 1.0, n);          // NaN on absent lanes
    const bool in = v > c;
    const unsigned shift = threadIdx.x & 24u;
    const unsigned support = (__ballot_sync(CDR_FULL_MASK, in) >> shift) & 0xffu;
    const int rho = __popc(support);
    const unsigned owners = (__ballot_sync(CDR_FULL_MASK, in && n == rho) >> shift) & 0xffu;
    double t = __shfl_sync(CDR_FULL_MASK, c, (int)shift + __ffs((int)owners) - 1);
    if (__any_sync(CDR_FULL_MASK, owners == 0u)) {
        double m = in ? c : ((v > -INFINITY && c == c) ? c : -INFINITY);
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(CDR_FULL_MASK, m, o, 8));
        if (owners == 0u) t = m;
    }
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
    for (int r = 0; r < KPL; ++r) {
        const double d = v[r] - t;
        out[r] = d > 0.0 ? d : 0.0;
    }
}

}  // namespace cdr
