// k x k pieces shared by small_ops.cu and the fused iteration kernels (iterate.cu): the
// GPNH dictionary solve matrix and the scalar bookkeeping of a finished sub-step.
#pragma once

#include "cdr_common.cuh"

namespace cdr {

// ======================================================================
// GPNH dictionary step: P = pinv(ZtZ / T + lambda * G_W) / T
// ======================================================================
// One CTA; cyclic Jacobi eigen-decomposition with a round-robin (parallel)
// ordering: k/2 disjoint rotations per round.  The pseudo-inverse drops
// eigenvalues below eps * k * max|eig|, the cut-off numpy.linalg.lstsq applies
// with rcond=None (gpnh_convex_coding.py:224).
constexpr int kJacLd = CDR_MAX_COMPONENTS + 1;

// In-place Gauss-Jordan inversion of the symmetric positive definite k x k matrix A (shared
// memory, leading dimension kJacLd) by one warp: lane c keeps column c in registers, pivot p
// is broadcast from lane p.  Writes scale * A^-1 to P (row-major k x k) and returns 1, or
// returns 0 (P untouched) when a pivot is not above 1e-10 * max diag.  All 32 lanes call this.
constexpr int CDR_MAX_GJ = 32;                   // k <= 32 on the fast path (one lane per column)
template <int KMAX>
__device__ __forceinline__ int spd_inverse_warp(const double* A, int k, double scale,
                                                double* __restrict__ P)
{
    const int lane = threadIdx.x & 31;
    if (k > KMAX) return 0;
    double col[KMAX];
#pragma unroll
    for (int i = 0; i < KMAX; ++i) col[i] = (i < k && lane < k) ? A[i * kJacLd + lane] : 0.0;
    double dmax = (lane < k) ? A[lane * kJacLd + lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dmax = fmax(dmax, __shfl_xor_sync(CDR_FULL_MASK, dmax, o));
    const double thr = 1e-10 * dmax;
    bool ok = dmax > 0.0;
#pragma unroll
    for (int p = 0; p < KMAX; ++p) {
        if (p < k && ok) {                                       // warp-uniform
            const double piv = __shfl_sync(CDR_FULL_MASK, col[p], p);
            if (!(piv > thr)) {
                ok = false;
            } else {
                const double inv = 1.0 / piv;
                const double rp = (lane == p) ? inv : col[p] * inv;       // row p of the result
#pragma unroll
                for (int i = 0; i < KMAX; ++i) {
                    if (i != p && i < k) {
                        const double f = __shfl_sync(CDR_FULL_MASK, col[i], p);
                        col[i] = (lane == p) ? -f * inv : fma(-f, rp, col[i]);
                    }
                }
                col[p] = rp;
            }
        }
    }
    if (!ok) return 0;
    if (lane < k) {
#pragma unroll
        for (int i = 0; i < KMAX; ++i)
            if (i < k) P[i * k + lane] = col[i] * scale;
    }
    return 1;
}

// All threads of the CTA call this together (blockDim.x a multiple of 32); jac_sm holds
// 2 * kmax * kJacLd doubles of shared memory, kmax >= k.  ZtZ may be global or shared.
// GJMAX: the widest Gauss-Jordan instantiation to compile in (8 / 16 in the fused iteration
// kernels, whose k never exceeds that; the fully unrolled 32-wide variant is 2000 shuffles of
// code).
template <int GJMAX = CDR_MAX_GJ>
__device__ __forceinline__ void solve_matrix_cta(const double* ZtZ, int k, int kmax, double inv_n,
                                                 double lambda_W, double gw_prefactor,
                                                 double* __restrict__ P, double* jac_sm)
{
    double* A = jac_sm;
    double* V = jac_sm + kmax * kJacLd;
    __shared__ double rc[CDR_MAX_COMPONENTS / 2], rs[CDR_MAX_COMPONENTS / 2];
    __shared__ int rp[CDR_MAX_COMPONENTS / 2], rq[CDR_MAX_COMPONENTS / 2];
    __shared__ int rotated;
    __shared__ double inv_eig[CDR_MAX_COMPONENTS];

    __shared__ int chol_ok;
    __shared__ double chol_tmp[CDR_MAX_COMPONENTS];

    const int tid = threadIdx.x;
    for (int idx = tid; idx < k * k; idx += blockDim.x) {
        const int i = idx / k, j = idx % k;
        double v = ZtZ[idx] * inv_n;
        if (k > 1) v += lambda_W * gw_prefactor * ((i == j ? (double)k : 0.0) - 1.0);
        A[i * kJacLd + j] = v;
    }
    __syncthreads();

    // ---- fast path: the matrix is symmetric positive definite and reasonably conditioned
    // (the usual case): P = A^-1 by Gauss-Jordan elimination without pivoting in one warp,
    // about a microsecond instead of ~50 for the Jacobi sweeps.  The pivots are those of the
    // Cholesky / LDL' factorisation; any of them below 1e-10 * max diag falls through to the
    // pseudo-inverse, which reproduces lstsq's minimum-norm solution for singular Z'Z.
    if (k <= GJMAX) {
        if (tid < 32) {
            int ok = 0;
            if (k <= 8) {
                ok = spd_inverse_warp<8>(A, k, inv_n, P);
            } else if constexpr (GJMAX > 8) {
                if (k <= 16) {
                    ok = spd_inverse_warp<16>(A, k, inv_n, P);
                } else if constexpr (GJMAX > 16) {
                    ok = spd_inverse_warp<GJMAX>(A, k, inv_n, P);
                }
            }
            if (tid == 0) chol_ok = ok;
        }
        __syncthreads();
        if (chol_ok) return;                      // CTA-uniform
    } else {
        // wider matrices: Cholesky in shared memory (V), same pivot test
        for (int idx = tid; idx < k * k; idx += blockDim.x)
            V[(idx / k) * kJacLd + idx % k] = A[(idx / k) * kJacLd + idx % k];
        __syncthreads();
        if (tid < 32) {
            const int lane = tid;
            double dmax = 0.0;
            for (int i = 0; i < k; ++i) dmax = fmax(dmax, V[i * kJacLd + i]);
            const double thr = 1e-10 * dmax;
            bool ok = dmax > 0.0;
            for (int j = 0; j < k && ok; ++j) {
                for (int i = j + lane; i < k; i += 32) {
                    double sacc = V[i * kJacLd + j];
                    for (int q = 0; q < j; ++q) sacc = fma(-V[i * kJacLd + q], V[j * kJacLd + q], sacc);
                    chol_tmp[i] = sacc;
                }
                __syncwarp();
                const double dj = chol_tmp[j];
                if (!(dj > thr)) {
                    ok = false;
                } else {
                    const double root = sqrt(dj);
                    for (int i = j + lane; i < k; i += 32)
                        V[i * kJacLd + j] = (i == j) ? root : chol_tmp[i] / root;
                }
                __syncwarp();
            }
            if (ok) {
                // columns of L^-1 by forward substitution (one column per lane), stored in A's
                // upper part is not safe (A may still be needed) -> reuse chol-free rows of V:
                // L^-1 overwrites the strictly-upper triangle + a separate diagonal pass
                for (int c = lane; c < k; c += 32) {
                    // y = L^-1 e_c, kept in the upper triangle V[c][i] (i >= c)
                    double ycc = 1.0 / V[c * kJacLd + c];
                    for (int i = c + 1; i < k; ++i) {
                        double sacc = V[i * kJacLd + c] * ycc;
                        for (int q = c + 1; q < i; ++q) sacc = fma(V[i * kJacLd + q], V[c * kJacLd + q], sacc);
                        V[c * kJacLd + i] = -sacc / V[i * kJacLd + i];
                    }
                    chol_tmp[c] = ycc;
                }
                __syncwarp();
            }
            if (lane == 0) chol_ok = ok ? 1 : 0;
        }
        __syncthreads();
        if (chol_ok) {
            // (L^-1)[i][c] = V[c][i] for i > c, chol_tmp[c] for i == c;  P = L^-T L^-1
            for (int idx = tid; idx < k * k; idx += blockDim.x) {
                const int a = idx / k, b = idx % k;
                const int lo = a > b ? a : b;
                double sacc = 0.0;
                for (int i = lo; i < k; ++i) {
                    const double la = (i == a) ? chol_tmp[a] : V[a * kJacLd + i];
                    const double lb = (i == b) ? chol_tmp[b] : V[b * kJacLd + i];
                    sacc = fma(la, lb, sacc);
                }
                P[idx] = sacc * inv_n;
            }
            return;                                   // CTA-uniform
        }
    }
    for (int idx = tid; idx < k * k; idx += blockDim.x) {
        const int i = idx / k, j = idx % k;
        V[i * kJacLd + j] = (i == j) ? 1.0 : 0.0;
    }
    __syncthreads();

    const int kk = (k + 1) & ~1;        // even number of players
    const int half = kk / 2;
    for (int sweep = 0; sweep < 40; ++sweep) {
        if (tid == 0) rotated = 0;
        __syncthreads();
        for (int round = 0; round < kk - 1; ++round) {
            if (tid < half) {
                int a, b;
                if (tid == 0) {
                    a = kk - 1;
                    b = round;
                } else {
                    a = (round + tid) % (kk - 1);
                    b = (round - tid + (kk - 1)) % (kk - 1);
                }
                const int p = min(a, b), q = max(a, b);
                double c = 1.0, s = 0.0;
                if (q < k) {
                    const double apq = A[p * kJacLd + q];
                    const double app = A[p * kJacLd + p], aqq = A[q * kJacLd + q];
                    if (apq != 0.0 && fabs(apq) > 2.220446049250313e-16 * sqrt(fabs(app * aqq))) {
                        const double theta = (aqq - app) / (2.0 * apq);
                        const double t = copysign(1.0, theta) / (fabs(theta) + sqrt(theta * theta + 1.0));
                        c = 1.0 / sqrt(t * t + 1.0);
                        s = t * c;
                        rotated = 1;
                    }
                }
                rp[tid] = p;
                rq[tid] = q;
                rc[tid] = c;
                rs[tid] = s;
            }
            __syncthreads();
            // columns: A <- A J, V <- V J
            for (int idx = tid; idx < half * k; idx += blockDim.x) {
                const int m = idx / k, i = idx % k;
                const int p = rp[m], q = rq[m];
                if (q < k && rs[m] != 0.0) {
                    const double c = rc[m], s = rs[m];
                    const double aip = A[i * kJacLd + p], aiq = A[i * kJacLd + q];
                    A[i * kJacLd + p] = c * aip - s * aiq;
                    A[i * kJacLd + q] = s * aip + c * aiq;
                    const double vip = V[i * kJacLd + p], viq = V[i * kJacLd + q];
                    V[i * kJacLd + p] = c * vip - s * viq;
                    V[i * kJacLd + q] = s * vip + c * viq;
                }
            }
            __syncthreads();
            // rows: A <- J' A
            for (int idx = tid; idx < half * k; idx += blockDim.x) {
                const int m = idx / k, j = idx % k;
                const int p = rp[m], q = rq[m];
                if (q < k && rs[m] != 0.0) {
                    const double c = rc[m], s = rs[m];
                    const double apj = A[p * kJacLd + j], aqj = A[q * kJacLd + j];
                    A[p * kJacLd + j] = c * apj - s * aqj;
                    A[q * kJacLd + j] = s * apj + c * aqj;
                }
            }
            __syncthreads();
        }
        if (rotated == 0) break;
        __syncthreads();
    }

    if (tid == 0) {
        double emax = 0.0;
        for (int i = 0; i < k; ++i) emax = fmax(emax, fabs(A[i * kJacLd + i]));
        const double cut = 2.220446049250313e-16 * (double)k * emax;
        for (int i = 0; i < k; ++i) {
            const double e = A[i * kJacLd + i];
            inv_eig[i] = (fabs(e) > cut) ? 1.0 / e : 0.0;
        }
    }
    __syncthreads();
    for (int idx = tid; idx < k * k; idx += blockDim.x) {
        const int a = idx / k, b = idx % k;
        double s = 0.0;
        for (int i = 0; i < k; ++i) s = fma(V[a * kJacLd + i] * inv_eig[i], V[b * kJacLd + i], s);
        P[idx] = s * inv_n;
    }
}

// ======================================================================
// GPNH cost bookkeeping (gpnh_convex_coding.py:352-399)
// ======================================================================
__device__ __forceinline__ bool cost_increased(double old_cost, double new_cost, double tol)
{
    // archetypal_analysis.py:167-174 / gpnh_convex_coding.py:146-153
    return (new_cost > old_cost) && (fabs(new_cost - old_cost) > tol);
}

__device__ __forceinline__ bool stop_rule(int rule, double old_cost, double new_cost, double tol)
{
    // archetypal_analysis.py:177-197
    const double delta = new_cost - old_cost;
    if (rule == 0) return fabs(delta) < tol;
    const double mx = fmax(fabs(new_cost), fabs(old_cost));
    return fabs(delta / mx) < tol;
}

__device__ __forceinline__ void finish_sub_step(cdr_loop_state* st, double* cost_deltas, double cost, int stage,
                                int end_of_iteration)
{
    st->cost = cost;
    if (stage == 0) return;                        // initial cost only
    if (st->require_monotone && cost_increased(st->old_cost, cost, st->tolerance)) {
        st->error_stage = stage;
        st->done = 1;
        return;
    }
    if (end_of_iteration) {
        const int it = st->n_iter;
        if (cost_deltas) cost_deltas[it] = cost - st->old_cost;
        st->n_iter = it + 1;
        if (stop_rule(st->stopping_rule, st->old_cost, cost, st->tolerance)) {
            st->converged = 1;
            st->done = 1;
        } else if (it + 1 >= st->max_iterations) {
            st->done = 1;
        }
    }
}

}  // namespace cdr
