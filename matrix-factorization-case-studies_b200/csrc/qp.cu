// Batched simplex-constrained QP:  for every sample t
//     min_z  1/2 z' A' z + b_t' z   s.t.  z in the probability simplex,
// solved with the non-monotone spectral projected gradient of the reference
// (quad_simplex_spg, spg.py:286-398) -- same step-length rules, same stopping
// tests, same iteration / evaluation limits.
//
// Replaces the serial per-sample loops _gu_update_kernel_aa_weights
// (archetypal_analysis.py:344-366) and _gu_update_gpnh_weights
// (gpnh_convex_coding.py:229-251).
//
// Mapping: an aligned group of 8 lanes owns one sample; each lane owns KPL
// consecutive components (k <= 8*KPL).  A' is kept transposed in shared memory
// (conflict-free, broadcast across the 4 groups of a warp) or, for k <= 8, as
// one row per lane in registers.  The k x k mat-vec, the projections and the
// scalar reductions are all done with 8-wide shuffles; nothing but the sample's
// own b_t and z_t (2*8k bytes) is read from HBM.  The kernel is latency bound
// (dependent fp64 chains), not bandwidth bound: see DESIGN.md.
#include "simplex.cuh"

namespace cdr {

__device__ __forceinline__ double group8_sum(double v)
{
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(CDR_FULL_MASK, v, o, 8);
    return v;
}

__device__ __forceinline__ double group8_max(double v)
{
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(CDR_FULL_MASK, v, o, 8));
    return v;
}

template <int KPL>
struct QpMatVec {
    // y = A' x for the lane's KPL components.  As: transposed A' in shared
    // memory, As[j * KP + c]; arow: the lane's row of A' when KPL == 1.
    static constexpr int KP = 8 * KPL;
    __device__ static __forceinline__ void apply(const double* As, const double (&arow)[8],
                                                 const double (&x)[KPL], double (&y)[KPL], int g)
    {
        if constexpr (KPL == 1) {
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                acc = fma(arow[j], __shfl_sync(CDR_FULL_MASK, x[0], j, 8), acc);
            y[0] = acc;
        } else {
#pragma unroll
            for (int r = 0; r < KPL; ++r) y[r] = 0.0;
#pragma unroll
            for (int src = 0; src < 8; ++src) {
#pragma unroll
                for (int rr = 0; rr < KPL; ++rr) {
                    const double xj = __shfl_sync(CDR_FULL_MASK, x[rr], src, 8);
                    const double* col = As + (src * KPL + rr) * KP + g * KPL;
#pragma unroll
                    for (int r = 0; r < KPL; ++r) y[r] = fma(col[r], xj, y[r]);
                }
            }
        }
    }
};

template <int KPL>
__global__ void __launch_bounds__(128)
qp_batched_kernel(const double* __restrict__ A, const double* __restrict__ alpha,
                  const double* __restrict__ B, long sb_t, long sb_c, double* Z, int T, int k,
                  int spw, cdr_spg_params p, int* n_iter_out, int* n_feval_out,
                  const cdr_flags* flags)
{
    if (is_done(flags)) return;
    constexpr int KP = 8 * KPL;
    extern __shared__ double As[];      // KP x KP, transposed and scaled (unused for KPL == 1)

    if constexpr (KPL > 1) {
        for (int idx = threadIdx.x; idx < KP * KP; idx += blockDim.x) {
            const int j = idx / KP, c = idx % KP;
            double v = 0.0;
            if (j < k && c < k) {
                v = A[(long)c * k + j];
                if (alpha) v *= alpha[c] * alpha[j];
            }
            As[idx] = v;
        }
        __syncthreads();
    }

    const int lane = threadIdx.x & 31;
    const int g = lane & 7;
    const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    // spw samples per warp (1, 2 or 4): small batches use one sample per warp so that no
    // sample waits in lock step for a slower neighbour (the kernel is latency bound)
    const int t_raw = warp_global * spw + (lane >> 3);
    const bool valid = ((lane >> 3) < spw) && (t_raw < T);
    const long t = valid ? t_raw : (T - 1);

    double arow[8];
    if constexpr (KPL == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            double v = 0.0;
            if (g < k && j < k) {
                v = A[(long)g * k + j];
                if (alpha) v *= alpha[g] * alpha[j];
            }
            arow[j] = v;
        }
    }

    double x[KPL], b[KPL], Ax[KPL], gk[KPL], dk[KPL], xo[KPL], tmp[KPL], prj[KPL];
    bool present[KPL];
#pragma unroll
    for (int r = 0; r < KPL; ++r) {
        const int c = g * KPL + r;
        present[r] = c < k;
        if (present[r]) {
            tmp[r] = Z[t * k + c];
            const double a = alpha ? alpha[c] : 1.0;
            b[r] = -a * B[t * sb_t + (long)c * sb_c];
        } else {
            tmp[r] = -INFINITY;
            b[r] = 0.0;
        }
    }

    // spg.py:300: make the initial guess feasible
    group8_project<KPL>(tmp, x);

    double f_mem[CDR_MAX_MEMORY];
#pragma unroll
    for (int i = 0; i < CDR_MAX_MEMORY; ++i) f_mem[i] = NAN;

    QpMatVec<KPL>::apply(As, arow, x, Ax, g);
    double f_old;
    {
        double s = 0.0;
#pragma unroll
        for (int r = 0; r < KPL; ++r) s += x[r] * (0.5 * Ax[r] + b[r]);
        f_old = group8_sum(s);
    }
    int n_feval = 1;
    int n_iter = 0;
    double alpha_s = 1.0;
    bool active = valid;     // group-uniform

    for (int it = 0; it < p.max_iterations; ++it) {
        if (!__any_sync(CDR_FULL_MASK, active)) break;
        // ---- gradient, first step length (spg.py:326-339)
#pragma unroll
        for (int r = 0; r < KPL; ++r) {
            xo[r] = x[r];
            gk[r] = Ax[r] + b[r];
        }
        if (it == 0) {
            if (p.alpha_min <= p.alpha0 && p.alpha0 <= p.alpha_max) {
                alpha_s = p.alpha0;
            } else {
#pragma unroll
                for (int r = 0; r < KPL; ++r) tmp[r] = present[r] ? x[r] - gk[r] : -INFINITY;
                group8_project<KPL>(tmp, prj);
                double m = 0.0;
#pragma unroll
                for (int r = 0; r < KPL; ++r) m = fmax(m, fabs(prj[r] - x[r]));
                double ainv = group8_max(m);
                if (fabs(ainv) < 1e-12) ainv = 1.0;
                alpha_s = fmin(fmax(p.alpha_min, 1.0 / ainv), p.alpha_max);
            }
        }
        // ---- projected-gradient direction (spg.py:341)
#pragma unroll
        for (int r = 0; r < KPL; ++r) tmp[r] = present[r] ? x[r] - alpha_s * gk[r] : -INFINITY;
        group8_project<KPL>(tmp, prj);
        double sd = 0.0, sdd = 0.0;
#pragma unroll
        for (int r = 0; r < KPL; ++r) {
            dk[r] = prj[r] - x[r];
            sd += dk[r] * gk[r];
            sdd += dk[r] * dk[r];
        }
        const double delta = group8_sum(sd);
        const double dkdk = group8_sum(sdd);

        // ---- non-monotone reference value (spg.py:343-347)
#pragma unroll
        for (int i = CDR_MAX_MEMORY - 1; i > 0; --i)
            if (i < p.memory) f_mem[i] = f_mem[i - 1];
        f_mem[0] = f_old;
        double f_max = -INFINITY;
#pragma unroll
        for (int i = 0; i < CDR_MAX_MEMORY; ++i)
            if (i < p.memory && !isnan(f_mem[i])) f_max = fmax(f_max, f_mem[i]);

        // ---- line search (spg.py:349-372).  Trial points are x_old + lam d; A x is linear
        // in lam, so one extra mat-vec A d makes every backtracking trial a reduction only.
        double Ad[KPL];
        QpMatVec<KPL>::apply(As, arow, dk, Ad, g);
        double lam = 1.0;
        double xn[KPL];
        double f_new;
        {
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < KPL; ++r) {
                xn[r] = xo[r] + dk[r];
                s += xn[r] * (0.5 * (Ax[r] + Ad[r]) + b[r]);
            }
            f_new = group8_sum(s);
        }
        int fe = 1;
        bool searching = active && (f_new > f_max + p.gamma * lam * delta);
        // Backtracking.  Once sigma_two * lam < sigma_one the safeguarded interpolation of
        // spg.py:19-33 can only return lam / 2 (its acceptance interval is empty), so the next
        // four trial steps are known in advance: they are evaluated together (four independent
        // 8-lane reductions in flight instead of one) and then examined in order, exactly as
        // the reference would.  This is where the samples with rounding-level Armijo failures
        // spend their time (~33 halvings per iteration down to lambda_min).
        while (__any_sync(CDR_FULL_MASK, searching)) {
            const bool halving = p.sigma_two * lam < p.sigma_one;
            double lt[4];
            lt[0] = searching ? spg_step_length(lam, delta, f_old, f_new, p.sigma_one, p.sigma_two)
                              : lam;
            lt[1] = 0.5 * lt[0];
            lt[2] = 0.5 * lt[1];
            lt[3] = 0.5 * lt[2];
            double st[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int r = 0; r < KPL; ++r) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const double xt = xo[r] + lt[j] * dk[r];
                    st[j] += xt * (0.5 * (Ax[r] + lt[j] * Ad[r]) + b[r]);
                }
            }
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
#pragma unroll
                for (int j = 0; j < 4; ++j) st[j] += __shfl_xor_sync(CDR_FULL_MASK, st[j], o, 8);
            }
            const int nvalid = halving ? 4 : 1;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (j < nvalid && searching) {
                    lam = lt[j];
                    f_new = st[j];
                    fe += 1;
                    if (fabs(lam) < p.lambda_min) searching = false;
                    else searching = f_new > f_max + p.gamma * lam * delta;
                }
            }
        }
#pragma unroll
        for (int r = 0; r < KPL; ++r) xn[r] = xo[r] + lam * dk[r];
        // exact A x at the accepted point (gradient and f_old as in spg.py:374-386)
        double Axn[KPL];
        QpMatVec<KPL>::apply(As, arow, xn, Axn, g);
        {
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < KPL; ++r) s += xn[r] * (0.5 * Axn[r] + b[r]);
            f_new = group8_sum(s);
        }

        // ---- accept, spectral step length (spg.py:374-386)
        double sy = 0.0;
#pragma unroll
        for (int r = 0; r < KPL; ++r) {
            const double gnew = Axn[r] + b[r];
            const double yk = gnew - gk[r];
            sy += dk[r] * yk;
            tmp[r] = present[r] ? xn[r] - (yk + gk[r]) : -INFINITY;   // x - g_new for the residual
        }
        const double dkyk = group8_sum(sy);
        const double sksk = lam * lam * dkdk;
        const double betak = lam * dkyk;
        const double alpha_next = spg_cauchy_step(betak, sksk, p.alpha_min, p.alpha_max);

        // ---- projected-gradient residual (spg.py:388-394)
        group8_project<KPL>(tmp, prj);
        double r2 = 0.0, rinf = 0.0;
#pragma unroll
        for (int r = 0; r < KPL; ++r) {
            const double res = prj[r] - xn[r];
            r2 += res * res;
            rinf = fmax(rinf, fabs(res));
        }
        r2 = group8_sum(r2);
        rinf = group8_max(rinf);

        if (active) {
#pragma unroll
            for (int r = 0; r < KPL; ++r) {
                x[r] = xn[r];
                Ax[r] = Axn[r];
            }
            alpha_s = alpha_next;
            f_old = f_new;                     // spg.py:386 re-evaluates the same expression
            n_feval += fe + 1;
            n_iter = it;
            const bool conv = (sqrt(r2) < p.epsilon_two) || (rinf < p.epsilon_one);
            if (conv || n_feval > p.max_feval) active = false;
        }
    }

    if (valid) {
#pragma unroll
        for (int r = 0; r < KPL; ++r)
            if (present[r]) Z[t * k + g * KPL + r] = x[r];
        if (g == 0) {
            if (n_iter_out) n_iter_out[t] = n_iter;
            if (n_feval_out) n_feval_out[t] = n_feval;
        }
    }
}

template <int KPL>
static int launch_qp(const double* A, const double* alpha, const double* B, long sb_t, long sb_c,
                     double* Z, int T, int k, const cdr_spg_params& p, int* n_iter, int* n_feval,
                     const cdr_flags* flags, cudaStream_t stream)
{
    constexpr int KP = 8 * KPL;
    // Samples per warp: one while every warp can still have (nearly) its own scheduler
    // slot, up to four for large batches.  Few warps per CTA so the samples spread over
    // all SMs (the kernel is latency bound).
    int spw = 4;
    if (T <= 148 * 4 * 6) spw = 1;
    else if (T <= 148 * 4 * 16) spw = 2;
    const int warps_needed = (T + spw - 1) / spw;
    int warps_per_block = 1;
    if (warps_needed > 148 * 16) warps_per_block = 2;
    if (warps_needed > 148 * 32) warps_per_block = 4;
    const int blocks = (warps_needed + warps_per_block - 1) / warps_per_block;
    const size_t smem = (KPL > 1) ? (size_t)KP * KP * sizeof(double) : 0;
    qp_batched_kernel<KPL><<<blocks, warps_per_block * 32, smem, stream>>>(
        A, alpha, B, sb_t, sb_c, Z, T, k, spw, p, n_iter, n_feval, flags);
    CDR_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

}  // namespace cdr

extern "C" int cdr_quad_simplex_spg_batched(const double* A, const double* alpha, const double* B,
                                            long sb_t, long sb_c, double* Z, int T, int k,
                                            const cdr_spg_params* params, int* n_iter,
                                            int* n_feval, const cdr_flags* flags,
                                            cdr_stream_t stream)
{
    CDR_CHECK_ARG(params != nullptr && T >= 0 && k >= 1);
    if (k > CDR_MAX_COMPONENTS) return CDR_ERR_UNSUPPORTED;
    if (params->memory < 1 || params->memory > CDR_MAX_MEMORY) return CDR_ERR_UNSUPPORTED;
    if (T == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const cdr_spg_params& p = *params;
    if (k <= 8) return cdr::launch_qp<1>(A, alpha, B, sb_t, sb_c, Z, T, k, p, n_iter, n_feval, flags, s);
    if (k <= 16) return cdr::launch_qp<2>(A, alpha, B, sb_t, sb_c, Z, T, k, p, n_iter, n_feval, flags, s);
    if (k <= 24) return cdr::launch_qp<3>(A, alpha, B, sb_t, sb_c, Z, T, k, p, n_iter, n_feval, flags, s);
    if (k <= 32) return cdr::launch_qp<4>(A, alpha, B, sb_t, sb_c, Z, T, k, p, n_iter, n_feval, flags, s);
    if (k <= 48) return cdr::launch_qp<6>(A, alpha, B, sb_t, sb_c, Z, T, k, p, n_iter, n_feval, flags, s);
    return cdr::launch_qp<8>(A, alpha, B, sb_t, sb_c, Z, T, k, p, n_iter, n_feval, flags, s);
}
