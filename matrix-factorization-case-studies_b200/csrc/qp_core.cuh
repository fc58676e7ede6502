// Solver core of the batched simplex-constrained QP (quad_simplex_spg, spg.py:286-398),
// shared by the stand-alone kernel (qp.cu) and the fused weights-update kernels
// (iterate.cu).  An aligned group of 8 lanes owns one sample, each lane KPL consecutive
// components; see qp.cu for the mapping.
#pragma once

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

// line-search trials a lane group evaluates per round in the halving regime.  Eight (one round
// of four replicas would cover the 32 halvings from 1 down past lambda_min = 1e-10) was
// measured against four on the same box: AA 0.4341 vs 0.4347 ms, GPNH 0.2445 vs 0.2369 ms per
// iteration (128 registers and a spill in the fused kernel) -> four.
#ifndef CDR_QP_TRIALS
#define CDR_QP_TRIALS 4
#endif
constexpr int kQpTrials = CDR_QP_TRIALS;

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

// Runs the SPG iteration for the sample of this 8-lane group.
//   As / arow : A' (transposed in shared memory, or the lane's row for KPL == 1)
//   z0        : the sample's initial guess (absent components = -inf), x receives the result
//   b         : linear term of the lane's components (0 on absent components)
//   valid     : group-uniform; groups without a sample run along with frozen state
//   spw       : samples per warp (1, 2 or 4).  Lane group q works on sample q % spw of the
//               warp; the 4 / spw groups with the same sample are exact replicas of each
//               other (the callers load them identically) and share the line-search trials.
// All 32 lanes of the warp must call this together.
template <int KPL>
__device__ __forceinline__ void qp_solve(const double* As, const double (&arow)[8],
                                         const double (&z0)[KPL], const double (&b)[KPL],
                                         const bool (&present)[KPL], const cdr_spg_params& p,
                                         bool valid, int g, int spw, double (&x)[KPL],
                                         int& n_iter_out, int& n_feval_out)
{
    const int lane_group = (threadIdx.x & 31) >> 3;
    const int sample_group = lane_group % spw;           // first lane group of this sample
    const int replica = lane_group / spw;
    const int replicas = 4 / spw;
    double Ax[KPL], gk[KPL], dk[KPL], xo[KPL], tmp[KPL], prj[KPL];
#pragma unroll
    for (int r = 0; r < KPL; ++r) tmp[r] = z0[r];
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
    // bracket of epsilon_two^2 for the convergence test (a tiny epsilon_two whose square
    // underflows always takes the square root)
    const double eps2_sq = p.epsilon_two > 0.0 ? p.epsilon_two * p.epsilon_two : 0.0;
    const bool eps2_ok = eps2_sq > 1e-290;
    const double eps2_lo = eps2_ok ? eps2_sq * (1.0 - 1e-12) : 0.0;
    const double eps2_hi = eps2_ok ? eps2_sq * (1.0 + 1e-12) : (p.epsilon_two > 0.0 ? INFINITY : 0.0);
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

        // ---- non-monotone reference value (spg.py:343-347); memory = 1 (the default of
        // quad_simplex_spg) is the monotone search on f_old alone
        double f_max = f_old;
        if (p.memory > 1) {
#pragma unroll
            for (int i = CDR_MAX_MEMORY - 1; i > 0; --i)
                if (i < p.memory) f_mem[i] = f_mem[i - 1];
            f_mem[0] = f_old;
            f_max = -INFINITY;
#pragma unroll
            for (int i = 0; i < CDR_MAX_MEMORY; ++i)
                if (i < p.memory && !isnan(f_mem[i])) f_max = fmax(f_max, f_mem[i]);
        }

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
        // spg.py:19-33 can only return lam / 2 (its acceptance interval is empty), so the
        // following trial steps l0, l0/2, l0/4, ... are known in advance.  They are evaluated
        // kQpTrials * replicas at a time -- four per lane group (four independent 8-lane
        // reductions in flight), and when several lane groups of the warp hold the same sample (small
        // batches: one sample per warp, see the callers) each replica takes its own four --
        // and then examined in order, exactly as the reference would.  This is where the
        // samples with rounding-level Armijo failures spend their time (~33 halvings per
        // iteration down to lambda_min); every replica forms its sums in the same order, so
        // the result does not depend on the number of replicas.
        while (__any_sync(CDR_FULL_MASK, searching)) {
            const bool halving = p.sigma_two * lam < p.sigma_one;
            const double l0 = searching ? spg_step_length(lam, delta, f_old, f_new, p.sigma_one,
                                                          p.sigma_two)
                                        : lam;
            const int first = halving ? kQpTrials * replica : 0;   // this group's first trial
            double lt[kQpTrials];
            lt[0] = l0 * (1.0 / (double)(1ull << first));          // exact power-of-two scaling
#pragma unroll
            for (int j = 1; j < kQpTrials; ++j) lt[j] = 0.5 * lt[j - 1];
            double st[kQpTrials];
#pragma unroll
            for (int j = 0; j < kQpTrials; ++j) st[j] = 0.0;
#pragma unroll
            for (int r = 0; r < KPL; ++r) {
#pragma unroll
                for (int j = 0; j < kQpTrials; ++j) {
                    const double xt = xo[r] + lt[j] * dk[r];
                    st[j] += xt * (0.5 * (Ax[r] + lt[j] * Ad[r]) + b[r]);
                }
            }
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
#pragma unroll
                for (int j = 0; j < kQpTrials; ++j) st[j] += __shfl_xor_sync(CDR_FULL_MASK, st[j], o, 8);
            }
            // first trial of this group at which the search ends (step below lambda_min, or
            // the Armijo condition holds): 99 = none
            const int nvalid = halving ? kQpTrials : 1;
            int stop = 99;
#pragma unroll
            for (int j = kQpTrials - 1; j >= 0; --j) {
                const bool ends = (fabs(lt[j]) < p.lambda_min) ||
                                  !(st[j] > f_max + p.gamma * lt[j] * delta);
                if (j < nvalid && ends) stop = first + j;
            }
            // earliest such trial over the replicas of the sample
            // (a fixed four shuffles: with fewer replicas some of them are read twice)
            int m = stop;
#pragma unroll
            for (int rho = 0; rho < 4; ++rho)
                m = min(m, __shfl_sync(CDR_FULL_MASK, stop,
                                       (sample_group + spw * (rho & (replicas - 1))) * 8 + g));
            const int round_trials = halving ? kQpTrials * replicas : 1;
            const int last = (m == 99) ? round_trials - 1 : m;   // the trial the search is at now
            const int jsel = last % kQpTrials;
            double fsel = st[0];
#pragma unroll
            for (int j = 1; j < kQpTrials; ++j) fsel = (jsel == j) ? st[j] : fsel;
            const double f_at = __shfl_sync(
                CDR_FULL_MASK, fsel, (sample_group + spw * (halving ? last / kQpTrials : 0)) * 8 + g);
            if (searching) {
                lam = l0 * (1.0 / (double)(1ull << last));
                f_new = f_at;
                fe += last + 1;
                searching = (m == 99);
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
        double r2 = 0.0;
        bool below = true;           // max |res| < epsilon_one  <=>  every |res| is
#pragma unroll
        for (int r = 0; r < KPL; ++r) {
            const double res = prj[r] - xn[r];
            r2 += res * res;
            below = below && (fabs(res) < p.epsilon_one);
        }
        r2 = group8_sum(r2);
        const bool rinf_small =
            ((__ballot_sync(CDR_FULL_MASK, below) >> ((threadIdx.x & 24u))) & 0xffu) == 0xffu;

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
            // sqrt(r2) < epsilon_two, with the square root taken only when r2 is within
            // rounding distance of epsilon_two^2
            bool r2_small = r2 < eps2_lo;
            if (!r2_small && r2 < eps2_hi) r2_small = sqrt(r2) < p.epsilon_two;
            const bool conv = r2_small || rinf_small;
            if (conv || n_feval > p.max_feval) active = false;
        }
    }

    n_iter_out = n_iter;
    n_feval_out = n_feval;
}

}  // namespace cdr
