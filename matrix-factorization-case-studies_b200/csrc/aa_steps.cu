// Device-side pieces of the AA dictionary update: the reference runs the
// generic spg() (spg.py:46-283) with Python callbacks
//   f  = _aa_dictionary_cost / _kernel_aa_dictionary_cost   (archetypal_analysis.py:261-281)
//   df = _aa_dictionary_gradient / _kernel_aa_dictionary_gradient (:284-301)
//   project = simplex_project_rows
// Here the same iteration is expressed in "Gram space" quantities that are all
// k x T or k x k:
//   CK  = C K            (K = X X'; in feature mode CK = (C X) X' by two streaming passes)
//   KZt = (K Z)'
// so that   df(C) = s_g (DZtZD CK - D KZt),
//           f(C + lam D) = s_f/2 (tr K - 2 (a0 + lam a1)
//                                 + tr(DZtZD (G00 + lam (G01 + G01') + lam^2 G11)))
// with G00 = CK C', G01 = CK D', G11 = (D K) D'.  One pass D K per SPG iteration
// is the only large product; every backtracking trial of the line search costs
// O(k^2) and runs inside a single-thread kernel with no host round trip.
//
// Row kernels use one CTA per dictionary row (length T) with the row staged in
// shared memory; the simplex threshold is the reduction-only Michelot iteration.
#include "simplex.cuh"

namespace cdr {

constexpr int kRowMaxT = 26000;

// slots of row_scratch (each k doubles)
enum { RS_A0 = 0, RS_ROWMAX = 1, RS_DELTA = 2, RS_DD = 3, RS_A1 = 4, RS_BETA = 5, RS_R2 = 6, RS_RINF = 7 };

__device__ __forceinline__ bool spg_idle(const cdr_loop_state* st)
{
    return *((volatile const int*)&st->done) != 0 || *((volatile const int*)&st->spg_active) == 0;
}

// gradient entry (j, t):  s_g * (sum_i a_j a_i ZtZ[j][i] CK[i][t] - a_j KZt[j][t])
__device__ __forceinline__ double grad_entry(const cdr_aa_buffers& b, const double* coef, int j, int t)
{
    double s = 0.0;
    for (int i = 0; i < b.k; ++i) s = fma(coef[i], b.CK[(long)i * b.ldt + t], s);
    return b.grad_scale * (s - b.alpha[j] * b.KZt[(long)j * b.ldt + t]);
}

// spg.py:146-148 -- x = project(x0); also the linear trace term a0 = tr(D x KZ)
__global__ void __launch_bounds__(1024) aa_spg_begin_kernel(cdr_aa_buffers b)
{
    if (*((volatile const int*)&b.state->done) != 0) return;
    extern __shared__ double sm[];
    double* scratch = sm;
    double* work = sm + 64;
    const int j = blockIdx.x;
    double* crow = b.C + (long)j * b.ldt;
    for (int t = threadIdx.x; t < b.T; t += blockDim.x) work[t] = crow[t];
    __syncthreads();
    const double th = block_simplex_threshold(work, 1, b.T, scratch);
    double a0[1] = {0.0};
    const double* kz = b.KZt + (long)j * b.ldt;
    for (int t = threadIdx.x; t < b.T; t += blockDim.x) {
        const double x = fmax(work[t] - th, 0.0);
        crow[t] = x;
        a0[0] = fma(x, kz[t], a0[0]);
    }
    block_sum<1>(a0, scratch);
    if (threadIdx.x == 0) b.row_scratch[RS_A0 * b.k + j] = b.alpha[j] * a0[0];
}

// f(x) at the start of spg() (spg.py:155-157) and counter reset
__global__ void aa_spg_f0_kernel(cdr_aa_buffers b, cdr_spg_params p)
{
    cdr_loop_state* st = b.state;
    if (st->done) return;
    const int k = b.k;
    double a0 = 0.0, quad = 0.0;
    for (int j = 0; j < k; ++j) a0 += b.row_scratch[RS_A0 * k + j];
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j)
            quad += b.alpha[i] * b.alpha[j] * b.ZtZ[i * k + j] * b.CKCt[j * k + i];
    st->a0 = a0;
    st->f_old = 0.5 * (st->trace_data - 2.0 * a0 + quad) * b.cost_scale;
    st->spg_iter = 0;
    st->spg_feval = 1;
    st->spg_active = (p.max_iterations > 0) ? 1 : 0;
    // spg_warnings is sticky over the outer iterations of a fit (the reference warns on every
    // occurrence; the host reports each kind once at the end)
    for (int i = 0; i < CDR_MAX_MEMORY; ++i) st->f_mem[i] = 0.0;      // zeros: spg.py:153
    if (p.alpha0 > 0.0) {
        // explicit alpha0 (spg.py:151); the reference uses it unclamped when not None
        st->alpha = p.alpha0;
        st->spg_alpha_set = 1;
    } else {
        st->spg_alpha_set = 0;
    }
}

// df(x) (spg.py:176) and, on the first iteration, max |P(x - g) - x| per row (spg.py:184)
__global__ void __launch_bounds__(1024) aa_spg_grad_kernel(cdr_aa_buffers b)
{
    if (spg_idle(b.state)) return;
    extern __shared__ double sm[];
    double* scratch = sm;
    double* coef = sm + 64;                 // k doubles
    double* work = sm + 64 + CDR_MAX_COMPONENTS;
    const int j = blockIdx.x, k = b.k;
    for (int i = threadIdx.x; i < k; i += blockDim.x)
        coef[i] = b.alpha[j] * b.alpha[i] * b.ZtZ[j * k + i];
    __syncthreads();
    const bool need_alpha = b.state->spg_alpha_set == 0;
    const double* crow = b.C + (long)j * b.ldt;
    double* grow = b.G + (long)j * b.ldt;
    for (int t = threadIdx.x; t < b.T; t += blockDim.x) {
        const double g = grad_entry(b, coef, j, t);
        grow[t] = g;
        work[t] = crow[t] - g;
    }
    if (!need_alpha) return;
    __syncthreads();
    const double th = block_simplex_threshold(work, 1, b.T, scratch);
    double m = 0.0;
    for (int t = threadIdx.x; t < b.T; t += blockDim.x)
        m = fmax(m, fabs(fmax(work[t] - th, 0.0) - crow[t]));
    m = block_max(m, scratch);
    if (threadIdx.x == 0) b.row_scratch[RS_ROWMAX * k + j] = m;
}

// stand-alone df(C) for the public _aa_dictionary_gradient / _kernel_aa_dictionary_gradient
__global__ void __launch_bounds__(1024) aa_gradient_kernel(cdr_aa_buffers b)
{
    extern __shared__ double sm[];
    double* coef = sm;
    const int j = blockIdx.x, k = b.k;
    for (int i = threadIdx.x; i < k; i += blockDim.x)
        coef[i] = b.alpha[j] * b.alpha[i] * b.ZtZ[j * k + i];
    __syncthreads();
    double* grow = b.G + (long)j * b.ldt;
    for (int t = threadIdx.x; t < b.T; t += blockDim.x) grow[t] = grad_entry(b, coef, j, t);
}

// stand-alone f(C) (archetypal_analysis.py:261-281) from CKZ = C (KZ) and CKCt
__global__ void aa_dictionary_cost_kernel(cdr_aa_buffers b, double trace_data, double* out)
{
    const int k = b.k;
    double t1 = 0.0, t2 = 0.0;
    for (int i = 0; i < k; ++i) t1 += b.alpha[i] * b.CKZ[i * k + i];
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j)
            t2 += b.alpha[i] * b.alpha[j] * b.ZtZ[i * k + j] * b.CKCt[j * k + i];
    *out = 0.5 * (trace_data - 2.0 * t1 + t2) * b.cost_scale;
}

// spg.py:178-189
__global__ void aa_spg_alpha_init_kernel(cdr_aa_buffers b)
{
    cdr_loop_state* st = b.state;
    if (st->done || !st->spg_active || st->spg_alpha_set) return;
    double m = 0.0;
    for (int j = 0; j < b.k; ++j) m = fmax(m, b.row_scratch[RS_ROWMAX * b.k + j]);
    st->alpha = (fabs(m) > 1e-12) ? 1.0 / m : 1.0;
    st->spg_alpha_set = 1;
}

// d = P(x - alpha g) - x (spg.py:191-194) with <d,g>, <d,d> and the linear term along d
__global__ void __launch_bounds__(1024) aa_spg_direction_kernel(cdr_aa_buffers b)
{
    if (spg_idle(b.state)) return;
    extern __shared__ double sm[];
    double* scratch = sm;
    double* work = sm + 64;
    const int j = blockIdx.x, k = b.k;
    const double alpha = b.state->alpha;
    const double* crow = b.C + (long)j * b.ldt;
    const double* grow = b.G + (long)j * b.ldt;
    const double* kz = b.KZt + (long)j * b.ldt;
    double* drow = b.D + (long)j * b.ldt;
    for (int t = threadIdx.x; t < b.T; t += blockDim.x) work[t] = crow[t] - alpha * grow[t];
    __syncthreads();
    const double th = block_simplex_threshold(work, 1, b.T, scratch);
    double r[3] = {0.0, 0.0, 0.0};
    for (int t = threadIdx.x; t < b.T; t += blockDim.x) {
        const double d = fmax(work[t] - th, 0.0) - crow[t];
        drow[t] = d;
        r[0] = fma(d, grow[t], r[0]);
        r[1] = fma(d, d, r[1]);
        r[2] = fma(d, kz[t], r[2]);
    }
    block_sum<3>(r, scratch);
    if (threadIdx.x == 0) {
        b.row_scratch[RS_DELTA * k + j] = r[0];
        b.row_scratch[RS_DD * k + j] = r[1];
        b.row_scratch[RS_A1 * k + j] = b.alpha[j] * r[2];
    }
}

// Non-monotone Armijo line search on scalars (spg.py:196-229).
__global__ void aa_spg_linesearch_kernel(cdr_aa_buffers b, cdr_spg_params p)
{
    cdr_loop_state* st = b.state;
    if (st->done || !st->spg_active) return;
    const int k = b.k;
    double delta = 0.0, dd = 0.0, a1 = 0.0;
    for (int j = 0; j < k; ++j) {
        delta += b.row_scratch[RS_DELTA * k + j];
        dd += b.row_scratch[RS_DD * k + j];
        a1 += b.row_scratch[RS_A1 * k + j];
    }
    // quadratic form coefficients  q(lam) = q0 + lam q1 + lam^2 q2
    double q0 = 0.0, q1 = 0.0, q2 = 0.0;
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) {
            const double w = b.alpha[i] * b.alpha[j] * b.ZtZ[i * k + j];   // DZtZD[i][j]
            q0 += w * b.CKCt[j * k + i];
            q1 += w * (b.G01[j * k + i] + b.G01[i * k + j]);
            q2 += w * b.G11[j * k + i];
        }
    const double a0 = st->a0;
    const double tr = st->trace_data;
    const double sf = b.cost_scale;

    // f_mem = roll(f_mem, 1); f_mem[0] = f_old; f_max = max(f_mem)  (spg.py:196-203)
    for (int i = p.memory - 1; i > 0; --i) st->f_mem[i] = st->f_mem[i - 1];
    st->f_mem[0] = st->f_old;
    double f_max = st->f_mem[0];
    for (int i = 1; i < p.memory; ++i) f_max = fmax(f_max, st->f_mem[i]);

    double lam = 1.0;
    double f_new = 0.5 * (tr - 2.0 * (a0 + lam * a1) + (q0 + lam * q1 + lam * lam * q2)) * sf;
    int feval = st->spg_feval + 1;
    while (f_new > f_max + p.gamma * lam * delta) {
        lam = spg_step_length(lam, delta, st->f_old, f_new, p.sigma_one, p.sigma_two);
        f_new = 0.5 * (tr - 2.0 * (a0 + lam * a1) + (q0 + lam * q1 + lam * lam * q2)) * sf;
        feval += 1;
        if (fabs(lam) < p.lambda_min) {
            st->spg_warnings |= 1;
            break;
        }
    }
    // the last inner iteration: the limit is reached (spg.py:277-281; the convergence test of
    // that iteration is not evaluated when its result could only suppress this warning)
    if (st->spg_iter + 1 >= p.max_iterations) st->spg_warnings |= 4;
    st->lam = lam;
    st->f_new = f_new;
    st->delta = delta;
    st->dd = dd;
    st->a1 = a1;
    st->a0 = a0 + lam * a1;
    st->spg_feval = feval;
    // C K C' of the accepted point
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j)
            b.CKCt[i * k + j] += lam * (b.G01[i * k + j] + b.G01[j * k + i]) + lam * lam * b.G11[i * k + j];
}

// x <- x + lam d,  CK <- CK + lam DK   (spg.py:219; linearity of C -> C K)
__global__ void __launch_bounds__(256) aa_spg_axpy_kernel(cdr_aa_buffers b)
{
    if (spg_idle(b.state)) return;
    const double lam = b.state->lam;
    const long n = (long)b.k * b.ldt;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
         idx += (long)gridDim.x * blockDim.x) {
        b.C[idx] = fma(lam, b.D[idx], b.C[idx]);
        b.CK[idx] = fma(lam, b.DK[idx], b.CK[idx]);
    }
}

// g_new = df(x_new), <d, g_new - g>, projected-gradient residual (spg.py:231-252)
__global__ void __launch_bounds__(1024) aa_spg_post_kernel(cdr_aa_buffers b, int compute_residual)
{
    if (spg_idle(b.state)) return;
    extern __shared__ double sm[];
    double* scratch = sm;
    double* coef = sm + 64;
    double* work = sm + 64 + CDR_MAX_COMPONENTS;
    const int j = blockIdx.x, k = b.k;
    for (int i = threadIdx.x; i < k; i += blockDim.x)
        coef[i] = b.alpha[j] * b.alpha[i] * b.ZtZ[j * k + i];
    __syncthreads();
    const double* crow = b.C + (long)j * b.ldt;
    const double* drow = b.D + (long)j * b.ldt;
    double* grow = b.G + (long)j * b.ldt;
    double beta[1] = {0.0};
    for (int t = threadIdx.x; t < b.T; t += blockDim.x) {
        const double g = grad_entry(b, coef, j, t);
        beta[0] = fma(drow[t], g - grow[t], beta[0]);
        grow[t] = g;
        work[t] = crow[t] - g;
    }
    block_sum<1>(beta, scratch);
    if (threadIdx.x == 0) b.row_scratch[RS_BETA * k + j] = beta[0];
    if (!compute_residual) return;
    __syncthreads();
    const double th = block_simplex_threshold(work, 1, b.T, scratch);
    double r2[1] = {0.0};
    double rinf = 0.0;
    for (int t = threadIdx.x; t < b.T; t += blockDim.x) {
        const double res = fmax(work[t] - th, 0.0) - crow[t];
        r2[0] = fma(res, res, r2[0]);
        rinf = fmax(rinf, fabs(res));
    }
    block_sum<1>(r2, scratch);
    rinf = block_max(rinf, scratch);
    if (threadIdx.x == 0) {
        b.row_scratch[RS_R2 * k + j] = r2[0];
        b.row_scratch[RS_RINF * k + j] = rinf;
    }
}

// spectral step, convergence tests and counters (spg.py:236-281)
__global__ void aa_spg_finish_kernel(cdr_aa_buffers b, cdr_spg_params p, int compute_residual)
{
    cdr_loop_state* st = b.state;
    if (st->done || !st->spg_active) return;
    const int k = b.k;
    double beta = 0.0, r2 = 0.0, rinf = 0.0;
    for (int j = 0; j < k; ++j) {
        beta += b.row_scratch[RS_BETA * k + j];
        if (compute_residual) {
            r2 += b.row_scratch[RS_R2 * k + j];
            rinf = fmax(rinf, b.row_scratch[RS_RINF * k + j]);
        }
    }
    const double sksk = st->lam * st->lam * st->dd;
    const double betak = st->lam * beta;
    st->beta = betak;
    st->alpha = spg_cauchy_step(betak, sksk, p.alpha_min, p.alpha_max);
    st->f_old = st->f_new;                    // spg.py:243 re-evaluates f at the same point
    st->spg_feval += 1;
    st->spg_iter += 1;
    st->res2 = r2;
    st->resinf = rinf;
    if (compute_residual) {
        bool conv = sqrt(r2) < p.epsilon_two;
        if (p.use_infinity_norm) conv = conv || (rinf < p.epsilon_one);
        if (conv) st->spg_active = 0;
    }
    if (st->spg_feval > p.max_feval) {
        st->spg_warnings |= 2;
        st->spg_active = 0;
    }
    if (st->spg_iter >= p.max_iterations) {
        if (st->spg_active) st->spg_warnings |= 4;
        st->spg_active = 0;
    }
}

__device__ __forceinline__ bool aa_cost_increased(double old_cost, double new_cost, double tol)
{
    return (new_cost > old_cost) && (fabs(new_cost - old_cost) > tol);
}

// cost from traces (archetypal_analysis.py:555-556, 623-652) + monotonicity / stopping tests
__global__ void __launch_bounds__(32) aa_cost_kernel(cdr_aa_buffers b, int stage, int end_of_iteration)
{
    cdr_loop_state* st = b.state;
    if (*((volatile int*)&st->done)) return;
    const int k = b.k, lane = threadIdx.x;
    double t1 = 0.0, t2 = 0.0;
    for (int idx = lane; idx < k * k; idx += 32) {
        const int i = idx / k, j = idx % k;
        if (i == j) t1 += b.alpha[i] * b.CKZ[idx];
        t2 += b.alpha[i] * b.alpha[j] * b.ZtZ[idx] * b.CKCt[j * k + i];
    }
    t1 = warp_sum(t1);
    t2 = warp_sum(t2);
    if (lane != 0) return;
    const double cost = 0.5 * (st->trace_data - 2.0 * t1 + t2) / (double)b.T;
    st->cost = cost;
    if (stage == 0) return;
    if (st->require_monotone && aa_cost_increased(st->old_cost, cost, st->tolerance)) {
        st->error_stage = stage;
        st->done = 1;
        return;
    }
    if (end_of_iteration) {
        const int it = st->n_iter;
        if (b.cost_deltas) b.cost_deltas[it] = cost - st->old_cost;
        st->n_iter = it + 1;
        const double delta = cost - st->old_cost;
        bool conv;
        if (st->stopping_rule == 0) conv = fabs(delta) < st->tolerance;
        else conv = fabs(delta / fmax(fabs(cost), fabs(st->old_cost))) < st->tolerance;
        if (conv) {
            st->converged = 1;
            st->done = 1;
        } else if (it + 1 >= st->max_iterations) {
            st->done = 1;
        }
    }
}

// ======================================================================
// Scale factors (delta != 0): _update_kernel_aa_scale_factors (archetypal_analysis.py:243-258)
// = the generic spg() (spg.py:46-283) on the k-vector alpha with the box [1 - delta, 1 + delta]
// as feasible set, objective and gradient from three k x k matrices
//   f(a)  = 1/2 (tr K - 2 a.diag(CKZ) + sum_ij a_i a_j ZtZ_ij CKCt_ij) / n     (:221-229)
//   df(a) = diag(ZtZ diag(a) CKCt - CKZ) / n                                    (:232-240)
// with n = CKZ.shape[1] = k (the reference's `n_samples` there is the number of components).
// One warp; lane l owns components l and l + 32.  M[i][j] = ZtZ_ij CKCt_ij and
// N[i][j] = ZtZ_ij CKCt_ji live in shared memory.
// ======================================================================
__device__ __forceinline__ double sf_objective(const double* M, const double* cz, const double* a,
                                               int k, double trace, int lane)
{
    double s = 0.0;
    for (int i = lane; i < k; i += 32) {
        double row = 0.0;
        for (int j = 0; j < k; ++j) row = fma(M[i * k + j], a[j], row);
        s += a[i] * (row - 2.0 * cz[i]);
    }
    s = warp_sum(s);
    return 0.5 * (trace + s) / (double)k;
}

__device__ __forceinline__ void sf_gradient(const double* N, const double* cz, const double* a, int k,
                                            double* g, int lane)
{
    for (int i = lane; i < k; i += 32) {
        double row = 0.0;
        for (int j = 0; j < k; ++j) row = fma(N[i * k + j], a[j], row);
        g[i] = (row - cz[i]) / (double)k;
    }
}

__global__ void __launch_bounds__(32)
aa_scale_factors_kernel(cdr_aa_buffers b, cdr_spg_params p, double delta)
{
    cdr_loop_state* st = b.state;
    if (*((volatile const int*)&st->done) != 0) return;
    extern __shared__ double sf_sm[];
    const int k = b.k, lane = threadIdx.x;
    double* M = sf_sm;                    // k x k
    double* N = M + k * k;                // k x k
    double* cz = N + k * k;               // diag(CKZ)
    double* x = cz + k;                   // iterate
    double* xo = x + k;
    double* g = xo + k;
    double* d = g + k;
    double* gn = d + k;
    double* f_mem = gn + k;               // CDR_MAX_MEMORY
    for (int idx = lane; idx < k * k; idx += 32) {
        const int i = idx / k, j = idx % k;
        M[idx] = b.ZtZ[idx] * b.CKCt[idx];
        N[idx] = b.ZtZ[idx] * b.CKCt[j * k + i];
    }
    const double lo = 1.0 - delta, hi = 1.0 + delta;
    for (int i = lane; i < k; i += 32) {
        cz[i] = b.CKZ[i * k + i];
        x[i] = fmin(fmax(lo, b.alpha[i]), hi);                 // spg.py:146-148
    }
    for (int i = lane; i < CDR_MAX_MEMORY; i += 32) f_mem[i] = 0.0;   // spg.py:153
    __syncwarp();
    const double trace = st->trace_data;
    double f_old = sf_objective(M, cz, x, k, trace, lane);
    int n_feval = 1;
    double alpha = p.alpha0;
    bool have_alpha = p.alpha0 > 0.0;                          // spg.py:151: alpha0 = None otherwise
    bool converged = false;
    int it = 0;
    for (; it < p.max_iterations; ++it) {
        for (int i = lane; i < k; i += 32) xo[i] = x[i];
        sf_gradient(N, cz, x, k, g, lane);
        __syncwarp();
        if (!have_alpha) {
            double m = 0.0;
            for (int i = lane; i < k; i += 32)
                m = fmax(m, fabs(fmin(fmax(lo, x[i] - g[i]), hi) - x[i]));
            m = warp_max(m);
            alpha = (fabs(m) > 1e-12) ? 1.0 / m : 1.0;
            have_alpha = true;
        }
        double sdg = 0.0, sdd = 0.0;
        for (int i = lane; i < k; i += 32) {
            const double di = fmin(fmax(lo, x[i] - alpha * g[i]), hi) - x[i];
            d[i] = di;
            sdg = fma(di, g[i], sdg);
            sdd = fma(di, di, sdd);
        }
        const double dlt = warp_sum(sdg);
        const double dd = warp_sum(sdd);
        // f_mem = roll(f_mem, 1); f_mem[0] = f_old; f_max = max(f_mem)   (spg.py:196-203)
        double f_max = f_old;
        if (lane == 0) {
            for (int i = p.memory - 1; i > 0; --i) f_mem[i] = f_mem[i - 1];
            f_mem[0] = f_old;
        }
        __syncwarp();
        for (int i = 1; i < p.memory; ++i) f_max = fmax(f_max, f_mem[i]);
        double lam = 1.0;
        for (int i = lane; i < k; i += 32) x[i] = xo[i] + d[i];
        __syncwarp();
        double f_new = sf_objective(M, cz, x, k, trace, lane);
        n_feval += 1;
        while (f_new > f_max + p.gamma * lam * dlt) {
            lam = spg_step_length(lam, dlt, f_old, f_new, p.sigma_one, p.sigma_two);
            for (int i = lane; i < k; i += 32) x[i] = xo[i] + lam * d[i];
            __syncwarp();
            f_new = sf_objective(M, cz, x, k, trace, lane);
            n_feval += 1;
            if (fabs(lam) < p.lambda_min) {
                if (lane == 0) st->spg_warnings |= 1;
                break;
            }
        }
        sf_gradient(N, cz, x, k, gn, lane);
        __syncwarp();
        double sdy = 0.0;
        for (int i = lane; i < k; i += 32) sdy = fma(d[i], gn[i] - g[i], sdy);
        const double sksk = lam * lam * dd;
        const double betak = lam * warp_sum(sdy);
        alpha = spg_cauchy_step(betak, sksk, p.alpha_min, p.alpha_max);
        f_old = f_new;                       // spg.py:243 re-evaluates f at the same point
        n_feval += 1;
        double r2 = 0.0, rinf = 0.0;
        for (int i = lane; i < k; i += 32) {
            const double res = fmin(fmax(lo, x[i] - gn[i]), hi) - x[i];
            r2 = fma(res, res, r2);
            rinf = fmax(rinf, fabs(res));
        }
        r2 = warp_sum(r2);
        rinf = warp_max(rinf);
        converged = sqrt(r2) < p.epsilon_two;
        if (p.use_infinity_norm) converged = converged || (rinf < p.epsilon_one);
        if (converged) break;
        if (n_feval > p.max_feval) {
            if (lane == 0) st->spg_warnings |= 2;
            break;
        }
    }
    if (it >= p.max_iterations - 1 && !converged && lane == 0 && p.max_iterations > 0)
        st->spg_warnings |= 4;               // spg.py:277-281
    __syncwarp();
    for (int i = lane; i < k; i += 32) b.alpha[i] = x[i];
}

static int row_threads(int T)
{
    if (T <= 2048) return 256;
    if (T <= 8192) return 512;
    return 1024;
}

// Raises the dynamic shared-memory limit of a kernel once per size (the call is
// not a stream operation, so it is kept out of CUDA-graph capture after warm-up).
template <auto Kern>
static int set_smem(size_t smem)
{
    static size_t configured = 48 * 1024;      // one instance per kernel
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(Kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured = smem;
    }
    return 0;
}

}  // namespace cdr

using namespace cdr;

#define CDR_AA_CHECK(b)                                                                  \
    CDR_CHECK_ARG((b) != nullptr && (b)->k >= 1 && (b)->T >= 1 && (b)->ldt >= (b)->T);  \
    if ((b)->k > CDR_MAX_COMPONENTS || (b)->T > kRowMaxT) return CDR_ERR_UNSUPPORTED

extern "C" int cdr_aa_spg_begin(const cdr_aa_buffers* b, const cdr_spg_params* p, cdr_stream_t stream)
{
    CDR_AA_CHECK(b);
    CDR_CHECK_ARG(p != nullptr);
    if (p->memory < 1 || p->memory > CDR_MAX_MEMORY) return CDR_ERR_UNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem = (64 + (size_t)b->T) * sizeof(double);
    int rc = set_smem<aa_spg_begin_kernel>(smem);
    if (rc) return rc;
    aa_spg_begin_kernel<<<b->k, row_threads(b->T), smem, s>>>(*b);
    CDR_RETURN_IF_LAUNCH_FAILED();
    aa_spg_f0_kernel<<<1, 1, 0, s>>>(*b, *p);
    CDR_RETURN_IF_LAUNCH_FAILED();
    const size_t smem2 = (64 + CDR_MAX_COMPONENTS + (size_t)b->T) * sizeof(double);
    rc = set_smem<aa_spg_grad_kernel>(smem2);
    if (rc) return rc;
    aa_spg_grad_kernel<<<b->k, row_threads(b->T), smem2, s>>>(*b);
    CDR_RETURN_IF_LAUNCH_FAILED();
    aa_spg_alpha_init_kernel<<<1, 1, 0, s>>>(*b);
    CDR_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

extern "C" int cdr_aa_spg_direction(const cdr_aa_buffers* b, const cdr_spg_params* p,
                                    cdr_stream_t stream)
{
    CDR_AA_CHECK(b);
    (void)p;
    const size_t smem = (64 + (size_t)b->T) * sizeof(double);
    int rc = set_smem<aa_spg_direction_kernel>(smem);
    if (rc) return rc;
    aa_spg_direction_kernel<<<b->k, row_threads(b->T), smem, (cudaStream_t)stream>>>(*b);
    CDR_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

extern "C" int cdr_aa_spg_linesearch(const cdr_aa_buffers* b, const cdr_spg_params* p,
                                     cdr_stream_t stream)
{
    CDR_AA_CHECK(b);
    CDR_CHECK_ARG(p != nullptr);
    cudaStream_t s = (cudaStream_t)stream;
    aa_spg_linesearch_kernel<<<1, 1, 0, s>>>(*b, *p);
    CDR_RETURN_IF_LAUNCH_FAILED();
    const long n = (long)b->k * b->ldt;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 592) blocks = 592;
    aa_spg_axpy_kernel<<<blocks, 256, 0, s>>>(*b);
    CDR_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

extern "C" int cdr_aa_spg_update(const cdr_aa_buffers* b, const cdr_spg_params* p,
                                 int compute_residual, cdr_stream_t stream)
{
    CDR_AA_CHECK(b);
    CDR_CHECK_ARG(p != nullptr);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem = (64 + CDR_MAX_COMPONENTS + (size_t)b->T) * sizeof(double);
    int rc = set_smem<aa_spg_post_kernel>(smem);
    if (rc) return rc;
    aa_spg_post_kernel<<<b->k, row_threads(b->T), smem, s>>>(*b, compute_residual);
    CDR_RETURN_IF_LAUNCH_FAILED();
    aa_spg_finish_kernel<<<1, 1, 0, s>>>(*b, *p, compute_residual);
    CDR_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

extern "C" int cdr_aa_scale_factors_step(const cdr_aa_buffers* b, const cdr_spg_params* p, double delta,
                                         cdr_stream_t stream)
{
    CDR_AA_CHECK(b);
    CDR_CHECK_ARG(p != nullptr && delta >= 0.0);
    if (p->memory < 1 || p->memory > CDR_MAX_MEMORY) return CDR_ERR_UNSUPPORTED;
    const size_t smem = (2 * (size_t)b->k * b->k + 6 * (size_t)b->k + CDR_MAX_MEMORY) * sizeof(double);
    int rc = set_smem<aa_scale_factors_kernel>(smem);
    if (rc) return rc;
    aa_scale_factors_kernel<<<1, 32, smem, (cudaStream_t)stream>>>(*b, *p, delta);
    CDR_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

extern "C" int cdr_aa_gradient(const cdr_aa_buffers* b, cdr_stream_t stream)
{
    CDR_AA_CHECK(b);
    aa_gradient_kernel<<<b->k, row_threads(b->T), CDR_MAX_COMPONENTS * sizeof(double),
                         (cudaStream_t)stream>>>(*b);
    CDR_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

extern "C" int cdr_aa_dictionary_cost(const cdr_aa_buffers* b, double trace_data, double* out,
                                      cdr_stream_t stream)
{
    CDR_AA_CHECK(b);
    aa_dictionary_cost_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(*b, trace_data, out);
    CDR_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

extern "C" int cdr_aa_cost_check(const cdr_aa_buffers* b, int stage, int end_of_iteration,
                                 cdr_stream_t stream)
{
    CDR_AA_CHECK(b);
    aa_cost_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(*b, stage, end_of_iteration);
    CDR_RETURN_IF_LAUNCH_FAILED();
    return 0;
}
