// Small dense pieces around the streaming passes: k x k products reduced over a
// long axis (Z'Z, C K C', W'W, ...), the k x k pseudo-inverse of the GPNH
// dictionary step and the scalar cost bookkeeping of the outer loops.
#include "cdr_common.cuh"

namespace cdr {

// ======================================================================
// batched small products
// ======================================================================
constexpr int kGramMaxBlocks = 512;
constexpr int kGramTile = 32;           // tile width for k > 16
constexpr int kGramTileWide = 128;      // tile width for k <= 16 (one tile per CTA)
constexpr int kGramTileDoubles = 2112;  // max(16 * (128 + 1), 64 * (32 + 1))
constexpr int kGramMaxPairs = CDR_MAX_COMPONENTS * CDR_MAX_COMPONENTS;

struct GramBatch {
    cdr_small_gram_desc d[CDR_GRAM_BATCH];
    int nblk[CDR_GRAM_BATCH];
};

// partial[desc][blk][pair].  256 threads = (pair slot, column sub-lane): with few pairs
// (k = 8 -> 64) four sub-lanes share the columns of a tile, so the whole CTA works.
__global__ void __launch_bounds__(256)
small_gram_partial_kernel(GramBatch batch, double* __restrict__ part, const cdr_flags* flags)
{
    if (is_done(flags)) return;
    const cdr_small_gram_desc& ds = batch.d[blockIdx.y];
    const int nblk = batch.nblk[blockIdx.y];
    if ((int)blockIdx.x >= nblk) return;
    __shared__ double As[kGramTileDoubles];
    __shared__ double Bs[kGramTileDoubles];
    __shared__ double red[256];

    const int ka = ds.ka, kb = ds.kb, npairs = ka * kb;
    const int tw = (ka <= 16 && kb <= 16) ? kGramTileWide : kGramTile;
    const int ld = tw + 1;
    const int nsub = (npairs <= 64) ? 4 : (npairs <= 128) ? 2 : 1;
    const int slots = 256 / nsub;                  // pair slots per pass
    const int slot = threadIdx.x % slots, sub = threadIdx.x / slots;
    int chunk = (ds.n + nblk - 1) / nblk;
    chunk = (chunk + tw - 1) / tw * tw;
    const int n0 = blockIdx.x * chunk;
    const int n1 = min(ds.n, n0 + chunk);

    double acc[kGramMaxPairs / 256];
#pragma unroll
    for (int s = 0; s < kGramMaxPairs / 256; ++s) acc[s] = 0.0;

    for (int nb = n0; nb < n1; nb += tw) {
        const int w = min(tw, n1 - nb);
        for (int idx = threadIdx.x; idx < ka * tw; idx += 256) {
            const int i = idx / tw, cix = idx % tw;
            As[i * ld + cix] = (cix < w) ? ds.A[(long)i * ds.sAi + (long)(nb + cix) * ds.sAn] : 0.0;
        }
        for (int idx = threadIdx.x; idx < kb * tw; idx += 256) {
            const int j = idx / tw, cix = idx % tw;
            Bs[j * ld + cix] = (cix < w) ? ds.B[(long)j * ds.sBj + (long)(nb + cix) * ds.sBn] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int s = 0; s < kGramMaxPairs / 256; ++s) {
            const int p = slot + s * slots;
            if (p < npairs && (s == 0 || nsub == 1)) {
                const int i = p / kb, j = p % kb;
                const double* ar = As + i * ld;
                const double* br = Bs + j * ld;
                double a = acc[s];
                if (ds.mode == 0) {
                    for (int cix = sub; cix < tw; cix += nsub) a = fma(ar[cix], br[cix], a);
                } else {
                    for (int cix = sub; cix < tw; cix += nsub) {
                        const double df = ar[cix] - br[cix];
                        a = fma(df, df, a);
                    }
                }
                acc[s] = a;
            }
        }
        __syncthreads();
    }
    double* dst = part + ((long)blockIdx.y * kGramMaxBlocks + blockIdx.x) * kGramMaxPairs;
    if (nsub == 1) {
#pragma unroll
        for (int s = 0; s < kGramMaxPairs / 256; ++s) {
            const int p = threadIdx.x + s * 256;
            if (p < npairs) dst[p] = acc[s];
        }
    } else {
        red[threadIdx.x] = acc[0];
        __syncthreads();
        if (sub == 0 && slot < npairs) {
            double v = 0.0;
            for (int q = 0; q < nsub; ++q) v += red[q * slots + slot];     // fixed order
            dst[slot] = v;
        }
    }
}

// out = scale * sum over blocks, fixed order.  256 threads = (pair, group of blocks).
__global__ void __launch_bounds__(256)
small_gram_final_kernel(GramBatch batch, const double* __restrict__ part, const cdr_flags* flags)
{
    if (is_done(flags)) return;
    const cdr_small_gram_desc& ds = batch.d[blockIdx.x];
    const int nblk = batch.nblk[blockIdx.x];
    const int npairs = ds.ka * ds.kb;
    __shared__ double red[256];
    const double* base = part + ((long)blockIdx.x * kGramMaxBlocks) * kGramMaxPairs;
    if (npairs <= 64) {
        const int groups = 256 / 64;
        const int p = threadIdx.x % 64, g = threadIdx.x / 64;
        const int per = (nblk + groups - 1) / groups;
        double s = 0.0;
        if (p < npairs)
            for (int b = g * per; b < min(nblk, (g + 1) * per); ++b) s += base[(long)b * kGramMaxPairs + p];
        red[threadIdx.x] = s;
        __syncthreads();
        if (g == 0 && p < npairs) {
            double v = 0.0;
            for (int q = 0; q < groups; ++q) v += red[q * 64 + p];
            ds.out[p] = ds.scale * v;
        }
    } else {
        for (int p = threadIdx.x; p < npairs; p += 256) {
            double s = 0.0;
            for (int b = 0; b < nblk; ++b) s += base[(long)b * kGramMaxPairs + p];
            ds.out[p] = ds.scale * s;
        }
    }
}

// ======================================================================
// GPNH dictionary step: P = pinv(ZtZ / T + lambda * G_W) / T
// ======================================================================
// One CTA; cyclic Jacobi eigen-decomposition with a round-robin (parallel)
// ordering: k/2 disjoint rotations per round.  The pseudo-inverse drops
// eigenvalues below eps * k * max|eig|, the cut-off numpy.linalg.lstsq applies
// with rcond=None (gpnh_convex_coding.py:224).
constexpr int kJacLd = CDR_MAX_COMPONENTS + 1;

__global__ void __launch_bounds__(256)
gpnh_solve_matrix_kernel(const double* __restrict__ ZtZ, int k, double inv_n, double lambda_W,
                         double gw_prefactor, double* __restrict__ P, const cdr_flags* flags)
{
    if (is_done(flags)) return;
    extern __shared__ double jac_sm[];            // 2 * k * kJacLd doubles
    double* A = jac_sm;
    double* V = jac_sm + CDR_MAX_COMPONENTS * kJacLd;
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
        V[i * kJacLd + j] = v;            // Cholesky works in V; Jacobi re-initialises it
    }
    __syncthreads();

    // ---- fast path: the matrix is symmetric positive definite and reasonably conditioned
    // (the usual case): P = A^-1 by Cholesky, a few microseconds instead of ~50 for the
    // Jacobi sweeps.  Any pivot below 1e-10 * max diag falls through to the
    // pseudo-inverse, which reproduces lstsq's minimum-norm solution for singular Z'Z.
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
        return;
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

__device__ void finish_sub_step(cdr_loop_state* st, double* cost_deltas, double cost, int stage,
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

__global__ void loop_begin_kernel(cdr_loop_state* st)
{
    if (st->done) return;
    if (st->n_iter >= st->max_iterations) {
        st->done = 1;
        return;
    }
    st->old_cost = st->cost;
}

__global__ void __launch_bounds__(32)
gpnh_cost_kernel(cdr_loop_state* st, double* cost_deltas, const double* __restrict__ XWtZ,
                 const double* __restrict__ ZtZ, const double* __restrict__ WtW,
                 const double* __restrict__ reg_pairs, int k, int n_samples, int n_features,
                 double lambda_W, int stage, int end_of_iteration)
{
    if (*((volatile int*)&st->done)) return;
    const int lane = threadIdx.x;
    double tr1 = 0.0, tr2 = 0.0, phi = 0.0;
    for (int idx = lane; idx < k * k; idx += 32) {
        const int i = idx / k, j = idx % k;
        if (i == j) tr1 += XWtZ[idx];
        tr2 += ZtZ[idx] * WtW[j * k + i];
        if (reg_pairs != nullptr && j > i) phi += reg_pairs[idx];
    }
    tr1 = warp_sum(tr1);
    tr2 = warp_sum(tr2);
    phi = warp_sum(phi);
    if (lane != 0) return;
    if (reg_pairs != nullptr) {
        // gpnh_convex_coding.py:179-196
        double pen = 0.0;
        if (lambda_W != 0.0 && k > 1)
            pen = lambda_W * phi * 2.0 / ((double)k * (double)n_features * ((double)k - 1.0));
        st->penalty = pen;
    }
    const double cost = 0.5 * (st->trace_data - 2.0 * tr1 + tr2) / (double)n_samples + st->penalty;
    finish_sub_step(st, cost_deltas, cost, stage, end_of_iteration);
}

}  // namespace cdr

using namespace cdr;

extern "C" size_t cdr_small_gram_workspace_bytes(void)
{
    return (size_t)CDR_GRAM_BATCH * kGramMaxBlocks * kGramMaxPairs * sizeof(double);
}

extern "C" int cdr_small_gram(const cdr_small_gram_desc* descs, int count, void* workspace,
                              size_t workspace_bytes, const cdr_flags* flags, cdr_stream_t stream)
{
    CDR_CHECK_ARG(descs != nullptr && count >= 1 && count <= CDR_GRAM_BATCH);
    if (workspace == nullptr || workspace_bytes < cdr_small_gram_workspace_bytes())
        return CDR_ERR_WORKSPACE;
    GramBatch batch;
    int max_blk = 1;
    for (int i = 0; i < count; ++i) {
        batch.d[i] = descs[i];
        CDR_CHECK_ARG(descs[i].ka >= 1 && descs[i].kb >= 1 && descs[i].n >= 0);
        if (descs[i].ka > CDR_MAX_COMPONENTS || descs[i].kb > CDR_MAX_COMPONENTS)
            return CDR_ERR_UNSUPPORTED;
        int nb = (descs[i].n + kGramTileWide - 1) / kGramTileWide;
        if (nb < 1) nb = 1;
        if (nb > kGramMaxBlocks) nb = kGramMaxBlocks;
        batch.nblk[i] = nb;
        if (nb > max_blk) max_blk = nb;
    }
    cudaStream_t s = (cudaStream_t)stream;
    small_gram_partial_kernel<<<dim3(max_blk, count), 256, 0, s>>>(batch, (double*)workspace, flags);
    CDR_RETURN_IF_LAUNCH_FAILED();
    small_gram_final_kernel<<<count, 256, 0, s>>>(batch, (const double*)workspace, flags);
    CDR_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

extern "C" size_t cdr_sym_pinv_workspace_bytes(int k)
{
    (void)k;
    return 0;
}

extern "C" int cdr_gpnh_solve_matrix(const double* ZtZ, int k, int n_samples, int n_features,
                                     double lambda_W, double* P, void* workspace,
                                     size_t workspace_bytes, const cdr_flags* flags,
                                     cdr_stream_t stream)
{
    (void)workspace;
    (void)workspace_bytes;
    CDR_CHECK_ARG(k >= 1 && n_samples >= 1 && n_features >= 1);
    if (k > CDR_MAX_COMPONENTS) return CDR_ERR_UNSUPPORTED;
    // gpnh_convex_coding.py:296-300
    const double pref = (k > 1) ? 4.0 / ((double)n_features * k * (k - 1)) : 0.0;
    const size_t smem = 2 * (size_t)CDR_MAX_COMPONENTS * kJacLd * sizeof(double);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gpnh_solve_matrix_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    // k/2 * k work items per Jacobi phase: one warp is enough (and cheapest to synchronise)
    // up to k = 16
    gpnh_solve_matrix_kernel<<<1, k <= 16 ? 32 : 256, smem, (cudaStream_t)stream>>>(
        ZtZ, k, 1.0 / (double)n_samples, lambda_W, pref, P, flags);
    CDR_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

extern "C" int cdr_sym_pinv(const double* S, int k, double* P, const cdr_flags* flags,
                            cdr_stream_t stream)
{
    CDR_CHECK_ARG(k >= 1);
    if (k > CDR_MAX_COMPONENTS) return CDR_ERR_UNSUPPORTED;
    const size_t smem = 2 * (size_t)CDR_MAX_COMPONENTS * kJacLd * sizeof(double);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gpnh_solve_matrix_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    gpnh_solve_matrix_kernel<<<1, k <= 16 ? 32 : 256, smem, (cudaStream_t)stream>>>(S, k, 1.0, 0.0, 0.0, P, flags);
    CDR_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

extern "C" int cdr_loop_begin(cdr_loop_state* state, cdr_stream_t stream)
{
    CDR_CHECK_ARG(state != nullptr);
    loop_begin_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state);
    CDR_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

extern "C" int cdr_gpnh_cost_check(cdr_loop_state* state, double* cost_deltas, const double* XWtZ,
                                   const double* ZtZ, const double* WtW, const double* reg_pairs,
                                   int k, int n_samples, int n_features, double lambda_W,
                                   int stage, int end_of_iteration, cdr_stream_t stream)
{
    CDR_CHECK_ARG(state != nullptr && k >= 1);
    gpnh_cost_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(state, cost_deltas, XWtZ, ZtZ, WtW,
                                                        reg_pairs, k, n_samples, n_features,
                                                        lambda_W, stage, end_of_iteration);
    CDR_RETURN_IF_LAUNCH_FAILED();
    return 0;
}
