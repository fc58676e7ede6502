// Small dense pieces around the streaming passes: k x k products reduced over a
// long axis (Z'Z, C K C', W'W, ...), the k x k pseudo-inverse of the GPNH
// dictionary step and the scalar cost bookkeeping of the outer loops.
#include "small_solve.cuh"

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
// GPNH dictionary step: P = pinv(ZtZ / T + lambda * G_W) / T   (small_solve.cuh)
// ======================================================================
__global__ void __launch_bounds__(256)
gpnh_solve_matrix_kernel(const double* __restrict__ ZtZ, int k, double inv_n, double lambda_W,
                         double gw_prefactor, double* __restrict__ P, const cdr_flags* flags)
{
    if (is_done(flags)) return;
    extern __shared__ double jac_sm[];            // 2 * CDR_MAX_COMPONENTS * kJacLd doubles
    solve_matrix_cta(ZtZ, k, CDR_MAX_COMPONENTS, inv_n, lambda_W, gw_prefactor, P, jac_sm);
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
