// One Lloyd iteration of k-means behind a single C call (cdr_kmeans_iterate_enqueue), with
// the stopping rule on the device, so that the host replays a captured CUDA graph and only
// looks at the 48-byte state block every few iterations.
//
// The reference's drivers call scikit-learn's KMeans (bin/run_hadisst_kmeans.py:128-131);
// semantics are those of scikit-learn 1.9.0's `_kmeans_single_lloyd` (_kmeans.py:620-760,
// _k_means_lloyd.pyx:193-218, _k_means_common.pyx:167-262), see kmeans.cu.  Per iteration:
//
//   1. reduce over features    centres . samples as per-strip partials (stream_tma.cu)
//   2. kmeans_assign_kernel    per sample: sum of the strip partials, score_j = ||c_j||^2 -
//                              2 x.c_j, FIRST minimum (strict <), one-hot row, cluster sizes,
//                              "a label changed" flag
//   3. reduce over samples     per-cluster sums = one_hot' X  (stream_tma.cu)
//   4. kmeans_update_kernel    (k x 16 CTAs) centres = sums / n_j, squared shift and the new
//                              ||c_j||^2 per slice; the last CTA sums the slices in fixed
//                              order, applies the stopping rule (labels unchanged -> strict
//                              convergence; total shift <= tol; iteration limit) and clears
//                              the counters for the next E step.
// cdr_kmeans_prepare_enqueue computes ||c_j||^2 of the initial centres and clears the counters.
//
// An empty cluster (rare) stops the device loop before the centre update with
// state->needs_relocation set; the host relocates (_k_means_common.pyx:167-212), finishes
// that iteration with the stand-alone calls of kmeans.cu and resumes.  Small shapes, which
// the strip kernels do not cover, run the stand-alone calls only.
#include "cdr_common.cuh"

namespace cdr {

__device__ __forceinline__ bool km_done(const cdr_kmeans_state* st)
{
    return *((volatile const int*)&st->done) != 0;
}

// prepare: squared norms of the centres (one CTA per centre); counters cleared
__global__ void __launch_bounds__(256)
kmeans_begin_kernel(const double* __restrict__ C, long ldc, int d, int k, double* cnorm, int* counts,
                    cdr_kmeans_state* st)
{
    if (km_done(st)) return;
    __shared__ double scratch[32];
    const double* row = C + (long)blockIdx.x * ldc;
    double s[1] = {0.0};
    for (int f = threadIdx.x; f < d; f += blockDim.x) s[0] = fma(row[f], row[f], s[0]);
    block_sum<1>(s, scratch);
    if (threadIdx.x == 0) {
        cnorm[blockIdx.x] = s[0];
        counts[blockIdx.x] = 0;
        if (blockIdx.x == 0) st->changed = 0;
    }
}

// 2. assignment from the per-strip partials part[strip][t][KP] (KP = 8 for k <= 8, 16 for
// k <= 16): an aligned group of KP lanes owns a sample, lane c component c; the 32 / KP groups
// of a warp split the strips of one sample and combine in fixed order
template <int KP>
__global__ void __launch_bounds__(256)
kmeans_assign_kernel(const double* __restrict__ part, int nstrips, const double* __restrict__ cnorm,
                     int T, int k, int32_t* labels, double* onehot, long ldo, int* counts,
                     cdr_kmeans_state* st)
{
    if (km_done(st)) return;
    constexpr int PHASES = 32 / KP;
    const int lane = threadIdx.x & 31, g = lane % KP, q = lane / KP;
    const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);      // one sample per warp
    if (t >= T) return;
    double a0 = 0.0, a1 = 0.0;
    {
        const double* base = part + ((long)q * T + t) * KP + g;
        const long stride = (long)PHASES * T * KP;
        const int n = (nstrips - q + PHASES - 1) / PHASES;
        int i = 0;
        for (; i + 8 <= n; i += 8) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldcg(base + (long)(i + u) * stride);
#pragma unroll
            for (int u = 0; u < 8; u += 2) {
                a0 += v[u];
                a1 += v[u + 1];
            }
        }
        for (; i < n; ++i) a0 += __ldcg(base + (long)i * stride);
    }
    const double v = a0 + a1;
    // x.c_j: the strip phases in fixed order
    double dot = __shfl_sync(CDR_FULL_MASK, v, g);
#pragma unroll
    for (int ph = 1; ph < PHASES; ++ph) dot += __shfl_sync(CDR_FULL_MASK, v, g + ph * KP);
    double score = (g < k) ? cnorm[g] - 2.0 * dot : INFINITY;
    int idx = g;
    // first minimum over the components: the smaller score wins, the lower index on a tie
#pragma unroll
    for (int o = KP / 2; o > 0; o >>= 1) {
        const double os = __shfl_xor_sync(CDR_FULL_MASK, score, o, KP);
        const int oi = __shfl_xor_sync(CDR_FULL_MASK, idx, o, KP);
        if (os < score || (os == score && oi < idx)) {
            score = os;
            idx = oi;
        }
    }
    if (q == 0) {
        if (g < k) onehot[(long)g * ldo + t] = (g == idx) ? 1.0 : 0.0;
        if (g == 0) {
            if (labels[t] != idx) {
                labels[t] = idx;
                atomicOr(&st->changed, 1);
            }
            atomicAdd(&counts[idx], 1);
        }
    }
}

// 4. centres <- sums * (1 / count); squared shift and squared norm of the new centre per slice
// (grid k x kUpdSlices); the last CTA sums the slices in fixed order and applies the stopping
// rule of _kmeans.py:703-743
constexpr int kUpdSlices = 16;

__global__ void __launch_bounds__(256)
kmeans_update_kernel(const double* __restrict__ sums, long lds, int* counts, double* centres, long ldc,
                     int d, int k, double* shift, double* cnorm, double* slice_part,
                     cdr_kmeans_state* st)
{
    if (km_done(st)) return;
    __shared__ double scratch[64];
    __shared__ int empty;
    __shared__ int is_last;
    if (threadIdx.x == 0) {
        int e = 0;
        for (int j = 0; j < k; ++j) e |= (counts[j] == 0);
        empty = e;
    }
    __syncthreads();
    if (empty) {
        // the host relocates the empty cluster(s) and finishes this iteration
        if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
            st->needs_relocation = 1;
            __threadfence();
            st->done = 1;
        }
        return;
    }
    const int j = blockIdx.x, sl = blockIdx.y;
    const int per = ((d + kUpdSlices - 1) / kUpdSlices + 1) & ~1;
    const int f0 = sl * per, f1 = min(d, f0 + per);
    const double alpha = 1.0 / (double)counts[j];
    double s[2] = {0.0, 0.0};
    for (int f = f0 + threadIdx.x; f < f1; f += blockDim.x) {
        const double nv = sums[(long)j * lds + f] * alpha;
        const double df = nv - centres[(long)j * ldc + f];
        s[0] = fma(df, df, s[0]);
        s[1] = fma(nv, nv, s[1]);
        centres[(long)j * ldc + f] = nv;
    }
    block_sum<2>(s, scratch);
    if (threadIdx.x == 0) {
        __stcg(slice_part + ((long)j * kUpdSlices + sl) * 2, s[0]);
        __stcg(slice_part + ((long)j * kUpdSlices + sl) * 2 + 1, s[1]);
        __threadfence();
        is_last = (atomicAdd(&st->ticket, 1u) == gridDim.x * gridDim.y - 1) ? 1 : 0;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // one lane per centre sums its slices in fixed order
    if (threadIdx.x < k) {
        double sh = 0.0, nr = 0.0;
        for (int q = 0; q < kUpdSlices; ++q) {
            sh += __ldcg(slice_part + ((long)threadIdx.x * kUpdSlices + q) * 2);
            nr += __ldcg(slice_part + ((long)threadIdx.x * kUpdSlices + q) * 2 + 1);
        }
        shift[threadIdx.x] = sh;
        cnorm[threadIdx.x] = nr;
        counts[threadIdx.x] = 0;                      // for the next E step
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    st->ticket = 0u;
    double total = 0.0;
    for (int i = 0; i < k; ++i) {
        const double r = sqrt(shift[i]);                // _kmeans.py:733: (center_shift ** 2).sum()
        total += r * r;
    }
    st->shift_total = total;
    st->n_iter += 1;
    const int changed = st->changed;
    st->changed = 0;
    if (changed == 0) {
        st->strict = 1;
        st->done = 1;
    } else if (total <= st->tol_abs || st->n_iter >= st->max_iter) {
        st->done = 1;
    }
}

}  // namespace cdr

using namespace cdr;

extern "C" size_t cdr_kmeans_workspace_bytes(int T, int d, int k)
{
    if (T < 1 || d < 1 || k < 1 || k > CDR_MAX_COMPONENTS) return 0;
    const size_t s1 = cdr_reduce_samples_workspace_bytes(T, d, k);
    const size_t s2 = cdr_reduce_features_workspace_bytes(T, d, k);
    // streaming scratch | per-slice partials of the update kernel
    return ((s1 > s2 ? s1 : s2) + 255) / 256 * 256 + (size_t)k * kUpdSlices * 2 * sizeof(double) + 256;
}

static double* kmeans_slice_part(const cdr_kmeans_problem* p)
{
    const size_t s1 = cdr_reduce_samples_workspace_bytes(p->T, p->d, p->k);
    const size_t s2 = cdr_reduce_features_workspace_bytes(p->T, p->d, p->k);
    return reinterpret_cast<double*>(static_cast<unsigned char*>(p->workspace) +
                                     ((s1 > s2 ? s1 : s2) + 255) / 256 * 256);
}

extern "C" int cdr_kmeans_fused_applicable(int T, int d, int k)
{
    if (k > 16) return 0;
    int out[12];
    tma_stream_plan(T, d, k, 0, out);
    int TC, nstrips;
    return (out[0] && out[5] && features_strip_geometry(T, d, k, &TC, &nstrips)) ? 1 : 0;
}

static int check_kmeans(const cdr_kmeans_problem* p)
{
    CDR_CHECK_ARG(p != nullptr && p->X && p->centres && p->labels && p->onehot && p->sums &&
                  p->cnorm && p->shift && p->counts && p->state);
    CDR_CHECK_ARG(p->T >= 1 && p->d >= 1 && p->k >= 1 && p->ldt >= p->T);
    return 0;
}

// ||c_j||^2 of the current centres and cleared counters: before the first iteration, and
// after the caller changed the centres itself (empty-cluster relocation)
extern "C" int cdr_kmeans_prepare_enqueue(const cdr_kmeans_problem* p, cdr_stream_t stream)
{
    const int rc = check_kmeans(p);
    if (rc) return rc;
    kmeans_begin_kernel<<<p->k, 256, 0, (cudaStream_t)stream>>>(p->centres, p->ldx, p->d, p->k, p->cnorm,
                                                                p->counts, p->state);
    CDR_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

extern "C" int cdr_kmeans_iterate_enqueue(const cdr_kmeans_problem* p, cdr_stream_t stream)
{
    {
        const int rc = check_kmeans(p);
        if (rc) return rc;
    }
    if (!cdr_kmeans_fused_applicable(p->T, p->d, p->k)) return CDR_ERR_NOT_APPLICABLE;
    if (p->workspace == nullptr || p->workspace_bytes < cdr_kmeans_workspace_bytes(p->T, p->d, p->k))
        return CDR_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    const int T = p->T, d = p->d, k = p->k;
    const cdr_flags* flags = reinterpret_cast<const cdr_flags*>(p->state);   // `done` comes first
    int TC = 0, nstrips = 0;
    features_strip_geometry(T, d, k, &TC, &nstrips);

    {
        const int rc = run_reduce_features_tma(p->centres, p->ldx, p->X, p->ldx, T, d, k, nullptr, 0,
                                               p->workspace, p->workspace_bytes, flags, s, nullptr);
        if (rc != 0) return rc == CDR_TMA_NOT_APPLICABLE ? CDR_ERR_NOT_APPLICABLE : rc;
    }
    if (k <= 8)
        kmeans_assign_kernel<8><<<(T + 7) / 8, 256, 0, s>>>((const double*)p->workspace, nstrips, p->cnorm,
                                                            T, k, p->labels, p->onehot, p->ldt,
                                                            p->counts, p->state);
    else
        kmeans_assign_kernel<16><<<(T + 7) / 8, 256, 0, s>>>((const double*)p->workspace, nstrips, p->cnorm,
                                                             T, k, p->labels, p->onehot, p->ldt,
                                                             p->counts, p->state);
    CDR_RETURN_IF_LAUNCH_FAILED();
    {
        const int rc = cdr_reduce_samples(p->onehot, p->ldt, 1, p->X, p->ldx, T, d, k, nullptr, p->sums,
                                          p->ldx, p->workspace, p->workspace_bytes, flags, s);
        if (rc != 0) return rc;
    }
    kmeans_update_kernel<<<dim3(k, kUpdSlices), 256, 0, s>>>(p->sums, p->ldx, p->counts, p->centres,
                                                             p->ldx, d, k, p->shift, p->cnorm,
                                                             kmeans_slice_part(p), p->state);
    CDR_RETURN_IF_LAUNCH_FAILED();
    return 0;
}
