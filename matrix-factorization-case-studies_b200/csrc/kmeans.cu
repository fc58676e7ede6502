// k-means (Lloyd) pieces.  The reference calls scikit-learn's KMeans
// (bin/run_hadisst_kmeans.py:128-131, bin/run_jra55_kmeans.py:115-132); the
// semantics restated here are those of scikit-learn 1.9.0
// (sklearn/cluster/_kmeans.py, _k_means_lloyd.pyx, _k_means_common.pyx):
//
//   E step   label_t = first argmin_j (||c_j||^2 - 2 x_t.c_j)       _k_means_lloyd.pyx:193-213
//   M step   c_j = (sum of the samples of cluster j) * (1 / n_j)   _k_means_common.pyx:215-240
//   shift    sum_j ||c_j_new - c_j_old||^2                          _k_means_common.pyx:243-262
//   inertia  sum_t ||x_t - c_label(t)||^2 (direct differences)      _k_means_common.pyx:118-150
//
// The two large products (x_t.c_j for all t, j and the per-cluster sums) are
// the streaming passes of stream_gemm.cu: reduce_features(centres, X) and
// reduce_samples(one_hot(labels), X); this file holds the small kernels
// around them.
#include "cdr_common.cuh"

namespace cdr {

// ||row||^2 of a k x d matrix, one CTA per row (row_norms(centers, squared=True))
__global__ void __launch_bounds__(256)
row_sqnorms_kernel(const double* __restrict__ C, long ldc, int d, double* out)
{
    __shared__ double scratch[32];
    const double* row = C + (long)blockIdx.x * ldc;
    double s[1] = {0.0};
    for (int f = threadIdx.x; f < d; f += blockDim.x) s[0] = fma(row[f], row[f], s[0]);
    block_sum<1>(s, scratch);
    if (threadIdx.x == 0) out[blockIdx.x] = s[0];
}

// labels, one-hot assignment matrix (k x ldo), cluster sizes and a "labels
// changed" flag.  xct is the k x T matrix of inner products centres . samples.
__global__ void __launch_bounds__(256)
kmeans_labels_kernel(const double* __restrict__ xct, long ldt, const double* __restrict__ cnorm,
                     int T, int k, int32_t* labels, double* onehot, long ldo, int* counts,
                     int* changed)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    double best = cnorm[0] - 2.0 * xct[t];
    int label = 0;
    for (int j = 1; j < k; ++j) {
        const double sc = cnorm[j] - 2.0 * xct[(long)j * ldt + t];
        if (sc < best) {        // strict: first minimum wins
            best = sc;
            label = j;
        }
    }
    if (labels[t] != label) {
        labels[t] = label;
        atomicOr(changed, 1);
    }
    for (int j = 0; j < k; ++j) onehot[(long)j * ldo + t] = (j == label) ? 1.0 : 0.0;
    atomicAdd(&counts[label], 1);
}

// squared distance of every sample to its own centre (direct differences),
// one warp per sample, fixed reduction order
__global__ void __launch_bounds__(256)
kmeans_sqdist_kernel(const double* __restrict__ X, long ldx, int T, int d,
                     const double* __restrict__ C, long ldc, const int32_t* __restrict__ labels,
                     double* out)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= T) return;
    const double* x = X + (long)warp * ldx;
    const double* c = C + (long)labels[warp] * ldc;
    double s = 0.0;
    for (int f = lane; f < d; f += 32) {
        const double df = x[f] - c[f];
        s = fma(df, df, s);
    }
    s = warp_sum(s);
    if (lane == 0) out[warp] = s;
}

// centres <- sums * (1 / count) and the squared shift per centre; one CTA per centre
__global__ void __launch_bounds__(256)
kmeans_update_kernel(const double* __restrict__ sums, long lds, const double* __restrict__ counts,
                     double* centres, long ldc, int d, double* shift)
{
    __shared__ double scratch[32];
    const int j = blockIdx.x;
    const double w = counts[j];
    const double alpha = (w > 0.0) ? 1.0 / w : 1.0;
    double s[1] = {0.0};
    for (int f = threadIdx.x; f < d; f += blockDim.x) {
        const double nv = sums[(long)j * lds + f] * alpha;
        const double df = nv - centres[(long)j * ldc + f];
        s[0] = fma(df, df, s[0]);
        centres[(long)j * ldc + f] = nv;
    }
    block_sum<1>(s, scratch);
    if (threadIdx.x == 0) shift[j] = s[0];
}

// column means and variances with the sequential row order NumPy uses for
// X.mean(axis=0) / np.var(X, axis=0) on a C-contiguous matrix
__global__ void __launch_bounds__(128)
column_moments_kernel(const double* __restrict__ X, long ldx, int T, int d, double* mean, double* var)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= d) return;
    double s = 0.0;
    for (int t = 0; t < T; ++t) s += X[(long)t * ldx + f];
    const double m = s / (double)T;
    mean[f] = m;
    if (var != nullptr) {
        double v = 0.0;
        for (int t = 0; t < T; ++t) {
            const double df = X[(long)t * ldx + f] - m;
            v += df * df;
        }
        var[f] = v / (double)T;
    }
}

__global__ void __launch_bounds__(256)
center_columns_kernel(double* X, long ldx, int T, int d, const double* __restrict__ mean, double sign)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    const int t = blockIdx.y;
    if (f >= d) return;
    X[(long)t * ldx + f] += sign * mean[f];
}

}  // namespace cdr

using namespace cdr;

extern "C" int cdr_row_sqnorms(const double* C, long ldc, int k, int d, double* out,
                               cdr_stream_t stream)
{
    CDR_CHECK_ARG(k >= 1 && d >= 1 && ldc >= d);
    row_sqnorms_kernel<<<k, 256, 0, (cudaStream_t)stream>>>(C, ldc, d, out);
    CDR_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

extern "C" int cdr_kmeans_labels(const double* xct, long ldt, const double* cnorm, int T, int k,
                                 int32_t* labels, double* onehot, long ldo, int* counts,
                                 int* changed, cdr_stream_t stream)
{
    CDR_CHECK_ARG(T >= 1 && k >= 1 && ldt >= T && ldo >= T);
    kmeans_labels_kernel<<<(T + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        xct, ldt, cnorm, T, k, labels, onehot, ldo, counts, changed);
    CDR_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

extern "C" int cdr_kmeans_sqdist(const double* X, long ldx, int T, int d, const double* C,
                                 long ldc, const int32_t* labels, double* out, cdr_stream_t stream)
{
    CDR_CHECK_ARG(T >= 1 && d >= 1 && ldx >= d && ldc >= d);
    const long threads = (long)T * 32;
    kmeans_sqdist_kernel<<<(int)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        X, ldx, T, d, C, ldc, labels, out);
    CDR_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

extern "C" int cdr_kmeans_update(const double* sums, long lds, const double* counts,
                                 double* centres, long ldc, int k, int d, double* shift,
                                 cdr_stream_t stream)
{
    CDR_CHECK_ARG(k >= 1 && d >= 1 && lds >= d && ldc >= d);
    kmeans_update_kernel<<<k, 256, 0, (cudaStream_t)stream>>>(sums, lds, counts, centres, ldc, d,
                                                             shift);
    CDR_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

extern "C" int cdr_column_moments(const double* X, long ldx, int T, int d, double* mean,
                                  double* var, cdr_stream_t stream)
{
    CDR_CHECK_ARG(T >= 1 && d >= 1 && ldx >= d);
    column_moments_kernel<<<(d + 127) / 128, 128, 0, (cudaStream_t)stream>>>(X, ldx, T, d, mean, var);
    CDR_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

extern "C" int cdr_center_columns(double* X, long ldx, int T, int d, const double* mean,
                                  double sign, cdr_stream_t stream)
{
    CDR_CHECK_ARG(T >= 1 && d >= 1 && ldx >= d);
    dim3 grid((d + 255) / 256, T);
    center_columns_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(X, ldx, T, d, mean, sign);
    CDR_RETURN_IF_LAUNCH_FAILED();
    return 0;
}
