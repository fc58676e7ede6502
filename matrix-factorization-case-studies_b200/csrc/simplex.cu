// Row / column projection onto the probability simplex.
// Replaces simplex_projection.py:30-47 (simplex_project_rows / _columns).
//
// One CTA per vector.  The vector is staged once in shared memory (coalesced
// read when the element stride is 1), the threshold is found with the
// reduction-only Michelot iteration (simplex.cuh) and the result is written
// with one coalesced pass: 8n bytes read + 8n bytes written per vector, which
// is the algorithmic minimum.
#include "simplex.cuh"

namespace cdr {

constexpr int kSimplexMaxSmemDoubles = 26000;   // 203 KB of the 227 KB a CTA may use

__global__ void simplex_vectors_kernel(const double* __restrict__ A, double* out, int n,
                                       long vec_stride_in, long elem_stride_in,
                                       long vec_stride_out, long elem_stride_out, int use_smem,
                                       const cdr_flags* flags)
{
    if (is_done(flags)) return;
    extern __shared__ double sm[];
    double* scratch = sm;                         // 64 doubles
    const long v = blockIdx.x;
    const double* src = A + v * vec_stride_in;
    double* dst = out + v * vec_stride_out;
    double* work = use_smem ? sm + 64 : dst;
    const long ws = use_smem ? 1 : elem_stride_out;

    for (int i = threadIdx.x; i < n; i += blockDim.x) work[(long)i * ws] = src[(long)i * elem_stride_in];
    __syncthreads();
    const double t = block_simplex_threshold(work, ws, n, scratch);
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        dst[(long)i * elem_stride_out] = fmax(work[(long)i * ws] - t, 0.0);
}

static int launch_simplex(const double* A, double* out, int nvec, int n, long vs_in, long es_in,
                          long vs_out, long es_out, const cdr_flags* flags, cudaStream_t stream)
{
    if (nvec == 0 || n == 0) return 0;
    const int use_smem = n <= kSimplexMaxSmemDoubles;
    const size_t smem = (64 + (use_smem ? (size_t)n : 0)) * sizeof(double);
    int threads = 128;
    if (n > 512) threads = 256;
    if (n > 4096) threads = 512;
    if (n > 16384) threads = 1024;
    static size_t configured = 48 * 1024;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(simplex_vectors_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured = smem;
    }
    simplex_vectors_kernel<<<nvec, threads, smem, stream>>>(A, out, n, vs_in, es_in, vs_out,
                                                            es_out, use_smem, flags);
    CDR_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

}  // namespace cdr

extern "C" int cdr_simplex_project_rows(const double* A, double* out, int m, int n, long lda,
                                        long ldo, const cdr_flags* flags, cdr_stream_t stream)
{
    CDR_CHECK_ARG(m >= 0 && n >= 0 && lda >= n && ldo >= n);
    return cdr::launch_simplex(A, out, m, n, lda, 1, ldo, 1, flags, (cudaStream_t)stream);
}

extern "C" int cdr_simplex_project_columns(const double* A, double* out, int m, int n, long lda,
                                           long ldo, const cdr_flags* flags, cdr_stream_t stream)
{
    CDR_CHECK_ARG(m >= 0 && n >= 0 && lda >= n && ldo >= n);
    // one vector per column: n vectors of length m, element stride = leading dimension
    return cdr::launch_simplex(A, out, n, m, 1, lda, 1, ldo, flags, (cudaStream_t)stream);
}
