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
#include "qp_core.cuh"

namespace cdr {

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
    // lane group q works on sample q % spw of the warp: with spw < 4 the remaining groups are
    // replicas, which share the line-search trials of their sample (qp_core.cuh)
    const int t_raw = warp_global * spw + ((lane >> 3) % spw);
    const bool has_sample = t_raw < T;
    const bool valid = ((lane >> 3) < spw) && has_sample;
    const long t = has_sample ? t_raw : (T - 1);

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

    double x[KPL], b[KPL], tmp[KPL];
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

    int n_iter = 0, n_feval = 0;
    qp_solve<KPL>(As, arow, tmp, b, present, p, has_sample, g, spw, x, n_iter, n_feval);

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
