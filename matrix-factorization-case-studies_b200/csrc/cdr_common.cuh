// Shared device/host helpers for the convex_dim_red B200 kernels (sm_100a).
//
// Everything here is internal to the shared library; the public surface is
// include/cdr_b200.h.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/cdr_b200.h"

#define CDR_FULL_MASK 0xffffffffu

// Launch-error check used by every C-ABI entry point (no synchronisation:
// all entry points are asynchronous on the caller's stream).
// Also counts kernel launches (cdr_launch_count): the macro follows every <<<>>>.
// CDR_DEBUG_SYNC=1 (debugging aid): synchronise after every launch and report the first
// failing one with its source line.
extern unsigned long long cdr_g_kernel_launches;
static inline bool cdr_debug_sync()
{
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("CDR_DEBUG_SYNC");
        v = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}
// CDR_TIME_LAUNCHES=1 (debugging aid, eager launches only): an event after every launch;
// cdr_debug_timing_report() prints the time between consecutive events per launch site.
void cdr_debug_note_launch(const char* file, int line);
static inline bool cdr_debug_timing()
{
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("CDR_TIME_LAUNCHES");
        v = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}
#define CDR_RETURN_IF_LAUNCH_FAILED()                                                       \
    do {                                                                                    \
        if (cdr_debug_timing()) cdr_debug_note_launch(__FILE__, __LINE__);                  \
        cudaError_t e__ = cudaGetLastError();                                               \
        if (e__ == cudaSuccess && cdr_debug_sync()) e__ = cudaDeviceSynchronize();          \
        if (e__ != cudaSuccess) {                                                           \
            if (cdr_debug_sync())                                                           \
                fprintf(stderr, "[cdr] %s:%d: %s\n", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return (int)e__;                                                                \
        }                                                                                   \
        ++cdr_g_kernel_launches;                                                            \
    } while (0)

#define CDR_CHECK_ARG(cond)                                   \
    do {                                                      \
        if (!(cond)) return CDR_ERR_INVALID_ARGUMENT;         \
    } while (0)

namespace cdr {

__device__ __forceinline__ bool is_done(const cdr_flags* flags)
{
    // Every kernel of an iteration checks this first: once the on-device
    // convergence / error test has fired, the rest of a captured CUDA graph
    // degenerates to empty launches and the state is left untouched.
    return flags != nullptr && *((volatile const int*)&flags->done) != 0;
}

// --------------------------------------------------------------------------
// warp / block reductions (fixed order => deterministic results)
// --------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(CDR_FULL_MASK, v, o);
    return v;
}

__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(CDR_FULL_MASK, v, o));
    return v;
}

__device__ __forceinline__ int warp_sum_int(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(CDR_FULL_MASK, v, o);
    return v;
}

// Second stage of the block reductions for eight warps (256 threads): every aligned group of
// eight lanes combines the eight warp results in three butterfly steps.  The full five-step
// butterfly over lanes padded with the neutral element forms the same tree, so the results
// are identical.
__device__ __forceinline__ double group8_tree_sum(double x)
{
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) x += __shfl_xor_sync(CDR_FULL_MASK, x, o);
    return x;
}
__device__ __forceinline__ double group8_tree_max(double x)
{
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(CDR_FULL_MASK, x, o));
    return x;
}

// Block-wide sum of up to N values per thread.  `scratch` must hold
// N * 32 doubles.  All threads receive the result.  blockDim.x must be a
// multiple of 32 and <= 1024.
template <int N>
__device__ __forceinline__ void block_sum(double (&v)[N], double* scratch)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = warp_sum(v[i]);
    __syncthreads();                      // scratch may still be read from a previous call
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < N; ++i) scratch[i * 32 + warp] = v[i];
    }
    __syncthreads();
    if (nwarps == 8) {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = group8_tree_sum(scratch[i * 32 + (lane & 7)]);
        return;
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double x = (lane < nwarps) ? scratch[i * 32 + lane] : 0.0;
        v[i] = warp_sum(x);
    }
}

// sum and maximum in one exchange (scratch: 64 doubles); same results as block_sum<1> and
// block_max
__device__ __forceinline__ void block_sum_and_max(double& sum, double& mx, double* scratch)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    sum = warp_sum(sum);
    mx = warp_max(mx);
    __syncthreads();                      // scratch may still be read from a previous call
    if (lane == 0) {
        scratch[warp] = sum;
        scratch[32 + warp] = mx;
    }
    __syncthreads();
    if (nwarps == 8) {
        sum = group8_tree_sum(scratch[lane & 7]);
        mx = group8_tree_max(scratch[32 + (lane & 7)]);
        return;
    }
    sum = warp_sum((lane < nwarps) ? scratch[lane] : 0.0);
    mx = warp_max((lane < nwarps) ? scratch[32 + lane] : -INFINITY);
}

__device__ __forceinline__ double block_max(double v, double* scratch)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    if (nwarps == 8) return group8_tree_max(scratch[lane & 7]);
    double x = (lane < nwarps) ? scratch[lane] : -INFINITY;
    return warp_max(x);
}

// --------------------------------------------------------------------------
// fp64 tensor-core tile: D(8x8) += A(8x4, row) * B(4x8, col).  SASS: DMMA.8x8x4
//   a : A[lane>>2][lane&3]        b : B[lane&3][lane>>2]
//   c0, c1 : C[lane>>2][2*(lane&3) + {0,1}]
// --------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b)
{
    asm volatile(
        "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

__device__ __forceinline__ double2 ldg_nc_d2(const double* p)
{
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];\n"
                 : "=d"(v.x), "=d"(v.y)
                 : "l"(p));
    return v;
}

// --------------------------------------------------------------------------
// SPG scalar helpers (reference spg.py:19-43)
// --------------------------------------------------------------------------
__host__ __device__ __forceinline__ double spg_step_length(double lam, double delta,
                                                           double f_old, double f_new,
                                                           double sigma_one, double sigma_two)
{
    // spg.py:23-31: safeguarded quadratic interpolation; note the lower
    // safeguard sigma_one is absolute, not relative to lam.
    const double cand = -0.5 * lam * lam * delta / (f_new - f_old - lam * delta);
    if (sigma_one <= cand && cand <= sigma_two * lam) return cand;
    return 0.5 * lam;
}

__host__ __device__ __forceinline__ double spg_cauchy_step(double beta, double sksk,
                                                           double alpha_min, double alpha_max)
{
    // spg.py:36-43
    if (beta <= 0.0) return alpha_max;
    return fmin(alpha_max, fmax(alpha_min, sksk / beta));
}

static inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }

// stream_tma.cu: bulk-copy pipelined variants; return CDR_TMA_NOT_APPLICABLE when the
// shape should be handled by the direct-load kernels of stream_gemm.cu
#define CDR_TMA_NOT_APPLICABLE (-100)
int run_reduce_samples_tma(const double* Lp, long sLi, long sLt, const double* X, long ldx, int T,
                           int d, int k, const double* E, double* out, long ldo,
                           const cdr_flags* flags, cudaStream_t stream);
// Optional by-product of the strip-owned reduce over features (the fused GPNH iteration,
// iterate.cu): the k x k Gram matrix M M' of the operand itself.  The B fragments of a strip
// serve as both DMMA operands (a fragment value M[lc][f(lr)] is at once A[lc][lr] and
// B[lr][lc]); the per-strip results go to part[strip][KP * KP] and the last CTA to finish sums
// them in strip order into out.
struct StripGram {
    double* part;            // nullptr: no by-product
    double* out;             // k x k, row-major
    unsigned int* ticket;    // zero before the first launch; left at zero
};
// out == nullptr: no finalize launch, the caller consumes the per-strip partials
// workspace[strip][t][KP] itself (KP = 8 or 16; geometry from features_strip_geometry)
int run_reduce_features_tma(const double* M, long ldm, const double* X, long ldx, int T, int d,
                            int k, double* out, long ldo, void* workspace, size_t workspace_bytes,
                            const cdr_flags* flags, cudaStream_t stream,
                            const StripGram* gram = nullptr);
bool features_strip_geometry(int T, int d, int k, int* TC, int* nstrips);
size_t reduce_features_tma_workspace_bytes(int T, int d, int k);
void tma_stream_plan(int T, int d, int k, int with_epilogue, int* out);
// stream_gemm64.cu: GEMM-shaped kernels for 16 < k <= 64 (tensor-bound shapes); return
// CDR_TMA_NOT_APPLICABLE for shapes they do not take
bool gemm64_applicable(int T, int d, int k);
size_t features64_workspace_bytes(int T, int d, int k);
size_t samples64_workspace_bytes(int T, int d, int k);
int run_reduce_features64(const double* M, long ldm, const double* X, long ldx, int T, int d, int k,
                          double* out, long ldo, void* workspace, size_t workspace_bytes,
                          const cdr_flags* flags, cudaStream_t stream);
int run_reduce_samples64(const double* Lp, long sLi, long sLt, const double* X, long ldx, int T, int d,
                         int k, const double* E, double* out, long ldo, void* workspace,
                         size_t workspace_bytes, const cdr_flags* flags, cudaStream_t stream);
int run_reduce_samples_exchange(const cdr_peer_group& g, size_t out_offset, const double* Lp,
                                long sLi, long sLt, const double* X, long ldx, int T, int T_min,
                                int d, int k, const double* E, long ldo, const cdr_flags* flags,
                                cudaStream_t stream);

}  // namespace cdr
