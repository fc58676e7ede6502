// Library-level entry points.
#include "cdr_common.cuh"

extern "C" const char* cdr_version(void) { return "cdr_b200 0.1 (sm_100a)"; }

extern "C" int cdr_device_check(void)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) return (int)e;
    return (prop.major == 10) ? 0 : CDR_ERR_UNSUPPORTED;
}

unsigned long long cdr_g_kernel_launches = 0;

// Number of kernels this library has launched (or recorded into a graph) so far.
extern "C" unsigned long long cdr_launch_count(void) { return cdr_g_kernel_launches; }


// fp64 tensor-pipe throughput probe (the roofline denominator of the tensor-bound shapes; not
// in MEASURED_PEAKS.json): every warp runs 8 independent chains of DMMA.8x8x4 on registers.
namespace cdr {
__global__ void __launch_bounds__(256) dmma_probe_kernel(double* out, int iters, double seed)
{
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
    const double a = seed + threadIdx.x, b = seed * 0.5 + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace cdr

// Launches the probe on `blocks` CTAs of 256 threads; the caller times it.  Flops of one launch:
// 2 * 8*8*4 * 8 chains * iters * 8 warps * blocks.  out: blocks * 256 doubles.
extern "C" int cdr_debug_dmma_probe(double* out, int blocks, int iters, cdr_stream_t stream)
{
    CDR_CHECK_ARG(out != nullptr && blocks >= 1 && iters >= 1);
    cdr::dmma_probe_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(out, iters, 1.0);
    CDR_RETURN_IF_LAUNCH_FAILED();
    return 0;
}
