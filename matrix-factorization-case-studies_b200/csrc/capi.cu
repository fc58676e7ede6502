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
