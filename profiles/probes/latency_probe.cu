// Dependent-issue latencies (cycles) of the instructions on the critical path of the per-sample
// SPG solver (qp_core.cuh), one warp, measured with clock64() over chains of 256 operations.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o latency_probe latency_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

#define N 256
template <int OP>
__global__ void probe(double* out, long long* cyc, double a, double b)
{
    double x = a + threadIdx.x * 1e-9, y = b;
    int iv = threadIdx.x;
    __shared__ double sm[64];
    sm[threadIdx.x] = x;
    sm[threadIdx.x + 32] = y;
    __syncwarp();
    const long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        if (OP == 0) x = fma(x, y, y);                                   // DFMA
        if (OP == 1) x = x + y;                                          // DADD
        if (OP == 2) x = __shfl_xor_sync(0xffffffffu, x, 1);             // 64-bit SHFL
        if (OP == 3) x = __shfl_xor_sync(0xffffffffu, x, 1) + y;         // SHFL + DADD (a reduction level)
        if (OP == 4) x = (x > y) ? x : y + x;                            // DSETP + select
        if (OP == 5) iv = __popc(__ballot_sync(0xffffffffu, iv & 1)) + iv;   // VOTE + POPC
        if (OP == 6) x = x / y;                                          // IEEE division
        if (OP == 7) x = sm[(__double2loint(x) & 31)] + y;               // LDS + DADD
        if (OP == 8) x = fmax(x, y * x);                                 // DMUL + fmax
        if (OP == 9) x = sqrt(x) + y;                                    // sqrt
        if (OP == 10) iv = __shfl_xor_sync(0xffffffffu, iv, 1) + 1;      // 32-bit SHFL + IADD
        if (OP == 11) x = x * y;                                         // DMUL
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[OP] = t1 - t0;
    out[OP * 32 + threadIdx.x] = x + iv;
}

int main()
{
    double* out;
    long long* cyc;
    cudaMalloc(&out, 16 * 32 * sizeof(double));
    cudaMallocManaged(&cyc, 16 * sizeof(long long));
    const char* names[] = {"DFMA", "DADD", "SHFL.64", "SHFL.64 + DADD", "DSETP + select (+DADD)",
                           "VOTE + POPC + IADD", "division", "LDS + DADD", "DMUL + fmax", "sqrt + DADD",
                           "SHFL.32 + IADD", "DMUL"};
    for (int rep = 0; rep < 2; ++rep) {
        probe<0><<<1, 32>>>(out, cyc, 1.0, 1.0000001);
        probe<1><<<1, 32>>>(out, cyc, 1.0, 1.0000001);
        probe<2><<<1, 32>>>(out, cyc, 1.0, 1.0000001);
        probe<3><<<1, 32>>>(out, cyc, 1.0, 1.0000001);
        probe<4><<<1, 32>>>(out, cyc, 1.0, 1.0000001);
        probe<5><<<1, 32>>>(out, cyc, 1.0, 1.0000001);
        probe<6><<<1, 32>>>(out, cyc, 1.0, 1.0000001);
        probe<7><<<1, 32>>>(out, cyc, 1.0, 1.0000001);
        probe<8><<<1, 32>>>(out, cyc, 1.0, 1.0000001);
        probe<9><<<1, 32>>>(out, cyc, 1.0, 1.0000001);
        probe<10><<<1, 32>>>(out, cyc, 1.0, 1.0000001);
        probe<11><<<1, 32>>>(out, cyc, 1.0, 1.0000001);
        cudaDeviceSynchronize();
    }
    for (int i = 0; i < 12; ++i) printf("%-26s %7.1f cycles per dependent step\n", names[i], (double)cyc[i] / N);
    return 0;
}
