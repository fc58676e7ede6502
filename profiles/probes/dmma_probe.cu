// Probe: fp64 throughput of the tensor pipe (DMMA.8x8x4) and of the FMA pipe (DFMA) on this GPU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters, double seed)
{
    double c[CHAINS][2];
    for (int i = 0; i < CHAINS; ++i) c[i][0] = c[i][1] = 0.0;
    double a = seed + threadIdx.x, b = seed * 0.5 + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < CHAINS; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double seed)
{
    double c[CHAINS];
    for (int i = 0; i < CHAINS; ++i) c[i] = i;
    double a = seed + threadIdx.x * 1e-9, b = seed * 0.5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
    for (int i = 0; i < CHAINS; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float timeit(F f)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();
    float best = 1e9;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    double* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(double) * 4);
    const int iters = 20000;
    for (int blocks_per_sm : {1, 2, 4}) {
        const int grid = 148 * blocks_per_sm;
        float ms = timeit([&] { dmma_kernel<8><<<grid, 256>>>(out, iters, 1.0); });
        double flops = 2.0 * 8 * 8 * 4 * 8.0 * iters * (256 / 32) * grid;
        printf("DMMA.8x8x4  8 chains, %d x 256 threads/SM: %.3f ms  %.2f TFLOP/s\n", blocks_per_sm, ms, flops / ms / 1e9);
        ms = timeit([&] { dmma_kernel<4><<<grid, 256>>>(out, iters, 1.0); });
        flops = 2.0 * 8 * 8 * 4 * 4.0 * iters * (256 / 32) * grid;
        printf("DMMA.8x8x4  4 chains, %d x 256 threads/SM: %.3f ms  %.2f TFLOP/s\n", blocks_per_sm, ms, flops / ms / 1e9);
        ms = timeit([&] { dfma_kernel<8><<<grid, 256>>>(out, iters, 1.0); });
        flops = 2.0 * 8.0 * iters * 256.0 * grid;
        printf("DFMA        8 chains, %d x 256 threads/SM: %.3f ms  %.2f TFLOP/s\n", blocks_per_sm, ms, flops / ms / 1e9);
    }
    printf("err=%d\n", (int)cudaGetLastError());
    return 0;
}
