// Probe: what read bandwidth does a cp.async.bulk + mbarrier ring reach on this GPU as a
// function of copy size, copies per stage, ring depth and access pattern?  (No compute.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_probe tma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mb_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mb_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mb_wait(uint64_t* b, uint32_t ph) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p; }" : "=r"(ok) : "r"(s32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk(void* d, const void* s, uint32_t n, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(d)), "l"(s), "r"(n), "r"(s32(b)) : "memory");
}

// pattern 0: CTA streams a contiguous chunk.  pattern 1: CTA owns a strip of `copy_bytes`
// per row, rows `row_stride` bytes apart (the reduce-over-samples access pattern).
__global__ void __launch_bounds__(64, 1)
probe(const char* __restrict__ src, size_t total, int copy_bytes, int copies_per_stage, int stages,
      int pattern, size_t row_stride, int nrows, double* sink, const char* __restrict__ msrc = nullptr,
      int m_rows = 0, int rows_per_cta = 11)
{
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t* full = (uint64_t*)sm;
    uint64_t* empty = full + 16;
    unsigned char* tiles = sm + 256;
    const size_t stage_bytes = (size_t)copy_bytes * copies_per_stage;
    // pattern 2: copies_per_stage = rows_per_cta + m_rows; the CTA owns rows_per_cta rows and walks
    // along them in chunks of copy_bytes, plus m_rows rows of a small second matrix
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mb_init(&full[s], 1); mb_init(&empty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long nst;
    if (pattern == 0) {
        const size_t chunk = total / gridDim.x / stage_bytes * stage_bytes;
        nst = chunk / stage_bytes;
    } else if (pattern == 1) {
        nst = nrows / copies_per_stage;
    } else {
        nst = row_stride / copy_bytes;
    }
    if (warp == 1) {
        for (long it = 0; it < nst; ++it) {
            const int s = it % stages; const uint32_t ph = (it / stages) & 1;
            if (lane == 0) { mb_wait(&empty[s], ph ^ 1); mb_expect(&full[s], (uint32_t)stage_bytes); }
            __syncwarp();
            for (int c = lane; c < copies_per_stage; c += 32) {
                const char* g;
                if (pattern == 0) {
                    const size_t chunk = total / gridDim.x / stage_bytes * stage_bytes;
                    g = src + (size_t)blockIdx.x * chunk + it * stage_bytes + (size_t)c * copy_bytes;
                } else if (pattern == 1) {
                    g = src + ((size_t)it * copies_per_stage + c) * row_stride + (size_t)blockIdx.x * copy_bytes;
                } else if (c < rows_per_cta) {
                    g = src + ((size_t)blockIdx.x * rows_per_cta + c) * row_stride + (size_t)it * copy_bytes;
                } else {
                    g = msrc + (size_t)(c - rows_per_cta) * row_stride + (size_t)it * copy_bytes;
                }
                bulk(tiles + s * stage_bytes + (size_t)c * copy_bytes, g, copy_bytes, &full[s]);
            }
        }
    } else {
        double acc = 0;
        for (long it = 0; it < nst; ++it) {
            const int s = it % stages; const uint32_t ph = (it / stages) & 1;
            mb_wait(&full[s], ph);
            acc += ((double*)(tiles + s * stage_bytes))[lane];
            __syncwarp();
            if (lane == 0) mb_arrive(&empty[s]);
        }
        if (acc == 1.2345) sink[0] = acc;
    }
}

int main()
{
    const size_t row_stride = 352000, nrows = 1620;
    const size_t total = row_stride * nrows;           // 570 MB
    char* d; double* sink;
    cudaMalloc(&d, total); cudaMemset(d, 0, total); cudaMalloc(&sink, 8);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    struct Cfg { int pattern, copy, cps, stages, grid; };
    Cfg cfgs[] = {
        {0, 512, 64, 5, 148}, {0, 1024, 32, 5, 148}, {0, 2048, 16, 5, 148}, {0, 4096, 8, 5, 148},
        {0, 8192, 4, 5, 148}, {0, 16384, 2, 5, 148}, {0, 32768, 1, 5, 148}, {0, 16384, 1, 8, 148},
        {0, 16384, 2, 6, 296}, {0, 32768, 2, 3, 148},
        {1, 2432, 8, 8, 144}, {1, 2432, 16, 5, 144}, {1, 2432, 4, 16, 144}, {1, 1216, 8, 16, 289},
        {1, 4864, 8, 5, 72}, {1, 2432, 27, 3, 144},
        {2, 3200, 11, 5, 147}, {2, 3200, 19, 3, 147}, {2, 4000, 11, 4, 147}, {2, 4000, 19, 2, 147},
        {2, 2000, 19, 5, 147}, {2, 3200, 16, 4, 101}, {2, 3200, 24, 2, 101},
    };
    char* dm; cudaMalloc(&dm, row_stride * 8); cudaMemset(dm, 0, row_stride * 8);
    for (const Cfg& c : cfgs) {
        const size_t smem = 256 + (size_t)c.copy * c.cps * c.stages;
        if (smem > 220 * 1024) { printf("skip\n"); continue; }
        const int nr = (int)(nrows / c.cps * c.cps);
        float best = 1e9;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0);
            if (c.pattern == 2) {
                const int rpc = (c.grid == 101) ? 16 : 11;
                probe<<<c.grid, 64, smem>>>(d, total, c.copy, c.cps, c.stages, 2, row_stride, nr, sink, dm, c.cps - rpc, rpc);
            } else
            probe<<<c.grid, 64, smem>>>(d, total, c.copy, c.cps, c.stages, c.pattern, row_stride, nr, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        if (c.pattern == 2) {
            const int rpc = (c.grid == 101) ? 16 : 11;
            size_t xb = (size_t)c.grid * rpc * (row_stride / c.copy * c.copy);
            printf("pattern 2 rows/cta %d + %d M rows, copy %d B, %d stages, grid %d: %.3f ms  %.0f GB/s of X\n", rpc, c.cps - rpc, c.copy, c.stages, c.grid, best, xb / best / 1e6);
            continue;
        }
        size_t bytes = c.pattern == 0 ? (total / c.grid / ((size_t)c.copy * c.cps) * ((size_t)c.copy * c.cps)) * c.grid
                                      : (size_t)nr * c.copy * c.grid;
        printf("pattern %d copy %6d B x %2d per stage, %2d stages (%3zu KB ring), grid %3d: %.3f ms  %.0f GB/s  err=%d\n",
               c.pattern, c.copy, c.cps, c.stages, (size_t)c.copy * c.cps * c.stages / 1024, c.grid, best,
               bytes / best / 1e6, (int)cudaGetLastError());
    }
    return 0;
}
