// Bulk-copy (TMA, cp.async.bulk -> SASS UBLKCP) + mbarrier producer/consumer
// pipeline used by the streaming contractions.
//
// One producer thread per CTA issues 1-D bulk copies of row segments of X
// straight into (padded) shared-memory tiles and arms a "full" mbarrier with the
// byte count; eight consumer warps wait on it, feed the DMMA fragments from
// shared memory and release the stage through an "empty" mbarrier.  This keeps
// ~150-200 KB per SM in flight with a handful of registers, which is what the
// HBM-bound passes need (the direct-LDG kernels of stream_gemm.cu stall at
// 50-65 % of the HBM peak for lack of memory-level parallelism).
#pragma once

#include "cdr_common.cuh"

namespace cdr {
namespace tma {

constexpr int kMaxStages = 8;
constexpr int kConsumerWarps = 8;
constexpr int kThreads = (kConsumerWarps + 1) * 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void fence_barrier_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}

// global -> shared bulk copy; dst, src and bytes must be multiples of 16
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::
            "r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// L2 eviction-priority hints for the bulk copies (createpolicy + .L2::cache_hint).  X is
// several times the L2 and is streamed once per pass: its copies are marked evict_first so
// that they do not displace what is re-read (see l2_resident_from_row in stream_tma.cu, which
// can also mark a band of rows evict_last for experiments).
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(p));
    return p;
}

__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(p));
    return p;
}

__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar,
                                              uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1], %2, [%3], %4;\n" ::"r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

struct Pipeline {
    uint64_t* full;
    uint64_t* empty;
    int stages;

    __device__ __forceinline__ void init(uint64_t* bars, int n_stages)
    {
        full = bars;
        empty = bars + kMaxStages;
        stages = n_stages;
        if (threadIdx.x == 0) {
            for (int s = 0; s < n_stages; ++s) {
                mbar_init(&full[s], 1);
                mbar_init(&empty[s], kConsumerWarps);
            }
            fence_barrier_init();
        }
    }
};

}  // namespace tma
}  // namespace cdr
