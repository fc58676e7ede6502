// Peer-memory collectives over NVLink / NVSwitch (one process per GPU).
//
// Every rank allocates one "symmetric region" of the same size with cudaMalloc, exports it
// with CUDA IPC and maps the regions of all other ranks; a buffer that takes part in a
// collective lives at the same offset in every region.  Kernels then exchange data with
// plain loads / stores on the mapped peer pointers and synchronise through 64-bit flags
// that carry a monotonically increasing epoch (no resets, so an iteration captured in a
// CUDA graph can be replayed):
//
//   writer : data stores, __threadfence_system(), CTA barrier, st.release.sys flag = epoch
//   reader : spin on ld.acquire.sys flag >= epoch, CTA barrier, ld.relaxed.sys data
//
// Epochs come from launch counters kept in the region header, which advance identically on
// every rank because all ranks issue the same sequence of collective launches with the same
// grids (the replicated k x k / k x T state is bit-identical across ranks, so even the
// device-side `done` early exit is taken by all ranks together).  The stand-alone collectives
// index both their flags and their counters by CTA; the fused kernel indexes its flags by
// strip, so it uses one counter per launch (bumped by the last CTA to leave), which stays
// monotonic for every strip whatever grid a later launch uses.
// Sums are formed by one owner per chunk in rank order 0..world-1 and pushed to every
// rank, which keeps the replicas bit-identical and the result run-to-run deterministic.
//
// All waits are bounded: a flag that does not arrive within ~2 s sets PeerHeader::error
// and the kernel carries on (wrong numbers, but no hung GPU); the host checks the flag.
#pragma once

#include "cdr_common.cuh"

namespace cdr {
namespace peer {

// The first CDR_PEER_HEADER_BYTES of every region.  Zero-initialised at allocation.
struct PeerHeader {
    unsigned long long fused_epoch;                      // completed launches of the fused kernel
    unsigned int fused_tickets;                          // CTAs of the current launch that left
    unsigned int reserved_;
    unsigned long long coll_seq[CDR_PEER_MAX_CTAS];     // launches of the stand-alone collectives
    unsigned long long coll_start[CDR_MAX_PEERS][CDR_PEER_MAX_CTAS];   // [from rank][cta]
    unsigned long long coll_finish[CDR_MAX_PEERS][CDR_PEER_MAX_CTAS];
    unsigned long long ready[CDR_PEER_MAX_STRIPS][CDR_MAX_PEERS];      // [strip][from rank]
    unsigned long long done[CDR_PEER_MAX_STRIPS];
    int error;                                           // 0, or the wait that timed out
};
static_assert(sizeof(PeerHeader) <= CDR_PEER_HEADER_BYTES, "peer header does not fit");

enum PeerWait { kWaitStart = 1, kWaitFinish = 2, kWaitReady = 3, kWaitDone = 4 };

__device__ __forceinline__ PeerHeader* header_of(const cdr_peer_group& g, int r)
{
    return reinterpret_cast<PeerHeader*>(g.region[r]);
}

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// data written by another GPU: read at the system coherence point, never from L1
__device__ __forceinline__ double2 ld_sys_d2(const double* p)
{
    double2 v;
    asm volatile("ld.relaxed.sys.global.v2.f64 {%0, %1}, [%2];\n"
                 : "=d"(v.x), "=d"(v.y)
                 : "l"(p)
                 : "memory");
    return v;
}

__device__ __forceinline__ double ld_sys_d(const double* p)
{
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];\n" : "=d"(v) : "l"(p) : "memory");
    return v;
}

constexpr long long kWaitCycles = 4000000000LL;      // ~2 s at 1.9 GHz

// Spin until *flag >= epoch.  On timeout records `what` in the local header and returns.
__device__ __forceinline__ void wait_flag(const unsigned long long* flag, unsigned long long epoch,
                                          PeerHeader* mine, int what)
{
    if (ld_acquire_sys(flag) >= epoch) return;
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < epoch) {
        // after the first time-out every later wait gives up at once: a dead peer costs ~2 s
        // in total, not 2 s per flag
        if (*((volatile int*)&mine->error) != 0) return;
        if (clock64() - t0 > kWaitCycles) {
            atomicExch(&mine->error, what);
            return;
        }
    }
}

// Barrier over CTA `blockIdx.x` of all ranks (every thread of the CTA must call it).
// `which` selects the flag array (start / finish) inside the headers.
template <int WHICH>
__device__ __forceinline__ void cta_barrier_all_ranks(const cdr_peer_group& g,
                                                      unsigned long long epoch)
{
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < g.world) {
        const int p = threadIdx.x;
        PeerHeader* theirs = header_of(g, p);
        PeerHeader* mine = header_of(g, g.rank);
        unsigned long long* remote = (WHICH == kWaitStart ? theirs->coll_start[g.rank]
                                                          : theirs->coll_finish[g.rank]) + blockIdx.x;
        const unsigned long long* local = (WHICH == kWaitStart ? mine->coll_start[p]
                                                               : mine->coll_finish[p]) + blockIdx.x;
        st_release_sys(remote, epoch);
        wait_flag(local, epoch, mine, WHICH);
    }
    __syncthreads();
}

// Arguments of the fused reduce-over-samples + all-reduce kernel (stream_tma.cu).
struct SamplesExchange {
    cdr_peer_group g;
    size_t out_offset;      // byte offset of `out` (k x ldo) in every region
};

struct NoExchange {};

}  // namespace peer
}  // namespace cdr
