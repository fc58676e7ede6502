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
    int reserved2_;
    // one-shot exchanges of small vectors by the tails of the fused iteration kernels
    // (cta_allreduce_small): two slot sets used alternately, one slot and one flag per peer
    unsigned long long small_epoch;
    unsigned long long small_flag[2][CDR_MAX_PEERS];
    double small_slot[2][CDR_MAX_PEERS][CDR_PEER_SMALL_MAX];
};
static_assert(sizeof(PeerHeader) <= CDR_PEER_HEADER_BYTES, "peer header does not fit");

enum PeerWait { kWaitStart = 1, kWaitFinish = 2, kWaitReady = 3, kWaitDone = 4, kWaitSmall = 5 };

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

// In-place sum over ranks of the n <= CDR_PEER_SMALL_MAX doubles vals[0..n) (shared or global
// memory of the calling CTA), executed by ONE CTA per rank -- the last CTA of a fused kernel,
// after it has reduced its own rank's partials.  One shot: every rank pushes its vector into
// slot[rank] of every peer, raises a flag there, waits for the flags of all peers and sums the
// slots in rank order (so every rank gets the same bits).  Doubles as a barrier between the
// ranks (n = 0): data pushed to peer memory by the calling kernel before this call -- by any
// of its CTAs, provided they fenced at system scope before the caller learnt it is last --
// has arrived when it returns.  Two slot sets alternate: a rank can be at most one exchange
// ahead of its peers (it cannot finish exchange e + 1 before every peer has entered it, i.e.
// has finished reading exchange e).  All threads of the CTA must call this.
__device__ __forceinline__ void cta_allreduce_small(const cdr_peer_group& g, double* vals, int n)
{
    PeerHeader* mine = header_of(g, g.rank);
    const unsigned long long epoch = *((volatile unsigned long long*)&mine->small_epoch) + 1;
    const int set = (int)(epoch & 1ull);
    for (int r = 0; r < g.world; ++r) {
        double* dst = header_of(g, r)->small_slot[set][g.rank];
        for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = vals[i];
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < g.world) {
        const int p = threadIdx.x;
        st_release_sys(&header_of(g, p)->small_flag[set][g.rank], epoch);
        wait_flag(&mine->small_flag[set][p], epoch, mine, kWaitSmall);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double s = ld_sys_d(&mine->small_slot[set][0][i]);
        for (int r = 1; r < g.world; ++r) s += ld_sys_d(&mine->small_slot[set][r][i]);
        vals[i] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) *((volatile unsigned long long*)&mine->small_epoch) = epoch;
}

// byte offset of a local pointer inside this rank's region
__device__ __host__ __forceinline__ size_t region_offset(const cdr_peer_group& g, const void* p)
{
    return (size_t)(static_cast<const unsigned char*>(p) -
                    static_cast<const unsigned char*>(g.region[g.rank]));
}

// the same location in the region of rank r
template <class T>
__device__ __forceinline__ T* peer_ptr(const cdr_peer_group& g, int r, size_t offset)
{
    return reinterpret_cast<T*>(static_cast<unsigned char*>(g.region[r]) + offset);
}

// Arguments of the fused reduce-over-samples + all-reduce kernel (stream_tma.cu).
struct SamplesExchange {
    cdr_peer_group g;
    size_t out_offset;      // byte offset of `out` (k x ldo) in every region
};

struct NoExchange {};

}  // namespace peer
}  // namespace cdr
