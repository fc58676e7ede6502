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
    // per double two 8-byte words {32 data bits | 32-bit epoch tag}: data and flag travel in
    // the same store, so no fence and no separate flag round trip is needed
    unsigned long long small_ll[2][CDR_MAX_PEERS][2 * CDR_PEER_SMALL_MAX];
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

// "LL" words: a double travels as two 8-byte words, each carrying 32 data bits and a 32-bit
// tag (the epoch of the exchange).  Data and flag arrive in the same store, so the writer
// needs no fence and no separate flag message, and the reader polls the words themselves.
__device__ __forceinline__ void st_ll(unsigned long long* dst, double v, unsigned long long tag)
{
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    const unsigned long long lo = (bits & 0xffffffffull) | tag;
    const unsigned long long hi = (bits >> 32) | tag;
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};\n" ::"l"(dst), "l"(lo), "l"(hi)
                 : "memory");
}

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

// Polls a word pair until both carry `tag` (bounded like wait_flag) and returns the double.
__device__ __forceinline__ double ld_ll(const unsigned long long* src, unsigned long long tag,
                                        PeerHeader* mine, int what)
{
    unsigned long long lo, hi;
    long long t0 = 0;
    bool timing = false;
    while (true) {
        asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];\n"
                     : "=l"(lo), "=l"(hi)
                     : "l"(src)
                     : "memory");
        if ((lo & 0xffffffff00000000ull) == tag && (hi & 0xffffffff00000000ull) == tag) break;
        if (!timing) {
            t0 = clock64();
            timing = true;
        }
        if (*((volatile int*)&mine->error) != 0) break;
        if (clock64() - t0 > kWaitCycles) {
            atomicExch(&mine->error, what);
            break;
        }
    }
    return __longlong_as_double((long long)((lo & 0xffffffffull) | (hi << 32)));
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
// after it has reduced its own rank's partials.  One shot, low latency: every rank writes its
// vector into slot[rank] of every peer as 8-byte words that carry 32 data bits and a 32-bit
// epoch tag (the "LL" scheme: the flag travels inside the store, so neither a fence nor a
// separate flag message is needed; one NVLink one-way latency in all), then polls the slots
// of all peers in its own region and sums them in rank order (every rank gets the same bits).
// Doubles as a barrier between the ranks (vals = nullptr): bulk data pushed to peer memory
// by the calling kernel before this call has arrived when it returns, PROVIDED every CTA
// that pushed fenced at system scope before the caller learnt it is the last CTA (the
// pushes are then performed before the tagged words are even sent).  Two slot sets alternate:
// a rank can be at most one exchange ahead of its peers (it cannot finish exchange e + 1
// before every peer has entered it, i.e. has finished reading exchange e).  All threads of
// the CTA must call this.
__device__ __forceinline__ void cta_allreduce_small(const cdr_peer_group& g, double* vals, int n)
{
    PeerHeader* mine = header_of(g, g.rank);
    const unsigned long long epoch = *((volatile unsigned long long*)&mine->small_epoch) + 1;
    const int set = (int)(epoch & 1ull);
    const unsigned long long tag = (epoch & 0xffffffffull) << 32;
    const int n_eff = (vals == nullptr || n < 1) ? 1 : n;         // a barrier sends one word pair
    for (int r = 0; r < g.world; ++r) {
        unsigned long long* dst = header_of(g, r)->small_ll[set][g.rank];
        for (int i = threadIdx.x; i < n_eff; i += blockDim.x) {
            const unsigned long long bits =
                (vals == nullptr) ? 0ull : (unsigned long long)__double_as_longlong(vals[i]);
            const unsigned long long lo = (bits & 0xffffffffull) | tag;
            const unsigned long long hi = (bits >> 32) | tag;
            asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};\n" ::"l"(dst + 2 * i), "l"(lo),
                         "l"(hi)
                         : "memory");
        }
    }
    for (int i = threadIdx.x; i < n_eff; i += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < g.world; ++r) {
            const unsigned long long* src = mine->small_ll[set][r] + 2 * i;
            unsigned long long lo, hi;
            long long t0 = 0;
            bool timing = false;
            while (true) {
                asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];\n"
                             : "=l"(lo), "=l"(hi)
                             : "l"(src)
                             : "memory");
                if ((lo & 0xffffffff00000000ull) == tag && (hi & 0xffffffff00000000ull) == tag) break;
                if (!timing) {
                    t0 = clock64();
                    timing = true;
                }
                if (*((volatile int*)&mine->error) != 0) break;
                if (clock64() - t0 > kWaitCycles) {
                    atomicExch(&mine->error, kWaitSmall);
                    break;
                }
            }
            const double v = __longlong_as_double((long long)((lo & 0xffffffffull) | (hi << 32)));
            s = (r == 0) ? v : s + v;
        }
        if (vals != nullptr && i < n) vals[i] = s;
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
