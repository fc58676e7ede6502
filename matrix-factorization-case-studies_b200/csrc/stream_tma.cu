// TMA-pipelined versions of the two streaming contractions (see stream_gemm.cu for
// the operation definitions and the fragment trick, stream_tma.cuh for the
// pipeline).  These are the production path at HadISST-like shapes; the
// direct-load kernels of stream_gemm.cu remain for shapes these do not cover
// (few features, k > 16 for the feature reduction).
//
//   reduce over samples  : a CTA owns a strip of TC ~ d/148 features and streams all
//                          T rows of it in 8-row stages; every consumer warp keeps
//                          the accumulators of its 16-feature units for the whole
//                          strip, so there are no partial sums, no workspace and no
//                          second kernel -- the k x k epilogue (E) is applied from
//                          shared memory before the single coalesced store.
//   reduce over features : the same strip-owned tile stream; the strip of M lives in
//                          registers as DMMA B fragments, the per-stage results of
//                          the four feature quarters are combined through shared
//                          memory and written as per-strip partials (a few % of the
//                          bytes of X), summed in fixed order by a finalize kernel.
#include <type_traits>

#include "stream_tma.cuh"
#include "peer.cuh"

namespace cdr {

using namespace tma;
using peer::NoExchange;
using peer::SamplesExchange;

// ======================================================================
// reduce over samples (strip owned)
// ======================================================================
// rows per stage (four DMMA k-steps).  Measured with profiles/probes/tma_probe.cu: the
// bulk-copy ring needs >= ~32 KB per stage to reach the HBM peak (8 rows x 2.4 KB stalls at
// 4.6 TB/s whatever the ring depth, 16 rows reach 6.5 TB/s).
constexpr int kSampTR = 16;
constexpr int kSampKS = kSampTR / 4;

// barrier over the eight consumer warps only (the producer warp runs ahead / has exited)
__device__ __forceinline__ void samples_consumer_barrier()
{
    asm volatile("bar.sync 2, %0;\n" ::"n"(kConsumerWarps * 32) : "memory");
}

// Exchange = NoExchange: `out` is the local result.
// Exchange = SamplesExchange (sample-sharded fit): the strip's tile is pushed into the inbox
// of the strip's owner rank (strip % world) instead; the owner sums the world tiles in rank
// order and pushes the sum into `out` of every rank (peer.cuh describes the flags).
template <int KT, int MAXU, bool FUSE_E, class Exchange = NoExchange>
__global__ void __launch_bounds__(kThreads, 1)
reduce_samples_tma_kernel(const double* __restrict__ Lp, long sLi, long sLt,
                          const double* __restrict__ X, long ldx, int T, int dpad, int k, int TC,
                          int nstrips, int stages, const double* __restrict__ E,
                          double* __restrict__ out, long ldo, const cdr_flags* flags, int keep_from,
                          Exchange xch = Exchange())
{
    if (is_done(flags)) return;
    constexpr int KP = 8 * KT;
    constexpr bool kExchange = !std::is_same<Exchange, NoExchange>::value;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
    double* Es = reinterpret_cast<double*>(smem_raw + 128);
    double* epi = Es + (FUSE_E ? KP * KP : 0);
    double* tiles = epi + ((FUSE_E || kExchange) ? kConsumerWarps * KP * 20 : 0);
    const int RS = TC + 4;                       // row stride (doubles): == 4 mod 16, conflict free
    const long stage_doubles = (long)kSampTR * RS;

    Pipeline pipe;
    pipe.init(bars, stages);
    if (FUSE_E) {
        for (int idx = threadIdx.x; idx < KP * KP; idx += blockDim.x) {
            const int j = idx / KP, i = idx % KP;
            Es[idx] = (j < k && i < k) ? E[j * k + i] : 0.0;
        }
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = (T + kSampTR - 1) / kSampTR;

    if (warp == kConsumerWarps) {
        // ------------------------------ producer warp: lane 0 owns the barriers, every
        // lane issues the bulk copy of one row (a single thread issuing all copies of a
        // stage is limited by the per-instruction issue latency, not by bytes)
        const uint64_t pol_last = l2_policy_evict_last(), pol_first = l2_policy_evict_first();
        int it = 0;
        for (int strip = blockIdx.x; strip < nstrips; strip += gridDim.x) {
            const int c0 = strip * TC;
            const int w = min(TC, dpad - c0);
            for (int rt = 0; rt < ntiles; ++rt, ++it) {
                const int s = it % stages;
                const uint32_t ph = (uint32_t)(it / stages) & 1u;
                const int rows = min(kSampTR, T - rt * kSampTR);
                if (lane == 0) {
                    mbar_wait(&pipe.empty[s], ph ^ 1u);
                    mbar_arrive_expect_tx(&pipe.full[s], (uint32_t)(rows * w * 8));
                }
                __syncwarp();
                if (lane < rows) {
                    double* dst = tiles + s * stage_doubles + (long)lane * RS;
                    const int row = rt * kSampTR + lane;
                    const double* src = X + (long)row * ldx + c0;
                    bulk_g2s_hint(dst, src, (uint32_t)(w * 8), &pipe.full[s],
                                  row >= keep_from ? pol_last : pol_first);
                }
            }
        }
        return;
    }

    // ---------------------------------- consumers
    const int lr = lane & 3, lc = lane >> 2;
    unsigned long long epoch = 0, tag = 0;
    size_t set_bytes = 0;                             // offset of this launch's set inside a slot
    if constexpr (kExchange) {
        epoch = *((volatile unsigned long long*)&peer::header_of(xch.g, xch.g.rank)->fused_epoch) + 1;
        tag = (epoch & 0xffffffffull) << 32;
        set_bytes = (size_t)(epoch & 1ull) * (xch.g.inbox_slot_bytes / 2);
    }
    unsigned long long* ll_dst = nullptr;             // tagged words of the strip owner's inbox
    int it = 0;
    for (int strip = blockIdx.x; strip < nstrips; strip += gridDim.x) {
        const int c0 = strip * TC;
        const int w = min(TC, dpad - c0);
        const int nunits = w / 16;
        if constexpr (kExchange) {
            // this rank's slot in the inbox of the strip's owner takes the place of `out`
            const cdr_peer_group& g = xch.g;
            ll_dst = reinterpret_cast<unsigned long long*>(
                static_cast<unsigned char*>(g.region[strip % g.world]) + g.inbox_offset +
                (size_t)g.rank * g.inbox_slot_bytes + set_bytes);
        }
        // one result element: a plain store, or -- sample-sharded -- two 8-byte words with 32
        // data bits and the 32-bit epoch tag each, pushed to the owner (peer.cuh: no fence, no
        // flag message; the reader polls the tags)
        auto emit = [&](long el, double v) {
            if constexpr (kExchange) {
                peer::st_ll(ll_dst + 2 * el, v, tag);
            } else {
                out[el] = v;
            }
        };

        double acc[MAXU][KT][2][2];
#pragma unroll
        for (int u = 0; u < MAXU; ++u)
#pragma unroll
            for (int mt = 0; mt < KT; ++mt)
                acc[u][mt][0][0] = acc[u][mt][0][1] = acc[u][mt][1][0] = acc[u][mt][1][1] = 0.0;

        // left-operand fragments L[i][t] are prefetched one stage ahead into registers: a
        // load issued right before use stalls the first DMMA of every stage on an L2 round
        // trip (27 % of the stall samples in profiles/r01b)
        double a_next[kSampKS][KT];
        auto load_left = [&](int rt_load) {
#pragma unroll
            for (int ks = 0; ks < kSampKS; ++ks) {
                const int tt = rt_load * kSampTR + 4 * ks + lr;
#pragma unroll
                for (int mt = 0; mt < KT; ++mt) {
                    const int i = mt * 8 + lc;
                    a_next[ks][mt] = (tt < T && i < k) ? Lp[(long)i * sLi + (long)tt * sLt] : 0.0;
                }
            }
        };
        load_left(0);
        for (int rt = 0; rt < ntiles; ++rt, ++it) {
            const int s = it % stages;
            const uint32_t ph = (uint32_t)(it / stages) & 1u;
            const int t0 = rt * kSampTR;
            double a[kSampKS][KT];
            bool rowok[kSampKS];
#pragma unroll
            for (int ks = 0; ks < kSampKS; ++ks) {
                rowok[ks] = (t0 + 4 * ks + lr) < T;
#pragma unroll
                for (int mt = 0; mt < KT; ++mt) a[ks][mt] = a_next[ks][mt];
            }
            if (rt + 1 < ntiles) load_left(rt + 1);
            mbar_wait(&pipe.full[s], ph);
            const double* tl = tiles + s * stage_doubles;
#pragma unroll
            for (int ks = 0; ks < kSampKS; ++ks) {
                const double* rowp = tl + (long)(4 * ks + lr) * RS + 2 * lc;
#pragma unroll
                for (int u = 0; u < MAXU; ++u) {
                    const int unit = warp + kConsumerWarps * u;
                    if (unit < nunits) {
                        double2 xv = make_double2(0.0, 0.0);
                        if (rowok[ks]) xv = *reinterpret_cast<const double2*>(rowp + unit * 16);
#pragma unroll
                        for (int mt = 0; mt < KT; ++mt) {
                            dmma884(acc[u][mt][0][0], acc[u][mt][0][1], a[ks][mt], xv.x);
                            dmma884(acc[u][mt][1][0], acc[u][mt][1][1], a[ks][mt], xv.y);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&pipe.empty[s]);
        }

        // ------------------------------ epilogue of the strip
        // lane holds, for row i = mt*8 + lc, features unit*16 + 4*lr + {0,1,2,3}
        //   = (acc[..][0][0], acc[..][1][0], acc[..][0][1], acc[..][1][1])
#pragma unroll
        for (int u = 0; u < MAXU; ++u) {
            const int unit = warp + kConsumerWarps * u;
            if (unit >= nunits) continue;
            const int fbase = c0 + unit * 16;
            if (!FUSE_E) {
#pragma unroll
                for (int mt = 0; mt < KT; ++mt) {
                    const int i = mt * 8 + lc;
                    if (i < k) {
                        const long el = (long)i * ldo + fbase + 4 * lr;
                        if constexpr (kExchange) {
                            // (handled below: the tile goes through shared memory so that the
                            // tagged words leave in 256-byte runs)
                        } else {
                            double* p = out + el;
                            *reinterpret_cast<double2*>(p) =
                                make_double2(acc[u][mt][0][0], acc[u][mt][1][0]);
                            *reinterpret_cast<double2*>(p + 2) =
                                make_double2(acc[u][mt][0][1], acc[u][mt][1][1]);
                        }
                    }
                }
                if constexpr (kExchange) {
                    // peer stores of 16 bytes with a 64-byte stride use a fraction of the link
                    // (r02n: +70 us per pass at 8 GPUs against +48 us for the variant below,
                    // whose lanes write consecutive words): same staging here
                    double* eb = epi + warp * (KP * 20);
#pragma unroll
                    for (int mt = 0; mt < KT; ++mt) {
                        double* p = eb + (mt * 8 + lc) * 20 + 4 * lr;
                        p[0] = acc[u][mt][0][0];
                        p[1] = acc[u][mt][1][0];
                        p[2] = acc[u][mt][0][1];
                        p[3] = acc[u][mt][1][1];
                    }
                    __syncwarp();
                    const int f = lane & 15, jh = lane >> 4;
                    for (int j = jh; j < k; j += 2) emit((long)j * ldo + fbase + f, eb[j * 20 + f]);
                    __syncwarp();
                }
            } else {
                double* eb = epi + warp * (KP * 20);
#pragma unroll
                for (int mt = 0; mt < KT; ++mt) {
                    double* p = eb + (mt * 8 + lc) * 20 + 4 * lr;
                    p[0] = acc[u][mt][0][0];
                    p[1] = acc[u][mt][1][0];
                    p[2] = acc[u][mt][0][1];
                    p[3] = acc[u][mt][1][1];
                }
                __syncwarp();
                const int f = lane & 15, jh = lane >> 4;
                double col[KP];
#pragma unroll
                for (int i = 0; i < KP; ++i) col[i] = eb[i * 20 + f];
                for (int j = jh; j < k; j += 2) {
                    double sacc = 0.0;
#pragma unroll
                    for (int i = 0; i < KP; ++i) sacc = fma(Es[j * KP + i], col[i], sacc);
                    emit((long)j * ldo + fbase + f, sacc);
                }
                __syncwarp();
            }
        }

        if constexpr (kExchange) {
            // the owner of the strip sums the tiles of all ranks in rank order as they arrive
            // (tagged words: no flag, no fence), stores the sum locally and pushes it, tagged,
            // into the result buffer of every other rank
            const cdr_peer_group& g = xch.g;
            if (strip % g.world == g.rank) {
                peer::PeerHeader* mine = peer::header_of(g, g.rank);
                unsigned char* base = static_cast<unsigned char*>(g.region[g.rank]) + g.inbox_offset;
                double* local_out = reinterpret_cast<double*>(
                    static_cast<unsigned char*>(g.region[g.rank]) + xch.out_offset);
                for (int idx = threadIdx.x; idx < k * w; idx += kConsumerWarps * 32) {
                    const int i = idx / w, c = idx - i * w;
                    const long el = (long)i * ldo + c0 + c;
                    // (requesting the word pairs of all ranks at once and polling only the
                    // late ones was measured at 8 GPUs: 162 us per pass against 148 us for
                    // this loop, profiles/r02/r02z_launch_timing_n8_poll_ab.txt -- the wait is
                    // for the data to arrive, not for the polls)
                    double sum = 0.0;
                    for (int r = 0; r < g.world; ++r) {
                        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(
                            base + (size_t)r * g.inbox_slot_bytes + set_bytes) + 2 * el;
                        const double v = peer::ld_ll(src, tag, mine, peer::kWaitReady);
                        sum = (r == 0) ? v : sum + v;
                    }
                    local_out[el] = sum;
                    for (int r = 0; r < g.world; ++r)
                        if (r != g.rank)
                            peer::st_ll(reinterpret_cast<unsigned long long*>(
                                            static_cast<unsigned char*>(g.region[r]) + g.inbox_offset +
                                            (size_t)g.world * g.inbox_slot_bytes + set_bytes) + 2 * el,
                                        sum, tag);
                }
            }
        }
    }

    if constexpr (kExchange) {
        // the sums of the strips other ranks own arrive, tagged, in this rank's result buffer
        const cdr_peer_group& g = xch.g;
        peer::PeerHeader* mine = peer::header_of(g, g.rank);
        const unsigned long long* res = reinterpret_cast<const unsigned long long*>(
            static_cast<unsigned char*>(g.region[g.rank]) + g.inbox_offset +
            (size_t)g.world * g.inbox_slot_bytes + set_bytes);
        double* local_out = reinterpret_cast<double*>(static_cast<unsigned char*>(g.region[g.rank]) +
                                                      xch.out_offset);
        for (int strip = blockIdx.x; strip < nstrips; strip += gridDim.x) {
            if (strip % g.world == g.rank) continue;
            const int c0 = strip * TC;
            const int w = min(TC, dpad - c0);
            for (int idx = threadIdx.x; idx < k * w; idx += kConsumerWarps * 32) {
                const int i = idx / w, c = idx - i * w;
                const long el = (long)i * ldo + c0 + c;
                local_out[el] = peer::ld_ll(res + 2 * el, tag, mine, peer::kWaitDone);
            }
        }
        samples_consumer_barrier();
        if (threadIdx.x == 0) {
            // the last CTA to leave publishes the epoch for the next launch (all CTAs of this
            // launch are resident and have read the old value long before)
            __threadfence();
            if (atomicAdd(&mine->fused_tickets, 1u) == gridDim.x - 1) {
                mine->fused_tickets = 0;
                __threadfence();
                *((volatile unsigned long long*)&mine->fused_epoch) = epoch;
            }
        }
    }
}

// ======================================================================
// reduce over features (strip owned)
// ======================================================================
// Same tile stream as the reduce over samples (the access pattern that reaches the HBM
// peak in profiles/probes/tma_probe.cu): a CTA owns a strip of TC features and walks
// down all T rows in 16-row stages.  The strip of M (k x TC) is held in registers as DMMA
// B fragments for the whole strip.  Warp w handles m-tile (w & 1) of the stage and the
// feature quarter (w >> 1); the four quarter results of a stage are combined through a
// double-buffered shared-memory exchange (one 256-thread named barrier per stage) and
// written as the strip's partial out[strip][t][j]; the strips are summed in fixed order by
// reduce_features_strip_finalize_kernel.
// The rows are walked BOTTOM-UP.  In an outer iteration the passes over X alternate -- reduce
// over samples (top-down; its summation order over t fixes the rounding, so it stays), then
// this kernel -- and X (570 MB at HadISST shape) is several times the 126 MB L2: walking the
// other way round, each pass starts on the rows the previous pass read last, which are still
// L2-resident, instead of on the rows that were evicted first.  The results do not depend on
// the row order (every sample is reduced on its own).
constexpr int kFsTR = 16;

__device__ __forceinline__ void consumer_barrier()
{
    asm volatile("bar.sync 1, %0;\n" ::"n"(kConsumerWarps * 32) : "memory");
}

// Optional by-product (the fused GPNH iteration, iterate.cu): the k x k Gram matrix M M' of the
// operand itself.  The B fragments of a strip serve as both DMMA operands (a fragment value
// M[lc][f(lr)] is at once A[lc][lr] and B[lr][lc]); the per-strip results go to
// gram.part[strip][KP * KP] and the last CTA to finish sums them in strip order into gram.out.

template <int KT, int MAXUQ>
__global__ void __launch_bounds__(kThreads, 1)
reduce_features_strip_kernel(const double* __restrict__ M, long ldm, const double* __restrict__ X,
                             long ldx, int T, int dpad, int k, int TC, int nstrips, int stages,
                             double* __restrict__ part, const cdr_flags* flags, StripGram gram,
                             int keep_from)
{
    if (is_done(flags)) return;
    constexpr int KP = 8 * KT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
    double* red = reinterpret_cast<double*>(smem_raw + 128);      // [2][4][2][8][KP]
    double* tiles = red + 2 * 4 * 2 * 8 * KP;
    const int RS = TC + 8;                       // == 8 mod 16: conflict-free row fragments
    const long stage_doubles = (long)kFsTR * RS;

    Pipeline pipe;
    pipe.init(bars, stages);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = (T + kFsTR - 1) / kFsTR;

    if (warp == kConsumerWarps) {
        const uint64_t pol_last = l2_policy_evict_last(), pol_first = l2_policy_evict_first();
        int it = 0;
        for (int strip = blockIdx.x; strip < nstrips; strip += gridDim.x) {
            const int c0 = strip * TC;
            const int w = min(TC, dpad - c0);
            for (int rt = 0; rt < ntiles; ++rt, ++it) {
                const int s = it % stages;
                const uint32_t ph = (uint32_t)(it / stages) & 1u;
                const int tile = ntiles - 1 - rt;          // bottom-up: see the kernel comment
                const int rows = min(kFsTR, T - tile * kFsTR);
                if (lane == 0) {
                    mbar_wait(&pipe.empty[s], ph ^ 1u);
                    mbar_arrive_expect_tx(&pipe.full[s], (uint32_t)(rows * w * 8));
                }
                __syncwarp();
                if (lane < rows)
                    bulk_g2s_hint(tiles + s * stage_doubles + (long)lane * RS,
                                  X + (long)(tile * kFsTR + lane) * ldx + c0, (uint32_t)(w * 8),
                                  &pipe.full[s],
                                  tile * kFsTR + lane >= keep_from ? pol_last : pol_first);
            }
        }
        return;
    }

    const int lr = lane & 3, lc = lane >> 2;
    const int mt = warp & 1, fq = warp >> 1;
    int it = 0;
    for (int strip = blockIdx.x; strip < nstrips; strip += gridDim.x) {
        const int c0 = strip * TC;
        const int w = min(TC, dpad - c0);
        const int nunits = w / 16;

        // B fragments of the strip of M: M[j = nt*8 + lc][unit*16 + 8p + 2*lr + {0,1}]
        double2 breg[MAXUQ][2][KT];
#pragma unroll
        for (int uq = 0; uq < MAXUQ; ++uq) {
            const int unit = fq + 4 * uq;
#pragma unroll
            for (int p = 0; p < 2; ++p)
#pragma unroll
                for (int nt = 0; nt < KT; ++nt) {
                    const int j = nt * 8 + lc;
                    double2 v = make_double2(0.0, 0.0);
                    if (unit < nunits && j < k)
                        v = *reinterpret_cast<const double2*>(M + (long)j * ldm + c0 + unit * 16 +
                                                              8 * p + 2 * lr);
                    breg[uq][p][nt] = v;
                }
        }

        if (gram.part != nullptr) {
            // M M' over the features of this strip: warps with mt == 0 cover the four
            // feature quarters, which are then combined in fixed order
            consumer_barrier();          // `red` may still be read by the previous strip's last tile
            if (mt == 0) {
                double wacc[KT][KT][2];
#pragma unroll
                for (int na = 0; na < KT; ++na)
#pragma unroll
                    for (int nb = 0; nb < KT; ++nb) wacc[na][nb][0] = wacc[na][nb][1] = 0.0;
#pragma unroll
                for (int uq = 0; uq < MAXUQ; ++uq)
#pragma unroll
                    for (int p = 0; p < 2; ++p)
#pragma unroll
                        for (int na = 0; na < KT; ++na)
#pragma unroll
                            for (int nb = 0; nb < KT; ++nb) {
                                dmma884(wacc[na][nb][0], wacc[na][nb][1], breg[uq][p][na].x,
                                        breg[uq][p][nb].x);
                                dmma884(wacc[na][nb][0], wacc[na][nb][1], breg[uq][p][na].y,
                                        breg[uq][p][nb].y);
                            }
#pragma unroll
                for (int na = 0; na < KT; ++na)
#pragma unroll
                    for (int nb = 0; nb < KT; ++nb)
                        *reinterpret_cast<double2*>(red + fq * KP * KP + (na * 8 + lc) * KP +
                                                    nb * 8 + 2 * lr) =
                            make_double2(wacc[na][nb][0], wacc[na][nb][1]);
            }
            consumer_barrier();
            for (int e = threadIdx.x; e < KP * KP; e += kConsumerWarps * 32)
                gram.part[(long)strip * KP * KP + e] =
                    ((red[e] + red[KP * KP + e]) + red[2 * KP * KP + e]) + red[3 * KP * KP + e];
            consumer_barrier();
        }

        for (int rt = 0; rt < ntiles; ++rt, ++it) {
            const int s = it % stages;
            const uint32_t ph = (uint32_t)(it / stages) & 1u;
            const int t = (ntiles - 1 - rt) * kFsTR + 8 * mt + lc;
            const bool rowok = t < T;
            double acc[KT][2];
#pragma unroll
            for (int nt = 0; nt < KT; ++nt) acc[nt][0] = acc[nt][1] = 0.0;

            mbar_wait(&pipe.full[s], ph);
            const double* rowp = tiles + s * stage_doubles + (long)(8 * mt + lc) * RS + 2 * lr;
#pragma unroll
            for (int uq = 0; uq < MAXUQ; ++uq) {
                const int unit = fq + 4 * uq;
                if (unit < nunits) {
#pragma unroll
                    for (int p = 0; p < 2; ++p) {
                        double2 xa = make_double2(0.0, 0.0);
                        if (rowok) xa = *reinterpret_cast<const double2*>(rowp + unit * 16 + 8 * p);
#pragma unroll
                        for (int nt = 0; nt < KT; ++nt) {
                            dmma884(acc[nt][0], acc[nt][1], xa.x, breg[uq][p][nt].x);
                            dmma884(acc[nt][0], acc[nt][1], xa.y, breg[uq][p][nt].y);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&pipe.empty[s]);

            // combine the four feature quarters of this stage (fixed order)
            const int buf = it & 1;
            double* mine = red + (((buf * 4 + fq) * 2 + mt) * 8 + lc) * KP + 2 * lr;
#pragma unroll
            for (int nt = 0; nt < KT; ++nt)
                *reinterpret_cast<double2*>(mine + nt * 8) = make_double2(acc[nt][0], acc[nt][1]);
            consumer_barrier();
            if (fq == 0 && rowok) {
                double* dst = part + ((long)strip * T + t) * KP + 2 * lr;
#pragma unroll
                for (int nt = 0; nt < KT; ++nt) {
                    double s0 = 0.0, s1 = 0.0;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const double2 v = *reinterpret_cast<const double2*>(
                            red + (((buf * 4 + q) * 2 + mt) * 8 + lc) * KP + 2 * lr + nt * 8);
                        s0 += v.x;
                        s1 += v.y;
                    }
                    *reinterpret_cast<double2*>(dst + nt * 8) = make_double2(s0, s1);
                }
            }
        }
    }

    if (gram.part != nullptr) {
        // the last CTA to get here sums the per-strip Gram matrices in strip order
        int* last = reinterpret_cast<int*>(red + 4 * KP * KP);
        __threadfence();
        consumer_barrier();
        if (threadIdx.x == 0) *last = (atomicAdd(gram.ticket, 1u) == gridDim.x - 1) ? 1 : 0;
        consumer_barrier();
        if (*last) {
            __threadfence();
            constexpr int E = KP * KP;                       // 64 or 256
            constexpr int NSUB = (kConsumerWarps * 32) / E;  // 4 or 1
            const int e = (int)threadIdx.x % E, sub = (int)threadIdx.x / E;
            double sacc = 0.0;
            for (int sidx = sub; sidx < nstrips; sidx += NSUB)
                sacc += __ldcg(gram.part + (long)sidx * E + e);
            red[sub * E + e] = sacc;
            consumer_barrier();
            if (sub == 0) {
                double tot = red[e];
#pragma unroll
                for (int q = 1; q < NSUB; ++q) tot += red[q * E + e];
                const int i = e / KP, j = e % KP;
                if (i < k && j < k) gram.out[i * k + j] = tot;
            }
            if (threadIdx.x == 0) *gram.ticket = 0u;
        }
    }
}

// out[j][t] = sum_strip part[strip][t][j].  A CTA handles 32 (t, component pair) items;
// its 8 warps sum disjoint groups of strips which are then combined in fixed order.
template <int KT>
__global__ void __launch_bounds__(256)
reduce_features_strip_finalize_kernel(const double* __restrict__ part, int T, int nstrips, int k,
                                      double* __restrict__ out, long ldo, const cdr_flags* flags)
{
    if (is_done(flags)) return;
    constexpr int KP = 8 * KT;
    constexpr int PAIRS = KP / 2;
    __shared__ double2 red[8][32];
    const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
    const long item = (long)blockIdx.x * 32 + lane;
    const long nitems = (long)T * PAIRS;
    const long stride = nitems;                                   // double2 per strip
    const int per = (nstrips + 7) / 8;
    double s0 = 0.0, s1 = 0.0;
    if (item < nitems) {
        const double2* src = reinterpret_cast<const double2*>(part) + item;
        const int hi = min(nstrips, (g + 1) * per);
        // 24 strips in flight (all of a group's strips at the usual widths), summed in order
        for (int s_lo = g * per; s_lo < hi; s_lo += 24) {
            double2 v[24];
#pragma unroll
            for (int q = 0; q < 24; ++q)
                v[q] = (s_lo + q < hi) ? src[(long)(s_lo + q) * stride] : make_double2(0.0, 0.0);
#pragma unroll
            for (int q = 0; q < 24; ++q) {
                s0 += v[q].x;
                s1 += v[q].y;
            }
        }
    }
    red[g][lane] = make_double2(s0, s1);
    __syncthreads();
    if (g == 0 && item < nitems) {
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            a0 += red[q][lane].x;
            a1 += red[q][lane].y;
        }
        const int t = (int)(item / PAIRS), j = 2 * (int)(item % PAIRS);
        if (j < k) out[(long)j * ldo + t] = a0;
        if (j + 1 < k) out[(long)(j + 1) * ldo + t] = a1;
    }
}

// ---------------------------------------------------------------------- host side
static int sm_count()
{
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// CDR_DISABLE_TMA=1 forces the direct-load kernels (A/B testing of the two paths)
static bool tma_disabled()
{
    const char* e = getenv("CDR_DISABLE_TMA");
    return e != nullptr && e[0] == '1';
}

constexpr size_t kSmemBudget = 220 * 1024;

// First row of the band of X that the strip kernels copy with evict_last (rows >= it; the
// others with evict_first).  Measured on B200 at HadISST shape (profiles/r02/
// r02s_l2_band_sweep.txt): marking ALL of X evict_first -- streaming data that will not be
// re-read before it is evicted anyway, so it should not displace the partials and the small
// replicated matrices -- makes the passes 2-8 % faster (100 / 102 -> 98 / 94 us); an
// evict_last band of 48-112 MB brought no L2 hits on top of that, so the default band is
// empty.  CDR_L2_RESIDENT_MB sets a band size for experiments.  A matrix of at most 96 MB is
// marked evict_last as a whole.
static int l2_resident_from_row(int T, int dpad)
{
    static long band_bytes = -1;
    if (band_bytes < 0) {
        const char* e = getenv("CDR_L2_RESIDENT_MB");
        band_bytes = (e != nullptr ? atol(e) : 0L) * 1000000L;
    }
    // a matrix that fits the L2 as a whole (a shard of a sample-sharded fit: 203 x 44 000 =
    // 71 MB) is kept resident entirely
    if ((long)T * dpad * 8 <= 96L * 1000000L) return 0;
    const long rows = band_bytes / ((long)dpad * 8);
    return rows >= T ? 0 : T - (int)rows;
}

template <auto Kern>
static int ensure_smem(size_t smem)
{
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(Kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured = smem;
    }
    return 0;
}

// strip width shared by the two strip-owned kernels: a multiple of 16 features, as close to
// dpad / (waves * #SM) as the unit allows
static void strip_geometry(int dpad, int tc_max, int* TC, int* nstrips)
{
    const int nsm = sm_count();
    int waves = (dpad + nsm * tc_max - 1) / (nsm * tc_max);
    if (waves < 1) waves = 1;
    int tc = (dpad + nsm * waves - 1) / (nsm * waves);
    tc = (tc + 15) / 16 * 16;
    if (tc > tc_max) tc = tc_max;
    *TC = tc;
    *nstrips = (dpad + tc - 1) / tc;
}

static int ring_stages(size_t fixed_bytes, size_t stage_bytes)
{
    int stages = (int)((kSmemBudget - fixed_bytes) / stage_bytes);
    return stages > kMaxStages ? kMaxStages : stages;
}

static size_t samples_fixed_smem(int kp, bool fuse_e, bool exchange = false)
{
    return 128 + (fuse_e ? (size_t)kp * kp * 8 : 0) +
           ((fuse_e || exchange) ? (size_t)kConsumerWarps * kp * 20 * 8 : 0);
}

// Strip plan of the reduce over samples; false when the direct-load (split-T) kernels should
// be used: k > 32, few or narrow strips (Gram-space "X" = K, PCA-reduced data), tiny T, or an
// epilogue matrix with k > 16.
static bool samples_strip_plan(int T, int dpad, int k, bool with_epilogue, int* TC, int* nstrips)
{
    if (k > 32 || (with_epilogue && k > 16)) return false;
    strip_geometry(dpad, (k <= 16) ? 512 : 256, TC, nstrips);
    return !(*nstrips < sm_count() / 2 || *TC < 64 || T < 64);
}

// returns CDR_TMA_NOT_APPLICABLE when the shape should use the direct-load kernels
template <int KT, int MAXU, bool FUSE_E>
static int launch_samples_tma(const double* Lp, long sLi, long sLt, const double* X, long ldx, int T,
                              int dpad, int k, int TC, int nstrips, const double* E, double* out,
                              long ldo, const cdr_flags* flags, cudaStream_t stream)
{
    constexpr int KP = 8 * KT;
    const size_t fixed = samples_fixed_smem(KP, FUSE_E);
    const size_t stage = (size_t)kSampTR * (TC + 4) * 8;
    const int stages = ring_stages(fixed, stage);
    if (stages < 2) return CDR_TMA_NOT_APPLICABLE;
    const size_t smem = fixed + stages * stage;
    int rc = ensure_smem<reduce_samples_tma_kernel<KT, MAXU, FUSE_E>>(smem);
    if (rc) return rc;
    const int grid = nstrips < sm_count() ? nstrips : sm_count();
    reduce_samples_tma_kernel<KT, MAXU, FUSE_E><<<grid, kThreads, smem, stream>>>(
        Lp, sLi, sLt, X, ldx, T, dpad, k, TC, nstrips, stages, E, out, ldo, flags,
        l2_resident_from_row(T, dpad));
    CDR_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

int run_reduce_samples_tma(const double* Lp, long sLi, long sLt, const double* X, long ldx, int T,
                           int d, int k, const double* E, double* out, long ldo,
                           const cdr_flags* flags, cudaStream_t stream)
{
    if (tma_disabled()) return CDR_TMA_NOT_APPLICABLE;
    if ((ldx % 2) != 0 || (((uintptr_t)X) & 15) != 0) return CDR_TMA_NOT_APPLICABLE;
    const int dpad = (d + 31) / 32 * 32;
    int TC, nstrips;
    if (!samples_strip_plan(T, dpad, k, E != nullptr, &TC, &nstrips)) return CDR_TMA_NOT_APPLICABLE;
    const int kt = (k + 7) / 8;
#define CDR_S(KT, MAXU)                                                                          \
    do {                                                                                         \
        if (E != nullptr && KT <= 2)                                                             \
            return launch_samples_tma<KT, MAXU, true>(Lp, sLi, sLt, X, ldx, T, dpad, k, TC,      \
                                                      nstrips, E, out, ldo, flags, stream);      \
        if (E != nullptr) return CDR_TMA_NOT_APPLICABLE; /* excluded by the plan */              \
        return launch_samples_tma<KT, MAXU, false>(Lp, sLi, sLt, X, ldx, T, dpad, k, TC, nstrips, \
                                                   nullptr, out, ldo, flags, stream);            \
    } while (0)
    if (kt == 1) CDR_S(1, 4);
    if (kt == 2) CDR_S(2, 4);
    if (kt == 3) CDR_S(3, 2);
    CDR_S(4, 2);
#undef CDR_S
}

// Fused with the sum over ranks (cdr_reduce_samples_allreduce).  The plan uses the smallest
// local T of any rank, so every rank takes the same decision and the same grid.
template <int KT, int MAXU, bool FUSE_E>
static int launch_samples_exchange(const cdr_peer_group& g, size_t out_offset, const double* Lp,
                                   long sLi, long sLt, const double* X, long ldx, int T, int dpad,
                                   int k, int TC, int nstrips, const double* E, long ldo,
                                   const cdr_flags* flags, cudaStream_t stream)
{
    constexpr int KP = 8 * KT;
    const size_t fixed = samples_fixed_smem(KP, FUSE_E, true);
    const size_t stage = (size_t)kSampTR * (TC + 4) * 8;
    const int stages = ring_stages(fixed, stage);
    if (stages < 2) return CDR_ERR_NOT_APPLICABLE;
    const size_t smem = fixed + stages * stage;
    int rc = ensure_smem<reduce_samples_tma_kernel<KT, MAXU, FUSE_E, SamplesExchange>>(smem);
    if (rc) return rc;
    const int grid = nstrips < sm_count() ? nstrips : sm_count();
    if (grid > CDR_PEER_MAX_CTAS || nstrips > CDR_PEER_MAX_STRIPS) return CDR_ERR_NOT_APPLICABLE;
    SamplesExchange xch;
    xch.g = g;
    xch.out_offset = out_offset;
    reduce_samples_tma_kernel<KT, MAXU, FUSE_E, SamplesExchange><<<grid, kThreads, smem, stream>>>(
        Lp, sLi, sLt, X, ldx, T, dpad, k, TC, nstrips, stages, E, nullptr, ldo, flags,
        l2_resident_from_row(T, dpad), xch);
    CDR_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

int run_reduce_samples_exchange(const cdr_peer_group& g, size_t out_offset, const double* Lp,
                                long sLi, long sLt, const double* X, long ldx, int T, int T_min,
                                int d, int k, const double* E, long ldo, const cdr_flags* flags,
                                cudaStream_t stream)
{
    if (tma_disabled() || k > 16) return CDR_ERR_NOT_APPLICABLE;
    if ((ldx % 2) != 0 || (((uintptr_t)X) & 15) != 0) return CDR_ERR_NOT_APPLICABLE;
    const int dpad = (d + 31) / 32 * 32;
    int TC, nstrips;
    if (!samples_strip_plan(T_min, dpad, k, E != nullptr, &TC, &nstrips))
        return CDR_ERR_NOT_APPLICABLE;
#define CDR_SX(KT)                                                                               \
    do {                                                                                         \
        if (E != nullptr)                                                                        \
            return launch_samples_exchange<KT, 4, true>(g, out_offset, Lp, sLi, sLt, X, ldx, T,  \
                                                        dpad, k, TC, nstrips, E, ldo, flags,     \
                                                        stream);                                 \
        return launch_samples_exchange<KT, 4, false>(g, out_offset, Lp, sLi, sLt, X, ldx, T,     \
                                                     dpad, k, TC, nstrips, nullptr, ldo, flags,  \
                                                     stream);                                    \
    } while (0)
    if (k <= 8) CDR_SX(1);
    CDR_SX(2);
#undef CDR_SX
}

constexpr int kFsTcMax = 384;

// Strip plan of the reduce over features (shape only): k <= 16, enough wide strips, and the
// per-strip partials must stay a small fraction of the bytes of X.
static bool features_strip_plan(int T, int dpad, int k, int* TC, int* nstrips)
{
    if (k > 16) return false;
    strip_geometry(dpad, kFsTcMax, TC, nstrips);
    if (*nstrips < sm_count() / 2 || *TC < 64 || T < 64) return false;
    const int kp = (k + 7) / 8 * 8;
    return (long)(*nstrips) * kp * 10 <= (long)dpad;
}

static bool features_strip_ok(const double* M, long ldm, const double* X, long ldx, int T, int d,
                              int k, int* TC, int* nstrips)
{
    if (tma_disabled()) return false;
    if ((ldx % 2) != 0 || (ldm % 2) != 0) return false;
    if ((((uintptr_t)X) & 15) != 0 || (((uintptr_t)M) & 15) != 0) return false;
    return features_strip_plan(T, (d + 31) / 32 * 32, k, TC, nstrips);
}

template <int KT>
static int launch_features_strip(const double* M, long ldm, const double* X, long ldx, int T,
                                 int dpad, int k, int TC, int nstrips, double* out, long ldo,
                                 void* workspace, size_t workspace_bytes, const cdr_flags* flags,
                                 cudaStream_t stream, const StripGram& gram)
{
    constexpr int KP = 8 * KT;
    constexpr int MAXUQ = kFsTcMax / 16 / 4;
    const size_t need = (size_t)nstrips * T * KP * sizeof(double);
    if (workspace == nullptr || workspace_bytes < need) return CDR_ERR_WORKSPACE;
    const size_t fixed = 128 + (size_t)2 * 4 * 2 * 8 * KP * 8;
    const size_t stage = (size_t)kFsTR * (TC + 8) * 8;
    const int stages = ring_stages(fixed, stage);
    if (stages < 2) return CDR_TMA_NOT_APPLICABLE;
    const size_t smem = fixed + stages * stage;
    int rc = ensure_smem<reduce_features_strip_kernel<KT, MAXUQ>>(smem);
    if (rc) return rc;
    const int grid = nstrips < sm_count() ? nstrips : sm_count();
    reduce_features_strip_kernel<KT, MAXUQ><<<grid, kThreads, smem, stream>>>(
        M, ldm, X, ldx, T, dpad, k, TC, nstrips, stages, (double*)workspace, flags, gram,
        l2_resident_from_row(T, dpad));
    CDR_RETURN_IF_LAUNCH_FAILED();
    if (out == nullptr) return 0;          // the caller consumes the per-strip partials itself
    const long nitems = (long)T * (KP / 2);
    reduce_features_strip_finalize_kernel<KT><<<(int)((nitems + 31) / 32), 256, 0, stream>>>(
        (const double*)workspace, T, nstrips, k, out, ldo, flags);
    CDR_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

size_t reduce_features_tma_workspace_bytes(int T, int d, int k)
{
    const int dpad = (d + 31) / 32 * 32;
    int TC, nstrips;
    strip_geometry(dpad, kFsTcMax, &TC, &nstrips);
    const int kp = (k + 7) / 8 * 8;
    if (k > 16 || (long)nstrips * kp * 10 > (long)dpad) return 0;
    return (size_t)nstrips * T * kp * sizeof(double);
}

// Host-only description of the strip plans (cdr_debug_stream_plan): out[0..4] samples
// {uses strips, TC, nstrips, stages, smem bytes}, out[5..9] the same for features.
void tma_stream_plan(int T, int d, int k, int with_epilogue, int* out)
{
    const int dpad = (d + 31) / 32 * 32;
    const int kp = (k + 7) / 8 * 8;
    for (int i = 0; i < 10; ++i) out[i] = 0;
    int TC, nstrips;
    if (samples_strip_plan(T, dpad, k, with_epilogue != 0, &TC, &nstrips)) {
        const size_t fixed = samples_fixed_smem(kp, with_epilogue != 0);
        const size_t stage = (size_t)kSampTR * (TC + 4) * 8;
        const int stages = ring_stages(fixed, stage);
        if (stages >= 2) {
            out[0] = 1; out[1] = TC; out[2] = nstrips; out[3] = stages;
            out[4] = (int)(fixed + stages * stage);
        }
    }
    if (features_strip_plan(T, dpad, k, &TC, &nstrips)) {
        const size_t fixed = 128 + (size_t)2 * 4 * 2 * 8 * kp * 8;
        const size_t stage = (size_t)kFsTR * (TC + 8) * 8;
        const int stages = ring_stages(fixed, stage);
        if (stages >= 2) {
            out[5] = 1; out[6] = TC; out[7] = nstrips; out[8] = stages;
            out[9] = (int)(fixed + stages * stage);
        }
    }
}

int run_reduce_features_tma(const double* M, long ldm, const double* X, long ldx, int T, int d,
                            int k, double* out, long ldo, void* workspace, size_t workspace_bytes,
                            const cdr_flags* flags, cudaStream_t stream, const StripGram* gram)
{
    int TC, nstrips;
    if (!features_strip_ok(M, ldm, X, ldx, T, d, k, &TC, &nstrips)) return CDR_TMA_NOT_APPLICABLE;
    const int dpad = (d + 31) / 32 * 32;
    StripGram g = {nullptr, nullptr, nullptr};
    if (gram != nullptr) g = *gram;
    if (k <= 8)
        return launch_features_strip<1>(M, ldm, X, ldx, T, dpad, k, TC, nstrips, out, ldo, workspace,
                                        workspace_bytes, flags, stream, g);
    return launch_features_strip<2>(M, ldm, X, ldx, T, dpad, k, TC, nstrips, out, ldo, workspace,
                                    workspace_bytes, flags, stream, g);
}

// Strip geometry of the reduce over features (for callers that consume the per-strip
// partials part[strip][t][KP] themselves): false when the strip kernel does not apply.
bool features_strip_geometry(int T, int d, int k, int* TC, int* nstrips)
{
    if (tma_disabled()) return false;
    return features_strip_plan(T, (d + 31) / 32 * 32, k, TC, nstrips);
}

}  // namespace cdr
