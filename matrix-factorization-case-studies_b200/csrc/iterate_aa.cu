// Whole outer iterations of archetypal analysis behind one C call
// (cdr_aa_prepare_enqueue / cdr_aa_iterate_enqueue).
//
// The reference's loop `_iterate_aa` (archetypal_analysis.py:534-670) is, per outer iteration:
// dictionary update by the generic spg() with Python callbacks (:324-341, spg.py:46-283; the
// drivers run it with max_iterations=1, bin/run_hadisst_aa.py:160-166), recomputes of C X,
// C X X', C X X' C' (:618-621), cost check, per-sample QPs (:636), recomputes of Z'Z, X'Z,
// X X'Z (:640-643), cost check.  With one inner SPG iteration, feature-space data and k <= 16
// at streaming shapes that is eight kernels here:
//
//   1. aa_head_kernel (one CTA per dictionary row, one grid barrier): x = P(C) (spg.py:148),
//      the linear trace term, the first step length 1 / max|P(x - g) - x| (spg.py:178-189),
//      d = P(x - alpha g) - x and the row partials of <d,g>, <d,d> (spg.py:191-206); the
//      gradient g = df(x) was left in G by kernel 8 of the previous iteration
//   2. reduce over samples   D X                        (stream_tma.cu)
//   3. reduce over features  (D X) X' as per-strip partials
//   4. aa_finalize_ls_kernel: D K = sum of the strip partials, and per 32-sample block the
//      k x k products CK C', CK D', DK D'; the last CTA sums them in fixed order, evaluates
//      f(x) (spg.py:156), runs the non-monotone Armijo search on scalars (spg.py:196-229: with
//      CK = C K maintained, f(x + lam d) is a quadratic in lam), forms C K C' of the accepted
//      point and applies the cost check after the dictionary update (:623-630)
//   5. aa_weights_fused_kernel: x += lam d, CK += lam DK for the sample's column, its QP
//      (:344-366), and the statistics z z', tr(C K Z); the last CTA sums them, applies the cost
//      check after the weights update and the stopping rule (:645-663) and starts the next
//      iteration
//   6.-8. (K Z)' = X (X' Z): reduce over samples, reduce over features, and
//      aa_kzt_gradient_kernel: sum of the strip partials (:641-642) plus the gradient of the
//      next dictionary step for the same columns (:293-299)
//
// Other configurations (more inner SPG iterations, k > 16, small or Gram-space problems) run
// the general kernel sequence of aa_steps.cu.  All reductions have a fixed order.
#include "fused_weights.cuh"
#include "peer.cuh"
#include "small_solve.cuh"

namespace cdr {

// slots of row_scratch (each k doubles); must match aa_steps.cu
enum { RS_A0 = 0, RS_ROWMAX = 1, RS_DELTA = 2, RS_DD = 3, RS_A1 = 4, RS_BETA = 5, RS_R2 = 6, RS_RINF = 7 };

constexpr int kAaRowMaxT = 26000;      // aa_steps.cu: rows are staged in shared memory
constexpr int kAaStagedMaxT = 6000;    // head kernel: four row buffers of T doubles (<= 188 KB)
constexpr int kAaStagedXMaxT = 13500;  // ... or two (the projected row only)

// gradient entry (j, t):  s_g * (sum_i a_j a_i ZtZ[j][i] CK[i][t] - a_j KZt[j][t]), k <= 16;
// coef[i] = a_j a_i ZtZ[j][i], 0 for i >= k (adding 0 * 0 leaves the sum -- formed in the order
// i = 0, 1, ... as in aa_steps.cu -- unchanged).  One expression for every kernel that forms the
// gradient, so they all round alike.
__device__ __forceinline__ double aa_grad_value(const double (&ck)[kFusedMaxK], const double* coef,
                                                double alpha_j, double kz, double grad_scale)
{
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < kFusedMaxK; ++i) s = fma(coef[i], ck[i], s);
    return grad_scale * (s - alpha_j * kz);
}

// The k loads of the column are issued together (a loop with a run-time trip count would
// serialise one L2 round trip per component).
__device__ __forceinline__ void aa_load_ck_column(const cdr_aa_buffers& b, int t, double (&ck)[kFusedMaxK])
{
#pragma unroll
    for (int i = 0; i < kFusedMaxK; ++i) ck[i] = (i < b.k) ? b.CK[(long)i * b.ldt + t] : 0.0;
}

__device__ __forceinline__ double aa_grad_entry16(const cdr_aa_buffers& b, const double* coef, int j, int t)
{
    double ck[kFusedMaxK];
    aa_load_ck_column(b, t, ck);
    return aa_grad_value(ck, coef, b.alpha[j], b.KZt[(long)j * b.ldt + t], b.grad_scale);
}

// coef[j][i] of all rows (see aa_grad_value) into shared memory: KP x kFusedMaxK doubles
__device__ __forceinline__ void aa_fill_coef(const cdr_aa_buffers& b, double* coef, int kp)
{
    for (int idx = threadIdx.x; idx < kp * kFusedMaxK; idx += blockDim.x) {
        const int j = idx / kFusedMaxK, i = idx % kFusedMaxK;
        coef[idx] = (j < b.k && i < b.k) ? b.alpha[j] * b.alpha[i] * b.ZtZ[j * b.k + i] : 0.0;
    }
}

#ifdef CDR_PROFILE_PHASES
// profiling build only (profiles/phase_profile.py)
__device__ unsigned long long cdr_head_ns[16 * 16];
#define CDR_HEAD_MARK(m)                                                                  \
    do {                                                                                  \
        if (threadIdx.x == 0) {                                                           \
            unsigned long long t__;                                                       \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__));                       \
            cdr_head_ns[blockIdx.x * 16 + (m)] = t__;                                     \
        }                                                                                 \
    } while (0)
#else
#define CDR_HEAD_MARK(m)
#endif

// ---------------------------------------------------------------------- kernel 1
// grad_ready: G already holds df(x) (aa_kzt_gradient_kernel / aa_gradient_columns_kernel wrote it with
// the (K Z)' it belongs to; the gradient uses the maintained C K, not the re-projected x).
// staged: 2 = the row's x, g and (K Z)' are kept in shared memory next to `work` (4 T
// doubles), 1 = x only (2 T doubles), 0 = long rows of a large sample-sharded fit: they are
// re-read from global memory.
__global__ void __launch_bounds__(1024)
aa_head_kernel(cdr_aa_buffers b, cdr_spg_params p, int grad_ready, int staged)
{
    cdr_loop_state* st = b.state;
    const int done = is_done(st) ? 1 : 0;
    CDR_HEAD_MARK(0);
    extern __shared__ double sm[];
    double* scratch = sm;
    double* coef = sm + 64;
    double* work = sm + 64 + CDR_MAX_COMPONENTS;
    const int j = blockIdx.x, k = b.k, T = b.T;
    double* crow = b.C + (long)j * b.ldt;
    double* grow = b.G + (long)j * b.ldt;
    double* drow = b.D + (long)j * b.ldt;
    const double* kzrow = b.KZt + (long)j * b.ldt;
    // generic pointers: shared-memory copies when staged, the global rows otherwise
    double* xs = staged ? work + T : crow;
    double* gs = (staged == 2) ? work + 2 * (long)T : grow;
    const double* kz = (staged == 2) ? work + 3 * (long)T : kzrow;

    // the loads below do not depend on `done` (issued together with its read); the three rows
    // of a staged launch go out together, eight elements per thread before the first store
    if (staged == 2) {
        double* xw = work;                              // known shared-memory pointers: the
        double* gw = work + 2 * (long)T;                // loads can be hoisted over the stores
        double* kw = work + 3 * (long)T;
        const double* __restrict__ cg = crow;
        const double* __restrict__ gg = grow;
        const double* __restrict__ kg = kzrow;
        for (int t0 = threadIdx.x; t0 < T; t0 += 8 * blockDim.x) {
            double c[8], g[8], z[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int t = t0 + u * blockDim.x;
                c[u] = (t < T) ? __ldcg(cg + t) : 0.0;
                z[u] = (t < T) ? __ldcg(kg + t) : 0.0;
                g[u] = (t < T && grad_ready) ? __ldcg(gg + t) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int t = t0 + u * blockDim.x;
                if (t < T) {
                    xw[t] = c[u];
                    kw[t] = z[u];
                    if (grad_ready) gw[t] = g[u];
                }
            }
        }
    } else {
        for (int t = threadIdx.x; t < T; t += blockDim.x) work[t] = crow[t];
    }
    if (!grad_ready)
        for (int i = threadIdx.x; i < kFusedMaxK; i += blockDim.x)
            coef[i] = (i < k) ? b.alpha[j] * b.alpha[i] * b.ZtZ[j * k + i] : 0.0;
    if (done) return;
    __syncthreads();
    CDR_HEAD_MARK(1);

    // x = project(x0) (spg.py:146-148) and the linear trace term a0 = a_j <x, (K Z)_j>
    double th = block_simplex_threshold(work, 1, T, scratch);
    CDR_HEAD_MARK(2);
    double a0[1] = {0.0};
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        const double x = fmax(work[t] - th, 0.0);
        crow[t] = x;
        if (staged) xs[t] = x;
        a0[0] = fma(x, kz[t], a0[0]);
    }
    block_sum<1>(a0, scratch);
    if (threadIdx.x == 0) b.row_scratch[RS_A0 * k + j] = b.alpha[j] * a0[0];
    CDR_HEAD_MARK(3);

    // g = df(x) (spg.py:176); work = x - g
    if (grad_ready) {
        // (each thread re-reads its own elements of xs / crow)
        for (int t = threadIdx.x; t < T; t += blockDim.x) work[t] = xs[t] - gs[t];
    } else {
        for (int t = threadIdx.x; t < T; t += 2 * blockDim.x) {
            const int t2 = t + blockDim.x;
            const double g = aa_grad_entry16(b, coef, j, t);
            const double g2 = (t2 < T) ? aa_grad_entry16(b, coef, j, t2) : 0.0;
            grow[t] = g;
            if (staged == 2) gs[t] = g;
            work[t] = xs[t] - g;
            if (t2 < T) {
                grow[t2] = g2;
                if (staged == 2) gs[t2] = g2;
                work[t2] = xs[t2] - g2;
            }
        }
    }
    const bool explicit_alpha = p.alpha0 > 0.0;        // spg.py:151; used unclamped when given
    if (!explicit_alpha) {
        __syncthreads();
        CDR_HEAD_MARK(4);
        th = block_simplex_threshold(work, 1, T, scratch);
        CDR_HEAD_MARK(5);
        double m = 0.0;
        for (int t = threadIdx.x; t < T; t += blockDim.x)
            m = fmax(m, fabs(fmax(work[t] - th, 0.0) - xs[t]));
        m = block_max(m, scratch);
        if (threadIdx.x == 0) b.row_scratch[RS_ROWMAX * k + j] = m;
    }

    // ---- barrier over the k row CTAs (all resident: k <= 16).  The counter only grows:
    // every launch adds exactly k, so the k arrivals of this launch see old values in
    // [n k, (n + 1) k) and wait for (n + 1) k.
    CDR_HEAD_MARK(6);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int old = atomicAdd(&st->tickets[2], 1u);
        const unsigned int target = (old / (unsigned int)k + 1u) * (unsigned int)k;
        while (*((volatile unsigned int*)&st->tickets[2]) < target) {
        }
        __threadfence();
    }
    __syncthreads();

    CDR_HEAD_MARK(7);
    // first step length (spg.py:178-189)
    double alpha;
    if (explicit_alpha) {
        alpha = p.alpha0;
    } else {
        double m = 0.0;
        for (int i = 0; i < k; ++i) m = fmax(m, __ldcg(b.row_scratch + RS_ROWMAX * k + i));
        alpha = (fabs(m) > 1e-12) ? 1.0 / m : 1.0;
    }
    if (j == 0 && threadIdx.x == 0) {
        st->alpha = alpha;
        st->spg_alpha_set = 1;
    }

    // d = P(x - alpha g) - x with <d,g>, <d,d> and the linear term along d (spg.py:191-206)
    for (int t = threadIdx.x; t < T; t += blockDim.x) work[t] = xs[t] - alpha * gs[t];
    __syncthreads();
    CDR_HEAD_MARK(8);
    th = block_simplex_threshold(work, 1, T, scratch);
    CDR_HEAD_MARK(9);
    double r[3] = {0.0, 0.0, 0.0};
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        const double d = fmax(work[t] - th, 0.0) - xs[t];
        drow[t] = d;
        r[0] = fma(d, gs[t], r[0]);
        r[1] = fma(d, d, r[1]);
        r[2] = fma(d, kz[t], r[2]);
    }
    block_sum<3>(r, scratch);
    if (threadIdx.x == 0) {
        b.row_scratch[RS_DELTA * k + j] = r[0];
        b.row_scratch[RS_DD * k + j] = r[1];
        b.row_scratch[RS_A1 * k + j] = b.alpha[j] * r[2];
    }
    CDR_HEAD_MARK(10);
}

#ifdef CDR_PROFILE_PHASES
}  // namespace cdr
extern "C" int cdr_debug_head_read(unsigned long long* out)
{
    return (int)cudaMemcpyFromSymbol(out, cdr::cdr_head_ns, sizeof(unsigned long long) * 16 * 16);
}
namespace cdr {
#endif

// ---------------------------------------------------------------------- kernel 4
constexpr int kFinTB = 32;             // samples per CTA
constexpr int kFinThreads = 1024;

// Sample-sharded fit (g.world > 1): the block's columns of D K are pushed to every rank (the
// k x T matrices are replicated), the k x k products are partial sums over this rank's
// samples and the last CTA exchanges them with the other ranks before the line search, which
// every rank then runs on identical numbers.
struct AaFinalizeArgs {
    cdr_aa_buffers b;        // b.T = all samples
    cdr_spg_params p;
    const double* part;      // [nstrips][Tl][KP] per-strip partials of (D X) X' for the local samples
    int nstrips;
    double* cta_part;
    int Tl, row0;            // local samples: columns row0 .. row0 + Tl of the k x T matrices
    cdr_peer_group g;
    size_t dk_offset;        // of b.DK in the symmetric region (world > 1)
};

template <int KT>
__global__ void __launch_bounds__(kFinThreads) aa_finalize_ls_kernel(AaFinalizeArgs a)
{
    const cdr_aa_buffers& b = a.b;
    const cdr_spg_params& p = a.p;
    const double* __restrict__ part = a.part;
    const int nstrips = a.nstrips;
    double* cta_part = a.cta_part;
    cdr_loop_state* st = b.state;
    if (is_done(st)) return;
    constexpr int KP = 8 * KT;
    constexpr int PAIRS = KP / 2;
    constexpr int ITEMS = kFinTB * PAIRS;              // 128 | 256
    constexpr int GROUPS = kFinThreads / ITEMS;        // 8 | 4
    constexpr int NOUT = 3 * KP * KP;                  // 192 | 768
    constexpr int PH = (kFinThreads / NOUT) > 0 ? (kFinThreads / NOUT) : 1;   // 5 | 1
    __shared__ double2 red[GROUPS][ITEMS];             // 16 KB; re-used by the tail
    __shared__ double tiles[4][KP][kFinTB + 1];        // C, D, CK, DK
    __shared__ int is_last;
    const int k = b.k, T = a.Tl;
    const long ldt = b.ldt;
    const int t0 = blockIdx.x * kFinTB;
    const bool sharded = a.g.world > 1;

    // the block's columns of C, D and C K: requested first, so that the loads overlap the
    // partial sums below (one L2 round trip less on the kernel's critical path)
    constexpr int kTileLoads = (3 * KP * kFinTB + kFinThreads - 1) / kFinThreads;    // 1 | 2
    double tile_v[kTileLoads];
#pragma unroll
    for (int u = 0; u < kTileLoads; ++u) {
        const int idx = threadIdx.x + u * kFinThreads;
        const int which = idx / (KP * kFinTB), rem = idx % (KP * kFinTB);
        const int i = rem / kFinTB, t = t0 + rem % kFinTB;
        const double* src = (which == 0) ? b.C : (which == 1) ? b.D : b.CK;
        tile_v[u] = (idx < 3 * KP * kFinTB && i < k && t < T) ? src[(long)i * ldt + a.row0 + t] : 0.0;
    }

    // ---- D K for this block of samples: sum of the per-strip partials, fixed order
    {
        const int item = threadIdx.x % ITEMS, grp = threadIdx.x / ITEMS;
        const int tl = item / PAIRS, pr = item % PAIRS;
        const int t = t0 + tl;
        double s0 = 0.0, s1 = 0.0;
        if (t < T) {
            const double2* src = reinterpret_cast<const double2*>(part) + ((long)t * PAIRS + pr);
            const long stride = (long)T * PAIRS;
            // all of a thread's strips in flight at once (<= 24 of them: nstrips <= 148 + ...)
            for (int s = grp; s < nstrips; s += 24 * GROUPS) {
                double2 v[24];
#pragma unroll
                for (int q = 0; q < 24; ++q) {
                    const int sq = s + q * GROUPS;
                    v[q] = (sq < nstrips) ? __ldcg(src + (long)sq * stride) : make_double2(0.0, 0.0);
                }
#pragma unroll
                for (int q = 0; q < 24; ++q) {
                    s0 += v[q].x;
                    s1 += v[q].y;
                }
            }
        }
        red[grp][item] = make_double2(s0, s1);
        __syncthreads();
        if (grp == 0) {
            double a0 = 0.0, a1 = 0.0;
#pragma unroll
            for (int q = 0; q < GROUPS; ++q) {
                a0 += red[q][item].x;
                a1 += red[q][item].y;
            }
            const int j0 = 2 * pr;
            const bool ok = t < T;
            const long col = a.row0 + t;
            if (!sharded) {
                if (ok && j0 < k) b.DK[(long)j0 * ldt + col] = a0;
                if (ok && j0 + 1 < k) b.DK[(long)(j0 + 1) * ldt + col] = a1;
            } else {
                for (int r = 0; r < a.g.world; ++r) {
                    double* dk = peer::peer_ptr<double>(a.g, r, a.dk_offset);
                    if (ok && j0 < k) dk[(long)j0 * ldt + col] = a0;
                    if (ok && j0 + 1 < k) dk[(long)(j0 + 1) * ldt + col] = a1;
                }
            }
            tiles[3][j0][tl] = (ok && j0 < k) ? a0 : 0.0;
            tiles[3][j0 + 1][tl] = (ok && j0 + 1 < k) ? a1 : 0.0;
        }
    }
#pragma unroll
    for (int u = 0; u < kTileLoads; ++u) {
        const int idx = threadIdx.x + u * kFinThreads;
        if (idx < 3 * KP * kFinTB) {
            const int which = idx / (KP * kFinTB), rem = idx % (KP * kFinTB);
            tiles[which][rem / kFinTB][rem % kFinTB] = tile_v[u];
        }
    }
    __syncthreads();

    // ---- k x k products over the block: CK C' (fresh C K C'), CK D', DK D'
    for (int o = threadIdx.x; o < NOUT; o += blockDim.x) {
        const int which = o / (KP * KP), i = (o % (KP * KP)) / KP, j = o % KP;
        const double* left = (which == 2) ? tiles[3][i] : tiles[2][i];
        const double* right = (which == 0) ? tiles[0][j] : tiles[1][j];
        double s = 0.0;
#pragma unroll
        for (int tl = 0; tl < kFinTB; ++tl) s = fma(left[tl], right[tl], s);
        __stcg(cta_part + (long)blockIdx.x * NOUT + o, s);
    }
    if (sharded) __threadfence_system();          // the pushed columns of D K
    else __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(&st->tickets[3], 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!is_last) return;
    __threadfence();

    // ------------------------------------------------------------------ last CTA
    double* phs = reinterpret_cast<double*>(&red[0][0]);     // [PH][NOUT], then fin[NOUT]
    double* fin = phs + PH * NOUT;
    static_assert((PH + 1) * NOUT <= GROUPS * ITEMS * 2, "tail scratch does not fit");
    {
        const int nblk = gridDim.x;
        const int e = (int)threadIdx.x % NOUT, ph = (int)threadIdx.x / NOUT;
        if (ph < PH) {
            const int n_mine = (nblk - ph + PH - 1) / PH;
            double s = 0.0;
            const double* src = cta_part + (long)ph * NOUT + e;
            for (int q0 = 0; q0 < n_mine; q0 += 16) {           // 16 loads in flight, same order
                double v[16];
#pragma unroll
                for (int q = 0; q < 16; ++q)
                    v[q] = (q0 + q < n_mine) ? __ldcg(src + (long)(q0 + q) * PH * NOUT) : 0.0;
#pragma unroll
                for (int q = 0; q < 16; ++q) s += v[q];
            }
            phs[ph * NOUT + e] = s;
        }
        __syncthreads();
        for (int e2 = threadIdx.x; e2 < NOUT; e2 += blockDim.x) {
            double s = phs[e2];
#pragma unroll
            for (int q = 1; q < PH; ++q) s += phs[q * NOUT + e2];
            fin[e2] = s;
        }
        __syncthreads();
    }
    // sample-sharded fit: sum over ranks (also the barrier after which every rank's columns
    // of D K have arrived here)
    if (sharded) peer::cta_allreduce_small(a.g, fin, NOUT);
    const double* G00 = fin;
    const double* G01 = fin + KP * KP;
    const double* G11 = fin + 2 * KP * KP;
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        // quadratic form coefficients  q(lam) = q0 + lam q1 + lam^2 q2  of tr(D Z'Z D . C K C')
        double q0 = 0.0, q1 = 0.0, q2 = 0.0;
        for (int idx = lane; idx < k * k; idx += 32) {
            const int i = idx / k, j = idx % k;
            const double w = b.alpha[i] * b.alpha[j] * b.ZtZ[i * k + j];
            q0 += w * G00[j * KP + i];
            q1 += w * (G01[j * KP + i] + G01[i * KP + j]);
            q2 += w * G11[j * KP + i];
        }
        q0 = warp_sum(q0);
        q1 = warp_sum(q1);
        q2 = warp_sum(q2);
        double lam = 1.0;
        // row partials of the head kernel: one row per lane, summed in row order
        const double ra0 = (lane < k) ? b.row_scratch[RS_A0 * k + lane] : 0.0;
        const double rdl = (lane < k) ? b.row_scratch[RS_DELTA * k + lane] : 0.0;
        const double rdd = (lane < k) ? b.row_scratch[RS_DD * k + lane] : 0.0;
        const double ra1 = (lane < k) ? b.row_scratch[RS_A1 * k + lane] : 0.0;
        double a0 = 0.0, delta = 0.0, dd = 0.0, a1 = 0.0;
        for (int j = 0; j < k; ++j) {
            a0 += __shfl_sync(CDR_FULL_MASK, ra0, j);
            delta += __shfl_sync(CDR_FULL_MASK, rdl, j);
            dd += __shfl_sync(CDR_FULL_MASK, rdd, j);
            a1 += __shfl_sync(CDR_FULL_MASK, ra1, j);
        }
        if (lane == 0) {
            const double tr = st->trace_data, sf = b.cost_scale;
            // f(x) at the start of spg() (spg.py:153-157); the memory starts as zeros
            const double f_old = 0.5 * (tr - 2.0 * a0 + q0) * sf;
            double f_max = f_old;
            for (int i = 1; i < p.memory; ++i) f_max = fmax(f_max, 0.0);
            // non-monotone Armijo search (spg.py:196-229)
            double f_new = 0.5 * (tr - 2.0 * (a0 + lam * a1) + (q0 + lam * q1 + lam * lam * q2)) * sf;
            int feval = 2;
            while (f_new > f_max + p.gamma * lam * delta) {
                lam = spg_step_length(lam, delta, f_old, f_new, p.sigma_one, p.sigma_two);
                f_new = 0.5 * (tr - 2.0 * (a0 + lam * a1) + (q0 + lam * q1 + lam * lam * q2)) * sf;
                feval += 1;
                if (fabs(lam) < p.lambda_min) {
                    st->spg_warnings |= 1;
                    break;
                }
            }
            // one inner iteration: the iteration limit is reached without a convergence test
            st->spg_warnings |= 4;
            st->lam = lam;
            st->f_old = f_old;
            st->f_new = f_new;
            st->delta = delta;
            st->dd = dd;
            st->a1 = a1;
            st->a0 = a0 + lam * a1;
            st->spg_iter = 1;
            st->spg_feval = feval + 1;
            // cost after the dictionary update (archetypal_analysis.py:623-630)
            const double cost = 0.5 * (tr - 2.0 * (a0 + lam * a1) + (q0 + lam * q1 + lam * lam * q2)) /
                                (double)b.T;
            finish_sub_step(st, b.cost_deltas, cost, 2, 0);
            st->tickets[3] = 0u;
        }
        lam = __shfl_sync(CDR_FULL_MASK, lam, 0);
        // C K C' of the accepted point
        for (int idx = lane; idx < k * k; idx += 32) {
            const int i = idx / k, j = idx % k;
            b.CKCt[idx] = G00[i * KP + j] + lam * (G01[i * KP + j] + G01[j * KP + i]) +
                          lam * lam * G11[i * KP + j];
        }
    }
}

// ---------------------------------------------------------------------- kernel 5
struct AaWeightsArgs {
    cdr_aa_buffers b;        // b.T = all samples
    double* Z;               // Tl x k: the local samples
    double* cta_part;
    int spw;
    int Tl, row0;
    cdr_spg_params p;
    cdr_peer_group g;        // world == 1: single GPU
};

template <int KPL>
__global__ void __launch_bounds__(kFusedThreads) aa_weights_fused_kernel(AaWeightsArgs a)
{
    const cdr_aa_buffers& b = a.b;
    cdr_loop_state* st = b.state;
    if (is_done(st)) return;
    constexpr int KP = 8 * KPL;
    constexpr int NST = KP * KP + 2;
    extern __shared__ double fsm[];
    double* As = fsm;                               // KP x KP (KPL > 1)
    double* wsum = fsm + (KPL > 1 ? KP * KP : 0);   // [kFusedWarps][NST]
    double* fin = wsum + kFusedWarps * NST;         // 4 * NST

    const int k = b.k, T = a.Tl;
    const long ldt = b.ldt;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane & 7, q = lane >> 3;
    const double lam = st->lam;

    // sample-sharded fit: the dictionary and C K are replicated, so the columns of the other
    // ranks' samples are updated here as well (their D K has arrived: barrier in kernel 4)
    if (a.g.world > 1) {
        const long total = (long)k * b.T;
        for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
             idx += (long)gridDim.x * blockDim.x) {
            const int i = (int)(idx / b.T), col = (int)(idx % b.T);
            if (col < a.row0 || col >= a.row0 + T) {
                const long e = (long)i * ldt + col;
                b.C[e] = fma(lam, b.D[e], b.C[e]);
                b.CK[e] = fma(lam, b.DK[e], b.CK[e]);
            }
        }
    }

    // A' = D (C K C') D  (archetypal_analysis.py:384-385)
    if constexpr (KPL > 1) {
        for (int idx = threadIdx.x; idx < KP * KP; idx += blockDim.x) {
            const int j = idx / KP, c = idx % KP;
            As[idx] = (j < k && c < k) ? b.CKCt[(long)c * k + j] * b.alpha[c] * b.alpha[j] : 0.0;
        }
        __syncthreads();
    }
    double arow[8];
    if constexpr (KPL == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            arow[j] = (g < k && j < k) ? b.CKCt[(long)g * k + j] * b.alpha[g] * b.alpha[j] : 0.0;
    }

    const int spw = a.spw;
    const int t_raw = (blockIdx.x * kFusedWarps + warp) * spw + (q % spw);
    const bool has_sample = t_raw < T;
    const bool valid = (q < spw) && has_sample;
    const long t = has_sample ? t_raw : (T - 1);
    const long col = a.row0 + t;

    // x <- x + lam d, C K <- C K + lam D K for this sample's column (spg.py:219; linearity
    // of C -> C K), then the linear term b = -D (C K)[:, t]
    double z0[KPL], x[KPL], bl[KPL], ck[KPL], cn[KPL], ckn[KPL];
    bool present[KPL];
#pragma unroll
    for (int r = 0; r < KPL; ++r) {
        const int c = g * KPL + r;
        present[r] = c < k;
        if (present[r]) {
            const long idx = (long)c * ldt + col;
            cn[r] = fma(lam, b.D[idx], b.C[idx]);
            ckn[r] = fma(lam, b.DK[idx], b.CK[idx]);
            ck[r] = b.alpha[c] * ckn[r];
            bl[r] = -ck[r];
            z0[r] = a.Z[t * k + c];
        } else {
            cn[r] = ckn[r] = ck[r] = 0.0;
            bl[r] = 0.0;
            z0[r] = -INFINITY;
        }
    }
    // the replicas of a sample (lane groups >= spw) have read the same column: store only
    // after every lane of the warp has loaded
    __syncwarp();
    if (valid) {
#pragma unroll
        for (int r = 0; r < KPL; ++r)
            if (present[r]) {
                const long idx = (long)(g * KPL + r) * ldt + col;
                b.C[idx] = cn[r];
                b.CK[idx] = ckn[r];
            }
    }

    int n_iter = 0, n_feval = 0;
    qp_solve<KPL>(As, arow, z0, bl, present, a.p, has_sample, g, spw, x, n_iter, n_feval);

    double tr_new = 0.0;
    if (valid) {
#pragma unroll
        for (int r = 0; r < KPL; ++r)
            if (present[r]) {
                a.Z[t * k + g * KPL + r] = x[r];
                tr_new = fma(ck[r], x[r], tr_new);
            }
    }
    tr_new = group8_sum(tr_new);

    if (!fused_sample_statistics<KPL>(x, present, valid, k, 0.0, tr_new, wsum, a.cta_part,
                                      &st->tickets[1]))
        return;
    fused_final_sum<KPL>(a.cta_part, fin);
    if (a.g.world > 1) peer::cta_allreduce_small(a.g, fin, NST);
    // fin[i * KP + j] = (Z'Z)[i][j]; fin[KP * KP + 1] = sum_i a_i (C K Z)[i][i]
    if (warp == 0) {
        double t2 = 0.0;
        for (int idx = lane; idx < k * k; idx += 32) {
            const int i = idx / k, j = idx % k;
            t2 += b.alpha[i] * b.alpha[j] * fin[i * KP + j] * b.CKCt[j * k + i];
        }
        t2 = warp_sum(t2);
        if (lane == 0) {
            // cost after the weights update (archetypal_analysis.py:645-652), stopping rule
            const double cost = 0.5 * (st->trace_data - 2.0 * fin[KP * KP + 1] + t2) / (double)b.T;
            finish_sub_step(st, b.cost_deltas, cost, 3, 1);
            if (!st->done) st->old_cost = st->cost;    // start of the next iteration
            st->tickets[1] = 0u;
        }
    }
    for (int idx = threadIdx.x; idx < k * k; idx += blockDim.x)
        b.ZtZ[idx] = fin[(idx / k) * KP + idx % k];
}

// ---------------------------------------------------------------------- kernel 8
// (K Z)'[j][t] = sum over strips of part[strip][t][j] -- the summation order of
// reduce_features_strip_finalize_kernel -- and, while the column is at hand, the gradient of
// the next dictionary step G[j][t] = df(x)[j][t] (archetypal_analysis.py:293-299): it needs
// only this column of (K Z)' and of the maintained C K, the new Z'Z and the scale factors, all
// final at this point, and saves the head kernel of the next iteration its (k + 1) k T loads.
template <int KT>
__global__ void __launch_bounds__(256)
aa_kzt_gradient_kernel(const double* __restrict__ part, int nstrips, cdr_aa_buffers b)
{
    if (is_done(b.state)) return;
    constexpr int KP = 8 * KT;
    constexpr int PAIRS = KP / 2;
    __shared__ double2 red[8][32];
    __shared__ double coef[KP * kFusedMaxK];
    const int k = b.k, T = b.T;
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const long item = (long)blockIdx.x * 32 + lane;
    const long nitems = (long)T * PAIRS;
    const long stride = nitems;
    const int per = (nstrips + 7) / 8;
    aa_fill_coef(b, coef, KP);
    double s0 = 0.0, s1 = 0.0;
    if (item < nitems) {
        const double2* src = reinterpret_cast<const double2*>(part) + item;
        const int hi = min(nstrips, (grp + 1) * per);
        // 24 strips in flight (all of a group's strips at the usual widths), summed in order
        for (int s_lo = grp * per; s_lo < hi; s_lo += 24) {
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
    red[grp][lane] = make_double2(s0, s1);
    __syncthreads();
    if (grp == 0 && item < nitems) {
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            a0 += red[q][lane].x;
            a1 += red[q][lane].y;
        }
        const int t = (int)(item / PAIRS), j = 2 * (int)(item % PAIRS);
        double ck[kFusedMaxK];
        aa_load_ck_column(b, t, ck);
        if (j < k) {
            b.KZt[(long)j * b.ldt + t] = a0;
            b.G[(long)j * b.ldt + t] = aa_grad_value(ck, coef + j * kFusedMaxK, b.alpha[j], a0, b.grad_scale);
        }
        if (j + 1 < k) {
            b.KZt[(long)(j + 1) * b.ldt + t] = a1;
            b.G[(long)(j + 1) * b.ldt + t] =
                aa_grad_value(ck, coef + (j + 1) * kFusedMaxK, b.alpha[j + 1], a1, b.grad_scale);
        }
    }
}

// The gradient alone, from (K Z)' in memory (after cdr_aa_prepare_enqueue; before the head kernel
// of a sample-sharded iteration): one sample per thread.
__global__ void __launch_bounds__(256) aa_gradient_columns_kernel(cdr_aa_buffers b)
{
    if (is_done(b.state)) return;
    __shared__ double coef[kFusedMaxK * kFusedMaxK];
    aa_fill_coef(b, coef, kFusedMaxK);
    __syncthreads();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= b.T) return;
    double ck[kFusedMaxK];
    aa_load_ck_column(b, t, ck);
    for (int j = 0; j < b.k; ++j)
        b.G[(long)j * b.ldt + t] = aa_grad_value(ck, coef + j * kFusedMaxK, b.alpha[j],
                                                 b.KZt[(long)j * b.ldt + t], b.grad_scale);
}

// ---------------------------------------------------------------------- kernel 8 (sharded)
// out[j][col0 + t] = sum over strips of part[strip][t][j] for the local samples t, written to
// EVERY rank (the k x T matrix is replicated); the last CTA then meets the other ranks, after
// which all columns of `out` are complete on this rank.  Same summation order as
// reduce_features_strip_finalize_kernel.
template <int KT>
__global__ void __launch_bounds__(256)
features_finalize_push_kernel(const double* __restrict__ part, int Tl, int nstrips, int k,
                              size_t out_offset, long ldo, int col0, cdr_peer_group g,
                              cdr_loop_state* st)
{
    if (is_done(st)) return;
    constexpr int KP = 8 * KT;
    constexpr int PAIRS = KP / 2;
    __shared__ double2 red[8][32];
    __shared__ int is_last;
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const long item = (long)blockIdx.x * 32 + lane;
    const long nitems = (long)Tl * PAIRS;
    const long stride = nitems;
    const int per = (nstrips + 7) / 8;
    double s0 = 0.0, s1 = 0.0;
    if (item < nitems) {
        const double2* src = reinterpret_cast<const double2*>(part) + item;
        const int hi = min(nstrips, (grp + 1) * per);
        // 24 strips in flight (all of a group's strips at the usual widths), summed in order
        for (int s_lo = grp * per; s_lo < hi; s_lo += 24) {
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
    red[grp][lane] = make_double2(s0, s1);
    __syncthreads();
    if (grp == 0 && item < nitems) {
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            a0 += red[q][lane].x;
            a1 += red[q][lane].y;
        }
        const int t = (int)(item / PAIRS), j = 2 * (int)(item % PAIRS);
        for (int r = 0; r < g.world; ++r) {
            double* out = peer::peer_ptr<double>(g, r, out_offset);
            if (j < k) out[(long)j * ldo + col0 + t] = a0;
            if (j + 1 < k) out[(long)(j + 1) * ldo + col0 + t] = a1;
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(&st->tickets[0], 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    peer::cta_allreduce_small(g, nullptr, 0);
    if (threadIdx.x == 0) st->tickets[0] = 0u;
}

static size_t aa_fused_smem_bytes(int kp)
{
    const int nst = kp * kp + 2;
    return ((kp > 8 ? (size_t)kp * kp : 0) + (size_t)kFusedWarps * nst + fused_fin_doubles(nst)) * sizeof(double);
}

static int aa_row_threads(int T)
{
    if (T <= 2048) return 256;
    if (T <= 8192) return 512;
    return 1024;
}

// ---------------------------------------------------------------------- workspace layout
struct AaWorkspace {
    double* stream;
    size_t stream_bytes;
    double* gram;
    size_t gram_bytes;
    double* fin_part;     // [ceil(T / 32)][3 * KP * KP]
    double* cta_part;     // [blocks][KP * KP + 2]
    size_t total;
};

static AaWorkspace carve_aa(void* base, int T, int d, int k)
{
    AaWorkspace w;
    const size_t s1 = cdr_reduce_samples_workspace_bytes(T, d, k);
    const size_t s2 = cdr_reduce_features_workspace_bytes(T, d, k);
    w.stream_bytes = align256((s1 > s2 ? s1 : s2) + 8);
    w.gram_bytes = align256(cdr_small_gram_workspace_bytes());
    const int kp = (k <= 8) ? 8 : 16;
    int spw = 1, blocks = 1;
    fused_grid(T, &spw, &blocks);
    const size_t fp = align256((size_t)((T + kFinTB - 1) / kFinTB) * 3 * kp * kp * sizeof(double));
    const size_t cp = align256((size_t)blocks * (kp * kp + 2) * sizeof(double));
    unsigned char* p = static_cast<unsigned char*>(base);
    w.stream = reinterpret_cast<double*>(p);
    w.gram = reinterpret_cast<double*>(p + w.stream_bytes);
    w.fin_part = reinterpret_cast<double*>(p + w.stream_bytes + w.gram_bytes);
    w.cta_part = reinterpret_cast<double*>(p + w.stream_bytes + w.gram_bytes + fp);
    w.total = w.stream_bytes + w.gram_bytes + fp + cp;
    return w;
}

static bool aa_fused_shape(int T, int d, int k, int dict_max_iterations, int* nstrips)
{
    if (k > kFusedMaxK || dict_max_iterations != 1 || T > kAaRowMaxT) return false;
    const char* e = getenv("CDR_DISABLE_FUSED");
    if (e != nullptr && e[0] == '1') return false;
    int out[12];
    tma_stream_plan(T, d, k, 0, out);
    if (!out[0] || !out[5]) return false;              // both passes on the strip kernels
    int TC;
    return features_strip_geometry(T, d, k, &TC, nstrips);
}

static bool aa_sharded(const cdr_aa_problem* p) { return p->peers != nullptr && p->peers->world > 1; }

static int check_aa(const cdr_aa_problem* p)
{
    CDR_CHECK_ARG(p != nullptr && p->X != nullptr && p->Z != nullptr && p->tmp_kd != nullptr);
    const cdr_aa_buffers& b = p->buf;
    CDR_CHECK_ARG(b.k >= 1 && b.T >= 1 && p->T >= 1 && p->d >= 1 && b.ldt >= b.T);
    if (aa_sharded(p)) {
        CDR_CHECK_ARG(p->row0 >= 0 && p->row0 + p->T <= b.T && p->T_min >= 1 && p->T_min <= p->T);
    } else {
        CDR_CHECK_ARG(p->T == b.T && p->row0 == 0);
    }
    CDR_CHECK_ARG(b.C && b.G && b.D && b.CK && b.DK && b.KZt && b.alpha && b.ZtZ && b.CKCt &&
                  b.CKZ && b.G01 && b.G11 && b.row_scratch && b.state);
    if (b.k > CDR_MAX_COMPONENTS || b.T > kAaRowMaxT) return CDR_ERR_UNSUPPORTED;
    if (p->dictionary_params.max_iterations < 1 || p->dictionary_params.max_iterations > 8)
        return CDR_ERR_UNSUPPORTED;                    // longer inner loops need host control
    if (p->workspace == nullptr || p->workspace_bytes < cdr_aa_workspace_bytes(p->T, p->d, b.k))
        return CDR_ERR_WORKSPACE;
    return 0;
}

// out (k x ldt) = (L X) X' for a k x T matrix L (row stride ldt)
static int aa_apply_left(const cdr_aa_problem* p, const AaWorkspace& w, const double* L, double* out,
                         cudaStream_t s)
{
    const cdr_aa_buffers& b = p->buf;
    CDR_TRY(cdr_reduce_samples(L, b.ldt, 1, p->X, p->ldx, p->T, p->d, b.k, nullptr, p->tmp_kd,
                               p->ldx, w.stream, w.stream_bytes, b.state, s));
    return cdr_reduce_features(p->tmp_kd, p->ldx, p->X, p->ldx, p->T, p->d, b.k, out, b.ldt,
                               w.stream, w.stream_bytes, b.state, s);
}

// KZt = (X (X' Z))'
static int aa_apply_right(const cdr_aa_problem* p, const AaWorkspace& w, cudaStream_t s)
{
    const cdr_aa_buffers& b = p->buf;
    CDR_TRY(cdr_reduce_samples(p->Z, 1, b.k, p->X, p->ldx, p->T, p->d, b.k, nullptr, p->tmp_kd,
                               p->ldx, w.stream, w.stream_bytes, b.state, s));
    return cdr_reduce_features(p->tmp_kd, p->ldx, p->X, p->ldx, p->T, p->d, b.k, b.KZt, b.ldt,
                               w.stream, w.stream_bytes, b.state, s);
}

static cdr_small_gram_desc kt_desc(const cdr_aa_buffers& b, const double* A, const double* B, double* out)
{
    return gram_desc(A, b.ldt, 1, b.k, B, b.ldt, 1, b.k, b.T, out, 0);
}

}  // namespace cdr

using namespace cdr;

extern "C" size_t cdr_aa_workspace_bytes(int T, int d, int k)
{
    if (T < 1 || d < 1 || k < 1 || k > CDR_MAX_COMPONENTS) return 0;
    return carve_aa(nullptr, T, d, k).total;
}

extern "C" int cdr_aa_fused_applicable(int T, int d, int k, int dictionary_max_iterations)
{
    int nstrips;
    return aa_fused_shape(T, d, k, dictionary_max_iterations, &nstrips) ? 1 : 0;
}

extern "C" int cdr_aa_prepare_enqueue(const cdr_aa_problem* p, cdr_stream_t stream)
{
    CDR_TRY(check_aa(p));
    // sample-sharded fits: the caller forms the initial products with its own collectives
    if (aa_sharded(p)) return CDR_ERR_UNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    const cdr_aa_buffers& b = p->buf;
    const int k = b.k;
    const AaWorkspace w = carve_aa(p->workspace, p->T, p->d, k);
    // archetypal_analysis.py:541-556
    CDR_TRY(aa_apply_left(p, w, b.C, b.CK, s));
    CDR_TRY(aa_apply_right(p, w, s));
    cdr_small_gram_desc ds[3];
    ds[0] = gram_desc(p->Z, 1, k, k, p->Z, 1, k, k, p->T, b.ZtZ, 0);
    ds[1] = kt_desc(b, b.CK, b.C, b.CKCt);
    ds[2] = kt_desc(b, b.C, b.KZt, b.CKZ);
    CDR_TRY(cdr_small_gram(ds, 3, w.gram, w.gram_bytes, b.state, s));
    CDR_TRY(cdr_aa_cost_check(&b, 0, 0, s));
    // spg() starts from project(x0) (spg.py:146-148).  A custom start may only be feasible to
    // np.isclose accuracy, so C K is rebuilt once for the projected iterate; later iterates
    // are feasible by construction.
    CDR_TRY(cdr_simplex_project_rows(b.C, b.C, k, b.T, b.ldt, b.ldt, b.state, s));
    CDR_TRY(aa_apply_left(p, w, b.C, b.CK, s));
    ds[0] = kt_desc(b, b.CK, b.C, b.CKCt);
    CDR_TRY(cdr_small_gram(ds, 1, w.gram, w.gram_bytes, b.state, s));
    // df(x) of the first dictionary step (the eight-kernel iteration expects it in G)
    if (k <= kFusedMaxK) {
        aa_gradient_columns_kernel<<<(b.T + 255) / 256, 256, 0, s>>>(b);
        CDR_RETURN_IF_LAUNCH_FAILED();
    }
    return cdr_loop_begin(b.state, s);
}

extern "C" int cdr_aa_iterate_enqueue(const cdr_aa_problem* p, cdr_stream_t stream)
{
    CDR_TRY(check_aa(p));
    cudaStream_t s = (cudaStream_t)stream;
    const cdr_aa_buffers& b = p->buf;
    const int k = b.k, T = p->T, d = p->d;
    const AaWorkspace w = carve_aa(p->workspace, T, d, k);
    const cdr_spg_params& dp = p->dictionary_params;
    if (dp.memory < 1 || dp.memory > CDR_MAX_MEMORY || p->weights_params.memory < 1 ||
        p->weights_params.memory > CDR_MAX_MEMORY)
        return CDR_ERR_UNSUPPORTED;
    int nstrips = 0;
    const bool sharded = aa_sharded(p);
    const bool fused = aa_fused_shape(sharded ? p->T_min : T, d, k, dp.max_iterations, &nstrips) &&
                       b.T <= kAaRowMaxT;
    // over peer memory only the eight-kernel path is offered
    if (sharded && !fused) return CDR_ERR_UNSUPPORTED;
    cdr_peer_group grp = cdr_peer_group();
    grp.world = 1;
    if (sharded) grp = *p->peers;

    if (!fused) {
        // general sequence (archetypal_analysis.py:586-663 with the pieces of aa_steps.cu)
        CDR_TRY(cdr_loop_begin(b.state, s));
        CDR_TRY(cdr_aa_spg_begin(&b, &dp, s));
        for (int n = 0; n < dp.max_iterations; ++n) {
            CDR_TRY(cdr_aa_spg_direction(&b, &dp, s));
            CDR_TRY(aa_apply_left(p, w, b.D, b.DK, s));
            cdr_small_gram_desc dg[2] = {kt_desc(b, b.CK, b.D, b.G01), kt_desc(b, b.DK, b.D, b.G11)};
            CDR_TRY(cdr_small_gram(dg, 2, w.gram, w.gram_bytes, b.state, s));
            CDR_TRY(cdr_aa_spg_linesearch(&b, &dp, s));
            if (n + 1 < dp.max_iterations) CDR_TRY(cdr_aa_spg_update(&b, &dp, 1, s));
        }
        cdr_small_gram_desc d1[2] = {kt_desc(b, b.CK, b.C, b.CKCt), kt_desc(b, b.C, b.KZt, b.CKZ)};
        CDR_TRY(cdr_small_gram(d1, 2, w.gram, w.gram_bytes, b.state, s));
        CDR_TRY(cdr_aa_cost_check(&b, 2, 0, s));
        CDR_TRY(cdr_quad_simplex_spg_batched(b.CKCt, b.alpha, b.CK, 1, b.ldt, p->Z, T, k,
                                             &p->weights_params, nullptr, nullptr, b.state, s));
        CDR_TRY(aa_apply_right(p, w, s));
        cdr_small_gram_desc d2[2] = {gram_desc(p->Z, 1, k, k, p->Z, 1, k, k, T, b.ZtZ, 0),
                                     kt_desc(b, b.C, b.KZt, b.CKZ)};
        CDR_TRY(cdr_small_gram(d2, 2, w.gram, w.gram_bytes, b.state, s));
        return cdr_aa_cost_check(&b, 3, 1, s);
    }

    // 1. head: projection, gradient, first step length, direction
    {
        // the rows of the (replicated) dictionary span all samples; short rows keep x, g and
        // (K Z)' in shared memory as well.  On one GPU the gradient was left in G by
        // cdr_aa_prepare_enqueue / the previous iteration's last kernel.
        const int staged = (b.T <= kAaStagedMaxT) ? 2 : (b.T <= kAaStagedXMaxT) ? 1 : 0;
        const size_t smem =
            (64 + CDR_MAX_COMPONENTS + (size_t)(staged == 2 ? 4 : staged == 1 ? 2 : 1) * b.T) * sizeof(double);
        CDR_TRY(ensure_dyn_smem<aa_head_kernel>(smem));
        if (sharded) {
            // the replicated (K Z)' was completed by the exchange that ended the previous
            // iteration: the gradient of all T_total columns as one wide launch (inside the
            // head kernel it is k CTAs x (k + 1) T_total loads: 35 us at 8 x 1620 samples)
            aa_gradient_columns_kernel<<<(b.T + 255) / 256, 256, 0, s>>>(b);
            CDR_RETURN_IF_LAUNCH_FAILED();
        }
        aa_head_kernel<<<k, aa_row_threads(b.T), smem, s>>>(b, dp, 1, staged);
        CDR_RETURN_IF_LAUNCH_FAILED();
    }
    // 2. D X (summed over ranks in the kernel's epilogue when sharded)
    // 3. (D X) X' as per-strip partials
    if (sharded) {
        CDR_TRY(cdr_reduce_samples_allreduce(p->peers, b.D + p->row0, b.ldt, 1, p->X, p->ldx, T,
                                             p->T_min, d, k, nullptr,
                                             peer::region_offset(grp, p->tmp_kd), p->ldx, b.state, s));
    } else {
        CDR_TRY(cdr_reduce_samples(b.D, b.ldt, 1, p->X, p->ldx, T, d, k, nullptr, p->tmp_kd, p->ldx,
                                   w.stream, w.stream_bytes, b.state, s));
    }
    {
        const int rc = run_reduce_features_tma(p->tmp_kd, p->ldx, p->X, p->ldx, T, d, k, nullptr, 0,
                                               w.stream, w.stream_bytes, b.state, s, nullptr);
        if (rc != 0) return rc == CDR_TMA_NOT_APPLICABLE ? CDR_ERR_UNSUPPORTED : rc;
    }
    // 4. D K, the k x k products, line search, cost check after the dictionary update
    {
        AaFinalizeArgs fa;
        fa.b = b;
        fa.p = dp;
        fa.part = w.stream;
        fa.nstrips = nstrips;
        fa.cta_part = w.fin_part;
        fa.Tl = T;
        fa.row0 = p->row0;
        fa.g = grp;
        fa.dk_offset = sharded ? peer::region_offset(grp, b.DK) : 0;
        const int blocks = (T + kFinTB - 1) / kFinTB;
        if (k <= 8) aa_finalize_ls_kernel<1><<<blocks, kFinThreads, 0, s>>>(fa);
        else aa_finalize_ls_kernel<2><<<blocks, kFinThreads, 0, s>>>(fa);
        CDR_RETURN_IF_LAUNCH_FAILED();
    }
    // 5. per-sample QPs with the update of C, C K and the statistics
    {
        AaWeightsArgs a;
        a.b = b;
        a.Z = p->Z;
        a.cta_part = w.cta_part;
        a.p = p->weights_params;
        a.Tl = T;
        a.row0 = p->row0;
        a.g = grp;
        int blocks;
        fused_grid(T, &a.spw, &blocks);
        if (k <= 8) {
            const size_t smem = aa_fused_smem_bytes(8);
            CDR_TRY(ensure_dyn_smem<aa_weights_fused_kernel<1>>(smem));
            aa_weights_fused_kernel<1><<<blocks, kFusedThreads, smem, s>>>(a);
        } else {
            const size_t smem = aa_fused_smem_bytes(16);
            CDR_TRY(ensure_dyn_smem<aa_weights_fused_kernel<2>>(smem));
            aa_weights_fused_kernel<2><<<blocks, kFusedThreads, smem, s>>>(a);
        }
        CDR_RETURN_IF_LAUNCH_FAILED();
    }
    // 6.-8. (K Z)' and the gradient for the next dictionary step
    if (!sharded) {
        CDR_TRY(cdr_reduce_samples(p->Z, 1, k, p->X, p->ldx, T, d, k, nullptr, p->tmp_kd, p->ldx,
                                   w.stream, w.stream_bytes, b.state, s));
        const int rc = run_reduce_features_tma(p->tmp_kd, p->ldx, p->X, p->ldx, T, d, k, nullptr, 0,
                                               w.stream, w.stream_bytes, b.state, s, nullptr);
        if (rc != 0) return rc == CDR_TMA_NOT_APPLICABLE ? CDR_ERR_UNSUPPORTED : rc;
        const int kp = (k <= 8) ? 8 : 16;
        const int blocks = (int)(((long)T * (kp / 2) + 31) / 32);
        if (k <= 8) aa_kzt_gradient_kernel<1><<<blocks, 256, 0, s>>>(w.stream, nstrips, b);
        else aa_kzt_gradient_kernel<2><<<blocks, 256, 0, s>>>(w.stream, nstrips, b);
        CDR_RETURN_IF_LAUNCH_FAILED();
        return 0;
    }
    CDR_TRY(cdr_reduce_samples_allreduce(p->peers, p->Z, 1, k, p->X, p->ldx, T, p->T_min, d, k, nullptr,
                                         peer::region_offset(grp, p->tmp_kd), p->ldx, b.state, s));
    {
        const int rc = run_reduce_features_tma(p->tmp_kd, p->ldx, p->X, p->ldx, T, d, k, nullptr, 0,
                                               w.stream, w.stream_bytes, b.state, s, nullptr);
        if (rc != 0) return rc == CDR_TMA_NOT_APPLICABLE ? CDR_ERR_UNSUPPORTED : rc;
    }
    {
        const int kp = (k <= 8) ? 8 : 16;
        const long nitems = (long)T * (kp / 2);
        const int blocks = (int)((nitems + 31) / 32);
        const size_t off = peer::region_offset(grp, b.KZt);
        if (k <= 8)
            features_finalize_push_kernel<1><<<blocks, 256, 0, s>>>(w.stream, T, nstrips, k, off, b.ldt,
                                                                    p->row0, grp, b.state);
        else
            features_finalize_push_kernel<2><<<blocks, 256, 0, s>>>(w.stream, T, nstrips, k, off, b.ldt,
                                                                    p->row0, grp, b.state);
        CDR_RETURN_IF_LAUNCH_FAILED();
    }
    return 0;
}
