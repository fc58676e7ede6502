// Whole outer iterations behind one C call (cdr_gpnh_prepare_enqueue /
// cdr_gpnh_iterate_enqueue; the AA twins are in iterate_aa.cu).
//
// The reference's loop `_iterate_gpnh_convex_coding` (gpnh_convex_coding.py:282-402) is, per
// outer iteration: k x k solve -> Z'X -> W -> W'X' -> W'W -> cost check -> per-sample QPs ->
// Z'Z -> cost check.  Here, for k <= 16 at streaming shapes, that is three kernels:
//
//   1. reduce over samples  W' = P (Z'X)          (stream_tma.cu; P applied in the epilogue)
//   2. reduce over features X W as per-strip partials, with W'W formed from the operand
//                           fragments of the same pass (last CTA sums the strips)
//   3. gpnh_weights_fused_kernel: every 8-lane group sums the strip partials of its sample
//      (the "finalize" of pass 2), solves its QP, and contributes z z' and the two trace terms
//      tr(W'X'Z_old), tr(W'X'Z_new); the last CTA to finish sums the per-CTA partials in fixed
//      order, runs both cost checks (gpnh_convex_coding.py:352-384), the stopping rule, and
//      -- unless the loop is done -- starts the next iteration: old_cost and the solve matrix
//      P = pinv(Z'Z / T + lambda G_W) / T.
//
// All reductions have a fixed order, so results are bit-reproducible run to run.
#include "fused_weights.cuh"
#include "peer.cuh"
#include "small_solve.cuh"

namespace cdr {

struct GpnhFusedArgs {
    const double* part;      // [nstrips][T][KP] per-strip partials of X W
    int nstrips;
    const double* WtW;       // k x k (final, from pass 2)
    const double* reg_pairs; // k x k or nullptr
    double* Z;               // T x k
    double* ZtZ;             // k x k: previous on entry, new on exit
    double* P;               // k x k: solve matrix of the next iteration
    double* cta_part;        // [grid][KP * KP + 2]
    cdr_loop_state* state;
    double* cost_deltas;
    int T, k, d, T_total, spw;
    double lambda_W;
    cdr_spg_params p;
    cdr_peer_group g;        // world == 1: single GPU
};

#ifdef CDR_PROFILE_PHASES
// profiling build only (profiles/phase_profile.py): per-warp time stamps of the phases
__device__ unsigned long long cdr_phase_ns[8 * 4096];
__device__ __forceinline__ unsigned long long phase_now()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define CDR_PHASE(slot)                                                                         \
    do {                                                                                        \
        if ((threadIdx.x & 31) == 0)                                                            \
            cdr_phase_ns[(slot) * 4096 + blockIdx.x * kFusedWarps + (threadIdx.x >> 5)] = phase_now(); \
    } while (0)
#define CDR_TAIL_MARK(m)                                                                        \
    do {                                                                                        \
        __syncthreads();                                                                        \
        if (threadIdx.x == 0) cdr_phase_ns[4 * 4096 + 4000 + (m)] = phase_now();                 \
    } while (0)
#else
#define CDR_PHASE(slot)
#define CDR_TAIL_MARK(m)
#endif

template <int KPL>
__global__ void __launch_bounds__(kFusedThreads)
gpnh_weights_fused_kernel(GpnhFusedArgs a)
{
    cdr_loop_state* st = a.state;
    if (is_done(st)) return;
    CDR_PHASE(0);
    constexpr int KP = 8 * KPL;
    constexpr int NST = KP * KP + 2;
    extern __shared__ double fsm[];
    double* As = fsm;                               // KP x KP (KPL > 1)
    double* wsum = fsm + (KPL > 1 ? KP * KP : 0);   // [kFusedWarps][NST]
    double* fin = wsum + kFusedWarps * NST;         // NST + per-warp sums, then Jacobi scratch
    const int k = a.k, T = a.T;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane & 7, q = lane >> 3;

    if constexpr (KPL > 1) {
        for (int idx = threadIdx.x; idx < KP * KP; idx += blockDim.x) {
            const int j = idx / KP, c = idx % KP;
            As[idx] = (j < k && c < k) ? a.WtW[(long)c * k + j] : 0.0;
        }
        __syncthreads();
    }
    double arow[8];
    if constexpr (KPL == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) arow[j] = (g < k && j < k) ? a.WtW[(long)g * k + j] : 0.0;
    }

    // spw samples per warp; the 4 / spw lane groups that share a sample split the strips
    const int spw = a.spw, helpers = 4 / spw;
    const int t_raw = (blockIdx.x * kFusedWarps + warp) * spw + (q % spw);
    const bool has_sample = t_raw < T;
    const bool valid = (q < spw) && has_sample;
    const long t = has_sample ? t_raw : (T - 1);
    const int phase = q / spw;

    // ---- linear term: b = -(X W)[t, :] = -sum over strips of the pass-2 partials
    double xw[KPL], z0[KPL], x[KPL], b[KPL];
    bool present[KPL];
    {
        const int n_mine = (a.nstrips - phase + helpers - 1) / helpers;
        const double* base = a.part + ((long)phase * T + t) * KP + g * KPL;
        const long stride = (long)helpers * T * KP;
#pragma unroll
        for (int r = 0; r < KPL; ++r) {
            double v = strided_sum_cg32(base + r, stride, n_mine);
            // combine the phases in fixed order (every lane group ends up with the total)
            double tot = __shfl_sync(CDR_FULL_MASK, v, (q % spw) * 8 + g);
            for (int ph = 1; ph < helpers; ++ph)
                tot += __shfl_sync(CDR_FULL_MASK, v, (ph * spw + (q % spw)) * 8 + g);
            xw[r] = tot;
        }
    }
    double tr_old = 0.0;
#pragma unroll
    for (int r = 0; r < KPL; ++r) {
        const int c = g * KPL + r;
        present[r] = c < k;
        if (present[r]) {
            z0[r] = a.Z[t * k + c];
            b[r] = -xw[r];
            tr_old = fma(xw[r], z0[r], tr_old);
        } else {
            z0[r] = -INFINITY;
            b[r] = 0.0;
        }
    }
    tr_old = group8_sum(tr_old);
    CDR_PHASE(1);

    int n_iter = 0, n_feval = 0;
    qp_solve<KPL>(As, arow, z0, b, present, a.p, has_sample, g, spw, x, n_iter, n_feval);
    CDR_PHASE(2);
#ifdef CDR_PROFILE_PHASES
    if (lane == 0) {
        cdr_phase_ns[6 * 4096 + blockIdx.x * kFusedWarps + warp] = (unsigned long long)n_iter;
        cdr_phase_ns[7 * 4096 + blockIdx.x * kFusedWarps + warp] = (unsigned long long)n_feval;
    }
#endif

    double tr_new = 0.0;
    if (valid) {
#pragma unroll
        for (int r = 0; r < KPL; ++r)
            if (present[r]) {
                a.Z[t * k + g * KPL + r] = x[r];
                tr_new = fma(xw[r], x[r], tr_new);
            }
    }
    tr_new = group8_sum(tr_new);

    if (!fused_sample_statistics<KPL>(x, present, valid, k, tr_old, tr_new, wsum, a.cta_part,
                                      &st->tickets[1])) {
        CDR_PHASE(3);
        return;
    }
    CDR_PHASE(3);
    CDR_TAIL_MARK(0);
    fused_final_sum<KPL>(a.cta_part, fin);
    CDR_PHASE(4);
    CDR_TAIL_MARK(1);
    // sample-sharded fit: the statistics of all ranks, summed in rank order on every rank
    if (a.g.world > 1) peer::cta_allreduce_small(a.g, fin, NST);
    // fin[i * KP + j] = (Z'Z)[i][j] of the new weights; fin[KP*KP], fin[KP*KP+1] the traces
    if (warp == 0) {
        double tp = 0.0, tn = 0.0, phi = 0.0;
        for (int idx = lane; idx < k * k; idx += 32) {
            const int i = idx / k, j = idx % k;
            const double w = a.WtW[j * k + i];
            tp += a.ZtZ[idx] * w;                      // Z'Z of the previous weights
            tn += fin[i * KP + j] * w;
            if (a.reg_pairs != nullptr && j > i) phi += a.reg_pairs[idx];
        }
        tp = warp_sum(tp);
        tn = warp_sum(tn);
        phi = warp_sum(phi);
        if (lane == 0) {
            if (a.reg_pairs != nullptr) {
                // gpnh_convex_coding.py:179-196
                double pen = 0.0;
                if (a.lambda_W != 0.0 && k > 1)
                    pen = a.lambda_W * phi * 2.0 / ((double)k * (double)a.d * ((double)k - 1.0));
                st->penalty = pen;
            }
            const double nT = (double)a.T_total;
            // after the dictionary update (gpnh_convex_coding.py:352-369)
            const double cost_d = 0.5 * (st->trace_data - 2.0 * fin[KP * KP] + tp) / nT + st->penalty;
            finish_sub_step(st, a.cost_deltas, cost_d, 2, 0);
            if (!st->done) {
                // after the weights update (:371-384)
                const double cost_w = 0.5 * (st->trace_data - 2.0 * fin[KP * KP + 1] + tn) / nT + st->penalty;
                finish_sub_step(st, a.cost_deltas, cost_w, 3, 1);
            }
            if (!st->done) st->old_cost = st->cost;    // start of the next iteration
            st->tickets[1] = 0u;
        }
    }
    __syncthreads();
    CDR_TAIL_MARK(2);
    for (int idx = threadIdx.x; idx < k * k; idx += blockDim.x)
        a.ZtZ[idx] = fin[(idx / k) * KP + idx % k];
    if (*((volatile int*)&st->done)) return;          // CTA-uniform
    // solve matrix of the next dictionary step (gpnh_convex_coding.py:221-226)
    {
        double* jac = fin + 4 * NST;
        double* S = jac + 2 * kFusedMaxK * kJacLd;     // dense k x k copy
        for (int idx = threadIdx.x; idx < k * k; idx += blockDim.x)
            S[idx] = fin[(idx / k) * KP + idx % k];
        __syncthreads();
        const double pref = (k > 1) ? 4.0 / ((double)a.d * k * (k - 1)) : 0.0;
        CDR_TAIL_MARK(3);
        solve_matrix_cta<KP>(S, k, kFusedMaxK, 1.0 / (double)a.T_total, a.lambda_W, pref, a.P, jac);
    }
    CDR_TAIL_MARK(4);
    CDR_PHASE(5);
}

#ifdef CDR_PROFILE_PHASES
}  // namespace cdr
extern "C" int cdr_debug_phase_read(unsigned long long* out)
{
    return (int)cudaMemcpyFromSymbol(out, cdr::cdr_phase_ns, sizeof(unsigned long long) * 8 * 4096);
}
namespace cdr {
#endif

static size_t fused_smem_bytes(int kp)
{
    const int nst = kp * kp + 2;
    return ((kp > 8 ? (size_t)kp * kp : 0) + (size_t)kFusedWarps * nst + fused_fin_doubles(nst) +
            2 * (size_t)kFusedMaxK * kJacLd + (size_t)kFusedMaxK * kFusedMaxK) * sizeof(double);
}

// ---------------------------------------------------------------------- workspace layout
struct GpnhWorkspace {
    double* stream;       // the two passes (partials of pass 2 live here)
    size_t stream_bytes;
    double* gram;         // cdr_small_gram scratch
    size_t gram_bytes;
    double* wtw_part;     // [nstrips][KP * KP]
    double* cta_part;     // [blocks][KP * KP + 2]
    size_t total;
};

static GpnhWorkspace carve_gpnh(void* base, int T, int d, int k)
{
    GpnhWorkspace w;
    const size_t s1 = cdr_reduce_samples_workspace_bytes(T, d, k);
    const size_t s2 = cdr_reduce_features_workspace_bytes(T, d, k);
    w.stream_bytes = align256((s1 > s2 ? s1 : s2) + 8);
    w.gram_bytes = align256(cdr_small_gram_workspace_bytes());
    const int kp = (k <= 8) ? 8 : 16;
    int TC = 0, nstrips = 0, spw = 1, blocks = 1;
    size_t wtw = 0, cta = 0;
    if (k <= kFusedMaxK && features_strip_geometry(T, d, k, &TC, &nstrips)) {
        fused_grid(T, &spw, &blocks);
        wtw = align256((size_t)nstrips * kp * kp * sizeof(double));
        cta = align256((size_t)blocks * (kp * kp + 2) * sizeof(double));
    }
    unsigned char* p = static_cast<unsigned char*>(base);
    w.stream = reinterpret_cast<double*>(p);
    w.gram = reinterpret_cast<double*>(p + w.stream_bytes);
    w.wtw_part = reinterpret_cast<double*>(p + w.stream_bytes + w.gram_bytes);
    w.cta_part = reinterpret_cast<double*>(p + w.stream_bytes + w.gram_bytes + wtw);
    w.total = w.stream_bytes + w.gram_bytes + wtw + cta;
    return w;
}

static bool gpnh_fused_applicable(const cdr_gpnh_problem* p, int* TC, int* nstrips)
{
    const bool sharded = p->peers != nullptr && p->peers->world > 1;
    const int T = sharded ? p->T_min : p->T;       // every rank must take the same decision
    if (p->k > kFusedMaxK) return false;
    const char* e = getenv("CDR_DISABLE_FUSED");
    if (e != nullptr && e[0] == '1') return false;
    if ((p->ldx % 2) != 0 || (((uintptr_t)p->X) & 15) != 0 || (((uintptr_t)p->WT) & 15) != 0)
        return false;
    int out[12];
    tma_stream_plan(T, p->d, p->k, 1, out);
    if (!out[0] || !out[5]) return false;              // both passes on the strip kernels
    return features_strip_geometry(T, p->d, p->k, TC, nstrips);
}

static int check_gpnh(const cdr_gpnh_problem* p)
{
    CDR_CHECK_ARG(p != nullptr && p->X != nullptr && p->Z != nullptr && p->WT != nullptr);
    CDR_CHECK_ARG(p->T >= 1 && p->d >= 1 && p->k >= 1 && p->T_total >= p->T);
    CDR_CHECK_ARG(p->XWt != nullptr && p->ldt >= p->T && p->state != nullptr);
    CDR_CHECK_ARG(p->ZtZ && p->XWtZ && p->WtW && p->REG && p->P);
    if (p->k > CDR_MAX_COMPONENTS) return CDR_ERR_UNSUPPORTED;
    if (p->workspace == nullptr || p->workspace_bytes < cdr_gpnh_workspace_bytes(p->T, p->d, p->k))
        return CDR_ERR_WORKSPACE;
    return 0;
}

static int gpnh_dictionary_products(const cdr_gpnh_problem* p, const GpnhWorkspace& w,
                                    bool with_ZtZ, cudaStream_t s)
{
    const int T = p->T, d = p->d, k = p->k;
    CDR_TRY(cdr_reduce_features(p->WT, p->ldx, p->X, p->ldx, T, d, k, p->XWt, p->ldt, w.stream,
                                w.stream_bytes, p->state, s));
    cdr_small_gram_desc ds[4];
    int n = 0;
    if (with_ZtZ) ds[n++] = gram_desc(p->Z, 1, k, k, p->Z, 1, k, k, T, p->ZtZ, 0);
    ds[n++] = gram_desc(p->WT, p->ldx, 1, k, p->WT, p->ldx, 1, k, d, p->WtW, 0);
    ds[n++] = gram_desc(p->XWt, p->ldt, 1, k, p->Z, 1, k, k, T, p->XWtZ, 0);
    if (p->lambda_W != 0.0) ds[n++] = gram_desc(p->WT, p->ldx, 1, k, p->WT, p->ldx, 1, k, d, p->REG, 1);
    return cdr_small_gram(ds, n, w.gram, w.gram_bytes, p->state, s);
}

}  // namespace cdr

using namespace cdr;

extern "C" size_t cdr_gpnh_workspace_bytes(int T, int d, int k)
{
    if (T < 1 || d < 1 || k < 1 || k > CDR_MAX_COMPONENTS) return 0;
    return carve_gpnh(nullptr, T, d, k).total;
}

extern "C" int cdr_gpnh_fused_applicable(int T, int d, int k)
{
    cdr_gpnh_problem p = {};
    p.T = T; p.d = d; p.k = k; p.ldx = (d + 31) / 32 * 32;
    int TC, nstrips;
    return gpnh_fused_applicable(&p, &TC, &nstrips) ? 1 : 0;
}

extern "C" int cdr_gpnh_prepare_enqueue(const cdr_gpnh_problem* p, cdr_stream_t stream)
{
    CDR_TRY(check_gpnh(p));
    // sample-sharded fits: the caller forms the initial statistics with its own collectives
    if (p->peers != nullptr && p->peers->world > 1) return CDR_ERR_UNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    const GpnhWorkspace w = carve_gpnh(p->workspace, p->T, p->d, p->k);
    // gpnh_convex_coding.py:292-314
    CDR_TRY(gpnh_dictionary_products(p, w, true, s));
    CDR_TRY(cdr_gpnh_cost_check(p->state, p->cost_deltas, p->XWtZ, p->ZtZ, p->WtW,
                                p->lambda_W != 0.0 ? p->REG : nullptr, p->k, p->T_total, p->d,
                                p->lambda_W, 0, 0, s));
    // start of the first iteration
    CDR_TRY(cdr_loop_begin(p->state, s));
    return cdr_gpnh_solve_matrix(p->ZtZ, p->k, p->T_total, p->d, p->lambda_W, p->P, nullptr, 0,
                                 p->state, s);
}

extern "C" int cdr_gpnh_iterate_enqueue(const cdr_gpnh_problem* p, cdr_stream_t stream)
{
    CDR_TRY(check_gpnh(p));
    cudaStream_t s = (cudaStream_t)stream;
    const int T = p->T, d = p->d, k = p->k;
    const GpnhWorkspace w = carve_gpnh(p->workspace, T, d, k);
    const double* reg = p->lambda_W != 0.0 ? p->REG : nullptr;
    int TC = 0, nstrips = 0;

    const bool sharded = p->peers != nullptr && p->peers->world > 1;
    const bool fused = gpnh_fused_applicable(p, &TC, &nstrips);
    if (sharded) {
        // W' = P sum_r Z_r' X_r: the strip kernel with the sum over ranks in its epilogue
        // (stream_tma.cu); only the three-kernel path is offered over peer memory
        if (!fused) return CDR_ERR_UNSUPPORTED;
        CDR_CHECK_ARG(p->T_min >= 1 && p->T_min <= T);
        CDR_TRY(cdr_reduce_samples_allreduce(p->peers, p->Z, 1, k, p->X, p->ldx, T, p->T_min, d, k,
                                             p->P, peer::region_offset(*p->peers, p->WT), p->ldx,
                                             p->state, s));
    } else {
        // W' = P Z'X  (gpnh_convex_coding.py:219-226)
        CDR_TRY(cdr_reduce_samples(p->Z, 1, k, p->X, p->ldx, T, d, k, p->P, p->WT, p->ldx, w.stream,
                                   w.stream_bytes, p->state, s));
    }

    if (!fused) {
        // general sequence (any k <= 64, any shape)
        CDR_TRY(gpnh_dictionary_products(p, w, false, s));
        CDR_TRY(cdr_gpnh_cost_check(p->state, p->cost_deltas, p->XWtZ, p->ZtZ, p->WtW, reg, k,
                                    p->T_total, d, p->lambda_W, 2, 0, s));
        CDR_TRY(cdr_quad_simplex_spg_batched(p->WtW, nullptr, p->XWt, 1, p->ldt, p->Z, T, k,
                                             &p->weights_params, nullptr, nullptr, p->state, s));
        cdr_small_gram_desc ds[2];
        ds[0] = gram_desc(p->Z, 1, k, k, p->Z, 1, k, k, T, p->ZtZ, 0);
        ds[1] = gram_desc(p->XWt, p->ldt, 1, k, p->Z, 1, k, k, T, p->XWtZ, 0);
        CDR_TRY(cdr_small_gram(ds, 2, w.gram, w.gram_bytes, p->state, s));
        CDR_TRY(cdr_gpnh_cost_check(p->state, p->cost_deltas, p->XWtZ, p->ZtZ, p->WtW, nullptr, k,
                                    p->T_total, d, p->lambda_W, 3, 1, s));
        CDR_TRY(cdr_loop_begin(p->state, s));
        return cdr_gpnh_solve_matrix(p->ZtZ, k, p->T_total, d, p->lambda_W, p->P, nullptr, 0,
                                     p->state, s);
    }

    // X W as per-strip partials + W'W from the same pass
    StripGram gram = {w.wtw_part, p->WtW, &p->state->tickets[0]};
    {
        const int rc = run_reduce_features_tma(p->WT, p->ldx, p->X, p->ldx, T, d, k, nullptr, 0,
                                               w.stream, w.stream_bytes, p->state, s, &gram);
        if (rc != 0) return rc == CDR_TMA_NOT_APPLICABLE ? CDR_ERR_UNSUPPORTED : rc;
    }
    if (reg != nullptr) {
        cdr_small_gram_desc ds = gram_desc(p->WT, p->ldx, 1, k, p->WT, p->ldx, 1, k, d, p->REG, 1);
        CDR_TRY(cdr_small_gram(&ds, 1, w.gram, w.gram_bytes, p->state, s));
    }
    GpnhFusedArgs a;
    a.part = w.stream;
    a.nstrips = nstrips;
    a.WtW = p->WtW;
    a.reg_pairs = reg;
    a.Z = p->Z;
    a.ZtZ = p->ZtZ;
    a.P = p->P;
    a.cta_part = w.cta_part;
    a.state = p->state;
    a.cost_deltas = p->cost_deltas;
    a.T = T; a.k = k; a.d = d; a.T_total = p->T_total;
    a.lambda_W = p->lambda_W;
    a.p = p->weights_params;
    if (sharded) {
        a.g = *p->peers;
    } else {
        a.g = cdr_peer_group();
        a.g.world = 1;
    }
    if (a.p.memory < 1 || a.p.memory > CDR_MAX_MEMORY) return CDR_ERR_UNSUPPORTED;
    int blocks;
    fused_grid(T, &a.spw, &blocks);
    if (k <= 8) {
        const size_t smem = fused_smem_bytes(8);
        CDR_TRY(ensure_dyn_smem<gpnh_weights_fused_kernel<1>>(smem));
        gpnh_weights_fused_kernel<1><<<blocks, kFusedThreads, smem, s>>>(a);
    } else {
        const size_t smem = fused_smem_bytes(16);
        CDR_TRY(ensure_dyn_smem<gpnh_weights_fused_kernel<2>>(smem));
        gpnh_weights_fused_kernel<2><<<blocks, kFusedThreads, smem, s>>>(a);
    }
    CDR_RETURN_IF_LAUNCH_FAILED();
    return 0;
}
