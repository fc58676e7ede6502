// Streaming fp64 contractions of the data matrix X (T x d, row-major, padded
// leading dimension) with k <= 64 vectors -- the passes over X that dominate the
// reference's run time:
//
//   reduce over samples   out(k x d) = E (L X)      dictionary.dot(X), weights.T.dot(X)
//                                                   archetypal_analysis.py:545,548,618,641
//                                                   gpnh_convex_coding.py:219-224
//   reduce over features  out(k x T) = M X'         CX.dot(X.T), X.dot(XtZ), X.dot(W)
//                                                   archetypal_analysis.py:546,549,619,642
//                                                   gpnh_convex_coding.py:271,352
//   gram                  K(T x T)  = X X'          archetypal_analysis.py:1032
//
// Both passes are HBM bound for k <= 20 (arithmetic intensity k/4 flop/B): the
// design goal is to read every byte of X exactly once with wide, fully used
// sectors and to do the 2kTd flops on the fp64 tensor pipe (DMMA.8x8x4) so the
// FMA pipe and the register file do not become the limiter.
//
// Fragment trick used throughout: the reduction index of an MMA may be permuted
// freely as long as A and B use the same permutation, and the n index of B maps
// to an arbitrary (but known) output column.  Each lane therefore loads a
// double2 (16 B) of X and feeds .x and .y to two different MMAs, which turns the
// 8-byte-per-lane DMMA fragment loads into 128-bit loads covering whole sectors.
#include "cdr_common.cuh"

namespace cdr {

// ======================================================================
// reduce over samples
// ======================================================================
// warp tile: 32 features x all rows of its split.  Lane l loads, for k-step t0,
//   xa = X[t0 + (l&3)][f0 + 2*(l>>2) + {0,1}],  xb = same + 16
// MMA q in {0,1,2,3} uses {xa.x, xa.y, xb.x, xb.y}: column n of that MMA is
// feature f0 + off_q + 2n with off = {0, 1, 16, 17}.  The C fragment of lane l
// then holds, for row (l>>2), features f0 + 4*(l&3) + {0,1,2,3} (q = 0,1) and
// f0 + 16 + 4*(l&3) + {0,1,2,3} (q = 2,3): two aligned 32-byte runs.
template <int KT, int U>
__global__ void __launch_bounds__(128)
reduce_samples_kernel(const double* __restrict__ Lp, long sLi, long sLt,
                      const double* __restrict__ X, long ldx, int T, int dpad, int k,
                      int rows_per_split, double* __restrict__ dst, long ldo, long split_stride,
                      const cdr_flags* flags)
{
    if (is_done(flags)) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f0 = (blockIdx.x * 4 + warp) * 32;
    if (f0 >= dpad) return;
    const int r0 = blockIdx.y * rows_per_split;
    const int r1 = min(T, r0 + rows_per_split);
    const int lr = lane & 3, lc = lane >> 2;

    double c[KT][4][2];
#pragma unroll
    for (int mt = 0; mt < KT; ++mt)
#pragma unroll
        for (int q = 0; q < 4; ++q) c[mt][q][0] = c[mt][q][1] = 0.0;

    const double* xcol = X + f0 + 2 * lc;
    for (int t0 = r0; t0 < r1; t0 += 4 * U) {
        double2 xa[U], xb[U];
        double a[U][KT];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int tt = t0 + 4 * u + lr;
            const bool in = tt < r1;
            const long row = in ? tt : (T - 1);
            const double* p = xcol + row * ldx;
            xa[u] = ldg_nc_d2(p);
            xb[u] = ldg_nc_d2(p + 16);
#pragma unroll
            for (int mt = 0; mt < KT; ++mt) {
                const int i = mt * 8 + lc;
                a[u][mt] = (in && i < k) ? Lp[(long)i * sLi + (long)tt * sLt] : 0.0;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
            for (int mt = 0; mt < KT; ++mt) {
                dmma884(c[mt][0][0], c[mt][0][1], a[u][mt], xa[u].x);
                dmma884(c[mt][1][0], c[mt][1][1], a[u][mt], xa[u].y);
                dmma884(c[mt][2][0], c[mt][2][1], a[u][mt], xb[u].x);
                dmma884(c[mt][3][0], c[mt][3][1], a[u][mt], xb[u].y);
            }
        }
    }

    double* base = dst + (long)blockIdx.y * split_stride + f0 + 4 * lr;
#pragma unroll
    for (int mt = 0; mt < KT; ++mt) {
        const int i = mt * 8 + lc;
        if (i < k) {
            double* p = base + (long)i * ldo;
            *reinterpret_cast<double2*>(p) = make_double2(c[mt][0][0], c[mt][1][0]);
            *reinterpret_cast<double2*>(p + 2) = make_double2(c[mt][0][1], c[mt][1][1]);
            *reinterpret_cast<double2*>(p + 16) = make_double2(c[mt][2][0], c[mt][3][0]);
            *reinterpret_cast<double2*>(p + 18) = make_double2(c[mt][2][1], c[mt][3][1]);
        }
    }
}

// out[j][f] = sum_i E[j][i] * (sum_s part[s][i][f]);  E == nullptr -> identity.
// Fixed summation order over the splits => deterministic.
template <int KT>
__global__ void __launch_bounds__(128)
reduce_samples_finalize_kernel(const double* __restrict__ part, long split_stride, int nsplit,
                               long ldp, const double* __restrict__ E, int k, int dpad,
                               double* __restrict__ out, long ldo, const cdr_flags* flags)
{
    if (is_done(flags)) return;
    constexpr int KP = 8 * KT;
    __shared__ double Es[KP * KP];
    if (E != nullptr) {
        for (int idx = threadIdx.x; idx < k * k; idx += blockDim.x) Es[idx] = E[idx];
    }
    __syncthreads();
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= dpad) return;
    double acc[KP];
#pragma unroll
    for (int i = 0; i < KP; ++i) {
        double s = 0.0;
        if (i < k)
            for (int sp = 0; sp < nsplit; ++sp) s += part[(long)sp * split_stride + (long)i * ldp + f];
        acc[i] = s;
    }
    if (E == nullptr) {
#pragma unroll
        for (int i = 0; i < KP; ++i)
            if (i < k) out[(long)i * ldo + f] = acc[i];
    } else {
        for (int j = 0; j < k; ++j) {
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < KP; ++i)
                if (i < k) s = fma(Es[j * k + i], acc[i], s);
            out[(long)j * ldo + f] = s;
        }
    }
}

static int samples_nsplit(int T, int dpad)
{
    // one warp per 32 features; split the sample axis only when the feature
    // axis alone cannot fill the machine (Gram-space "X" = K, PCA-reduced data)
    const int warps = dpad / 32;
    const int target = 148 * 12;
    int ns = 1;
    if (warps < target) ns = (target + warps - 1) / warps;
    const int max_ns = (T + 63) / 64;
    if (ns > max_ns) ns = max_ns;
    if (ns < 1) ns = 1;
    return ns;
}

template <int KT>
static int run_reduce_samples(const double* Lp, long sLi, long sLt, const double* X, long ldx,
                              int T, int d, int k, const double* E, double* out, long ldo,
                              void* workspace, size_t workspace_bytes, const cdr_flags* flags,
                              cudaStream_t stream)
{
    const int dpad = (d + 31) / 32 * 32;
    const int ns = samples_nsplit(T, dpad);
    int rows_per_split = (T + ns - 1) / ns;
    rows_per_split = (rows_per_split + 3) / 4 * 4;
    const bool direct = (ns == 1 && E == nullptr);
    double* dst = out;
    long dst_ld = ldo, split_stride = 0;
    if (!direct) {
        const size_t need = (size_t)ns * k * dpad * sizeof(double);
        if (workspace == nullptr || workspace_bytes < need) return CDR_ERR_WORKSPACE;
        dst = (double*)workspace;
        dst_ld = dpad;
        split_stride = (long)k * dpad;
    }
    dim3 grid((dpad / 32 + 3) / 4, ns);
    constexpr int U = (KT <= 2) ? 8 : (KT <= 4 ? 4 : 2);
    reduce_samples_kernel<KT, U><<<grid, 128, 0, stream>>>(Lp, sLi, sLt, X, ldx, T, dpad, k,
                                                          rows_per_split, dst, dst_ld,
                                                          split_stride, flags);
    CDR_RETURN_IF_LAUNCH_FAILED();
    if (!direct) {
        reduce_samples_finalize_kernel<KT><<<(dpad + 127) / 128, 128, 0, stream>>>(
            dst, split_stride, ns, dst_ld, E, k, dpad, out, ldo, flags);
        CDR_RETURN_IF_LAUNCH_FAILED();
    }
    return 0;
}

// ======================================================================
// reduce over features
// ======================================================================
// warp tile: 32 rows x one feature chunk.  A = X rows (m = sample, k = feature),
// B = M (k = feature, n = component).  Lane l loads double2 at feature
// f + 2*(l&3) of row (l>>2) for both operands; .x and .y feed two MMAs.
template <int KT, int U>
__global__ void __launch_bounds__(128)
reduce_features_kernel(const double* __restrict__ M, long ldm, const double* __restrict__ X,
                       long ldx, int T, int dpad, int k, int chunk, double* __restrict__ part,
                       long Tp, const cdr_flags* flags)
{
    if (is_done(flags)) return;
    constexpr int KP = 8 * KT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rb0 = (blockIdx.x * 4 + warp) * 32;
    if (rb0 >= T) return;
    const int c0 = blockIdx.y * chunk;
    const int c1 = min(dpad, c0 + chunk);
    const int lr = lane & 3, lc = lane >> 2;

    double c[4][KT][2];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < KT; ++nt) c[mt][nt][0] = c[mt][nt][1] = 0.0;

    const double* xrow[4];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
        const int row = min(rb0 + mt * 8 + lc, T - 1);
        xrow[mt] = X + (long)row * ldx + 2 * lr;
    }
    const double* mrow[KT];
    bool mok[KT];
#pragma unroll
    for (int nt = 0; nt < KT; ++nt) {
        const int j = nt * 8 + lc;
        mok[nt] = j < k;
        mrow[nt] = M + (long)(mok[nt] ? j : 0) * ldm + 2 * lr;
    }

    for (int f = c0; f < c1; f += 8 * U) {
        double2 xa[U][4], mb[U][KT];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int ff = f + 8 * u;
            const bool in = ff < c1;
            const int fc = in ? ff : c0;           // in-range address, value discarded
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) xa[u][mt] = ldg_nc_d2(xrow[mt] + fc);
#pragma unroll
            for (int nt = 0; nt < KT; ++nt) {
                double2 v = make_double2(0.0, 0.0);
                if (in && mok[nt]) v = *reinterpret_cast<const double2*>(mrow[nt] + fc);
                mb[u][nt] = v;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) {
#pragma unroll
                for (int nt = 0; nt < KT; ++nt) {
                    dmma884(c[mt][nt][0], c[mt][nt][1], xa[u][mt].x, mb[u][nt].x);
                    dmma884(c[mt][nt][0], c[mt][nt][1], xa[u][mt].y, mb[u][nt].y);
                }
            }
        }
    }

    // partials: part[chunk][t][KP]
    double* base = part + ((long)blockIdx.y * Tp) * KP + 2 * lr;
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
        const int t = rb0 + mt * 8 + lc;
        if (t < T) {
#pragma unroll
            for (int nt = 0; nt < KT; ++nt)
                *reinterpret_cast<double2*>(base + (long)t * KP + nt * 8) =
                    make_double2(c[mt][nt][0], c[mt][nt][1]);
        }
    }
}

template <int KT>
__global__ void __launch_bounds__(128)
reduce_features_finalize_kernel(const double* __restrict__ part, long Tp, int nchunk, int T, int k,
                                double* __restrict__ out, long ldo, const cdr_flags* flags)
{
    if (is_done(flags)) return;
    constexpr int KP = 8 * KT;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    double acc[KP];
#pragma unroll
    for (int j = 0; j < KP; ++j) acc[j] = 0.0;
    for (int ch = 0; ch < nchunk; ++ch) {
        const double2* p = reinterpret_cast<const double2*>(part + ((long)ch * Tp + t) * KP);
#pragma unroll
        for (int j = 0; j < KP / 2; ++j) {
            const double2 v = p[j];
            acc[2 * j] += v.x;
            acc[2 * j + 1] += v.y;
        }
    }
#pragma unroll
    for (int j = 0; j < KP; ++j)
        if (j < k) out[(long)j * ldo + t] = acc[j];
}

static void features_split(int T, int dpad, int* nchunk, int* chunk)
{
    const int row_warps = (T + 31) / 32;
    const int target = 148 * 16;
    int nc = (target + row_warps - 1) / row_warps;
    const int max_nc = (dpad + 255) / 256;
    if (nc > max_nc) nc = max_nc;
    if (nc < 1) nc = 1;
    int ch = (dpad + nc - 1) / nc;
    ch = (ch + 7) / 8 * 8;
    nc = (dpad + ch - 1) / ch;
    *nchunk = nc;
    *chunk = ch;
}

template <int KT>
static int run_reduce_features(const double* M, long ldm, const double* X, long ldx, int T, int d,
                               int k, double* out, long ldo, void* workspace,
                               size_t workspace_bytes, const cdr_flags* flags, cudaStream_t stream)
{
    constexpr int KP = 8 * KT;
    const int dpad = (d + 31) / 32 * 32;
    int nchunk, chunk;
    features_split(T, dpad, &nchunk, &chunk);
    const long Tp = T;
    const size_t need = (size_t)nchunk * Tp * KP * sizeof(double);
    if (workspace == nullptr || workspace_bytes < need) return CDR_ERR_WORKSPACE;
    double* part = (double*)workspace;
    dim3 grid(((T + 31) / 32 + 3) / 4, nchunk);
    constexpr int U = (KT <= 2) ? 4 : (KT <= 4 ? 2 : 1);
    reduce_features_kernel<KT, U><<<grid, 128, 0, stream>>>(M, ldm, X, ldx, T, dpad, k, chunk,
                                                           part, Tp, flags);
    CDR_RETURN_IF_LAUNCH_FAILED();
    reduce_features_finalize_kernel<KT><<<(T + 127) / 128, 128, 0, stream>>>(part, Tp, nchunk, T,
                                                                             k, out, ldo, flags);
    CDR_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

// ======================================================================
// Frobenius norm squared (trace of the Gram matrix)
// ======================================================================
constexpr int kFrobBlocks = 592;

__global__ void __launch_bounds__(256)
frobenius_partial_kernel(const double* __restrict__ X, long ldx, int T, int d, double* part)
{
    __shared__ double scratch[32];
    double s[1] = {0.0};
    // one row at a time per CTA: coalesced 16-byte loads, fixed assignment => deterministic
    // (rows are 256-byte aligned and zero padded, so the padded tail may be read)
    const int d2 = (d + 1) / 2;
    for (int t = blockIdx.x; t < T; t += gridDim.x) {
        const double2* row = reinterpret_cast<const double2*>(X + (long)t * ldx);
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int f = threadIdx.x;
        for (; f + 256 < d2; f += 512) {
            const double2 v = row[f], w = row[f + 256];
            a0 = fma(v.x, v.x, a0);
            a1 = fma(v.y, v.y, a1);
            a2 = fma(w.x, w.x, a2);
            a3 = fma(w.y, w.y, a3);
        }
        for (; f < d2; f += 256) {
            const double2 v = row[f];
            a0 = fma(v.x, v.x, a0);
            if (2 * f + 1 < d) a1 = fma(v.y, v.y, a1);
        }
        s[0] += (a0 + a1) + (a2 + a3);
    }
    block_sum<1>(s, scratch);
    if (threadIdx.x == 0) part[blockIdx.x] = s[0];
}

__global__ void __launch_bounds__(256) frobenius_final_kernel(const double* part, int n, double* out)
{
    __shared__ double scratch[32];
    double s[1] = {0.0};
    for (int i = threadIdx.x; i < n; i += blockDim.x) s[0] += part[i];
    block_sum<1>(s, scratch);
    if (threadIdx.x == 0) *out = s[0];
}

// ======================================================================
// ||X - Z A||_F^2 with direct differences (gpnh_convex_coding.py:199-210,
// archetypal_analysis.py:1196-1197).  One CTA per sample row; fixed order.
// ======================================================================
__global__ void __launch_bounds__(256)
residual_partial_kernel(const double* __restrict__ X, long ldx, int d, const double* __restrict__ Z,
                        int k, const double* __restrict__ A, long lda, double* part)
{
    __shared__ double scratch[32];
    __shared__ double z[CDR_MAX_COMPONENTS];
    const int t = blockIdx.x;
    for (int j = threadIdx.x; j < k; j += blockDim.x) z[j] = Z[(long)t * k + j];
    __syncthreads();
    const double* row = X + (long)t * ldx;
    double s[1] = {0.0};
    for (int f = threadIdx.x; f < d; f += blockDim.x) {
        double r = 0.0;
        for (int j = 0; j < k; ++j) r = fma(z[j], A[(long)j * lda + f], r);
        const double df = row[f] - r;
        s[0] = fma(df, df, s[0]);
    }
    block_sum<1>(s, scratch);
    if (threadIdx.x == 0) part[t] = s[0];
}

}  // namespace cdr

using namespace cdr;

extern "C" int cdr_residual_sq(const double* X, long ldx, int T, int d, const double* Z, int k,
                               const double* A, long lda, double* out, double* part,
                               cdr_stream_t stream)
{
    CDR_CHECK_ARG(T >= 1 && d >= 1 && k >= 1 && ldx >= d && lda >= d);
    if (k > CDR_MAX_COMPONENTS) return CDR_ERR_UNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    residual_partial_kernel<<<T, 256, 0, s>>>(X, ldx, d, Z, k, A, lda, part);
    CDR_RETURN_IF_LAUNCH_FAILED();
    frobenius_final_kernel<<<1, 256, 0, s>>>(part, T, out);
    CDR_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

extern "C" int cdr_sum_vector(const double* v, int n, double* out, cdr_stream_t stream)
{
    CDR_CHECK_ARG(n >= 0);
    frobenius_final_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(v, n, out);
    CDR_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

#define CDR_DISPATCH_KT(k, CALL)                    \
    do {                                            \
        if ((k) <= 8) { CALL(1); }                  \
        else if ((k) <= 16) { CALL(2); }            \
        else if ((k) <= 24) { CALL(3); }            \
        else if ((k) <= 32) { CALL(4); }            \
        else if ((k) <= 48) { CALL(6); }            \
        else { CALL(8); }                           \
    } while (0)

extern "C" size_t cdr_reduce_samples_workspace_bytes(int T, int d, int k)
{
    const int dpad = (d + 31) / 32 * 32;
    const int ns = samples_nsplit(T, dpad);
    const size_t direct = (size_t)ns * k * dpad * sizeof(double);
    const size_t gemm = samples64_workspace_bytes(T, d, k);
    return direct > gemm ? direct : gemm;
}

extern "C" int cdr_reduce_samples(const double* Lp, long sLi, long sLt, const double* X, long ldx,
                                  int T, int d, int k, const double* E, double* out, long ldo,
                                  void* workspace, size_t workspace_bytes, const cdr_flags* flags,
                                  cdr_stream_t stream)
{
    CDR_CHECK_ARG(T >= 1 && d >= 1 && k >= 1);
    if (k > CDR_MAX_COMPONENTS) return CDR_ERR_UNSUPPORTED;
    const int dpad = (d + 31) / 32 * 32;
    CDR_CHECK_ARG(ldx >= dpad && ldo >= dpad && ldx % 2 == 0 && ldo % 2 == 0);
    cudaStream_t s = (cudaStream_t)stream;
    {
        // 16 < k <= 64: the tensor-bound GEMM-shaped kernel
        const int rc = run_reduce_samples64(Lp, sLi, sLt, X, ldx, T, d, k, E, out, ldo, workspace,
                                            workspace_bytes, flags, s);
        if (rc != CDR_TMA_NOT_APPLICABLE) return rc;
    }
    {
        const int rc = run_reduce_samples_tma(Lp, sLi, sLt, X, ldx, T, d, k, E, out, ldo, flags, s);
        if (rc != CDR_TMA_NOT_APPLICABLE) return rc;
    }
#define CALL(KT) return run_reduce_samples<KT>(Lp, sLi, sLt, X, ldx, T, d, k, E, out, ldo, workspace, workspace_bytes, flags, s)
    CDR_DISPATCH_KT(k, CALL);
#undef CALL
    return 0;
}

extern "C" size_t cdr_reduce_features_workspace_bytes(int T, int d, int k)
{
    const int dpad = (d + 31) / 32 * 32;
    int nchunk, chunk;
    features_split(T, dpad, &nchunk, &chunk);
    const int kp = (k <= 8) ? 8 : (k <= 16) ? 16 : (k <= 24) ? 24 : (k <= 32) ? 32 : (k <= 48) ? 48 : 64;
    const size_t direct = (size_t)nchunk * T * kp * sizeof(double);
    const size_t piped = reduce_features_tma_workspace_bytes(T, d, k);
    const size_t gemm = features64_workspace_bytes(T, d, k);
    const size_t a = direct > piped ? direct : piped;
    return a > gemm ? a : gemm;
}

extern "C" int cdr_reduce_features(const double* M, long ldm, const double* X, long ldx, int T,
                                   int d, int k, double* out, long ldo, void* workspace,
                                   size_t workspace_bytes, const cdr_flags* flags,
                                   cdr_stream_t stream)
{
    CDR_CHECK_ARG(T >= 1 && d >= 1 && k >= 1);
    if (k > CDR_MAX_COMPONENTS) return CDR_ERR_UNSUPPORTED;
    const int dpad = (d + 31) / 32 * 32;
    CDR_CHECK_ARG(ldx >= dpad && ldm >= dpad && ldo >= T && ldx % 2 == 0 && ldm % 2 == 0);
    cudaStream_t s = (cudaStream_t)stream;
    {
        // 16 < k <= 64: the tensor-bound GEMM-shaped kernel
        const int rc = run_reduce_features64(M, ldm, X, ldx, T, d, k, out, ldo, workspace,
                                             workspace_bytes, flags, s);
        if (rc != CDR_TMA_NOT_APPLICABLE) return rc;
    }
    {
        const int rc = run_reduce_features_tma(M, ldm, X, ldx, T, d, k, out, ldo, workspace,
                                               workspace_bytes, flags, s);
        if (rc != CDR_TMA_NOT_APPLICABLE) return rc;
    }
#define CALL(KT) return run_reduce_features<KT>(M, ldm, X, ldx, T, d, k, out, ldo, workspace, workspace_bytes, flags, s)
    CDR_DISPATCH_KT(k, CALL);
#undef CALL
    return 0;
}

// Host-only: which kernels and tile geometry the two passes would use for a shape (no GPU
// needed; used by the CPU-side tests).  out: 12 ints = {samples: strips?, TC, nstrips, stages,
// smem bytes; features: strips?, TC, nstrips, stages, smem bytes; direct-kernel sample splits;
// direct-kernel feature chunks}.
extern "C" int cdr_debug_stream_plan(int T, int d, int k, int with_epilogue, int* out)
{
    CDR_CHECK_ARG(T >= 1 && d >= 1 && k >= 1 && out != nullptr);
    if (k > CDR_MAX_COMPONENTS) return CDR_ERR_UNSUPPORTED;
    const int dpad = (d + 31) / 32 * 32;
    tma_stream_plan(T, d, k, with_epilogue, out);
    out[10] = samples_nsplit(T, dpad);
    int nchunk, chunk;
    features_split(T, dpad, &nchunk, &chunk);
    out[11] = nchunk;
    return 0;
}

extern "C" size_t cdr_gram_workspace_bytes(int T, int d)
{
    // 64-row slabs use the direct-load kernel; a short last slab (<= 16 rows) may take the
    // strip-owned kernel, whose partials are sized differently
    size_t need = cdr_reduce_features_workspace_bytes(T, d, 64);
    const int tail = T % 64;
    if (tail != 0) {
        const size_t t = cdr_reduce_features_workspace_bytes(T, d, tail);
        if (t > need) need = t;
    }
    return need;
}

// First version: the Gram matrix as ceil(T/64) feature-reductions with M = a
// 64-row slab of X itself (each slab is one DMMA pass over X).
extern "C" int cdr_gram(const double* X, long ldx, int T, int d, double* K, long ldk,
                        void* workspace, size_t workspace_bytes, cdr_stream_t stream)
{
    CDR_CHECK_ARG(T >= 1 && d >= 1 && ldk >= T);
    for (int r0 = 0; r0 < T; r0 += 64) {
        const int rows = (T - r0 < 64) ? (T - r0) : 64;
        int rc = cdr_reduce_features(X + (long)r0 * ldx, ldx, X, ldx, T, d, rows,
                                     K + (long)r0 * ldk, ldk, workspace, workspace_bytes, nullptr,
                                     stream);
        if (rc != 0) return rc;
    }
    return 0;
}

extern "C" size_t cdr_frobenius_workspace_bytes(void) { return kFrobBlocks * sizeof(double); }

extern "C" int cdr_frobenius_sq(const double* X, long ldx, int T, int d, double* out,
                                void* workspace, size_t workspace_bytes, cdr_stream_t stream)
{
    CDR_CHECK_ARG(T >= 1 && d >= 1 && ldx >= d && ldx % 2 == 0);
    if (workspace == nullptr || workspace_bytes < kFrobBlocks * sizeof(double))
        return CDR_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    frobenius_partial_kernel<<<kFrobBlocks, 256, 0, s>>>(X, ldx, T, d, (double*)workspace);
    CDR_RETURN_IF_LAUNCH_FAILED();
    frobenius_final_kernel<<<1, 256, 0, s>>>((double*)workspace, kFrobBlocks, out);
    CDR_RETURN_IF_LAUNCH_FAILED();
    return 0;
}
