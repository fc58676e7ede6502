// The two streaming contractions for 16 < k <= 64 components (BASELINE.json configs[4]: k = 64
// on an 18 000 x 44 000 matrix).  At k = 64 the arithmetic intensity (k / 4 = 16 flop/B) is
// above the ridge of the machine: these shapes are bound by the fp64 tensor pipe, not by HBM,
// and are GEMMs rather than skinny passes.  Both kernels follow the SYRK kernel (syrk.cu):
// 16-byte cp.async (LDGSTS) multi-stage pipelines into padded shared-memory tiles feeding
// mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4), eight warps, the work split so that whole waves of SMs
// are busy, per-split partial tiles summed in fixed order by a second kernel.
//
//   reduce over features   out (k x T) = M X'       CTA tile 64 components x 128 samples,
//                          reduction over d: both operands are row blocks with the reduction
//                          index contiguous -- exactly the SYRK inner loop with M as the left
//                          operand;
//   reduce over samples    out (k x d) = E (L X)    CTA tile 64 components x 128 features,
//                          reduction over T: X enters as B[t][f] (row = reduction index), L
//                          either row-major k x T (the AA dictionary, direction) or as the
//                          T x k weights (L[i][t] = Z[t][i]).
//
// Replaces, for these k, the direct-load kernels of stream_gemm.cu (0.40-0.42 of the tensor
// peak at k = 64).
#include "cdr_common.cuh"

namespace cdr {

constexpr int kG64Threads = 256;
constexpr int kG64KC = 16;                  // reduction-index elements per stage
constexpr int kG64Cols = 128;               // samples / features per CTA tile

__device__ __forceinline__ void g64_cp16(void* smem, const void* gmem)
{
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}

__device__ __forceinline__ void g64_cp8(void* smem, const void* gmem)
{
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem) : "memory");
}

__device__ __forceinline__ void g64_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }

template <int N>
__device__ __forceinline__ void g64_wait()
{
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// ======================================================================
// reduce over features: part[item][64][128] = M[:, chunk range] X[t tile, chunk range]'
// ======================================================================
constexpr int kF64RS = kG64KC + 8;          // == 8 mod 16: conflict-free 16-byte fragment reads
constexpr int kF64Stages = 5;
constexpr size_t kF64StageDoubles = (size_t)(64 + kG64Cols) * kF64RS;
constexpr size_t kF64Smem = kF64Stages * kF64StageDoubles * sizeof(double);      // 180 KB
static_assert(kF64Smem <= 227 * 1024, "feature pipeline does not fit");

__global__ void __launch_bounds__(kG64Threads, 1)
features64_kernel(const double* __restrict__ M, long ldm, const double* __restrict__ X, long ldx, int T,
                  int k, int dpad, int nsplit, int chunks_per_split, double* __restrict__ part,
                  const cdr_flags* flags)
{
    if (is_done(flags)) return;
    extern __shared__ __align__(16) double gsm[];
    const int tile = blockIdx.x / nsplit, sp = blockIdx.x % nsplit;
    const int nchunks_total = dpad / kG64KC;
    const int c_begin = sp * chunks_per_split;
    int c_end = c_begin + chunks_per_split;
    if (c_end > nchunks_total) c_end = nchunks_total;
    const int nchunks = c_end > c_begin ? c_end - c_begin : 0;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lr = lane & 3, lc = lane >> 2;
    const int wm = warp >> 2, wn = warp & 3;          // warp tile: rows wm*32, columns wn*32

    // loaders: 8 threads cover the 128 bytes of a row segment, 32 rows per pass
    const int lrow = tid >> 3, lchunk = tid & 7;
    const double* srcA[2];
    const double* srcB[4];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        int r = p * 32 + lrow;
        if (r >= k) r = k - 1;                         // clamped rows: results are discarded
        srcA[p] = M + (long)r * ldm + (long)c_begin * kG64KC + lchunk * 2;
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        int r = tile * kG64Cols + p * 32 + lrow;
        if (r >= T) r = T - 1;
        srcB[p] = X + (long)r * ldx + (long)c_begin * kG64KC + lchunk * 2;
    }
    auto load_stage = [&](int chunk, int stage) {
        double* a = gsm + (size_t)stage * kF64StageDoubles;
        double* b = a + 64 * kF64RS;
        const long off = (long)chunk * kG64KC;
#pragma unroll
        for (int p = 0; p < 2; ++p) g64_cp16(a + (p * 32 + lrow) * kF64RS + lchunk * 2, srcA[p] + off);
#pragma unroll
        for (int p = 0; p < 4; ++p) g64_cp16(b + (p * 32 + lrow) * kF64RS + lchunk * 2, srcB[p] + off);
    };

    double acc[4][4][2];
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int n = 0; n < 4; ++n) acc[m][n][0] = acc[m][n][1] = 0.0;

#pragma unroll
    for (int s = 0; s < kF64Stages - 1; ++s) {
        if (s < nchunks) load_stage(s, s);
        g64_commit();
    }
    for (int c = 0; c < nchunks; ++c) {
        g64_wait<kF64Stages - 2>();
        __syncthreads();
        {
            const int nc = c + kF64Stages - 1;
            if (nc < nchunks) load_stage(nc, nc % kF64Stages);
            g64_commit();
        }
        const double* a = gsm + (size_t)(c % kF64Stages) * kF64StageDoubles;
        const double* b = a + 64 * kF64RS;
        const double* ap = a + (wm * 32 + lc) * kF64RS + 2 * lr;
        const double* bp = b + (wn * 32 + lc) * kF64RS + 2 * lr;
#pragma unroll
        for (int k8 = 0; k8 < kG64KC / 8; ++k8) {
            double2 af[4], bf[4];
#pragma unroll
            for (int m = 0; m < 4; ++m)
                af[m] = *reinterpret_cast<const double2*>(ap + m * 8 * kF64RS + k8 * 8);
#pragma unroll
            for (int n = 0; n < 4; ++n)
                bf[n] = *reinterpret_cast<const double2*>(bp + n * 8 * kF64RS + k8 * 8);
#pragma unroll
            for (int m = 0; m < 4; ++m)
#pragma unroll
                for (int n = 0; n < 4; ++n) {
                    dmma884(acc[m][n][0], acc[m][n][1], af[m].x, bf[n].x);
                    dmma884(acc[m][n][0], acc[m][n][1], af[m].y, bf[n].y);
                }
        }
    }
    g64_wait<0>();

    double* dst = part + (size_t)blockIdx.x * 64 * kG64Cols;
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            const int row = wm * 32 + m * 8 + lc, col = wn * 32 + n * 8 + 2 * lr;
            *reinterpret_cast<double2*>(dst + row * kG64Cols + col) =
                make_double2(acc[m][n][0], acc[m][n][1]);
        }
}

// out[i][t] = sum over splits (fixed order) of part[tile * nsplit + s][i][t - tile * 128]
__global__ void __launch_bounds__(256)
features64_reduce_kernel(const double* __restrict__ part, int T, int k, int nsplit,
                         double* __restrict__ out, long ldo, const cdr_flags* flags)
{
    if (is_done(flags)) return;
    const int tile = blockIdx.x;
    const double* base = part + (size_t)tile * nsplit * 64 * kG64Cols;
    for (int e = threadIdx.x; e < 64 * kG64Cols; e += blockDim.x) {
        const int i = e / kG64Cols, c = e % kG64Cols;
        const int t = tile * kG64Cols + c;
        if (i >= k || t >= T) continue;
        double s = 0.0;
        for (int q = 0; q < nsplit; ++q) s += base[(size_t)q * 64 * kG64Cols + e];
        out[(long)i * ldo + t] = s;
    }
}

// ======================================================================
// reduce over samples: part[item][64][128] = L[:, t range] X[t range, f tile]
// ======================================================================
constexpr int kS64RSX = kG64Cols + 4;       // X tile [t][f]: == 4 mod 16 (double2 reads at 2*lc)
constexpr int kS64RSA = kG64KC + 4;         // L tile [i][t] (row-major L): == 4 mod 16
constexpr int kS64RSZ = 64 + 4;             // L tile [t][i] (weights layout): == 4 mod 16
constexpr int kS64Stages = 6;
constexpr size_t kS64StageDoubles = (size_t)kG64KC * kS64RSX + (size_t)64 * kS64RSA;   // >= 16 * 68
constexpr size_t kS64Smem = kS64Stages * kS64StageDoubles * sizeof(double);             // 160 KB
static_assert((size_t)kG64KC * kS64RSZ <= (size_t)64 * kS64RSA, "weights tile must fit the L slot");
static_assert(kS64Smem <= 227 * 1024, "sample pipeline does not fit");

// ZLAYOUT: L[i][t] = Lp[t * k + i] (the T x k weights); otherwise L[i][t] = Lp[i * sLi + t]
template <bool ZLAYOUT>
__global__ void __launch_bounds__(kG64Threads, 1)
samples64_kernel(const double* __restrict__ Lp, long sLi, const double* __restrict__ X, long ldx, int T,
                 int k, int dpad, int nsplit, int rows_per_split, double* __restrict__ part,
                 const cdr_flags* flags)
{
    if (is_done(flags)) return;
    extern __shared__ __align__(16) double gsm[];
    const int strip = blockIdx.x / nsplit, sp = blockIdx.x % nsplit;
    const int f0 = strip * kG64Cols;
    const int t_begin = sp * rows_per_split;
    int t_end = t_begin + rows_per_split;
    if (t_end > T) t_end = T;
    const int nrows = t_end > t_begin ? t_end - t_begin : 0;
    const int nchunks = (nrows + kG64KC - 1) / kG64KC;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lr = lane & 3, lc = lane >> 2;
    const int wm = warp >> 2, wn = warp & 3;          // warp tile: rows wm*32, features wn*32

    auto load_stage = [&](int chunk, int stage) {
        double* xs = gsm + (size_t)stage * kS64StageDoubles;
        double* ls = xs + kG64KC * kS64RSX;
        const int t0 = t_begin + chunk * kG64KC;
        // X tile: 16 rows x 128 features = 1024 16-byte pieces, 4 per thread; rows beyond the
        // range are clamped and their L entries zeroed below
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int piece = p * kG64Threads + tid;
            const int r = piece >> 6, cpos = (piece & 63) * 2;
            int t = t0 + r;
            if (t >= t_end) t = t_end - 1;
            int f = f0 + cpos;
            if (f >= dpad) f = dpad - 2;               // clamped columns: results are discarded
            g64_cp16(xs + r * kS64RSX + cpos, X + (long)t * ldx + f);
        }
        if (ZLAYOUT) {
            // weights rows t0 .. t0+16, k doubles each: 16-byte pieces when k is even (sLi is
            // abused as that flag here: the row stride of this layout is always 1), 8-byte
            // pieces otherwise
            if (sLi == 2) {
                for (int e = tid; e < kG64KC * 32; e += kG64Threads) {
                    const int r = e >> 5, i = (e & 31) * 2;
                    const int t = t0 + r;
                    if (i < k && t < t_end) g64_cp16(ls + r * kS64RSZ + i, Lp + (long)t * k + i);
                }
            } else {
                for (int e = tid; e < kG64KC * 64; e += kG64Threads) {
                    const int r = e >> 6, i = e & 63;
                    const int t = t0 + r;
                    if (i < k && t < t_end) g64_cp8(ls + r * kS64RSZ + i, Lp + (long)t * k + i);
                }
            }
        } else {
            // rows i of L, 16 doubles (128 bytes) of the t axis each: 512 16-byte pieces
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const int piece = p * kG64Threads + tid;
                const int i = piece >> 3, cpos = (piece & 7) * 2;
                // (the pair may straddle the end of the range or of the row: rows are padded to
                // an even length and entries beyond t_end are masked when they are used)
                if (i < k && t0 + cpos < T)
                    g64_cp16(ls + i * kS64RSA + cpos, Lp + (long)i * sLi + t0 + cpos);
            }
        }
    };
    // entries of the L tile that are never loaded (i >= k, t beyond the range) must be zero:
    // clear every stage once
    for (size_t e = tid; e < kS64Stages * kS64StageDoubles; e += kG64Threads) gsm[e] = 0.0;
    __syncthreads();

    double acc[4][2][2][2];                             // [m tile][unit][.x/.y column][c0/c1]
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int u = 0; u < 2; ++u)
            acc[m][u][0][0] = acc[m][u][0][1] = acc[m][u][1][0] = acc[m][u][1][1] = 0.0;

#pragma unroll
    for (int s = 0; s < kS64Stages - 1; ++s) {
        if (s < nchunks) load_stage(s, s);
        g64_commit();
    }
    for (int c = 0; c < nchunks; ++c) {
        g64_wait<kS64Stages - 2>();
        __syncthreads();
        {
            const int nc = c + kS64Stages - 1;
            if (nc < nchunks) load_stage(nc, nc % kS64Stages);
            g64_commit();
        }
        const double* xs = gsm + (size_t)(c % kS64Stages) * kS64StageDoubles;
        const double* ls = xs + kG64KC * kS64RSX;
        const int t0 = t_begin + c * kG64KC;
#pragma unroll
        for (int ks = 0; ks < kG64KC / 4; ++ks) {
            const int tl = ks * 4 + lr;
            const bool rowok = t0 + tl < t_end;
            double a[4];
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const int i = wm * 32 + m * 8 + lc;
                const double v = ZLAYOUT ? ls[tl * kS64RSZ + i] : ls[i * kS64RSA + tl];
                a[m] = rowok ? v : 0.0;                 // a stale row of a recycled stage
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const double2 xv = *reinterpret_cast<const double2*>(
                    xs + tl * kS64RSX + wn * 32 + u * 16 + 2 * lc);
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    dmma884(acc[m][u][0][0], acc[m][u][0][1], a[m], xv.x);
                    dmma884(acc[m][u][1][0], acc[m][u][1][1], a[m], xv.y);
                }
            }
        }
    }
    g64_wait<0>();

    // lane holds, for row i = wm*32 + m*8 + lc, features wn*32 + u*16 + 4*lr + {0,1,2,3}
    //   = (acc[..][0][0], acc[..][1][0], acc[..][0][1], acc[..][1][1])
    double* dst = part + (size_t)blockIdx.x * 64 * kG64Cols;
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int row = wm * 32 + m * 8 + lc, col = wn * 32 + u * 16 + 4 * lr;
            double* p = dst + row * kG64Cols + col;
            *reinterpret_cast<double2*>(p) = make_double2(acc[m][u][0][0], acc[m][u][1][0]);
            *reinterpret_cast<double2*>(p + 2) = make_double2(acc[m][u][0][1], acc[m][u][1][1]);
        }
}

// out[j][f] = sum_i E[j][i] * (sum over splits of part[strip * nsplit + s][i][f - strip * 128])
// (E = nullptr: identity).  blockIdx.x: strip, blockIdx.y: block of 16 features of the strip.
constexpr int kS64RedCols = 16;

__global__ void __launch_bounds__(256)
samples64_reduce_kernel(const double* __restrict__ part, int d, int k, int nsplit,
                        const double* __restrict__ E, double* __restrict__ out, long ldo,
                        const cdr_flags* flags)
{
    if (is_done(flags)) return;
    __shared__ double tile[64][kS64RedCols + 1];
    __shared__ double Es[64 * 64];
    const int strip = blockIdx.x, c0 = blockIdx.y * kS64RedCols;
    const int dpad = (d + 31) / 32 * 32;
    const double* base = part + (size_t)strip * nsplit * 64 * kG64Cols;
    for (int e = threadIdx.x; e < 64 * kS64RedCols; e += blockDim.x) {
        const int i = e / kS64RedCols, c = e % kS64RedCols;
        double s = 0.0;
        if (i < k)
            for (int q = 0; q < nsplit; ++q)
                s += base[(size_t)q * 64 * kG64Cols + i * kG64Cols + c0 + c];
        tile[i][c] = s;
    }
    if (E != nullptr)
        for (int e = threadIdx.x; e < k * k; e += blockDim.x) Es[e] = E[e];
    __syncthreads();
    for (int e = threadIdx.x; e < k * kS64RedCols; e += blockDim.x) {
        const int j = e / kS64RedCols, c = e % kS64RedCols;
        const int f = strip * kG64Cols + c0 + c;
        if (f >= dpad) continue;
        double v;
        if (E == nullptr) {
            v = tile[j][c];
        } else {
            v = 0.0;
            for (int i = 0; i < k; ++i) v = fma(Es[j * k + i], tile[i][c], v);
        }
        out[(long)j * ldo + f] = v;
    }
}

// ---------------------------------------------------------------------- host side
static int g64_sm_count()
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

// smallest split count (<= 32) whose (tiles x splits) wastes the fewest SM slots in its last
// wave, every split keeping at least `min_units` units of the reduction axis
static int g64_splits(int ntiles, int units, int min_units)
{
    const int nsm = g64_sm_count();
    int best = 1;
    double best_eff = 0.0;
    for (int s = 1; s <= 32; ++s) {
        if (s > 1 && units / s < min_units) break;
        const long items = (long)ntiles * s;
        const long waves = (items + nsm - 1) / nsm;
        const double eff = (double)items / (double)(waves * nsm);
        if (eff > best_eff + 0.02) {
            best_eff = eff;
            best = s;
        }
    }
    return best;
}

static bool g64_disabled()
{
    const char* e = getenv("CDR_DISABLE_GEMM64");
    return e != nullptr && e[0] == '1';
}

// shapes taken by these kernels: 16 < k <= 64, large enough to fill the machine
bool gemm64_applicable(int T, int d, int k)
{
    return !g64_disabled() && k > 16 && k <= 64 && T >= 256 && d >= 2048;
}

size_t features64_workspace_bytes(int T, int d, int k)
{
    if (!gemm64_applicable(T, d, k)) return 0;
    const int ntiles = (T + kG64Cols - 1) / kG64Cols;
    const int nchunks = ((d + 31) / 32 * 32) / kG64KC;
    return (size_t)ntiles * g64_splits(ntiles, nchunks, 32) * 64 * kG64Cols * sizeof(double);
}

size_t samples64_workspace_bytes(int T, int d, int k)
{
    if (!gemm64_applicable(T, d, k)) return 0;
    const int dpad = (d + 31) / 32 * 32;
    const int nstrips = (dpad + kG64Cols - 1) / kG64Cols;
    return (size_t)nstrips * g64_splits(nstrips, T / kG64KC, 8) * 64 * kG64Cols * sizeof(double);
}

int run_reduce_features64(const double* M, long ldm, const double* X, long ldx, int T, int d, int k,
                          double* out, long ldo, void* workspace, size_t workspace_bytes,
                          const cdr_flags* flags, cudaStream_t stream)
{
    if (!gemm64_applicable(T, d, k)) return CDR_TMA_NOT_APPLICABLE;
    if ((ldx % 2) != 0 || (ldm % 2) != 0 || (((uintptr_t)X) & 15) != 0 || (((uintptr_t)M) & 15) != 0)
        return CDR_TMA_NOT_APPLICABLE;
    if (workspace == nullptr || workspace_bytes < features64_workspace_bytes(T, d, k))
        return CDR_ERR_WORKSPACE;
    const int dpad = (d + 31) / 32 * 32;
    const int ntiles = (T + kG64Cols - 1) / kG64Cols;
    const int nchunks = dpad / kG64KC;
    const int nsplit = g64_splits(ntiles, nchunks, 32);
    const int per = (nchunks + nsplit - 1) / nsplit;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(features64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)kF64Smem);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    features64_kernel<<<ntiles * nsplit, kG64Threads, kF64Smem, stream>>>(
        M, ldm, X, ldx, T, k, dpad, nsplit, per, (double*)workspace, flags);
    CDR_RETURN_IF_LAUNCH_FAILED();
    features64_reduce_kernel<<<ntiles, 256, 0, stream>>>((const double*)workspace, T, k, nsplit, out, ldo,
                                                        flags);
    CDR_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

int run_reduce_samples64(const double* Lp, long sLi, long sLt, const double* X, long ldx, int T, int d,
                         int k, const double* E, double* out, long ldo, void* workspace,
                         size_t workspace_bytes, const cdr_flags* flags, cudaStream_t stream)
{
    if (!gemm64_applicable(T, d, k)) return CDR_TMA_NOT_APPLICABLE;
    const bool zlayout = (sLi == 1 && sLt == k);
    const bool rowmajor = (sLt == 1 && sLi >= T && sLi % 2 == 0 && (((uintptr_t)Lp) & 15) == 0);
    if (!zlayout && !rowmajor) return CDR_TMA_NOT_APPLICABLE;
    if ((ldx % 2) != 0 || (((uintptr_t)X) & 15) != 0) return CDR_TMA_NOT_APPLICABLE;
    if (workspace == nullptr || workspace_bytes < samples64_workspace_bytes(T, d, k))
        return CDR_ERR_WORKSPACE;
    const int dpad = (d + 31) / 32 * 32;
    const int nstrips = (dpad + kG64Cols - 1) / kG64Cols;
    const int nsplit = g64_splits(nstrips, T / kG64KC, 8);
    int per = (T + nsplit - 1) / nsplit;
    per = (per + kG64KC - 1) / kG64KC * kG64KC;        // splits start on even rows (16-byte pieces of L)
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(samples64_kernel<true>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kS64Smem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(samples64_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)kS64Smem);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    if (zlayout)
        samples64_kernel<true><<<nstrips * nsplit, kG64Threads, kS64Smem, stream>>>(
            Lp, (k % 2 == 0 && (((uintptr_t)Lp) & 15) == 0) ? 2 : 1, X, ldx, T, k, dpad, nsplit, per,
            (double*)workspace, flags);
    else
        samples64_kernel<false><<<nstrips * nsplit, kG64Threads, kS64Smem, stream>>>(
            Lp, sLi, X, ldx, T, k, dpad, nsplit, per, (double*)workspace, flags);
    CDR_RETURN_IF_LAUNCH_FAILED();
    samples64_reduce_kernel<<<dim3(nstrips, kG64Cols / kS64RedCols), 256, 0, stream>>>(
        (const double*)workspace, d, k, nsplit, E, out, ldo, flags);
    CDR_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

}  // namespace cdr
