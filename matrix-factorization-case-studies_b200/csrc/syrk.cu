// Gram matrix K = X X' (archetypal_analysis.py:1032, the dissimilarities of the FurthestSum
// initialisation :96-100, PCA) as a SYRK: only the tiles on and above the diagonal are
// computed (T^2 d flops instead of 2 T^2 d) and mirrored.
//
// This is the one tensor-bound contraction of the path (arithmetic intensity ~T/8 flop/B).
// fp64 tensor cores are reached through mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4; tcgen05 has no
// f64 kind), so the kernel is a classic multi-stage pipeline around it:
//   * CTA tile 128 x 128, eight warps as 2 (rows) x 4 (columns), warp tile 64 x 32
//     = 8 x 4 DMMA tiles, 64 fp64 accumulators per thread;
//   * both operands are row blocks of X; a stage holds 128 x 16 doubles of each (four stages,
//     192 KB), copied global -> shared with 16-byte cp.async (LDGSTS) -- 128-byte row segments
//     are far too short for bulk copies to pay (profiles/probes/tma_probe.cu) and the pipe
//     only needs ~8 B / cycle / SM; rows are padded to 24 doubles so that the 16-byte fragment
//     reads of a quarter warp fall into distinct banks; diagonal tiles load one operand only;
//   * fragment trick of the streaming kernels: the reduction index of an MMA can be permuted
//     freely, so each lane loads a double2 and feeds .x / .y to two MMAs;
//   * the feature axis is split so that (tiles x splits) fills whole waves of SMs; every
//     split writes its 128 x 128 partial, a second kernel sums the splits in fixed order and
//     writes K[i][j] and K[j][i]  (deterministic, no atomics).
#include "cdr_common.cuh"

namespace cdr {

constexpr int kSyrkTile = 128;
constexpr int kSyrkKC = 16;                   // doubles of the feature axis per stage
constexpr int kSyrkRS = kSyrkKC + 8;          // padded row stride: == 8 mod 16
constexpr int kSyrkStages = 4;
constexpr int kSyrkThreads = 256;
constexpr size_t kSyrkStageDoubles = 2 * (size_t)kSyrkTile * kSyrkRS;
constexpr size_t kSyrkSmem = kSyrkStages * kSyrkStageDoubles * sizeof(double);   // 192 KB
static_assert(kSyrkSmem <= 227 * 1024, "SYRK pipeline does not fit in shared memory");

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem)
{
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }

template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// upper-triangle tile index -> (tile row ti <= tile column tj)
__device__ __forceinline__ void syrk_tile(int tile, int ntile, int* ti, int* tj)
{
    // tiles of the upper triangle in row-major order: row r holds ntile - r tiles
    int r = 0, rem = tile;
    while (rem >= ntile - r) {
        rem -= ntile - r;
        ++r;
    }
    *ti = r;
    *tj = r + rem;
}

__global__ void __launch_bounds__(kSyrkThreads, 1)
syrk_tile_kernel(const double* __restrict__ X, long ldx, int T, int dpad, int ntile, int nsplit,
                 int chunks_per_split, int part_index, int part_count, double* __restrict__ part)
{
    extern __shared__ __align__(16) double ssm[];
    int ti, tj;
    const int sp = blockIdx.x % nsplit;
    syrk_tile((blockIdx.x / nsplit) * part_count + part_index, ntile, &ti, &tj);
    const bool diag = ti == tj;
    const int nchunks_total = dpad / kSyrkKC;
    const int c_begin = sp * chunks_per_split;
    int c_end = c_begin + chunks_per_split;
    if (c_end > nchunks_total) c_end = nchunks_total;
    const int nchunks = c_end > c_begin ? c_end - c_begin : 0;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lr = lane & 3, lc = lane >> 2;
    const int wm = warp >> 2, wn = warp & 3;         // warp tile origin: rows wm*64, columns wn*32

    // loader mapping: 8 threads cover the 128 bytes of a row segment; 32 rows per pass
    constexpr int kPasses = 4;
    const int lrow = tid >> 3, lchunk = tid & 7;
    const double* srcA[kPasses];
    const double* srcB[kPasses];
#pragma unroll
    for (int p = 0; p < kPasses; ++p) {
        int ra = ti * kSyrkTile + p * 32 + lrow;
        int rb = tj * kSyrkTile + p * 32 + lrow;
        if (ra >= T) ra = T - 1;                      // clamped rows: results are discarded
        if (rb >= T) rb = T - 1;
        srcA[p] = X + (long)ra * ldx + (long)c_begin * kSyrkKC + lchunk * 2;
        srcB[p] = X + (long)rb * ldx + (long)c_begin * kSyrkKC + lchunk * 2;
    }
    auto load_stage = [&](int chunk, int stage) {
        double* a = ssm + (size_t)stage * kSyrkStageDoubles;
        double* b = a + (size_t)kSyrkTile * kSyrkRS;
        const long off = (long)chunk * kSyrkKC;
#pragma unroll
        for (int p = 0; p < kPasses; ++p) {
            cp_async16(a + (p * 32 + lrow) * kSyrkRS + lchunk * 2, srcA[p] + off);
            if (!diag) cp_async16(b + (p * 32 + lrow) * kSyrkRS + lchunk * 2, srcB[p] + off);
        }
    };

    double acc[8][4][2];
#pragma unroll
    for (int m = 0; m < 8; ++m)
#pragma unroll
        for (int n = 0; n < 4; ++n) acc[m][n][0] = acc[m][n][1] = 0.0;

    // prologue
#pragma unroll
    for (int s = 0; s < kSyrkStages - 1; ++s) {
        if (s < nchunks) load_stage(s, s);
        cp_async_commit();
    }
    for (int c = 0; c < nchunks; ++c) {
        cp_async_wait<kSyrkStages - 2>();
        __syncthreads();
        // the stage consumed in iteration c - 1 is free now: refill it
        {
            const int nc = c + kSyrkStages - 1;
            if (nc < nchunks) load_stage(nc, nc % kSyrkStages);
            cp_async_commit();
        }
        const double* a = ssm + (size_t)(c % kSyrkStages) * kSyrkStageDoubles;
        const double* b = diag ? a : a + (size_t)kSyrkTile * kSyrkRS;
        const double* ap = a + (wm * 64 + lc) * kSyrkRS + 2 * lr;
        const double* bp = b + (wn * 32 + lc) * kSyrkRS + 2 * lr;
#pragma unroll
        for (int k8 = 0; k8 < kSyrkKC / 8; ++k8) {
            double2 af[8], bf[4];
#pragma unroll
            for (int m = 0; m < 8; ++m)
                af[m] = *reinterpret_cast<const double2*>(ap + m * 8 * kSyrkRS + k8 * 8);
#pragma unroll
            for (int n = 0; n < 4; ++n)
                bf[n] = *reinterpret_cast<const double2*>(bp + n * 8 * kSyrkRS + k8 * 8);
#pragma unroll
            for (int m = 0; m < 8; ++m)
#pragma unroll
                for (int n = 0; n < 4; ++n) {
                    dmma884(acc[m][n][0], acc[m][n][1], af[m].x, bf[n].x);
                    dmma884(acc[m][n][0], acc[m][n][1], af[m].y, bf[n].y);
                }
        }
    }
    cp_async_wait<0>();

    // partial tile of this split: part[item][128][128]
    double* dst = part + (size_t)blockIdx.x * kSyrkTile * kSyrkTile;
#pragma unroll
    for (int m = 0; m < 8; ++m)
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            const int row = wm * 64 + m * 8 + lc, col = wn * 32 + n * 8 + 2 * lr;
            *reinterpret_cast<double2*>(dst + row * kSyrkTile + col) =
                make_double2(acc[m][n][0], acc[m][n][1]);
        }
}

// K[i][j] = K[j][i] = sum over splits (fixed order) of the partial tiles
__global__ void __launch_bounds__(256)
syrk_reduce_kernel(const double* __restrict__ part, int T, int ntile, int nsplit, int part_index,
                   int part_count, double* __restrict__ K, long ldk)
{
    __shared__ double tile[32][33];
    // blockIdx.x: this part's upper-triangle tile, blockIdx.y: 32 x 32 sub-block of the tile
    int ti, tj;
    syrk_tile(blockIdx.x * part_count + part_index, ntile, &ti, &tj);
    const int sb_r = (blockIdx.y / 4) * 32, sb_c = (blockIdx.y % 4) * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
    const double* base = part + (size_t)blockIdx.x * nsplit * kSyrkTile * kSyrkTile;
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
        const int r = sb_r + ty + 8 * rr, c = sb_c + tx;
        double s = 0.0;
        for (int q = 0; q < nsplit; ++q) s += base[(size_t)q * kSyrkTile * kSyrkTile + r * kSyrkTile + c];
        tile[ty + 8 * rr][tx] = s;
        const int gi = ti * kSyrkTile + r, gj = tj * kSyrkTile + c;
        if (gi < T && gj < T) K[(long)gi * ldk + gj] = s;
    }
    if (ti == tj) return;
    __syncthreads();
    // mirrored block, written with coalesced rows
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
        const int c = sb_c + ty + 8 * rr, r = sb_r + tx;          // K[gj][gi]
        const int gi = ti * kSyrkTile + r, gj = tj * kSyrkTile + c;
        if (gi < T && gj < T) K[(long)gj * ldk + gi] = tile[tx][ty + 8 * rr];
    }
}

static int syrk_sm_count()
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

// number of feature splits: the smallest split count (<= 16) whose (tiles x splits) wastes the
// fewest SM slots in its last wave; at least 16 chunks of 16 features per split
static int syrk_splits(int ntri, int nchunks)
{
    const int nsm = syrk_sm_count();
    int best = 1;
    double best_eff = 0.0;
    for (int s = 1; s <= 16; ++s) {
        if (s > 1 && nchunks / s < 16) break;
        const long items = (long)ntri * s;
        const long waves = (items + nsm - 1) / nsm;
        const double eff = (double)items / (double)(waves * nsm);
        if (eff > best_eff + 0.02) {
            best_eff = eff;
            best = s;
        }
    }
    return best;
}

}  // namespace cdr

using namespace cdr;

extern "C" size_t cdr_syrk_workspace_bytes(int T, int d)
{
    if (T < 1 || d < 1) return 0;
    const int ntile = (T + kSyrkTile - 1) / kSyrkTile;
    const int ntri = ntile * (ntile + 1) / 2;
    const int nchunks = ((d + 31) / 32 * 32) / kSyrkKC;
    const int nsplit = syrk_splits(ntri, nchunks);
    return (size_t)ntri * nsplit * kSyrkTile * kSyrkTile * sizeof(double);
}

extern "C" int cdr_syrk(const double* X, long ldx, int T, int d, double* K, long ldk, int part_index,
                        int part_count, void* workspace, size_t workspace_bytes, cdr_stream_t stream)
{
    CDR_CHECK_ARG(X != nullptr && K != nullptr && T >= 1 && d >= 1 && ldk >= T);
    CDR_CHECK_ARG(part_count >= 1 && part_index >= 0 && part_index < part_count);
    const int dpad = (d + 31) / 32 * 32;
    CDR_CHECK_ARG(ldx >= dpad && ldx % 2 == 0 && (((uintptr_t)X) & 15) == 0);
    if (workspace == nullptr || workspace_bytes < cdr_syrk_workspace_bytes(T, d)) return CDR_ERR_WORKSPACE;
    const int ntile = (T + kSyrkTile - 1) / kSyrkTile;
    const int ntri = ntile * (ntile + 1) / 2;
    const int nchunks = dpad / kSyrkKC;
    const int nsplit = syrk_splits(ntri, nchunks);
    const int per = (nchunks + nsplit - 1) / nsplit;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(syrk_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)kSyrkSmem);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    cudaStream_t s = (cudaStream_t)stream;
    // the tiles of this part: part_index, part_index + part_count, ...
    const int mine = (ntri - part_index + part_count - 1) / part_count;
    if (mine <= 0) return 0;
    syrk_tile_kernel<<<mine * nsplit, kSyrkThreads, kSyrkSmem, s>>>(
        X, ldx, T, dpad, ntile, nsplit, per, part_index, part_count, (double*)workspace);
    CDR_RETURN_IF_LAUNCH_FAILED();
    syrk_reduce_kernel<<<dim3(mine, 16), 256, 0, s>>>((const double*)workspace, T, ntile, nsplit,
                                                      part_index, part_count, K, ldk);
    CDR_RETURN_IF_LAUNCH_FAILED();
    return 0;
}
