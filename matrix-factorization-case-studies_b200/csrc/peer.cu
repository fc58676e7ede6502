// Symmetric-region management (CUDA IPC) and the stand-alone peer-memory collectives.
// The fused reduce-over-samples + all-reduce kernel lives in stream_tma.cu next to the
// kernel it extends.  See peer.cuh for the synchronisation protocol.
#include "peer.cuh"

#include <string.h>

namespace cdr {
namespace peer {

constexpr int kCollThreads = 256;
constexpr int kCollChunk = 2048;        // doubles per ownership chunk (16 KB)
constexpr int kCollMaxGrid = 64;        // plenty for the 2.8 MB k x d buffer; <= CDR_PEER_MAX_CTAS

// ----------------------------------------------------------------------------------------
// all-reduce (sum), in place, two-shot
// ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCollThreads)
peer_allreduce_kernel(cdr_peer_group g, size_t offset, size_t n, const cdr_flags* flags)
{
    if (is_done(flags)) return;
    PeerHeader* mine = header_of(g, g.rank);
    const unsigned long long epoch = mine->coll_seq[blockIdx.x] + 1;

    // every rank's partial is complete (its producer kernel precedes this one in stream order)
    cta_barrier_all_ranks<kWaitStart>(g, epoch);

    const size_t nchunks = (n + kCollChunk - 1) / kCollChunk;
    for (size_t c = (size_t)g.rank + (size_t)g.world * blockIdx.x; c < nchunks;
         c += (size_t)g.world * gridDim.x) {
        const size_t base = c * kCollChunk;
        const int len = (int)((n - base < (size_t)kCollChunk) ? (n - base) : (size_t)kCollChunk);
        for (int i = 2 * threadIdx.x; i < len; i += 2 * kCollThreads) {
            const size_t byte = offset + (base + i) * sizeof(double);
            double2 part[CDR_MAX_PEERS];
#pragma unroll
            for (int r = 0; r < CDR_MAX_PEERS; ++r)
                if (r < g.world)
                    part[r] = ld_sys_d2(reinterpret_cast<const double*>(
                        static_cast<const unsigned char*>(g.region[r]) + byte));
            double2 sum = part[0];
#pragma unroll
            for (int r = 1; r < CDR_MAX_PEERS; ++r)
                if (r < g.world) {
                    sum.x += part[r].x;
                    sum.y += part[r].y;
                }
#pragma unroll
            for (int r = 0; r < CDR_MAX_PEERS; ++r)
                if (r < g.world)
                    *reinterpret_cast<double2*>(static_cast<unsigned char*>(g.region[r]) + byte) = sum;
        }
    }

    // all sums pushed everywhere before any rank's next kernel reads (or rewrites) the buffer
    cta_barrier_all_ranks<kWaitFinish>(g, epoch);
    if (threadIdx.x == 0) mine->coll_seq[blockIdx.x] = epoch;
}

// ----------------------------------------------------------------------------------------
// all-gather of column blocks into a replicated k x ldd matrix (push)
// ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCollThreads)
peer_allgather_columns_kernel(cdr_peer_group g, const double* __restrict__ src, long lds,
                              size_t dst_offset, long ldd, int k, int col0, int ncols,
                              const cdr_flags* flags)
{
    if (is_done(flags)) return;
    PeerHeader* mine = header_of(g, g.rank);
    const unsigned long long epoch = mine->coll_seq[blockIdx.x] + 1;

    // nobody is still reading the previous contents of the destination
    cta_barrier_all_ranks<kWaitStart>(g, epoch);

    const long total = (long)k * ncols;
    for (long idx = (long)blockIdx.x * kCollThreads + threadIdx.x; idx < total;
         idx += (long)gridDim.x * kCollThreads) {
        const int i = (int)(idx / ncols), j = (int)(idx % ncols);
        const double v = src[(long)i * lds + j];
        const size_t byte = dst_offset + ((size_t)i * ldd + col0 + j) * sizeof(double);
#pragma unroll
        for (int r = 0; r < CDR_MAX_PEERS; ++r)
            if (r < g.world)
                *reinterpret_cast<double*>(static_cast<unsigned char*>(g.region[r]) + byte) = v;
    }

    cta_barrier_all_ranks<kWaitFinish>(g, epoch);
    if (threadIdx.x == 0) mine->coll_seq[blockIdx.x] = epoch;
}

__global__ void peer_read_error_kernel(cdr_peer_group g, int* out)
{
    PeerHeader* mine = header_of(g, g.rank);
    *out = atomicExch(&mine->error, 0);
}

static bool group_ok(const cdr_peer_group* g)
{
    if (g == nullptr || g->world < 1 || g->world > CDR_MAX_PEERS) return false;
    if (g->rank < 0 || g->rank >= g->world) return false;
    for (int r = 0; r < g->world; ++r)
        if (g->region[r] == nullptr) return false;
    return g->region_bytes >= CDR_PEER_HEADER_BYTES;
}

}  // namespace peer
}  // namespace cdr

using namespace cdr;
using namespace cdr::peer;

static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");

extern "C" int cdr_peer_region_alloc(size_t bytes, void** region)
{
    CDR_CHECK_ARG(region != nullptr && bytes >= CDR_PEER_HEADER_BYTES);
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset(p, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        cudaFree(p);
        return (int)e;
    }
    *region = p;
    return 0;
}

extern "C" int cdr_peer_region_free(void* region)
{
    if (region == nullptr) return 0;
    return (int)cudaFree(region);
}

extern "C" int cdr_peer_export(void* region, unsigned char handle[64])
{
    CDR_CHECK_ARG(region != nullptr && handle != nullptr);
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, region);
    if (e != cudaSuccess) return (int)e;
    memcpy(handle, &h, sizeof(h));
    return 0;
}

extern "C" int cdr_peer_import(const unsigned char handle[64], void** region)
{
    CDR_CHECK_ARG(region != nullptr && handle != nullptr);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return (int)e;
    *region = p;
    return 0;
}

extern "C" int cdr_peer_release(void* imported_region)
{
    if (imported_region == nullptr) return 0;
    return (int)cudaIpcCloseMemHandle(imported_region);
}

extern "C" int cdr_peer_error(const cdr_peer_group* group, int* error_out, cdr_stream_t stream)
{
    CDR_CHECK_ARG(group_ok(group) && error_out != nullptr);
    cudaStream_t s = (cudaStream_t)stream;
    int* dev = nullptr;
    cudaError_t e = cudaMalloc(&dev, sizeof(int));
    if (e != cudaSuccess) return (int)e;
    peer_read_error_kernel<<<1, 1, 0, s>>>(*group, dev);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(error_out, dev, sizeof(int), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(dev);
    if (e != cudaSuccess) return (int)e;
    ++cdr_g_kernel_launches;
    return 0;
}

extern "C" int cdr_peer_allreduce(const cdr_peer_group* group, size_t offset, size_t n,
                                  const cdr_flags* flags, cdr_stream_t stream)
{
    CDR_CHECK_ARG(group_ok(group) && n >= 2 && n % 2 == 0 && offset % 16 == 0);
    CDR_CHECK_ARG(offset >= CDR_PEER_HEADER_BYTES &&
                  offset + n * sizeof(double) <= group->region_bytes);
    const size_t nchunks = (n + kCollChunk - 1) / kCollChunk;
    // one CTA per chunk this rank may own, computed from sizes only (identical on all ranks)
    size_t grid = (nchunks + group->world - 1) / group->world;
    if (grid > (size_t)kCollMaxGrid) grid = kCollMaxGrid;
    if (grid < 1) grid = 1;
    peer_allreduce_kernel<<<(int)grid, kCollThreads, 0, (cudaStream_t)stream>>>(*group, offset, n,
                                                                              flags);
    CDR_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

extern "C" int cdr_peer_allgather_columns(const cdr_peer_group* group, const double* src, long lds,
                                          size_t dst_offset, long ldd, int k, int col0, int ncols,
                                          int max_cols, const cdr_flags* flags, cdr_stream_t stream)
{
    CDR_CHECK_ARG(group_ok(group) && src != nullptr && k >= 1 && ncols >= 0 && col0 >= 0);
    CDR_CHECK_ARG(max_cols >= ncols && max_cols >= 1 && lds >= ncols && ldd >= col0 + ncols);
    CDR_CHECK_ARG(dst_offset % 8 == 0 && dst_offset >= CDR_PEER_HEADER_BYTES &&
                  dst_offset + (size_t)k * ldd * sizeof(double) <= group->region_bytes);
    long grid = ((long)k * max_cols + 4 * kCollThreads - 1) / (4 * kCollThreads);
    if (grid > kCollMaxGrid) grid = kCollMaxGrid;
    if (grid < 1) grid = 1;
    peer_allgather_columns_kernel<<<(int)grid, kCollThreads, 0, (cudaStream_t)stream>>>(
        *group, src, lds, dst_offset, ldd, k, col0, ncols, flags);
    CDR_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

extern "C" int cdr_reduce_samples_allreduce(const cdr_peer_group* group, const double* Lp, long sLi,
                                            long sLt, const double* X, long ldx, int T, int T_min,
                                            int d, int k, const double* E, size_t out_offset,
                                            long ldo, const cdr_flags* flags, cdr_stream_t stream)
{
    CDR_CHECK_ARG(group_ok(group) && T >= 1 && T_min >= 1 && T_min <= T && d >= 1 && k >= 1);
    if (k > CDR_MAX_COMPONENTS) return CDR_ERR_UNSUPPORTED;
    const int dpad = (d + 31) / 32 * 32;
    CDR_CHECK_ARG(ldx >= dpad && ldo >= dpad && ldx % 2 == 0 && ldo % 2 == 0);
    const size_t bytes = (size_t)k * ldo * sizeof(double);
    CDR_CHECK_ARG(out_offset % 16 == 0 && out_offset >= CDR_PEER_HEADER_BYTES &&
                  out_offset + bytes <= group->region_bytes);
    // world inbox slots + one result slot; a slot holds two alternating sets of tagged words
    // (16 bytes per double): 4 x the bytes of the k x ldo matrix
    CDR_CHECK_ARG(group->inbox_offset % 16 == 0 && group->inbox_slot_bytes % 32 == 0 &&
                  group->inbox_offset >= CDR_PEER_HEADER_BYTES &&
                  group->inbox_offset + (size_t)(group->world + 1) * group->inbox_slot_bytes <=
                      group->region_bytes);
    if (4 * bytes > group->inbox_slot_bytes) return CDR_ERR_NOT_APPLICABLE;
    return run_reduce_samples_exchange(*group, out_offset, Lp, sLi, sLt, X, ldx, T, T_min, d, k, E,
                                       ldo, flags, (cudaStream_t)stream);
}
