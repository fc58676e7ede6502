// FurthestSum initialisation (furthest_sum.py:23-127) and the dissimilarity
// matrix the estimators build from the Gram matrix
// (archetypal_analysis.py:96-100, gpnh_convex_coding.py:58-71).
//
// The reference keeps a Python list of [index, running distance sum] pairs, and
// every pick is `list.sort(key=dist)` (stable) followed by `pop(-1)`.  The pick
// is therefore the candidate with the largest sum, and ties are broken by the
// list order left behind by the previous sorts.  That order is reproduced
// exactly without sorting: after a stable sort the relative order of two
// candidates is the lexicographic order of (sum now, sum at the previous sort,
// ..., original position), so it is enough to keep the history of the sums at
// every sort event and walk it backwards on a tie (ties are rare, the walk is
// almost never taken).  A candidate appended during the replacement steps sits
// at the end of the list, i.e. it compares as +inf for all earlier events.
//
// The running sums are updated with the same += / -= sequence per candidate as
// the reference, so given the same dissimilarity matrix the picks are bit-exact.
#include "cdr_common.cuh"

namespace cdr {

__global__ void __launch_bounds__(256)
dissimilarity_kernel(const double* __restrict__ K, long ldk, int T, double* __restrict__ D, long ldd)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= T) return;
    // (diag[j] - 2 K[i][j]) + diag[i], the evaluation order of
    // np.tile(kd,(n,1)) - 2*kernel + np.tile(kd[:,None],(1,n))
    const double v = (K[(long)j * ldk + j] - 2.0 * K[(long)i * ldk + j]) + K[(long)i * ldk + i];
    D[(long)i * ldd + j] = sqrt(v);
}

struct FsState {
    const double* D;
    long ldd;
    int T;
    double* dist;   // running sums, T
    int* inlist;    // candidate still in the list, T
    int* pos;       // original list position (index; T + n for the n-th appended candidate)
    double* H;      // H[event][candidate]
};

// true if candidate a sits after candidate b in the (virtually) sorted list
__device__ bool fs_after(const FsState& s, int a, int b, int ev)
{
    const double da = s.dist[a], db = s.dist[b];
    if (da != db) return da > db;
    for (int e = ev - 1; e >= 0; --e) {
        const double ha = s.H[(long)e * s.T + a], hb = s.H[(long)e * s.T + b];
        if (ha != hb) return ha > hb;
    }
    return s.pos[a] > s.pos[b];
}

__device__ int fs_pick(const FsState& s, int ev, int* sh_best)
{
    int best = -1;
    for (int j = threadIdx.x; j < s.T; j += blockDim.x) {
        if (s.inlist[j]) {
            s.H[(long)ev * s.T + j] = s.dist[j];
            if (best < 0 || fs_after(s, j, best, ev)) best = j;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int other = __shfl_xor_sync(CDR_FULL_MASK, best, o);
        if (other >= 0 && (best < 0 || fs_after(s, other, best, ev))) best = other;
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh_best[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        int b = -1;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
            const int other = sh_best[w];
            if (other >= 0 && (b < 0 || fs_after(s, other, b, ev))) b = other;
        }
        sh_best[32] = b;
    }
    __syncthreads();
    return sh_best[32];
}

__global__ void __launch_bounds__(1024)
furthest_sum_kernel(const double* __restrict__ D, long ldd, int T, int k, int start,
                    const int64_t* __restrict__ exclude, int n_exclude, int extra_steps,
                    int64_t* selected, double* dist, int* inlist, int* pos, double* H)
{
    __shared__ int sh_best[33];
    FsState s{D, ldd, T, dist, inlist, pos, H};

    // furthest_sum.py:79-100: candidates in index order with their distance to the start
    for (int i = threadIdx.x; i < T; i += blockDim.x) {
        bool ok = i != start;
        for (int e = 0; e < n_exclude && ok; ++e) ok = (exclude[e] != (int64_t)i);
        inlist[i] = ok ? 1 : 0;
        pos[i] = i;
        dist[i] = D[(long)i * ldd + start];
    }
    for (int i = threadIdx.x; i < k; i += blockDim.x) selected[i] = start;
    __syncthreads();

    int ev = 0;
    // furthest_sum.py:102-104
    for (int i = 1; i < k; ++i) {
        const int pick = fs_pick(s, ev, sh_best);
        if (pick < 0) break;                       // cannot happen after host validation
        if (threadIdx.x == 0) {
            selected[i] = pick;
            inlist[pick] = 0;
        }
        __syncthreads();
        for (int j = threadIdx.x; j < T; j += blockDim.x)
            if (inlist[j]) dist[j] += D[(long)pick * ldd + j];
        __syncthreads();
        ++ev;
    }
    // furthest_sum.py:106-125: replacement passes
    for (int step = 0; step < extra_steps; ++step) {
        const int u = step % k;
        const int old = (int)selected[u];
        for (int j = threadIdx.x; j < T; j += blockDim.x)
            if (inlist[j]) dist[j] -= D[(long)j * ldd + old];
        __syncthreads();
        if (threadIdx.x == 0) {
            double qi = 0.0;
            for (int m = 0; m < k; ++m) {
                const int idx = (int)selected[m];
                if (idx != old) qi += D[(long)old * ldd + idx];
            }
            dist[old] = qi;
            inlist[old] = 1;
            pos[old] = T + step;              // appended at the end of the list
        }
        for (int e = threadIdx.x; e < ev; e += blockDim.x) H[(long)e * T + old] = INFINITY;
        __syncthreads();
        const int pick = fs_pick(s, ev, sh_best);
        if (pick < 0) break;
        if (threadIdx.x == 0) {
            selected[u] = pick;
            inlist[pick] = 0;
        }
        __syncthreads();
        for (int j = threadIdx.x; j < T; j += blockDim.x)
            if (inlist[j]) dist[j] += D[(long)pick * ldd + j];
        __syncthreads();
        ++ev;
    }
}

}  // namespace cdr

using namespace cdr;

extern "C" int cdr_dissimilarity_from_gram(const double* K, long ldk, int T, double* D, long ldd,
                                           cdr_stream_t stream)
{
    CDR_CHECK_ARG(T >= 1 && ldk >= T && ldd >= T);
    dim3 grid((T + 255) / 256, T);
    dissimilarity_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(K, ldk, T, D, ldd);
    CDR_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

static size_t fs_events(int k, int extra_steps)
{
    return (size_t)(k > 0 ? k - 1 : 0) + (size_t)(extra_steps > 0 ? extra_steps : 0) + 1;
}

extern "C" size_t cdr_furthest_sum_workspace_bytes(int T, int k, int extra_steps)
{
    const size_t Tp = ((size_t)T + 1) / 2 * 2;
    return Tp * sizeof(double) + 2 * Tp * sizeof(int) + fs_events(k, extra_steps) * (size_t)T * sizeof(double);
}

extern "C" int cdr_furthest_sum(const double* D, long ldd, int T, int k, int start_index,
                                const int64_t* exclude, int n_exclude, int extra_steps,
                                int64_t* selected, void* workspace, size_t workspace_bytes,
                                cdr_stream_t stream)
{
    CDR_CHECK_ARG(T >= 1 && k >= 1 && ldd >= T && start_index >= 0 && start_index < T &&
                  n_exclude >= 0);
    if (workspace == nullptr || workspace_bytes < cdr_furthest_sum_workspace_bytes(T, k, extra_steps))
        return CDR_ERR_WORKSPACE;
    const size_t Tp = ((size_t)T + 1) / 2 * 2;
    double* dist = (double*)workspace;
    int* inlist = (int*)(dist + Tp);
    int* pos = inlist + Tp;
    double* H = (double*)(pos + Tp);
    furthest_sum_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(
        D, ldd, T, k, start_index, exclude, n_exclude, extra_steps < 0 ? 0 : extra_steps, selected,
        dist, inlist, pos, H);
    CDR_RETURN_IF_LAUNCH_FAILED();
    return 0;
}
