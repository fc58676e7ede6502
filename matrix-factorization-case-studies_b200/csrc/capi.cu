// Library-level entry points.
#include "cdr_common.cuh"

#include <map>
#include <string>
#include <vector>

extern "C" const char* cdr_version(void) { return "cdr_b200 0.1 (sm_100a)"; }

extern "C" int cdr_device_check(void)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) return (int)e;
    return (prop.major == 10) ? 0 : CDR_ERR_UNSUPPORTED;
}

unsigned long long cdr_g_kernel_launches = 0;

// Number of kernels this library has launched (or recorded into a graph) so far.
extern "C" unsigned long long cdr_launch_count(void) { return cdr_g_kernel_launches; }


// fp64 tensor-pipe throughput probe (the roofline denominator of the tensor-bound shapes; not
// in MEASURED_PEAKS.json): every warp runs 8 independent chains of DMMA.8x8x4 on registers.
namespace cdr {
__global__ void __launch_bounds__(256) dmma_probe_kernel(double* out, int iters, double seed)
{
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
    const double a = seed + threadIdx.x, b = seed * 0.5 + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace cdr

// Launches the probe on `blocks` CTAs of 256 threads; the caller times it.  Flops of one launch:
// 2 * 8*8*4 * 8 chains * iters * 8 warps * blocks.  out: blocks * 256 doubles.
extern "C" int cdr_debug_dmma_probe(double* out, int blocks, int iters, cdr_stream_t stream)
{
    CDR_CHECK_ARG(out != nullptr && blocks >= 1 && iters >= 1);
    cdr::dmma_probe_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(out, iters, 1.0);
    CDR_RETURN_IF_LAUNCH_FAILED();
    return 0;
}


// ---------------------------------------------------------------------- CDR_TIME_LAUNCHES
namespace {
struct LaunchNote {
    std::string site;
    cudaEvent_t ev;
};
std::vector<LaunchNote> g_notes;
}  // namespace

// The event goes to the legacy default stream, which is where the debugging runs launch
// everything (eager, no CUDA graph, torch's default stream).
void cdr_debug_note_launch(const char* file, int line)
{
    const char* base = file;
    for (const char* c = file; *c; ++c)
        if (*c == '/') base = c + 1;
    LaunchNote n;
    n.site = std::string(base) + ":" + std::to_string(line);
    if (cudaEventCreate(&n.ev) != cudaSuccess) return;
    cudaEventRecord(n.ev, 0);
    g_notes.push_back(n);
}

// Prints, per launch site, the number of launches and the mean / total time between the
// previous launch's event and this one's (i.e. the kernel's duration when the stream never
// runs dry), then forgets the events.  skip: leading events to ignore (warm-up).
extern "C" int cdr_debug_timing_report(int skip)
{
    cudaDeviceSynchronize();
    std::map<std::string, std::pair<double, int>> agg;
    std::vector<std::string> order;
    for (size_t i = (size_t)(skip > 0 ? skip : 1); i < g_notes.size(); ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, g_notes[i - 1].ev, g_notes[i].ev) != cudaSuccess) continue;
        auto it = agg.find(g_notes[i].site);
        if (it == agg.end()) {
            agg[g_notes[i].site] = std::make_pair((double)ms, 1);
            order.push_back(g_notes[i].site);
        } else {
            it->second.first += ms;
            it->second.second += 1;
        }
    }
    for (const std::string& site : order) {
        const auto& v = agg[site];
        fprintf(stderr, "[cdr timing] %-28s n=%5d mean=%9.2f us total=%10.2f us\n", site.c_str(),
                v.second, 1e3 * v.first / v.second, 1e3 * v.first);
    }
    for (LaunchNote& n : g_notes) cudaEventDestroy(n.ev);
    g_notes.clear();
    return 0;
}
