// C-ABI plumbing shared by all plans + the FP64 roofline probes.
#include "pcb_cbank.cuh"

namespace pcb {

thread_local std::string g_last_error;
std::atomic<int64_t> g_launches{0};

uint64_t next_plan_id() {
    static std::atomic<uint64_t> next{1};
    return next.fetch_add(1);
}

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

int device_props(int dev, int *sm_count, int *smem_optin, int *cc) {
    int ndev = 0;
    PCB_CUDA(cudaGetDeviceCount(&ndev));
    PCB_REQUIRE(dev >= 0 && dev < ndev, "device %d out of range (%d visible)", dev, ndev);
    int major = 0, minor = 0;
    PCB_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
    PCB_CUDA(cudaDeviceGetAttribute(smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    PCB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    PCB_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    *cc = major * 10 + minor;
    // this library ships sm_100a code only: fail loudly anywhere else
    if (*cc != 100)
        return fail(PCB_EUNSUPPORTED, "device %d has compute capability %d.%d; libpcb_b200 is built "
                    "for sm_100a (B200) only", dev, major, minor);
    return PCB_OK;
}

// ---- roofline probes --------------------------------------------------------------------------
// Register-resident dependent-chain kernels: 8 independent chains per thread so the FP64 pipe is
// issue-bound, not latency-bound.  Results are written so the compiler keeps the work.

constexpr int PROBE_ITERS = 4096;

__global__ void __launch_bounds__(256) probe_dfma_kernel(double *out, double seed) {
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + i + threadIdx.x;
    const double m = 1.0000001, c = 1e-9;
    for (int it = 0; it < PROBE_ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma(a[i], m, c);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) probe_dmma_kernel(double *out, double seed) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
    const double a = seed + (threadIdx.x & 31) * 1e-3, b = 1e-3 * (threadIdx.x & 7);
    for (int it = 0; it < PROBE_ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1])
                         : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Both at once: 8 DMMA + 8 DFMA per iteration in every warp.  If the FP64 tensor sub-pipe and the
// FP64 FMA pipe are separate, this runs faster than the sum of the two alone.
__global__ void __launch_bounds__(256) probe_mixed_kernel(double *out, double seed) {
    double c[8][2], a8[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        c[i][0] = c[i][1] = 0.0;
        a8[i] = seed + i + threadIdx.x;
    }
    const double a = seed + (threadIdx.x & 31) * 1e-3, b = 1e-3 * (threadIdx.x & 7);
    const double m = 1.0000001, k = 1e-9;
    for (int it = 0; it < PROBE_ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1])
                         : "d"(a), "d"(b));
            a8[i] = fma(a8[i], m, k);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + a8[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace pcb

using namespace pcb;

extern "C" PCB_API int pcb_version(void) { return PCB_ABI_VERSION; }

extern "C" PCB_API const char *pcb_last_error(void) { return g_last_error.c_str(); }

extern "C" PCB_API int64_t pcb_launch_count(void) { return g_launches.load(); }

extern "C" PCB_API int pcb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" PCB_API int pcb_device_info(int dev, int *sm_count, int *smem_optin, int *cc) {
    PCB_REQUIRE(sm_count && smem_optin && cc, "null argument");
    return device_props(dev, sm_count, smem_optin, cc);
}

extern "C" PCB_API int pcb_plan_destroy(void *plan) {
    if (!plan) return PCB_OK;
    PlanBase *pl = static_cast<PlanBase *>(plan);
    PCB_REQUIRE(pl->kind == PLAN_TT || pl->kind == PLAN_FULL || pl->kind == PLAN_SPLINE ||
                    pl->kind == PLAN_SLIDER, "not a plan");
    DeviceGuard guard(pl->dev);
    delete pl;
    return PCB_OK;
}

extern "C" PCB_API int pcb_probe_fp64_peak(int dev, int kind, double *tflops, double *ms) {
    PCB_REQUIRE(tflops && ms, "null argument");
    PCB_REQUIRE(kind >= 0 && kind <= 2, "probe kind %d unknown", kind);
    int sm = 0, smem = 0, cc = 0;
    if (int rc = device_props(dev, &sm, &smem, &cc)) return rc;
    DeviceGuard guard(dev);
    const int blocks = sm * 8, threads = 256;
    double *buf = nullptr;
    PCB_CUDA(cudaMalloc(&buf, (size_t)blocks * threads * sizeof(double)));
    cudaEvent_t e0, e1;
    PCB_CUDA(cudaEventCreate(&e0));
    PCB_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        if (kind == 0)
            probe_dfma_kernel<<<blocks, threads>>>(buf, 1.0 + rep);
        else if (kind == 1)
            probe_dmma_kernel<<<blocks, threads>>>(buf, 1.0 + rep);
        else
            probe_mixed_kernel<<<blocks, threads>>>(buf, 1.0 + rep);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float t = 0.f;
        cudaEventElapsedTime(&t, e0, e1);
        if (rep >= 2 && t < best) best = t;
    }
    g_launches.fetch_add(6);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    PCB_CUDA(cudaGetLastError());
    const double total_threads = (double)blocks * threads;
    // DFMA: 8 FMA per thread per iteration; DMMA: 8 MMAs per warp per iteration, 8*8*4 FMA each
    const double fma_dfma = total_threads * 8.0 * PROBE_ITERS;
    const double fma_dmma = (total_threads / 32.0) * 8.0 * 256.0 * PROBE_ITERS;
    const double fma = kind == 0 ? fma_dfma : (kind == 1 ? fma_dmma : fma_dfma + fma_dmma);
    *ms = best;
    *tflops = 2.0 * fma / (best * 1e-3) / 1e12;
    return PCB_OK;
}
