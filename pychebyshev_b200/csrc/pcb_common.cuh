// Shared host/device helpers of libpcb_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pcb_b200.h"

namespace pcb {

// ---- error plumbing -------------------------------------------------------------------------
extern thread_local std::string g_last_error;
extern std::atomic<int64_t> g_launches;

int fail(int code, const char *fmt, ...);

#define PCB_CUDA(expr)                                                                      \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess)                                                              \
            return pcb::fail(PCB_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                             __FILE__, __LINE__);                                           \
    } while (0)

#define PCB_REQUIRE(cond, ...)                                   \
    do {                                                         \
        if (!(cond)) return pcb::fail(PCB_EINVAL, __VA_ARGS__);  \
    } while (0)

// ---- plans ------------------------------------------------------------------------------------
enum PlanKind : int { PLAN_TT = 0x7454, PLAN_FULL = 0x4655, PLAN_SPLINE = 0x5350, PLAN_SLIDER = 0x534c };

struct PlanBase {
    int kind;
    int dev;
    int sm_count;
    int smem_optin;
    virtual ~PlanBase() {}
};

int device_props(int dev, int *sm_count, int *smem_optin, int *cc);

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// pcb_tensor.cu: out[o, j, i] = sum_k src[o, k, i] * M[j * n + k] (sequential FMA chain over k),
// src (outer, n, inner) -> dst (outer, rows, inner), everything in device memory
int tensor_mode_launch(int dev, const double *d_src, double *d_dst, long long outer, int n,
                       long long inner, int rows, const double *d_M, cudaStream_t st);

// Raise a kernel's dynamic shared-memory limit.  The attribute is per-function state shared by all
// host threads, so it is always set to the same value (the device's opt-in maximum): a per-launch
// value would race with a concurrent launch of the same kernel that needs more.
template <typename K>
inline cudaError_t allow_dynamic_smem(K kernel, size_t smem, int smem_optin) {
    if (smem <= 40 * 1024) return cudaSuccess;  // safely within the default 48 KB incl. static
    cudaFuncAttributes attr;
    const cudaError_t e = cudaFuncGetAttributes(&attr, reinterpret_cast<const void *>(kernel));
    if (e != cudaSuccess) return e;
    if (smem + attr.sharedSizeBytes <= 48 * 1024) return cudaSuccess;  // static + dynamic fit
    // the opt-in maximum covers static + dynamic shared memory
    return cudaFuncSetAttribute(reinterpret_cast<const void *>(kernel),
                                cudaFuncAttributeMaxDynamicSharedMemorySize,
                                smem_optin - (int)attr.sharedSizeBytes);
}

// RAII device switch: evaluation calls must not change the caller's current device.
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = (cudaSetDevice(dev) == cudaSuccess);
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// ---- mbarrier + TMA bulk copy (cp.async.bulk) -------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

}  // namespace pcb
