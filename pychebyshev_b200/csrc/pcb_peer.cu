// Peer-memory result gather for query-sharded evaluation ("optional result gather").
//
// The data path has no collective: every rank evaluates its own row shard.  When the caller wants
// every rank to hold ALL results, the usual recipe is kernel -> ncclAllGather, serialised.  On an
// NVSwitch box that wastes the copy engines and the NVLink ports while the SMs are busy with FP64
// work, so this file offers the B200 form instead:
//
//   * each rank allocates ONE replicated result tensor (world x rows x G doubles) with cudaMalloc
//     and exports it as a CUDA IPC handle (pcb_peer_alloc);
//   * every rank maps the peers' tensors into its own address space (pcb_peer_open -- peer access
//     over NVLink is enabled by the mapping);
//   * the evaluation kernels write this rank's slice of ITS OWN replica directly (no staging
//     buffer), chunk by chunk; after each chunk pcb_peer_push copies that chunk into the same slice
//     of every peer's replica with the copy engines, one internal stream per peer so all NVLink
//     ports are driven at once, while the SMs go on with the next chunk; pcb_peer_join makes a
//     stream wait for every outstanding push (before the slice is overwritten or read remotely).
//
// Bound: NVLink egress, (world-1) x 8 G bytes per query per rank -- for C2 at 8 GPUs 224 B/query x
// 1.55e9 q/s = 347 GB/s of a 900 GB/s port budget, i.e. hidden behind the FP64-bound kernel.
// There are no kernels in this file: copy engines only, the SMs belong to the evaluators.
#include "pcb_common.cuh"

#include <mutex>

namespace pcb {

constexpr int PEER_MAX = 16;

struct PushStreams {
    cudaStream_t s[PEER_MAX] = {};
    int n = 0;
};

static std::mutex g_peer_mu;
static PushStreams g_push[64];

static int push_streams(int dev, int n, PushStreams **out) {
    PCB_REQUIRE(dev >= 0 && dev < 64, "device %d out of range", dev);
    std::lock_guard<std::mutex> lock(g_peer_mu);
    PushStreams &p = g_push[dev];
    while (p.n < n) {
        PCB_CUDA(cudaStreamCreateWithFlags(&p.s[p.n], cudaStreamNonBlocking));
        ++p.n;
    }
    *out = &p;
    return PCB_OK;
}

}  // namespace pcb

using namespace pcb;

static_assert(sizeof(cudaIpcMemHandle_t) == PCB_PEER_HANDLE_BYTES, "IPC handle size");

extern "C" PCB_API int pcb_peer_alloc(int dev, uint64_t bytes, void **d_ptr, unsigned char *handle) {
    PCB_REQUIRE(d_ptr && handle && bytes > 0, "null argument or zero size");
    int sm = 0, smem = 0, cc = 0;
    if (int rc = device_props(dev, &sm, &smem, &cc)) return rc;
    DeviceGuard guard(dev);
    void *p = nullptr;
    PCB_CUDA(cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return fail(PCB_ECUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    }
    memcpy(handle, &h, sizeof(h));
    *d_ptr = p;
    return PCB_OK;
}

extern "C" PCB_API int pcb_peer_free(int dev, void *d_ptr) {
    if (!d_ptr) return PCB_OK;
    DeviceGuard guard(dev);
    PCB_CUDA(cudaFree(d_ptr));
    return PCB_OK;
}

extern "C" PCB_API int pcb_peer_open(int dev, const unsigned char *handle, void **d_ptr) {
    PCB_REQUIRE(handle && d_ptr, "null argument");
    DeviceGuard guard(dev);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    PCB_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return PCB_OK;
}

extern "C" PCB_API int pcb_peer_close(int dev, void *d_ptr) {
    if (!d_ptr) return PCB_OK;
    DeviceGuard guard(dev);
    PCB_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return PCB_OK;
}

extern "C" PCB_API int pcb_peer_push(int dev, int n_peers, void *const *peer_ptrs, uint64_t offset_bytes,
                                     const void *d_src, uint64_t bytes, void *stream) {
    PCB_REQUIRE(n_peers >= 0 && n_peers <= PEER_MAX, "n_peers %d outside [0, %d]", n_peers, PEER_MAX);
    if (n_peers == 0 || bytes == 0) return PCB_OK;
    PCB_REQUIRE(peer_ptrs && d_src, "null argument");
    for (int i = 0; i < n_peers; ++i) PCB_REQUIRE(peer_ptrs[i], "peer pointer %d is null", i);
    DeviceGuard guard(dev);
    PushStreams *ps = nullptr;
    if (int rc = push_streams(dev, n_peers, &ps)) return rc;
    // fork: the copies start once everything already enqueued on `stream` (the kernel that made
    // d_src) is done; `stream` itself does not wait -- the next chunk's kernel runs under them
    cudaEvent_t ready;
    PCB_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    cudaError_t e = cudaEventRecord(ready, static_cast<cudaStream_t>(stream));
    for (int i = 0; i < n_peers && e == cudaSuccess; ++i) {
        e = cudaStreamWaitEvent(ps->s[i], ready, 0);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(static_cast<char *>(peer_ptrs[i]) + offset_bytes, d_src, bytes,
                                cudaMemcpyDefault, ps->s[i]);
    }
    cudaEventDestroy(ready);
    if (e != cudaSuccess) return fail(PCB_ECUDA, "peer push failed: %s", cudaGetErrorString(e));
    return PCB_OK;
}

extern "C" PCB_API int pcb_peer_join(int dev, void *stream) {
    DeviceGuard guard(dev);
    PushStreams *ps = nullptr;
    if (int rc = push_streams(dev, 0, &ps)) return rc;
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < ps->n && e == cudaSuccess; ++i) {
        cudaEvent_t done;
        e = cudaEventCreateWithFlags(&done, cudaEventDisableTiming);
        if (e != cudaSuccess) break;
        e = cudaEventRecord(done, ps->s[i]);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), done, 0);
        cudaEventDestroy(done);
    }
    if (e != cudaSuccess) return fail(PCB_ECUDA, "peer join failed: %s", cudaGetErrorString(e));
    return PCB_OK;
}
