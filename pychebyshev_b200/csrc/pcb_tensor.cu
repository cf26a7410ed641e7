// Mode contractions of a C-order node-value tensor held in DEVICE memory.
//
//   pcb_tensor_deriv     one pass of _apply_derivative_passes (reference barycentric.py:982-989):
//                        T <- T x_axis D^T, i.e. out[o, j, i] = sum_k T[o, k, i] * Dm[j, k]
//   pcb_tensor_contract  _slice_tensor (reference _extrude_slice.py:79-92): tensordot of one axis with
//                        a weight vector, out[o, i] = sum_k T[o, k, i] * v[k]
//   pcb_tensor_extrude   _extrude_tensor (_extrude_slice.py:73-76): replicate along a new axis
//
// Bound: HBM.  A pass reads and writes the tensor once (16 B per element) for 2 n flop per element:
// 2 flop/B at n = 16, far below the FP64 roofline's ~5.5 flop/B ridge, so these are plain
// coalesced DFMA kernels -- the FP64 tensor cores would add nothing but a second rounding order.
//
// Rounding contract (parity finding 2 of SURVEY.md: the derivative tensors must be the reference's):
// every output element is ONE sequential fused-multiply-add chain over k = 0..n-1 starting from 0.
// That is exactly what OpenBLAS' dgemm micro-kernels do for the reference's
// `np.moveaxis(T, d, -1) @ D_d.T` on FMA-capable x86 (one accumulator per C element, k ascending),
// so the device tensors are BIT-IDENTICAL to the host recipe (tests/test_device_tensor.py checks
// this against _grid.differentiate_tensor on the 11^5 and 16^6 configs and on the same box against
// the reference itself).  pcb_tensor_contract uses the same chain; the reference's tensordot goes
// through dgemv, whose lane-split partial sums differ in the last bits (<= 1e-15 of the tensor's
// scale, tested).
#include "pcb_common.cuh"

namespace pcb {

constexpr int TEN_THREADS = 128;
constexpr int TEN_MAX_N = 64;

// out[o, j, i] = sum_k src[o, k, i] * M[j * n + k]    (rows = n for deriv; rows = 1 for contract)
// One thread per fibre (o, i); the fibre sits in a conflict-free shared-memory column, M in shared
// memory (broadcast reads).  Adjacent threads walk adjacent i, so for inner >= 32 every load and
// store is coalesced; for the last axis (inner = 1) a warp reads 32 consecutive fibres = one
// contiguous 32 n doubles.
__global__ void __launch_bounds__(TEN_THREADS)
tensor_mode_kernel(const double *__restrict__ src, double *__restrict__ dst, long long outer, int n,
                   long long inner, int rows, const double *__restrict__ M) {
    extern __shared__ __align__(16) double smem[];
    double *m_s = smem;                              // rows * n
    double *fib = smem + rows * n + threadIdx.x;     // n * TEN_THREADS, column per thread
    for (int e = threadIdx.x; e < rows * n; e += blockDim.x) m_s[e] = M[e];
    __syncthreads();
    const long long total = outer * inner;
    for (long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x; f < total;
         f += (long long)gridDim.x * blockDim.x) {
        const long long o = f / inner, i = f - o * inner;
        const double *s = src + o * n * inner + i;
        for (int k = 0; k < n; ++k) fib[k * TEN_THREADS] = s[(long long)k * inner];
        double *d = dst + o * rows * inner + i;
        for (int j = 0; j < rows; ++j) {
            const double *mj = m_s + j * n;
            double acc = 0.0;
            for (int k = 0; k < n; ++k) acc = fma(fib[k * TEN_THREADS], mj[k], acc);  // k ascending
            d[(long long)j * inner] = acc;
        }
    }
}

// dst[o, r, i] = src[o, i] for r < reps
__global__ void __launch_bounds__(256)
tensor_extrude_kernel(const double *__restrict__ src, double *__restrict__ dst, long long outer,
                      int reps, long long inner) {
    const long long total = outer * reps * inner;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const long long i = e % inner;
        const long long o = e / (inner * reps);
        dst[e] = src[o * inner + i];
    }
}

int tensor_mode_launch(int dev, const double *d_src, double *d_dst, long long outer, int n,
                       long long inner, int rows, const double *d_M, cudaStream_t st) {
    PCB_REQUIRE(n >= 1 && n <= TEN_MAX_N, "axis length %d outside [1, %d]", n, TEN_MAX_N);
    const size_t smem = ((size_t)rows * n + (size_t)n * TEN_THREADS) * sizeof(double);
    int sm = 0, optin = 0, cc = 0;
    if (int rc = device_props(dev, &sm, &optin, &cc)) return rc;
    PCB_CUDA(allow_dynamic_smem(tensor_mode_kernel, smem, optin));
    const long long total = outer * inner;
    const long long want = (total + TEN_THREADS - 1) / TEN_THREADS;
    const long long cap = (long long)sm * 16;
    tensor_mode_kernel<<<(int)(want < cap ? want : cap), TEN_THREADS, smem, st>>>(d_src, d_dst, outer, n,
                                                                               inner, rows, d_M);
    g_launches.fetch_add(1);
    PCB_CUDA(cudaGetLastError());
    return PCB_OK;
}

static int split_axis(int D, const int32_t *n, int axis, long long *outer, long long *inner) {
    PCB_REQUIRE(D >= 1 && D <= 64 && n, "bad tensor shape");
    PCB_REQUIRE(axis >= 0 && axis < D, "axis %d outside [0, %d)", axis, D);
    *outer = *inner = 1;
    for (int d = 0; d < D; ++d) {
        PCB_REQUIRE(n[d] >= 1, "n[%d] must be >= 1", d);
        if (d < axis) *outer *= n[d];
        if (d > axis) *inner *= n[d];
    }
    return PCB_OK;
}

// Stream-ordered upload of a small host matrix (freed after the kernel, on the same stream).
struct StagedMatrix {
    double *p = nullptr;
    cudaStream_t st;
    StagedMatrix(const double *host, size_t count, cudaStream_t s) : st(s) {
        if (cudaMallocAsync(&p, count * sizeof(double), s) != cudaSuccess) {
            cudaGetLastError();
            p = nullptr;
            return;
        }
        // pageable source: the runtime stages it before returning, so `host` may die afterwards
        if (cudaMemcpyAsync(p, host, count * sizeof(double), cudaMemcpyHostToDevice, s) != cudaSuccess) {
            cudaGetLastError();
            cudaFreeAsync(p, s);
            p = nullptr;
        }
    }
    ~StagedMatrix() {
        if (p) cudaFreeAsync(p, st);
    }
};

}  // namespace pcb

using namespace pcb;

extern "C" PCB_API int pcb_tensor_deriv(int dev, int D, const int32_t *n, int axis,
                                        const double *dmat_host, const double *d_src, double *d_dst,
                                        void *stream) {
    PCB_REQUIRE(dmat_host && d_src && d_dst, "null argument");
    PCB_REQUIRE(d_src != d_dst, "pcb_tensor_deriv cannot run in place");
    long long outer, inner;
    if (int rc = split_axis(D, n, axis, &outer, &inner)) return rc;
    DeviceGuard guard(dev);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    StagedMatrix M(dmat_host, (size_t)n[axis] * n[axis], st);
    if (!M.p) return fail(PCB_ENOMEM, "cannot stage the differentiation matrix");
    return tensor_mode_launch(dev, d_src, d_dst, outer, n[axis], inner, n[axis], M.p, st);
}

extern "C" PCB_API int pcb_tensor_contract(int dev, int D, const int32_t *n, int axis,
                                           const double *vec_host, const double *d_src,
                                           double *d_dst, void *stream) {
    PCB_REQUIRE(vec_host && d_src && d_dst, "null argument");
    long long outer, inner;
    if (int rc = split_axis(D, n, axis, &outer, &inner)) return rc;
    DeviceGuard guard(dev);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    StagedMatrix M(vec_host, (size_t)n[axis], st);
    if (!M.p) return fail(PCB_ENOMEM, "cannot stage the weight vector");
    return tensor_mode_launch(dev, d_src, d_dst, outer, n[axis], inner, 1, M.p, st);
}

extern "C" PCB_API int pcb_tensor_extrude(int dev, int D, const int32_t *n, int axis, int n_new,
                                          const double *d_src, double *d_dst, void *stream) {
    PCB_REQUIRE(d_src && d_dst && n, "null argument");
    PCB_REQUIRE(D >= 1 && D <= 64 && axis >= 0 && axis <= D, "axis %d outside [0, %d]", axis, D);
    PCB_REQUIRE(n_new >= 1, "n_new must be >= 1");
    long long outer = 1, inner = 1;
    for (int d = 0; d < D; ++d) {
        PCB_REQUIRE(n[d] >= 1, "n[%d] must be >= 1", d);
        if (d < axis) outer *= n[d];
        else inner *= n[d];
    }
    DeviceGuard guard(dev);
    const long long total = outer * n_new * inner;
    const long long want = (total + 255) / 256;
    tensor_extrude_kernel<<<(int)(want < 65535 * 16 ? want : 65535 * 16), 256, 0,
                            static_cast<cudaStream_t>(stream)>>>(d_src, d_dst, outer, n_new, inner);
    g_launches.fetch_add(1);
    PCB_CUDA(cudaGetLastError());
    return PCB_OK;
}
