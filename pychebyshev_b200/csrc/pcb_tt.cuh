// ChebyshevTT batch evaluation on sm_100a: shared definitions of the three TT kernels.
//
// Replaces ChebyshevTT.eval_batch (reference tensor_train.py:2217-2265) and loops of
// ChebyshevTT.eval_multi (tensor_train.py:2267-2463).
//
// Design: one query per THREAD slot (QPT slots per thread), coefficient cores
// resident in shared memory and read with warp-uniform (broadcast) 128-bit LDS, so one LDS feeds
// 2*QPT DFMAs in each of 32 lanes.  Per dimension k a thread updates its chain vector
//     v'[l] = sum_i sum_j (v[i] * T_j(s_k)) * G_k[i, j, l]
// with the accumulators v'[l] in registers (compile-time chunk width W <= LC), c = v[i]*T_j
// advanced by the Chebyshev three-term recurrence in registers, and v[] parked in a conflict-free
// per-thread shared-memory column.  No cross-lane traffic, no atomics, no per-tile barrier: a
// thread reads its own query coordinates straight from global memory (L1-prefetched one tile
// ahead).
#pragma once

#include <map>
#include <mutex>

#include "pcb_common.cuh"

namespace pcb {

constexpr int TT_MAX_ACTIVE = 3;
constexpr int TT_MAX_G = 16;

enum TTMode : int { TT_RESIDENT = 0, TT_STREAM = 1, TT_GLOBAL = 2 };

struct TTParams {
    int D;
    int rmaxp;    // max padded rank (even)
    int total;    // doubles in the packed forward cores
    int totalT;   // doubles in the packed transposed cores (stored right after the forward ones)
    int maxcore;  // largest single packed core (either layout)
    int n[PCB_MAX_DIMS];
    int r[PCB_MAX_DIMS + 1];
    int rp[PCB_MAX_DIMS];    // r[k+1] rounded up to even: row stride of forward core k [i][j][rp]
    int off[PCB_MAX_DIMS];   // offset (doubles) of forward core k
    int rpT[PCB_MAX_DIMS];   // r[k] rounded up to even: row stride of transposed core k [l][j][rpT]
    int offT[PCB_MAX_DIMS];  // offset (doubles, from the buffer start) of transposed core k
    int perm[PCB_MAX_DIMS];  // storage position k -> user column
    int coff[PCB_MAX_DIMS];   // constant bank: offset of unpadded forward core k  [i][j][r_k+1]
    int coffT[PCB_MAX_DIMS];  // constant bank: offset of unpadded transposed core k [l][j][r_k]
    double lo[PCB_MAX_DIMS];
    double hi[PCB_MAX_DIMS];
};

// Uniform-datapath kernels (pcb_tt_const.cu) read the cores of the resident plan from the 64 KB
// __constant__ bank.
constexpr int TT_CONST_MAX = 8192;
constexpr int TT_CONST_MAX_RANK = 16;

// One output row of pcb_tt_eval_fd, storage frame.
struct TTFdRow {
    int m;                   // differentiated dims
    int dim[TT_MAX_ACTIVE];  // storage positions, ascending
    int ord[TT_MAX_ACTIVE];  // 1 or 2
};

struct TTFdProgram {
    int G;
    TTFdRow row[TT_MAX_G];
};

// algo 2: every row differentiates at most one dim.
struct TTSharedProgram {
    int G;
    int n_slots;             // distinct differentiated storage dims, ascending
    int slot_dim[TT_MAX_G];
    int row_slot[TT_MAX_G];  // -1 value row, else index into slot_dim
    int row_ord[TT_MAX_G];   // 1 or 2
};

// Launch configuration of one kernel family for one plan.
struct TTCfg {
    int qpt = 0, threads = 0, lc = 0, mode = 0, pingpong = 0;
    size_t smem = 0;
};

struct TTPlan : PlanBase {
    TTParams P;
    double *d_cores = nullptr;
    TTCfg cfg_chain;   // value + general finite-difference kernels
    TTCfg cfg_shared;  // shared partial-product finite-difference kernel
    // uniform-datapath path (cores in the constant bank), when the train is small enough:
    // values need the forward cores, the shared-FD kernel the transposed copies as well
    bool const_value_ok = false, const_shared_ok = false;
    int const_qpt_value = 2, const_qpt_shared = 2, const_threads_value = 512, const_threads_shared = 512;
    // Constant-bank images (pcb_tt_const.cu).  The 8192-double bank holds only the cores a launch
    // reads: all forward cores for values; for price+Greeks the forward cores left of the last
    // differentiated dim, the transposed cores right of the first, and each differentiated core in
    // the orientation its coefficient pass uses.  One image per distinct need, built on first use.
    struct ConstImage {
        uint64_t id = 0;  // ConstBank residency key
        std::vector<double> data;
        int coff[PCB_MAX_DIMS], coffT[PCB_MAX_DIMS];
    };
    bool const_enabled = false;
    std::vector<double> h_fwd, h_T;  // unpadded cores, forward and transposed, same offsets
    int core_off[PCB_MAX_DIMS + 1] = {0};
    std::mutex image_mutex;
    std::map<uint64_t, ConstImage> images;  // key: need_fwd mask | need_T mask << 32
    // per-core launches with the chain state in global memory (large trains, pcb_tt_const.cu)
    bool gstream_ok = false, gstream_fd_ok = false;  // values / price+Greeks (shared-memory limits)
    std::map<int, ConstImage> gimages;  // key 2 k + orientation; coff[c] chunk bases, coffT[c] widths
    const ConstImage *const_image(uint64_t need_fwd, uint64_t need_T);
    ~TTPlan() override;
};

// Shared-memory carve-up (doubles), identical on host and device:
//   [cores area][nbuf chain-vector buffers of rmaxp * qpt * threads]
__host__ __device__ inline int tt_core_area(const TTParams &P, int mode, bool with_transposed) {
    if (mode == TT_RESIDENT) return P.total + (with_transposed ? P.totalT : 0);
    return mode == TT_STREAM ? P.maxcore : 0;
}
__host__ __device__ inline int tt_smem_doubles(const TTParams &P, int mode, bool with_transposed,
                                               int nbuf, int qpt, int threads) {
    return ((tt_core_area(P, mode, with_transposed) + 1) & ~1) + nbuf * P.rmaxp * qpt * threads;
}

int tt_pick_cfg(const TTPlan *pl, bool shared, TTCfg *cfg);

int tt_launch_value(const TTPlan *pl, const double *d_points, int64_t N, double *d_out, cudaStream_t st);
int tt_launch_general(const TTPlan *pl, const TTFdProgram &prog, const double *d_points, int64_t N,
                      double *d_out, cudaStream_t st);
int tt_launch_shared(const TTPlan *pl, const TTSharedProgram &prog, const double *d_points,
                     int64_t N, double *d_out, cudaStream_t st);
void ttc_forget(const TTPlan *pl);
// constant-bank launches; *fits = false (and nothing launched) when the needed cores exceed the bank
int ttc_launch_value(TTPlan *pl, const double *d_points, int64_t N, double *d_out, cudaStream_t st,
                     bool *fits);
int ttc_launch_shared(TTPlan *pl, const TTSharedProgram &prog, const double *d_points, int64_t N,
                      double *d_out, cudaStream_t st, bool *fits);
int ttc_shared_fits(TTPlan *pl, const TTSharedProgram &prog);  // 0 no, 1 one launch, 2 per dim
bool ttg_shared_fits(const TTPlan *pl, const TTSharedProgram &prog);
// one launch per core, chain state in global memory (trains whose single cores fit in the bank)
int ttg_launch_value(TTPlan *pl, const double *d_points, int64_t N, double *d_out, cudaStream_t st);
int ttg_launch_shared(TTPlan *pl, const TTSharedProgram &prog, const double *d_points, int64_t N,
                      double *d_out, cudaStream_t st, bool *fits);

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// device code
// ---------------------------------------------------------------------------------------------

// reference tensor_train.py:2254 -- exactly this operation order
__device__ __forceinline__ double tt_scale(double x, double a, double b) {
    return 2.0 * (x - a) / (b - a) - 1.0;
}

// reference tensor_train.py:2361-2370
__device__ __forceinline__ double tt_nudge(double x, double a, double b, double h) {
    const double need = h * 1.5;
    if (x - a < need) x = a + need;
    if (b - x < need) x = b - need;
    return x;
}

// Central-difference reduction of one nesting level (tensor_train.py:2387-2401, 2446-2461).
__device__ __forceinline__ double tt_fd_reduce(int ord, double fp, double fc, double fm, double h) {
    return ord == 1 ? (fp - fm) / (2.0 * h) : (fp - 2.0 * fc + fm) / (h * h);
}

__device__ __forceinline__ void tt_prefetch_l1(const void *p) {
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

// One chunk of W output columns of one core:  acc[qq][l] = sum_{i,j} v[i]*T_j(s) * g[i][j][l].
template <int W, int QPT>
__device__ __forceinline__ void tt_chunk(const double *__restrict__ g, int rp, int r_in, int n,
                                         const double *v_in, int vstride, const double (&s)[QPT],
                                         double *v_out) {
    double acc[QPT][W];
    double twos[QPT];
#pragma unroll
    for (int qq = 0; qq < QPT; ++qq) {
        twos[qq] = 2.0 * s[qq];
#pragma unroll
        for (int l = 0; l < W; ++l) acc[qq][l] = 0.0;
    }
    // Software pipeline over the rows (i, j) of the chunk, which are contiguous in memory with
    // stride rp: the 128-bit loads of row t+1 are issued before the FMAs of row t so that LDS and
    // DFMA issue slots interleave.  Rows are zero-padded to an even stride, so every 128-bit load
    // is in bounds; an odd W simply skips the FMA on the pad column.  The load after the last row
    // reads (and discards) the doubles that follow the core, which always exist (next core /
    // chain vectors in shared memory, allocation padding in global memory).
    constexpr int H = (W + 1) / 2;
    const double *gp = g;
    double2 gcur[H];
#pragma unroll
    for (int l = 0; l < H; ++l) gcur[l] = *reinterpret_cast<const double2 *>(gp + 2 * l);
    for (int i = 0; i < r_in; ++i) {
        double c0[QPT], c1[QPT];
#pragma unroll
        for (int qq = 0; qq < QPT; ++qq) {
            const double vi = v_in[(i * QPT + qq) * vstride];
            c0[qq] = vi;          // v[i] * T_0
            c1[qq] = vi * s[qq];  // v[i] * T_1
        }
#pragma unroll 2
        for (int j = 0; j < n; ++j) {
            gp += rp;
            double2 gnext[H];
#pragma unroll
            for (int l = 0; l < H; ++l) gnext[l] = *reinterpret_cast<const double2 *>(gp + 2 * l);
#pragma unroll
            for (int l = 0; l < H; ++l) {
#pragma unroll
                for (int qq = 0; qq < QPT; ++qq) {
                    acc[qq][2 * l] = fma(c0[qq], gcur[l].x, acc[qq][2 * l]);
                    if (2 * l + 1 < W) acc[qq][2 * l + 1] = fma(c0[qq], gcur[l].y, acc[qq][2 * l + 1]);
                }
            }
#pragma unroll
            for (int qq = 0; qq < QPT; ++qq) {
                const double c2 = fma(twos[qq], c1[qq], -c0[qq]);  // T_{j+2} = 2 s T_{j+1} - T_j
                c0[qq] = c1[qq];
                c1[qq] = c2;
            }
#pragma unroll
            for (int l = 0; l < H; ++l) gcur[l] = gnext[l];
        }
    }
#pragma unroll
    for (int l = 0; l < W; ++l)
#pragma unroll
        for (int qq = 0; qq < QPT; ++qq) v_out[(l * QPT + qq) * vstride] = acc[qq][l];
}

#define TT_CHUNK_CASE(FN, WW, ...)                    \
    case WW:                                          \
        if constexpr (WW <= LC) FN<WW, QPT>(__VA_ARGS__); \
        break;

#define TT_CHUNK_SWITCH(FN, ...)                    \
    switch (w) {                                    \
        TT_CHUNK_CASE(FN, 1, __VA_ARGS__)           \
        TT_CHUNK_CASE(FN, 2, __VA_ARGS__)           \
        TT_CHUNK_CASE(FN, 3, __VA_ARGS__)           \
        TT_CHUNK_CASE(FN, 4, __VA_ARGS__)           \
        TT_CHUNK_CASE(FN, 5, __VA_ARGS__)           \
        TT_CHUNK_CASE(FN, 6, __VA_ARGS__)           \
        TT_CHUNK_CASE(FN, 7, __VA_ARGS__)           \
        TT_CHUNK_CASE(FN, 8, __VA_ARGS__)           \
        TT_CHUNK_CASE(FN, 9, __VA_ARGS__)           \
        TT_CHUNK_CASE(FN, 10, __VA_ARGS__)          \
        TT_CHUNK_CASE(FN, 11, __VA_ARGS__)          \
        TT_CHUNK_CASE(FN, 12, __VA_ARGS__)          \
        TT_CHUNK_CASE(FN, 13, __VA_ARGS__)          \
        TT_CHUNK_CASE(FN, 14, __VA_ARGS__)          \
        TT_CHUNK_CASE(FN, 15, __VA_ARGS__)          \
        TT_CHUNK_CASE(FN, 16, __VA_ARGS__)          \
    }

// v_out = v_in . core (r_out true output columns, row stride rp), in chunks of at most LC columns.
// v_out may alias v_in only when the core has a single chunk (r_out <= LC).
template <int QPT, int LC>
__device__ __forceinline__ void tt_apply_core(const double *__restrict__ g, int rp, int r_out,
                                              int r_in, int n, const double *v_in, double *v_out,
                                              int vstride, const double (&s)[QPT]) {
    for (int l0 = 0; l0 < r_out; l0 += LC) {
        const int w = min(LC, r_out - l0);
        const double *gc = g + l0;
        double *vo = v_out + (size_t)l0 * QPT * vstride;
        TT_CHUNK_SWITCH(tt_chunk, gc, rp, r_in, n, v_in, vstride, s, vo)
    }
}

// Fetch one packed core for the whole CTA (STREAM mode copies it into the core area).
template <int MODE>
__device__ __forceinline__ const double *tt_core_ptr(const double *__restrict__ g_cores, double *smem,
                                                     int off, int count) {
    if (MODE == TT_RESIDENT) return smem + off;
    if (MODE == TT_STREAM) {
        __syncthreads();  // previous core fully consumed by every warp
        const double2 *src = reinterpret_cast<const double2 *>(g_cores + off);
        double2 *dst = reinterpret_cast<double2 *>(smem);
        for (int e = threadIdx.x; e < count / 2; e += blockDim.x) dst[e] = src[e];
        __syncthreads();
        return smem;
    }
    return g_cores + off;
}

// Resident mode: copy the packed cores into shared memory once per CTA.
template <int MODE>
__device__ __forceinline__ void tt_load_resident(const double *__restrict__ cores, double *smem,
                                                 int count) {
    if (MODE == TT_RESIDENT) {
        const double2 *src = reinterpret_cast<const double2 *>(cores);
        double2 *dst = reinterpret_cast<double2 *>(smem);
        for (int e = threadIdx.x; e < count / 2; e += blockDim.x) dst[e] = src[e];
        __syncthreads();
    }
}

// Rows of this thread's QPT query slots (tail tile: clamp; results of clamped slots are dropped),
// and an L1 prefetch of the rows the thread will own in its next tile.
template <int QPT>
__device__ __forceinline__ void tt_query_rows(const double *__restrict__ pts, int64_t N, int D,
                                              int64_t q0, int64_t next_q0, const double *(&xrow)[QPT]) {
#pragma unroll
    for (int qq = 0; qq < QPT; ++qq) {
        int64_t q = q0 + qq * (int64_t)blockDim.x + threadIdx.x;
        if (q >= N) q = N - 1;
        xrow[qq] = pts + q * D;
        const int64_t qn = next_q0 + qq * (int64_t)blockDim.x + threadIdx.x;
        if (qn < N) {
            const char *p = reinterpret_cast<const char *>(pts + qn * D);
            tt_prefetch_l1(p);
            tt_prefetch_l1(p + D * 8 - 1);
        }
    }
}

// Full chain for the QPT query slots of this thread.  `od/ox` override the coordinate of up to
// `m` storage dims (finite-difference stencil points).
template <int QPT, int MODE, int LC>
__device__ __forceinline__ void tt_chain(const TTParams &P, const double *__restrict__ g_cores,
                                         double *smem, int v_off, int pingpong,
                                         const double *const (&xrow)[QPT], int m, const int *od,
                                         const double (*ox)[QPT], double (&res)[QPT]) {
    const int vstride = blockDim.x;
    double *vin = smem + v_off + threadIdx.x;
    double *vout = vin + (pingpong ? P.rmaxp * QPT * vstride : 0);
#pragma unroll
    for (int qq = 0; qq < QPT; ++qq) vin[qq * vstride] = 1.0;
    for (int k = 0; k < P.D; ++k) {
        double s[QPT];
        const double a = P.lo[k], b = P.hi[k];
#pragma unroll
        for (int qq = 0; qq < QPT; ++qq) {
            double x = __ldg(xrow[qq] + P.perm[k]);
            for (int t = 0; t < m; ++t)
                if (od[t] == k) x = ox[t][qq];
            s[qq] = tt_scale(x, a, b);
        }
        const double *g = tt_core_ptr<MODE>(g_cores, smem, P.off[k], P.r[k] * P.n[k] * P.rp[k]);
        tt_apply_core<QPT, LC>(g, P.rp[k], P.r[k + 1], P.r[k], P.n[k], vin, vout, vstride, s);
        double *t = vin;
        vin = vout;
        vout = t;
    }
#pragma unroll
    for (int qq = 0; qq < QPT; ++qq) res[qq] = vin[qq * vstride];
}

// Host-side launch helper shared by the three kernel families.
template <typename K, typename... Args>
static int tt_launch_kernel(K kernel, const TTPlan *pl, const TTCfg &cfg, int64_t N, cudaStream_t st,
                            Args... args) {
    PCB_CUDA(allow_dynamic_smem(kernel, cfg.smem, pl->smem_optin));
    int per_sm = 0;
    PCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, cfg.threads, cfg.smem));
    if (per_sm < 1)
        return fail(PCB_ECUDA, "TT kernel (%d threads, %zu B smem) does not fit on an SM",
                    cfg.threads, cfg.smem);
    const int64_t tile = (int64_t)cfg.threads * cfg.qpt;
    const int64_t ntiles = (N + tile - 1) / tile;
    const int64_t cap = (int64_t)pl->sm_count * per_sm;
    kernel<<<(int)(ntiles < cap ? ntiles : cap), cfg.threads, cfg.smem, st>>>(args...);
    g_launches.fetch_add(1);
    PCB_CUDA(cudaGetLastError());
    return PCB_OK;
}

// Instantiated (QPT, LC, MAXT) combinations: resident cores get the tuned set, streamed / global
// cores the generic one.  X(QPT, LC, MAXT)
// (a launch with `threads` picks the first entry whose MAXT >= threads: keep MAXT ascending)
#define TT_RESIDENT_CONFIGS(X) \
    X(2, 8, 256) X(2, 12, 256) X(2, 16, 256) \
    X(1, 8, 512) X(1, 12, 512) X(1, 16, 512) X(2, 8, 512) X(2, 12, 512) X(2, 16, 512)
#define TT_GENERIC_CONFIGS(X) X(2, 16, 256) X(1, 16, 512) X(2, 16, 512)
#define TT_SHARED_RESIDENT_CONFIGS(X) \
    X(2, 8, 256) X(2, 12, 256) X(2, 16, 256) \
    X(2, 8, 384) X(2, 12, 384) X(2, 16, 384) \
    X(1, 8, 512) X(1, 12, 512) X(1, 16, 512) X(2, 12, 512)

#endif  // __CUDACC__

}  // namespace pcb
