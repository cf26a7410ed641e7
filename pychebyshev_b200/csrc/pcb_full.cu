// ChebyshevApproximation batch evaluation on sm_100a.
//
// Replaces ChebyshevApproximation.vectorized_eval_batch (reference barycentric.py:992-1047) for G
// pre-differentiated tensors at once (the derivative passes of barycentric.py:951-990 are applied
// on the host with the reference's own recipe and uploaded; SURVEY.md finding 2).
//
// Two kernels:
//   full_fma_kernel   thread-per-query depth-first FMA evaluator (any D <= 8; small tensors)
//   full_dmma_kernel  mode-1 contraction as an FP64 tensor-core GEMM (mma.sync m8n8k4 -> DMMA.8x8x4)
//                     with the remaining mode contractions fused on chip:
//                       M = 8 queries per MMA tile, K = last tensor axis (4 per k-step),
//                       N = 8 positions of the flattened leading axes; the two axes in between are
//                       folded with per-query weights into register accumulators; the leading
//                       axes are folded at the end of each 8-column group and the four lanes of a
//                       quad are reduced with warp shuffles.  The prepared tensor is streamed
//                       through shared memory with bulk async copies (TMA, cp.async.bulk) behind
//                       an mbarrier full/empty ring fed by a dedicated producer warp.
#include "pcb_grid.cuh"

namespace pcb {

constexpr int FULL_FMA_THREADS = 128;

constexpr int DM_WARPS = 8;                       // consumer warps per CTA
constexpr int DM_MT = 4;                          // 8-query MMA row tiles per warp
constexpr int DM_QT = DM_WARPS * DM_MT * 8;       // queries per CTA tile (256)
constexpr int DM_THREADS = (DM_WARPS + 1) * 32;   // + 1 producer warp
constexpr int DM_MAX_KB = 8;                      // last axis up to 32 nodes
constexpr int DM_MAX_STAGES = 4;

struct DmmaParams {
    int D;
    int G;
    int n[PCB_MAX_DIMS];
    int woff[PCB_MAX_DIMS];  // offset of dim d's weight rows in the smem weight table (rows)
    int n_lead_dims;         // dims [0, n_lead_dims) are flattened into the MMA column axis
    int L;                   // prod n[lead dims]
    int LG;                  // ceil(L / 8)
    int nc, nd;              // extents of the two middle axes (1 when absent)
    int dim_c, dim_d;        // their dim indices (-1 when absent)
    int KB;                  // ceil(n_last / 4)
    int wrows;               // rows of the smem weight table (sum of n over dims < D-1)
    int slab;                // doubles per (lead group, c) slab = nd * KB * 32
    int stages;
    long long gstride;       // doubles per prepared tensor
    int node_off[PCB_MAX_DIMS];
};

// Joint-K variant (round 2): the LAST TWO axes (d, e) are flattened into the MMA K dimension,
// K = n_d * n_e padded to a multiple of 4, with the A operand A[q, (d, e)] = w_d(q) * w_e(q) built
// once per query tile and held in REGISTERS (KB x 2 row tiles per lane; a first version that re-read
// A from shared memory for every slab measured 55 % on 11^5 -- one LDS per DMMA with four consumer
// warps -- against 66 % for the per-row variant).  Against the variant above this (i) pads K once
// (11 x 11 = 121 -> 124, 97.6 %, instead of 11 -> 12, 91.7 %, per row) and (ii) needs no per-d
// weighted fold on the FP64 pipe (8 DFMA per 12 DMMA), which together cap the 11^5 case at 79.5 % of
// the pipe; joint-K caps at 0.976 x 0.945 (121 leading positions in 128 columns) = 92 %.
constexpr int DM2_WARPS = 8;                        // consumer warps per CTA
constexpr int DM2_MT = 2;                           // 8-query row tiles per warp (A fragments in registers)
constexpr int DM2_QT = DM2_WARPS * DM2_MT * 8;      // 128 queries per CTA tile
constexpr int DM2_THREADS = (DM2_WARPS + 1) * 32;   // + 1 producer warp
constexpr int DM2_MAX_KB = 36;                      // K <= 144 (12 x 12); larger K keeps the per-row variant

struct Dmma2Params {
    int D, G;
    int n[PCB_MAX_DIMS];
    int woff[PCB_MAX_DIMS];   // rows of the weight table for dims < D - 2
    int node_off[PCB_MAX_DIMS];
    int n_lead_dims, L, LG;   // dims [0, n_lead_dims) flattened into the MMA column axis
    int nc, dim_c;            // the folded middle axis (-1 / 1 when absent)
    int K, KB;                // n[D-2] * n[D-1], ceil(K / 4)
    int wrows;                // rows of the weight table
    int slab;                 // doubles per (lead group, c) slab = KB * 32
    int stages;
    long long gstride;        // doubles per prepared tensor
};

struct FullPlan : PlanBase {
    void *small = nullptr;  // one-piece spline plan on the constant bank (small tensors), or null
    GridDesc gd;
    int G = 0;
    double *d_nodes = nullptr;    // nodes then weights, dims concatenated
    double *d_weights = nullptr;
    double *d_tensors = nullptr;  // G tensors interleaved in blocks of GB outputs: [block][elem][GB]
    int GB = 1;
    double *d_prepared = nullptr; // G prepared (fragment-ordered, zero-padded) tensors
    bool dmma_ok = false;
    DmmaParams dm;
    size_t dm_smem = 0;
    double *d_prepared2 = nullptr;  // joint-K fragment order
    bool dmma2_ok = false;
    Dmma2Params dm2;
    size_t dm2_smem = 0;
    ~FullPlan() override {
        if (d_prepared2) cudaFree(d_prepared2);
        if (small) pcb_plan_destroy(small);
        if (d_nodes) cudaFree(d_nodes);
        if (d_tensors) cudaFree(d_tensors);
        if (d_prepared) cudaFree(d_prepared);
    }
};

// ---------------------------------------------------------------------------------------------
// FMA evaluator
// ---------------------------------------------------------------------------------------------

template <int GB, int DM>
__global__ void __launch_bounds__(FULL_FMA_THREADS)
full_fma_kernel(const __grid_constant__ GridDesc gd, int G, const double *__restrict__ nodes,
                const double *__restrict__ weights, const double *__restrict__ tensors,
                const double *__restrict__ pts, int64_t N, double *__restrict__ out) {
    extern __shared__ __align__(16) double smem[];
    double *ws = smem + threadIdx.x;
    const int stride = blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < N;
         q += (int64_t)gridDim.x * blockDim.x) {
        const double *x = pts + q * gd.D;
        double wl[GRID_NL];
        const bool regs = grid_weights(gd, nodes, weights, [&](int d) { return __ldg(x + d); }, ws,
                                       stride, wl);
        grid_eval_outputs<GB, DM>(gd, tensors, G, ws, stride, wl, regs, out + q * G, 1);
    }
}

// ---------------------------------------------------------------------------------------------
// DMMA kernel
// ---------------------------------------------------------------------------------------------

// D(8x8) += A(8x4) * B(4x8), fp64 tensor core (SASS: DMMA.8x8x4)
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    // not volatile: a pure function of its operands, so the scheduler may interleave the
    // independent accumulator chains and the weighted folds freely
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

struct DmmaSmem {
    int w;      // weight table [wrows][DM_QT]
    int ring;   // stages * slab doubles (also scratch for the last-axis weights in the prologue)
    int bars;   // 2 * stages uint64
    int total;  // doubles
};

__host__ __device__ inline DmmaSmem dmma_smem_layout(const DmmaParams &P) {
    DmmaSmem L;
    L.w = 0;
    L.ring = L.w + P.wrows * DM_QT;
    int ring = P.stages * P.slab;
    const int scratch = P.KB * 4 * DM_QT;  // last-axis weight rows, padded to KB*4
    if (ring < scratch) ring = scratch;
    L.bars = L.ring + ring;
    L.total = L.bars + 2 * DM_MAX_STAGES;
    return L;
}

template <int KB>
__global__ void __launch_bounds__(DM_THREADS, 1)
full_dmma_kernel(const __grid_constant__ DmmaParams P, const double *__restrict__ nodes,
                 const double *__restrict__ weights, const double *__restrict__ prepared,
                 const double *__restrict__ pts, int64_t N, double *__restrict__ out) {
    extern __shared__ __align__(128) double smem[];
    const DmmaSmem L = dmma_smem_layout(P);
    double *w_s = smem + L.w;
    double *ring = smem + L.ring;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + L.bars);
    uint64_t *empty = full + DM_MAX_STAGES;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int D = P.D;
    const int nlast = P.n[D - 1];
    const int64_t ntiles = (N + DM_QT - 1) / DM_QT;
    const uint32_t slab_bytes = (uint32_t)P.slab * 8u;
    const int slabs_per_g = P.LG * P.nc;

    if (tid == 0) {
        for (int s = 0; s < P.stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], DM_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // ring position persists across tiles (producer and consumers advance in lock step)
    uint32_t it = 0;  // slabs handled so far by this thread's role

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t q0 = tile * DM_QT;
        // ---- prologue: barycentric weight rows of the tile's queries (K-A0) -----------------
        // one (query, dim) pair per thread iteration; reference barycentric.py:1038-1045
        __syncthreads();  // every warp finished the previous tile (ring drained, w_s free)
        for (int e = tid; e < DM_QT * D; e += DM_THREADS) {
            const int ql = e % DM_QT;
            const int d = e / DM_QT;
            int64_t q = q0 + ql;
            if (q >= N) q = N - 1;
            const double x = __ldg(pts + q * D + d);
            const int n = P.n[d];
            double *dst;
            int rows;
            if (d == D - 1) {
                dst = ring + ql;
                rows = KB * 4;
            } else {
                dst = w_s + (size_t)P.woff[d] * DM_QT + ql;
                rows = n;
            }
            grid_weight_row(x, n, nodes + P.node_off[d], weights + P.node_off[d], dst, DM_QT);
            for (int i = n; i < rows; ++i) dst[i * DM_QT] = 0.0;
        }
        __syncthreads();

        if (warp == DM_WARPS) {
            // ---- producer warp: stream the prepared tensors through the ring ------------------
            __syncthreads();  // matches the consumers' "A fragments loaded" barrier
            if (lane == 0) {
                for (int g = 0; g < P.G; ++g) {
                    const double *src = prepared + (size_t)g * P.gstride;
                    for (int sidx = 0; sidx < slabs_per_g; ++sidx, ++it) {
                        const int s = it % P.stages;
                        const uint32_t round = it / P.stages;
                        if (round > 0) mbar_wait(&empty[s], (round - 1) & 1);
                        mbar_arrive_expect_tx(&full[s], slab_bytes);
                        bulk_g2s(ring + (size_t)s * P.slab, src + (size_t)sidx * P.slab, slab_bytes,
                                 &full[s]);
                    }
                }
            }
            __syncwarp();
        } else {
            // ---- consumer warps ---------------------------------------------------------------
            const int qrow = warp * (DM_MT * 8) + (lane >> 2);  // + mt*8: this lane's MMA row
            const int kcol = lane & 3;
            double afrag[KB][DM_MT];
#pragma unroll
            for (int kb = 0; kb < KB; ++kb)
#pragma unroll
                for (int mt = 0; mt < DM_MT; ++mt)
                    afrag[kb][mt] = ring[(size_t)(kb * 4 + kcol) * DM_QT + qrow + mt * 8];
            __syncthreads();  // scratch consumed: the producer may overwrite the ring

            for (int g = 0; g < P.G; ++g) {
                double outacc[DM_MT];
#pragma unroll
                for (int mt = 0; mt < DM_MT; ++mt) outacc[mt] = 0.0;
                for (int lg = 0; lg < P.LG; ++lg) {
                    double acc3[DM_MT][2];
#pragma unroll
                    for (int mt = 0; mt < DM_MT; ++mt) acc3[mt][0] = acc3[mt][1] = 0.0;
                    for (int c = 0; c < P.nc; ++c, ++it) {
                        const int s = it % P.stages;
                        mbar_wait(&full[s], (it / P.stages) & 1);
                        const double *slab = ring + (size_t)s * P.slab + lane;
                        double acc2[DM_MT][2];
#pragma unroll
                        for (int mt = 0; mt < DM_MT; ++mt) acc2[mt][0] = acc2[mt][1] = 0.0;
                        // d-loop, software pipelined over two accumulator sets: the MMAs of step d
                        // are issued before the weighted fold of step d-1, so the fold never waits
                        // on the tensor pipe.
                        const double *wd_base =
                            P.dim_d >= 0 ? w_s + (size_t)P.woff[P.dim_d] * DM_QT + qrow : nullptr;
                        auto mma_step = [&](int d, double (&cf)[DM_MT][2]) {
#pragma unroll
                            for (int mt = 0; mt < DM_MT; ++mt) cf[mt][0] = cf[mt][1] = 0.0;
#pragma unroll
                            for (int kb = 0; kb < KB; ++kb) {
                                const double b = slab[(d * KB + kb) * 32];
#pragma unroll
                                for (int mt = 0; mt < DM_MT; ++mt)
                                    dmma884(cf[mt][0], cf[mt][1], afrag[kb][mt], b);
                            }
                        };
                        auto fold_step = [&](int d, const double (&cf)[DM_MT][2]) {
                            if (wd_base) {
                                const double *wd = wd_base + (size_t)d * DM_QT;
#pragma unroll
                                for (int mt = 0; mt < DM_MT; ++mt) {
                                    const double w = wd[mt * 8];
                                    acc2[mt][0] = fma(w, cf[mt][0], acc2[mt][0]);
                                    acc2[mt][1] = fma(w, cf[mt][1], acc2[mt][1]);
                                }
                            } else {
#pragma unroll
                                for (int mt = 0; mt < DM_MT; ++mt) {
                                    acc2[mt][0] = cf[mt][0];
                                    acc2[mt][1] = cf[mt][1];
                                }
                            }
                        };
                        double cfa[DM_MT][2], cfb[DM_MT][2];
                        mma_step(0, cfa);
                        int d = 1;
                        for (; d + 1 < P.nd; d += 2) {
                            mma_step(d, cfb);
                            fold_step(d - 1, cfa);
                            mma_step(d + 1, cfa);
                            fold_step(d, cfb);
                        }
                        if (d < P.nd) {
                            mma_step(d, cfb);
                            fold_step(d - 1, cfa);
                            fold_step(d, cfb);
                        } else {
                            fold_step(d - 1, cfa);
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&empty[s]);  // slab consumed by this warp
                        if (P.dim_c >= 0) {
                            const double *wc = w_s + (size_t)(P.woff[P.dim_c] + c) * DM_QT + qrow;
#pragma unroll
                            for (int mt = 0; mt < DM_MT; ++mt) {
                                const double w = wc[mt * 8];
                                acc3[mt][0] = fma(w, acc2[mt][0], acc3[mt][0]);
                                acc3[mt][1] = fma(w, acc2[mt][1], acc3[mt][1]);
                            }
                        } else {
#pragma unroll
                            for (int mt = 0; mt < DM_MT; ++mt) {
                                acc3[mt][0] = acc2[mt][0];
                                acc3[mt][1] = acc2[mt][1];
                            }
                        }
                    }
                    // fold the leading axes: this lane owns columns p0, p0 + 1 of the group
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int p = lg * 8 + kcol * 2 + e;
                        if (p < P.L) {
#pragma unroll
                            for (int mt = 0; mt < DM_MT; ++mt) {
                                double wl = 1.0;
                                int rem = p;
                                for (int dd = P.n_lead_dims - 1; dd >= 0; --dd) {
                                    const int idx = rem % P.n[dd];
                                    rem /= P.n[dd];
                                    wl *= w_s[(size_t)(P.woff[dd] + idx) * DM_QT + qrow + mt * 8];
                                }
                                outacc[mt] = fma(wl, acc3[mt][e], outacc[mt]);
                            }
                        }
                    }
                }
                // reduce the four lanes of each quad (they hold different columns of one query)
#pragma unroll
                for (int mt = 0; mt < DM_MT; ++mt) {
                    double v = outacc[mt];
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    const int64_t q = q0 + qrow + mt * 8;
                    if (kcol == 0 && q < N) out[q * P.G + g] = v;
                }
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------
// joint-K DMMA kernel
// ---------------------------------------------------------------------------------------------
struct Dmma2Smem {
    int w, ring, bars, total;  // doubles
};
__host__ __device__ inline Dmma2Smem dmma2_smem_layout(const Dmma2Params &P) {
    Dmma2Smem L;
    L.w = 0;
    L.ring = L.w + P.wrows * DM2_QT;
    int ring = P.stages * P.slab;
    const int scratch = (P.n[P.D - 2] + P.n[P.D - 1]) * DM2_QT;  // the two last weight rows (prologue)
    if (ring < scratch) ring = scratch;
    L.bars = L.ring + ring;
    L.total = L.bars + 2 * DM_MAX_STAGES;
    return L;
}

template <int KB>
__global__ void __launch_bounds__(DM2_THREADS, 1)
full_dmma2_kernel(const __grid_constant__ Dmma2Params P, const double *__restrict__ nodes,
                  const double *__restrict__ weights, const double *__restrict__ prepared,
                  const double *__restrict__ pts, int64_t N, double *__restrict__ out) {
    extern __shared__ __align__(128) double smem[];
    const Dmma2Smem L = dmma2_smem_layout(P);
    double *w_s = smem + L.w;
    double *ring = smem + L.ring;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + L.bars);
    uint64_t *empty = full + DM_MAX_STAGES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int D = P.D, nd = P.n[D - 2], ne = P.n[D - 1];
    const int64_t ntiles = (N + DM2_QT - 1) / DM2_QT;
    const uint32_t slab_bytes = (uint32_t)P.slab * 8u;
    const int slabs_per_g = P.LG * P.nc;
    if (tid == 0) {
        for (int s = 0; s < P.stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], DM2_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t it = 0;  // slabs handled so far by this thread's role (ring position persists across tiles)
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t q0 = tile * DM2_QT;
        __syncthreads();  // previous tile done: ring drained, tables free
        // ---- prologue 1: barycentric weight rows (K-A0); the two last dims go to the ring scratch
        for (int e = tid; e < DM2_QT * D; e += DM2_THREADS) {
            const int ql = e % DM2_QT, d = e / DM2_QT;
            int64_t q = q0 + ql;
            if (q >= N) q = N - 1;
            const double x = __ldg(pts + q * D + d);
            double *dst = d == D - 1 ? ring + (size_t)nd * DM2_QT + ql
                                     : (d == D - 2 ? ring + ql : w_s + (size_t)P.woff[d] * DM2_QT + ql);
            grid_weight_row(x, P.n[d], nodes + P.node_off[d], weights + P.node_off[d], dst, DM2_QT);
        }
        __syncthreads();
        // ---- prologue 2 (consumers): A fragments A[q, k] = w_d(q)[k / ne] * w_e(q)[k % ne] into
        //      registers, k = 4 kb + (lane & 3), q = this lane's rows; zero beyond K
        const int qrow = (warp < DM2_WARPS ? warp : 0) * (DM2_MT * 8) + (lane >> 2);
        const int kcol = lane & 3;
        double afrag[KB][DM2_MT];
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
            const int k = 4 * kb + kcol;
            const bool live_k = k < P.K;
            const int kd = live_k ? k / ne : 0, ke = live_k ? k - kd * ne : 0;
#pragma unroll
            for (int mt = 0; mt < DM2_MT; ++mt) {
                const int ql = qrow + mt * 8;
                const double v = ring[(size_t)kd * DM2_QT + ql] * ring[(size_t)(nd + ke) * DM2_QT + ql];
                afrag[kb][mt] = live_k ? v : 0.0;
            }
        }
        __syncthreads();  // scratch consumed: the producer may overwrite the ring
        if (warp == DM2_WARPS) {
            if (lane == 0) {
                for (int g = 0; g < P.G; ++g) {
                    const double *src = prepared + (size_t)g * P.gstride;
                    for (int sidx = 0; sidx < slabs_per_g; ++sidx, ++it) {
                        const int s = it % P.stages;
                        const uint32_t round = it / P.stages;
                        if (round > 0) mbar_wait(&empty[s], (round - 1) & 1);
                        mbar_arrive_expect_tx(&full[s], slab_bytes);
                        bulk_g2s(ring + (size_t)s * P.slab, src + (size_t)sidx * P.slab, slab_bytes, &full[s]);
                    }
                }
            }
            __syncwarp();
        } else {
            for (int g = 0; g < P.G; ++g) {
                double outacc[DM2_MT];
#pragma unroll
                for (int mt = 0; mt < DM2_MT; ++mt) outacc[mt] = 0.0;
                for (int lg = 0; lg < P.LG; ++lg) {
                    double acc3[DM2_MT][2];
#pragma unroll
                    for (int mt = 0; mt < DM2_MT; ++mt) acc3[mt][0] = acc3[mt][1] = 0.0;
                    for (int c = 0; c < P.nc; ++c, ++it) {
                        const int s = it % P.stages;
                        mbar_wait(&full[s], (it / P.stages) & 1);
                        const double *slab = ring + (size_t)s * P.slab + lane;
                        // two accumulator sets (even / odd k-blocks): four independent MMA chains
                        // per warp, so a chain's latency does not gate the tensor pipe
                        double acc2[DM2_MT][2], accb[DM2_MT][2];
#pragma unroll
                        for (int mt = 0; mt < DM2_MT; ++mt)
                            acc2[mt][0] = acc2[mt][1] = accb[mt][0] = accb[mt][1] = 0.0;
#pragma unroll
                        for (int kb = 0; kb < KB; ++kb) {
                            const double b = slab[kb * 32];
#pragma unroll
                            for (int mt = 0; mt < DM2_MT; ++mt) {
                                if (kb & 1)
                                    dmma884(accb[mt][0], accb[mt][1], afrag[kb][mt], b);
                                else
                                    dmma884(acc2[mt][0], acc2[mt][1], afrag[kb][mt], b);
                            }
                        }
#pragma unroll
                        for (int mt = 0; mt < DM2_MT; ++mt) {
                            acc2[mt][0] += accb[mt][0];
                            acc2[mt][1] += accb[mt][1];
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&empty[s]);
                        if (P.dim_c >= 0) {
                            const double *wc = w_s + (size_t)(P.woff[P.dim_c] + c) * DM2_QT + qrow;
#pragma unroll
                            for (int mt = 0; mt < DM2_MT; ++mt) {
                                const double w = wc[mt * 8];
                                acc3[mt][0] = fma(w, acc2[mt][0], acc3[mt][0]);
                                acc3[mt][1] = fma(w, acc2[mt][1], acc3[mt][1]);
                            }
                        } else {
#pragma unroll
                            for (int mt = 0; mt < DM2_MT; ++mt) {
                                acc3[mt][0] = acc2[mt][0];
                                acc3[mt][1] = acc2[mt][1];
                            }
                        }
                    }
                    // fold the leading axes: this lane owns columns 2 kcol, 2 kcol + 1 of the group
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int p = lg * 8 + kcol * 2 + e;
                        if (p < P.L) {
#pragma unroll
                            for (int mt = 0; mt < DM2_MT; ++mt) {
                                double wl = 1.0;
                                int rem = p;
                                for (int dd = P.n_lead_dims - 1; dd >= 0; --dd) {
                                    const int idx = rem % P.n[dd];
                                    rem /= P.n[dd];
                                    wl *= w_s[(size_t)(P.woff[dd] + idx) * DM2_QT + qrow + mt * 8];
                                }
                                outacc[mt] = fma(wl, acc3[mt][e], outacc[mt]);
                            }
                        }
                    }
                }
#pragma unroll
                for (int mt = 0; mt < DM2_MT; ++mt) {
                    double v = outacc[mt];
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    const int64_t q = q0 + qrow + mt * 8;
                    if (kcol == 0 && q < N) out[q * P.G + g] = v;
                }
            }
        }
    }
}

// prepared2[((lg * nc + c) * KB + kb) * 32 + lane] = T[lead = 8 lg + lane/4][c][k = 4 kb + lane%4]
__global__ void __launch_bounds__(256)
dmma2_prepare_kernel(const __grid_constant__ Dmma2Params P, const double *__restrict__ t,
                     double *__restrict__ dst) {
    const long long s_c = P.K;
    const long long s_lead = s_c * P.nc;
    for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < P.gstride;
         o += (long long)gridDim.x * blockDim.x) {
        const int lane = (int)(o & 31);
        long long r = o >> 5;
        const int kb = (int)(r % P.KB);
        r /= P.KB;
        const int c = (int)(r % P.nc);
        const long long lg = r / P.nc;
        const long long lead = lg * 8 + lane / 4;
        const int k = kb * 4 + lane % 4;
        dst[o] = (lead < P.L && k < P.K) ? t[lead * s_lead + c * s_c + k] : 0.0;
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------

// Reorder one C-order tensor (device) into MMA B-fragment order (device):
//   prepared[((lg * nc + c) * nd + d) * KB + kb][lane] = T[lead = 8 lg + lane/4][c][d][e = 4 kb + lane%4]
// One thread per output element: coalesced stores, gathered loads (each 32-lane group reads 8 rows
// of 4 consecutive doubles).  Runs once per tensor at plan creation.
__global__ void __launch_bounds__(256)
dmma_prepare_kernel(const __grid_constant__ DmmaParams P, const double *__restrict__ t,
                    double *__restrict__ dst) {
    const int nlast = P.n[P.D - 1];
    const long long s_d = nlast;
    const long long s_c = s_d * P.nd;
    const long long s_lead = s_c * P.nc;
    for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < P.gstride;
         o += (long long)gridDim.x * blockDim.x) {
        const int lane = (int)(o & 31);
        long long r = o >> 5;
        const int kb = (int)(r % P.KB);
        r /= P.KB;
        const int d = (int)(r % P.nd);
        r /= P.nd;
        const int c = (int)(r % P.nc);
        const long long lg = r / P.nc;
        const long long lead = lg * 8 + lane / 4;
        const int e = kb * 4 + lane % 4;
        dst[o] = (lead < P.L && e < nlast) ? t[lead * s_lead + c * s_c + d * s_d + e] : 0.0;
    }
}

// dst[e * GB + slot] = src[e]: one output of an interleaved block [elem][GB]
__global__ void __launch_bounds__(256)
interleave_kernel(const double *__restrict__ src, double *__restrict__ dst, long long size, int GB,
                  int slot) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < size;
         e += (long long)gridDim.x * blockDim.x)
        dst[e * GB + slot] = src[e];
}

static bool dmma_configure(FullPlan *pl, const int32_t *n) {
    const int D = pl->gd.D;
    if (D < 2) return false;
    DmmaParams &P = pl->dm;
    memset(&P, 0, sizeof(P));
    P.D = D;
    P.G = pl->G;
    int off = 0, woff = 0;
    for (int d = 0; d < D; ++d) {
        P.n[d] = n[d];
        P.node_off[d] = off;
        off += n[d];
        P.woff[d] = woff;
        if (d < D - 1) woff += n[d];
    }
    P.wrows = woff;
    P.KB = (n[D - 1] + 3) / 4;
    if (P.KB > DM_MAX_KB) return false;
    // number of middle axes: as many as possible (<= 2) while the flattened leading axes still
    // fill the 8-wide MMA column tiles well
    int best_m = -1;
    double best_eff = -1.0;
    for (int m = (D - 2 < 2 ? D - 2 : 2); m >= 0; --m) {
        long long L = 1;
        for (int d = 0; d < D - 1 - m; ++d) L *= n[d];
        const double eff = (double)L / (double)(((L + 7) / 8) * 8);
        if (eff >= 0.9) {  // deepest nesting whose column tiles are at least 90 % full
            best_m = m;
            break;
        }
        if (eff > best_eff) {
            best_eff = eff;
            best_m = m;
        }
    }
    const int m = best_m;
    P.n_lead_dims = D - 1 - m;
    long long L = 1;
    for (int d = 0; d < P.n_lead_dims; ++d) L *= n[d];
    if (L > (1 << 28)) return false;
    P.L = (int)L;
    P.LG = (P.L + 7) / 8;
    P.dim_d = m >= 1 ? D - 2 : -1;
    P.dim_c = m >= 2 ? D - 3 : -1;
    P.nd = P.dim_d >= 0 ? n[P.dim_d] : 1;
    P.nc = P.dim_c >= 0 ? n[P.dim_c] : 1;
    P.slab = P.nd * P.KB * 32;
    P.gstride = (long long)P.LG * P.nc * P.slab;
    // stages: as many as fit (<= 4)
    for (int st = DM_MAX_STAGES; st >= 2; --st) {
        P.stages = st;
        const DmmaSmem Ls = dmma_smem_layout(P);
        if ((size_t)Ls.total * 8 <= (size_t)pl->smem_optin) {
            pl->dm_smem = (size_t)Ls.total * 8;
            return true;
        }
    }
    return false;
}


// Joint-K variant: worth it when the last axis is not a multiple of 4 (or the per-d fold hurts) and
// the A table fits in shared memory.
static bool dmma2_configure(FullPlan *pl, const int32_t *n) {
    const int D = pl->gd.D;
    if (D < 3 || getenv("PCB_NO_DMMA2")) return false;
    Dmma2Params &P = pl->dm2;
    memset(&P, 0, sizeof(P));
    P.D = D;
    P.G = pl->G;
    int off = 0, woff = 0;
    for (int d = 0; d < D; ++d) {
        P.n[d] = n[d];
        P.node_off[d] = off;
        off += n[d];
        P.woff[d] = woff;
        if (d < D - 2) woff += n[d];
    }
    P.wrows = woff;
    P.K = n[D - 2] * n[D - 1];
    P.KB = (P.K + 3) / 4;
    if (P.KB > DM2_MAX_KB) return false;
    // the kernel is instantiated for a few K-block counts (A fragments live in registers, so the kb
    // loop must unroll); round up to the next one (the extra rows are zero in A and in the slabs)
    static const int kb_classes[] = {16, 21, 25, 31, 36};
    for (int kc : kb_classes)
        if (P.KB <= kc) {
            P.KB = kc;
            break;
        }
    // one folded middle axis when the flattened leading axes still fill the 8-wide column tiles
    P.dim_c = -1;
    P.nc = 1;
    P.n_lead_dims = D - 2;
    if (D >= 4) {
        long long Lc = 1;
        for (int d = 0; d < D - 3; ++d) Lc *= n[d];
        const double eff = (double)Lc / (double)(((Lc + 7) / 8) * 8);
        if (eff >= 0.9) {
            P.dim_c = D - 3;
            P.nc = n[D - 3];
            P.n_lead_dims = D - 3;
        }
    }
    long long L = 1;
    for (int d = 0; d < P.n_lead_dims; ++d) L *= n[d];
    if (L > (1 << 28)) return false;
    P.L = (int)L;
    P.LG = (P.L + 7) / 8;
    P.slab = P.KB * 32;
    P.gstride = (long long)P.LG * P.nc * P.slab;
    for (int st = DM_MAX_STAGES; st >= 2; --st) {
        P.stages = st;
        const Dmma2Smem Ls = dmma2_smem_layout(P);
        if ((size_t)Ls.total * 8 <= (size_t)pl->smem_optin) {
            pl->dm2_smem = (size_t)Ls.total * 8;
            return true;
        }
    }
    return false;
}

// Predicted FP64-pipe efficiency of the two tensor-core variants (useful work / pipe time).
static double dmma_eff(const FullPlan *pl) {
    const DmmaParams &P = pl->dm;
    const int ne = P.n[P.D - 1];
    const double kpad = (double)ne / (4.0 * P.KB);
    const double lpad = (double)P.L / (8.0 * P.LG);
    const double fold = P.dim_d >= 0 ? (double)(P.KB * 8) / (double)(P.KB * 8 + 1) : 1.0;  // 1 DFMA-pair per KB DMMAs
    return kpad * lpad * fold;
}
static double dmma2_eff(const FullPlan *pl) {
    const Dmma2Params &P = pl->dm2;
    return ((double)P.K / (4.0 * P.KB)) * ((double)P.L / (8.0 * P.LG));
}

}  // namespace pcb

using namespace pcb;

namespace pcb {

// Source of the G C-order tensors of a plan, as DEVICE pointers valid until the next call.
struct TensorSource {
    virtual const double *get(int g) = 0;   // nullptr + last error on failure
    virtual const double *host(int g) { return nullptr; }  // host copy if there is one
    virtual ~TensorSource() {}
};

// Tensors given on the host (made with the reference's NumPy recipe): staged through one buffer.
struct HostTensorSource : TensorSource {
    const double *const *tensors;
    long long size;
    double *d_stage = nullptr;
    HostTensorSource(const double *const *t, long long sz) : tensors(t), size(sz) {}
    const double *get(int g) override {
        if (!d_stage && cudaMalloc(&d_stage, (size_t)size * sizeof(double)) != cudaSuccess) {
            fail(PCB_ENOMEM, "cannot allocate the tensor staging buffer");
            return nullptr;
        }
        if (cudaMemcpy(d_stage, tensors[g], (size_t)size * sizeof(double), cudaMemcpyHostToDevice) !=
            cudaSuccess) {
            fail(PCB_ECUDA, "upload of tensor %d failed", g);
            return nullptr;
        }
        return d_stage;
    }
    const double *host(int g) override { return tensors[g]; }
    ~HostTensorSource() override {
        if (d_stage) cudaFree(d_stage);
    }
};

// N3: the value tensor is uploaded ONCE and every derivative tensor is made on the device by the
// passes of _apply_derivative_passes (barycentric.py:982-989: for d = D-1..0, order[d] times
// T <- T x_d D_d^T), bit-identical to the host recipe (pcb_tensor.cu).
struct DerivTensorSource : TensorSource {
    int dev, D;
    const int32_t *n;
    const int32_t *orders;  // G x D
    long long size;
    double *d_val = nullptr, *d_buf[2] = {nullptr, nullptr}, *d_dm = nullptr;
    std::vector<size_t> dm_off;
    std::vector<double> h_copy;
    bool ok = false;
    DerivTensorSource(int dev_, int D_, const int32_t *n_, const double *dmats_cat, const double *values,
                      const int32_t *orders_, long long sz)
        : dev(dev_), D(D_), n(n_), orders(orders_), size(sz) {
        size_t total = 0;
        for (int d = 0; d < D; ++d) {
            dm_off.push_back(total);
            total += (size_t)n[d] * n[d];
        }
        const size_t tb = (size_t)size * sizeof(double);
        ok = cudaMalloc(&d_val, tb) == cudaSuccess && cudaMalloc(&d_buf[0], tb) == cudaSuccess &&
             cudaMalloc(&d_buf[1], tb) == cudaSuccess &&
             cudaMalloc(&d_dm, total * sizeof(double)) == cudaSuccess &&
             cudaMemcpy(d_val, values, tb, cudaMemcpyHostToDevice) == cudaSuccess &&
             cudaMemcpy(d_dm, dmats_cat, total * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess;
        if (!ok) {
            cudaGetLastError();
            fail(PCB_ENOMEM, "device allocation / upload for the derivative passes failed");
        }
    }
    const double *get(int g) override {
        if (!ok) return nullptr;
        const double *cur = d_val;
        int flip = 0;
        for (int d = D - 1; d >= 0; --d) {
            long long outer = 1, inner = 1;
            for (int e = 0; e < d; ++e) outer *= n[e];
            for (int e = d + 1; e < D; ++e) inner *= n[e];
            for (int rep = 0; rep < orders[(size_t)g * D + d]; ++rep) {
                if (tensor_mode_launch(dev, cur, d_buf[flip], outer, n[d], inner, n[d], d_dm + dm_off[d],
                                       nullptr) != PCB_OK)
                    return nullptr;
                cur = d_buf[flip];
                flip ^= 1;
            }
        }
        return cur;
    }
    const double *host(int g) override {  // only used for small tensors (constant-bank plan)
        const double *d = get(g);
        if (!d) return nullptr;
        h_copy.resize((size_t)size);
        if (cudaMemcpy(h_copy.data(), d, (size_t)size * sizeof(double), cudaMemcpyDeviceToHost) !=
            cudaSuccess)
            return nullptr;
        return h_copy.data();
    }
    ~DerivTensorSource() override {
        if (d_val) cudaFree(d_val);
        if (d_buf[0]) cudaFree(d_buf[0]);
        if (d_buf[1]) cudaFree(d_buf[1]);
        if (d_dm) cudaFree(d_dm);
    }
};

static int full_plan_build(int dev, int D, const int32_t *n, const double *nodes_cat,
                           const double *weights_cat, int G, TensorSource &src, void **plan) {
    FullPlan *pl = new FullPlan();
    pl->kind = PLAN_FULL;
    pl->dev = dev;
    pl->G = G;
    int cc = 0;
    if (int rc = device_props(dev, &pl->sm_count, &pl->smem_optin, &cc)) {
        delete pl;
        return rc;
    }
    GridDesc &gd = pl->gd;
    memset(&gd, 0, sizeof(gd));
    gd.D = D;
    gd.size = 1;
    for (int d = 0; d < D; ++d) {
        gd.n[d] = n[d];
        gd.sum_n += n[d];
        gd.size *= n[d];
    }
    const size_t nb = (size_t)gd.sum_n * sizeof(double);
    pl->GB = grid_pick_gb(G);
    const int nblk = (G + pl->GB - 1) / pl->GB;
    const size_t tcount = (size_t)gd.size * nblk * pl->GB;
    if (cudaMalloc(&pl->d_nodes, 2 * nb) != cudaSuccess ||
        cudaMalloc(&pl->d_tensors, tcount * sizeof(double)) != cudaSuccess) {
        cudaGetLastError();
        delete pl;
        return fail(PCB_ENOMEM, "device allocation for the full-tensor plan failed");
    }
    pl->d_weights = pl->d_nodes + gd.sum_n;
    if (cudaMemcpy(pl->d_nodes, nodes_cat, nb, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(pl->d_weights, weights_cat, nb, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemset(pl->d_tensors, 0, tcount * sizeof(double)) != cudaSuccess) {
        delete pl;
        return fail(PCB_ECUDA, "upload of the full-tensor plan failed");
    }
    // tensor-core path: prepared copies in fragment order
    pl->dmma_ok = dmma_configure(pl, n);
    if (pl->dmma_ok && cudaMalloc(&pl->d_prepared, (size_t)pl->dm.gstride * sizeof(double) * G) != cudaSuccess) {
        cudaGetLastError();
        pl->dmma_ok = false;
    }
    pl->dmma2_ok = dmma2_configure(pl, n);
    if (pl->dmma2_ok && pl->dmma_ok && dmma2_eff(pl) <= dmma_eff(pl) + 0.02 && !getenv("PCB_FORCE_DMMA2"))
        pl->dmma2_ok = false;  // the per-row variant is as good (last axis a multiple of 4)
    if (pl->dmma2_ok &&
        cudaMalloc(&pl->d_prepared2, (size_t)pl->dm2.gstride * sizeof(double) * G) != cudaSuccess) {
        cudaGetLastError();
        pl->dmma2_ok = false;
    }
    const bool want_small = D <= 4 && gd.size * G <= 8192;
    std::vector<std::vector<double>> small_t;
    const int grid = (int)std::min<long long>((gd.size + 255) / 256, (long long)pl->sm_count * 32);
    for (int g = 0; g < G; ++g) {
        const double *t = src.get(g);
        if (!t) {
            delete pl;
            return PCB_ECUDA;  // src.get recorded the message
        }
        interleave_kernel<<<grid, 256>>>(t, pl->d_tensors + (size_t)(g / pl->GB) * gd.size * pl->GB,
                                         gd.size, pl->GB, g % pl->GB);
        if (pl->dmma_ok) {
            const int pg = (int)std::min<long long>((pl->dm.gstride + 255) / 256, (long long)pl->sm_count * 32);
            dmma_prepare_kernel<<<pg, 256>>>(pl->dm, t, pl->d_prepared + (size_t)g * pl->dm.gstride);
        }
        if (pl->dmma2_ok) {
            const int pg = (int)std::min<long long>((pl->dm2.gstride + 255) / 256, (long long)pl->sm_count * 32);
            dmma2_prepare_kernel<<<pg, 256>>>(pl->dm2, t, pl->d_prepared2 + (size_t)g * pl->dm2.gstride);
        }
        if (want_small) {
            const double *h = src.host(g);
            if (h) small_t.emplace_back(h, h + gd.size);
        }
        // the source reuses its buffers for the next tensor
        if (cudaDeviceSynchronize() != cudaSuccess || cudaGetLastError() != cudaSuccess) {
            delete pl;
            return fail(PCB_ECUDA, "preparing tensor %d on the device failed", g);
        }
    }
    // small tensors: the constant-bank evaluator of pcb_piecewise.cu through a one-piece spline plan
    // (no knots); kept only if that plan really runs from the bank
    if (want_small && (int)small_t.size() == G) {
        std::vector<int32_t> no_knots(D, 0);
        std::vector<const double *> ptrs;
        for (auto &v : small_t) ptrs.push_back(v.data());
        if (pcb_spline_plan_create(dev, D, no_knots.data(), nullptr, 1, n, nodes_cat, weights_cat, G,
                                   ptrs.data(), &pl->small) != PCB_OK)
            pl->small = nullptr;
        else if (!spline_plan_uses_bank(pl->small)) {
            pcb_plan_destroy(pl->small);
            pl->small = nullptr;
        }
    }
    *plan = pl;
    return PCB_OK;
}

static int full_check_shape(int D, const int32_t *n, int G) {
    PCB_REQUIRE(D >= 1 && D <= GRID_MAXD, "num_dimensions %d outside [1, %d]", D, GRID_MAXD);
    PCB_REQUIRE(G >= 1 && G <= 64, "number of derivative tensors %d outside [1, 64]", G);
    for (int d = 0; d < D; ++d) PCB_REQUIRE(n[d] >= 1, "n_nodes[%d] must be >= 1", d);
    return PCB_OK;
}

}  // namespace pcb

extern "C" PCB_API int pcb_full_plan_create(int dev, int D, const int32_t *n, const double *nodes_cat,
                                    const double *weights_cat, int G,
                                    const double *const *tensors_host, void **plan) {
    PCB_REQUIRE(plan && n && nodes_cat && weights_cat && tensors_host, "null argument");
    if (int rc = full_check_shape(D, n, G)) return rc;
    DeviceGuard guard(dev);
    if (!guard.ok) return fail(PCB_ECUDA, "cannot select device %d", dev);
    long long size = 1;
    for (int d = 0; d < D; ++d) size *= n[d];
    HostTensorSource src(tensors_host, size);
    return full_plan_build(dev, D, n, nodes_cat, weights_cat, G, src, plan);
}

extern "C" PCB_API int pcb_full_plan_create_from_values(int dev, int D, const int32_t *n,
                                                        const double *nodes_cat,
                                                        const double *weights_cat,
                                                        const double *diffmats_cat,
                                                        const double *values_host, int G,
                                                        const int32_t *orders, void **plan) {
    PCB_REQUIRE(plan && n && nodes_cat && weights_cat && diffmats_cat && values_host && orders,
                "null argument");
    if (int rc = full_check_shape(D, n, G)) return rc;
    for (int d = 0; d < D; ++d) PCB_REQUIRE(n[d] <= 64, "n_nodes[%d] = %d exceeds 64", d, n[d]);
    for (int i = 0; i < G * D; ++i)
        PCB_REQUIRE(orders[i] >= 0 && orders[i] <= 8, "derivative order %d outside [0, 8]", orders[i]);
    DeviceGuard guard(dev);
    if (!guard.ok) return fail(PCB_ECUDA, "cannot select device %d", dev);
    long long size = 1;
    for (int d = 0; d < D; ++d) size *= n[d];
    DerivTensorSource src(dev, D, n, diffmats_cat, values_host, orders, size);
    if (!src.ok) return PCB_ENOMEM;
    return full_plan_build(dev, D, n, nodes_cat, weights_cat, G, src, plan);
}

template <int KB>
static int launch_dmma(FullPlan *pl, const double *d_points, int64_t N, double *d_out, cudaStream_t st) {
    PCB_CUDA(allow_dynamic_smem(full_dmma_kernel<KB>, pl->dm_smem, pl->smem_optin));
    const int64_t ntiles = (N + DM_QT - 1) / DM_QT;
    const int grid = (int)(ntiles < pl->sm_count ? ntiles : pl->sm_count);
    full_dmma_kernel<KB><<<grid, DM_THREADS, pl->dm_smem, st>>>(pl->dm, pl->d_nodes, pl->d_weights,
                                                              pl->d_prepared, d_points, N, d_out);
    g_launches.fetch_add(1);
    PCB_CUDA(cudaGetLastError());
    return PCB_OK;
}

template <int KB>
static int launch_dmma2(FullPlan *pl, const double *d_points, int64_t N, double *d_out, cudaStream_t st) {
    PCB_CUDA(allow_dynamic_smem(full_dmma2_kernel<KB>, pl->dm2_smem, pl->smem_optin));
    const int64_t ntiles = (N + DM2_QT - 1) / DM2_QT;
    const int grid = (int)(ntiles < pl->sm_count ? ntiles : pl->sm_count);
    full_dmma2_kernel<KB><<<grid, DM2_THREADS, pl->dm2_smem, st>>>(pl->dm2, pl->d_nodes, pl->d_weights,
                                                                 pl->d_prepared2, d_points, N, d_out);
    g_launches.fetch_add(1);
    PCB_CUDA(cudaGetLastError());
    return PCB_OK;
}

extern "C" PCB_API int pcb_full_eval(void *plan, const double *d_points, int64_t N, double *d_out, int algo,
                             void *stream) {
    FullPlan *pl = static_cast<FullPlan *>(plan);
    PCB_REQUIRE(pl && pl->kind == PLAN_FULL, "not a full-tensor plan");
    PCB_REQUIRE(N >= 0, "negative N");
    PCB_REQUIRE(algo >= 0 && algo <= 3, "algo %d not available", algo);
    if (N == 0) return PCB_OK;
    PCB_REQUIRE(d_points && d_out, "null device pointer");
    DeviceGuard guard(pl->dev);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (algo == 2 && !pl->dmma_ok)
        return fail(PCB_EUNSUPPORTED, "tensor-core path not available for this shape");
    if (algo == 3 && !pl->dmma2_ok)
        return fail(PCB_EUNSUPPORTED, "joint-K tensor-core path not available for this shape");
    if (algo == 3 || (algo == 0 && pl->dmma2_ok && pl->gd.size >= 4096 && N >= 64)) {
        switch (pl->dm2.KB) {
            case 16: return launch_dmma2<16>(pl, d_points, N, d_out, st);
            case 21: return launch_dmma2<21>(pl, d_points, N, d_out, st);
            case 25: return launch_dmma2<25>(pl, d_points, N, d_out, st);
            case 31: return launch_dmma2<31>(pl, d_points, N, d_out, st);
            case 36: return launch_dmma2<36>(pl, d_points, N, d_out, st);
        }
        return fail(PCB_EUNSUPPORTED, "joint-K tensor-core path: no instantiation for %d K-blocks", pl->dm2.KB);
    }
    // auto: the tensor-core GEMM once the tensor is big enough to amortise a 256-query tile
    const bool use_dmma = algo == 2 || (algo == 0 && pl->dmma_ok && pl->gd.size >= 4096 && N >= 64);
    if (use_dmma) {
        switch (pl->dm.KB) {
            case 1: return launch_dmma<1>(pl, d_points, N, d_out, st);
            case 2: return launch_dmma<2>(pl, d_points, N, d_out, st);
            case 3: return launch_dmma<3>(pl, d_points, N, d_out, st);
            case 4: return launch_dmma<4>(pl, d_points, N, d_out, st);
            case 5: return launch_dmma<5>(pl, d_points, N, d_out, st);
            case 6: return launch_dmma<6>(pl, d_points, N, d_out, st);
            case 7: return launch_dmma<7>(pl, d_points, N, d_out, st);
            case 8: return launch_dmma<8>(pl, d_points, N, d_out, st);
            default:
                if (algo == 2)
                    return fail(PCB_EUNSUPPORTED, "last axis with %d nodes exceeds the register "
                                "budget of the tensor-core path", pl->gd.n[pl->gd.D - 1]);
        }
    }
    if (algo == 0 && pl->small) return pcb_spline_eval(pl->small, d_points, N, d_out, nullptr, stream);
    const size_t smem = (size_t)pl->gd.sum_n * FULL_FMA_THREADS * sizeof(double);
    if (smem > (size_t)pl->smem_optin)
        return fail(PCB_EUNSUPPORTED, "weight rows of %d nodes do not fit in shared memory",
                    pl->gd.sum_n);
    const void *kernel = GRID_KERNEL_TABLE(full_fma_kernel, pl->GB, grid_pick_dm(pl->gd.D));
    PCB_CUDA(allow_dynamic_smem(kernel, smem, pl->smem_optin));
    int per_sm = 0;
    PCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, FULL_FMA_THREADS, smem));
    const int64_t want = (N + FULL_FMA_THREADS - 1) / FULL_FMA_THREADS;
    const int64_t cap = (int64_t)pl->sm_count * (per_sm > 0 ? per_sm : 1);
    const int grid = (int)(want < cap ? want : cap);
    int G = pl->G;
    void *args[] = {(void *)&pl->gd, (void *)&G, (void *)&pl->d_nodes, (void *)&pl->d_weights,
                    (void *)&pl->d_tensors, (void *)&d_points, (void *)&N, (void *)&d_out};
    PCB_CUDA(cudaLaunchKernel(kernel, dim3(grid), dim3(FULL_FMA_THREADS), args, smem, st));
    g_launches.fetch_add(1);
    PCB_CUDA(cudaGetLastError());
    return PCB_OK;
}
