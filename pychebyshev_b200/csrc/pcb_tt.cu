// ChebyshevTT batch evaluation on sm_100a: batched core-chain contraction.
//
// Replaces ChebyshevTT.eval_batch (reference tensor_train.py:2217-2265) and loops of
// ChebyshevTT.eval_multi (tensor_train.py:2267-2463).
//
// Design: one query per THREAD slot (QPT slots per thread), coefficient cores
// resident in shared memory and read with warp-uniform (broadcast) 128-bit LDS, so one LDS feeds
// 2*QPT DFMAs in each of 32 lanes.  Per dimension k a thread updates its chain vector
//     v'[l] = sum_i sum_j (v[i] * T_j(s_k)) * G_k[i, j, l]
// with the accumulators v'[l] in registers (compile-time chunk width W), c = v[i]*T_j advanced by
// the Chebyshev three-term recurrence in registers, and v[] parked in a conflict-free per-thread
// shared-memory column.  No cross-lane traffic, no atomics.
#include "pcb_common.cuh"

namespace pcb {

constexpr int TT_THREADS = 256;
constexpr int TT_LCMAX = 16;  // widest register accumulator chunk (doubles per query slot)
constexpr int TT_MAX_ACTIVE = 3;
constexpr int TT_MAX_G = 16;

enum TTMode : int { TT_RESIDENT = 0, TT_STREAM = 1, TT_GLOBAL = 2 };

struct TTParams {
    int D;
    int rmaxp;  // max padded rank (even)
    int total;  // doubles in the packed forward cores
    int maxcore;
    int n[PCB_MAX_DIMS];
    int r[PCB_MAX_DIMS + 1];
    int rp[PCB_MAX_DIMS];    // r[k+1] rounded up to even: row stride of packed core k
    int off[PCB_MAX_DIMS];   // offset (doubles) of packed core k: [i][j][rp]
    int perm[PCB_MAX_DIMS];  // storage position k -> user column
    int totalT;              // doubles in the packed transposed cores (stored after the forward ones)
    int rpT[PCB_MAX_DIMS];   // r[k] rounded up to even: row stride of transposed core k
    int offT[PCB_MAX_DIMS];  // offset (doubles, from the start of the buffer) of transposed core k:
                             // [l][j][rpT] with GT[l][j][i] = G[i][j][l]
    double lo[PCB_MAX_DIMS];
    double hi[PCB_MAX_DIMS];
};

// One output row of pcb_tt_eval_fd, storage frame.
struct TTFdRow {
    int m;                     // active dims (order > 0)
    int dim[TT_MAX_ACTIVE];    // storage positions, ascending
    int ord[TT_MAX_ACTIVE];    // 1 or 2
};

struct TTFdProgram {
    int G;
    TTFdRow row[TT_MAX_G];
};

// pcb_tt_eval_fd algo 2: every row differentiates at most one dim.
struct TTSharedProgram {
    int G;
    int n_slots;                 // distinct differentiated storage dims, ascending
    int slot_dim[TT_MAX_G];
    int row_slot[TT_MAX_G];      // -1 value row, else index into slot_dim
    int row_ord[TT_MAX_G];       // 1 or 2
};

struct TTPlan : PlanBase {
    TTParams P;
    double *d_cores = nullptr;
    int mode = TT_RESIDENT;
    int mode_shared = TT_RESIDENT;
    int pingpong = 0;
    ~TTPlan() override {
        if (d_cores) cudaFree(d_cores);
    }
};

// ---------------------------------------------------------------------------------------------
// device code
// ---------------------------------------------------------------------------------------------

// reference tensor_train.py:2254 -- exactly this operation order
__device__ __forceinline__ double tt_scale(double x, double a, double b) {
    return 2.0 * (x - a) / (b - a) - 1.0;
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
    for (int i = 0; i < r_in; ++i) {
        double c0[QPT], c1[QPT];
#pragma unroll
        for (int qq = 0; qq < QPT; ++qq) {
            const double vi = v_in[(i * QPT + qq) * vstride];
            c0[qq] = vi;           // v[i] * T_0
            c1[qq] = vi * s[qq];   // v[i] * T_1
        }
        const double *gi = g + (size_t)i * n * rp;
#pragma unroll 2
        for (int j = 0; j < n; ++j) {
#pragma unroll
            for (int l = 0; l < W; l += 2) {
                const double2 gg = *reinterpret_cast<const double2 *>(gi + l);
#pragma unroll
                for (int qq = 0; qq < QPT; ++qq) {
                    acc[qq][l] = fma(c0[qq], gg.x, acc[qq][l]);
                    acc[qq][l + 1] = fma(c0[qq], gg.y, acc[qq][l + 1]);
                }
            }
#pragma unroll
            for (int qq = 0; qq < QPT; ++qq) {
                const double c2 = fma(twos[qq], c1[qq], -c0[qq]);  // T_{j+2} = 2 s T_{j+1} - T_j
                c0[qq] = c1[qq];
                c1[qq] = c2;
            }
            gi += rp;
        }
    }
#pragma unroll
    for (int l = 0; l < W; ++l)
#pragma unroll
        for (int qq = 0; qq < QPT; ++qq) v_out[(l * QPT + qq) * vstride] = acc[qq][l];
}

template <int QPT>
__device__ __forceinline__ void tt_apply_core(const double *__restrict__ g, int rp, int r_in, int n,
                                              const double *v_in, double *v_out, int vstride,
                                              const double (&s)[QPT]) {
    for (int l0 = 0; l0 < rp; l0 += TT_LCMAX) {
        const int w = min(TT_LCMAX, rp - l0);
        const double *gc = g + l0;
        double *vo = v_out + (size_t)l0 * QPT * vstride;
        switch (w) {
            case 2: tt_chunk<2, QPT>(gc, rp, r_in, n, v_in, vstride, s, vo); break;
            case 4: tt_chunk<4, QPT>(gc, rp, r_in, n, v_in, vstride, s, vo); break;
            case 6: tt_chunk<6, QPT>(gc, rp, r_in, n, v_in, vstride, s, vo); break;
            case 8: tt_chunk<8, QPT>(gc, rp, r_in, n, v_in, vstride, s, vo); break;
            case 10: tt_chunk<10, QPT>(gc, rp, r_in, n, v_in, vstride, s, vo); break;
            case 12: tt_chunk<12, QPT>(gc, rp, r_in, n, v_in, vstride, s, vo); break;
            case 14: tt_chunk<14, QPT>(gc, rp, r_in, n, v_in, vstride, s, vo); break;
            default: tt_chunk<16, QPT>(gc, rp, r_in, n, v_in, vstride, s, vo); break;
        }
    }
}

// Shared-memory carve-up (doubles), identical on host and device.
struct TTSmem {
    int cores;  // offset of the core area
    int v;      // offset of the chain-vector area: [vbufs][rmaxp][QPT][threads]
    int pts;    // offset of the staged query tile: [threads*QPT][D]
    int total;  // doubles
};

__host__ __device__ inline TTSmem tt_smem_layout(const TTParams &P, int mode, int pingpong, int qpt,
                                                 int threads) {
    TTSmem L;
    L.cores = 0;
    const int core_area = mode == TT_RESIDENT ? P.total : (mode == TT_STREAM ? P.maxcore : 0);
    L.v = (core_area + 1) & ~1;
    L.pts = L.v + (pingpong ? 2 : 1) * P.rmaxp * qpt * threads;
    L.total = L.pts + threads * qpt * P.D;
    L.total = (L.total + 1) & ~1;
    return L;
}

// Full chain for the QPT query slots of this thread.  `od/ox` override the coordinate of up to
// `m` storage dims (finite-difference stencil points).  Returns the values in res[].
template <int QPT, int MODE>
__device__ __forceinline__ void tt_chain(const TTParams &P, const double *__restrict__ g_cores,
                                         double *smem, const TTSmem &L, int pingpong, int m,
                                         const int *od, const double (*ox)[QPT], double (&res)[QPT]) {
    const int tid = threadIdx.x;
    const int vstride = blockDim.x;
    double *vbuf0 = smem + L.v + tid;
    double *vbuf1 = vbuf0 + (pingpong ? P.rmaxp * QPT * vstride : 0);
    const double *spts = smem + L.pts;
#pragma unroll
    for (int qq = 0; qq < QPT; ++qq) vbuf0[qq * vstride] = 1.0;
    double *vin = vbuf0, *vout = vbuf1;
    for (int k = 0; k < P.D; ++k) {
        double s[QPT];
        const double a = P.lo[k], b = P.hi[k];
#pragma unroll
        for (int qq = 0; qq < QPT; ++qq) {
            double x = spts[(qq * vstride + tid) * P.D + P.perm[k]];
            for (int t = 0; t < m; ++t)
                if (od[t] == k) x = ox[t][qq];
            s[qq] = tt_scale(x, a, b);
        }
        const double *g;
        if (MODE == TT_RESIDENT) {
            g = smem + L.cores + P.off[k];
        } else if (MODE == TT_STREAM) {
            __syncthreads();  // previous core fully consumed
            const int cnt = P.r[k] * P.n[k] * P.rp[k];
            const double2 *src = reinterpret_cast<const double2 *>(g_cores + P.off[k]);
            double2 *dst = reinterpret_cast<double2 *>(smem + L.cores);
            for (int e = tid; e < cnt / 2; e += vstride) dst[e] = src[e];
            __syncthreads();
            g = smem + L.cores;
        } else {
            g = g_cores + P.off[k];
        }
        tt_apply_core<QPT>(g, P.rp[k], P.r[k], P.n[k], vin, vout, vstride, s);
        double *t = vin;
        vin = vout;
        vout = t;
    }
#pragma unroll
    for (int qq = 0; qq < QPT; ++qq) res[qq] = vin[qq * vstride];
}

// Cooperative, coalesced staging of one query tile (rows [q0, q0+rows) of the (N, D) table).
__device__ __forceinline__ void tt_stage_points(const double *__restrict__ pts, int64_t q0, int rows,
                                                int tile_rows, int D, double *spts) {
    const int cnt = rows * D;
    const double *src = pts + q0 * D;
    for (int e = threadIdx.x; e < cnt; e += blockDim.x) spts[e] = __ldg(src + e);
    // tail tile: replicate the last valid row so every slot computes something finite
    for (int e = cnt + threadIdx.x; e < tile_rows * D; e += blockDim.x)
        spts[e] = __ldg(src + (size_t)(rows - 1) * D + (e % D));
}

template <int QPT, int MODE>
__device__ __forceinline__ void tt_load_resident(const TTParams &P, const double *__restrict__ cores,
                                                 double *smem, const TTSmem &L) {
    if (MODE == TT_RESIDENT) {
        const double2 *src = reinterpret_cast<const double2 *>(cores);
        double2 *dst = reinterpret_cast<double2 *>(smem + L.cores);
        for (int e = threadIdx.x; e < P.total / 2; e += blockDim.x) dst[e] = src[e];
    }
}

template <int QPT, int MODE>
__global__ void __launch_bounds__(TT_THREADS)
tt_value_kernel(const __grid_constant__ TTParams P, const double *__restrict__ cores,
                const double *__restrict__ pts, int64_t N, double *__restrict__ out, int pingpong) {
    extern __shared__ __align__(16) double smem[];
    const TTSmem L = tt_smem_layout(P, MODE, pingpong, QPT, blockDim.x);
    tt_load_resident<QPT, MODE>(P, cores, smem, L);
    const int tile_rows = blockDim.x * QPT;
    const int64_t ntiles = (N + tile_rows - 1) / tile_rows;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t q0 = tile * tile_rows;
        const int rows = (int)min((int64_t)tile_rows, N - q0);
        __syncthreads();  // previous tile's points no longer needed (and resident cores landed)
        tt_stage_points(pts, q0, rows, tile_rows, P.D, smem + L.pts);
        __syncthreads();
        double res[QPT];
        tt_chain<QPT, MODE>(P, cores, smem, L, pingpong, 0, nullptr, nullptr, res);
#pragma unroll
        for (int qq = 0; qq < QPT; ++qq) {
            const int64_t q = q0 + qq * blockDim.x + threadIdx.x;
            if (q < N) out[q] = res[qq];
        }
    }
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

// algo 1: one full chain per stencil point, the reference's own evaluation count and formulas.
template <int QPT, int MODE>
__global__ void __launch_bounds__(TT_THREADS)
tt_fd_general_kernel(const __grid_constant__ TTParams P, const __grid_constant__ TTFdProgram prog,
                     const double *__restrict__ cores, const double *__restrict__ pts, int64_t N,
                     double *__restrict__ out, int pingpong) {
    extern __shared__ __align__(16) double smem[];
    const TTSmem L = tt_smem_layout(P, MODE, pingpong, QPT, blockDim.x);
    tt_load_resident<QPT, MODE>(P, cores, smem, L);
    const int tile_rows = blockDim.x * QPT;
    const int64_t ntiles = (N + tile_rows - 1) / tile_rows;
    const int G = prog.G;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t q0 = tile * tile_rows;
        const int rows = (int)min((int64_t)tile_rows, N - q0);
        __syncthreads();
        tt_stage_points(pts, q0, rows, tile_rows, P.D, smem + L.pts);
        __syncthreads();
        const double *spts = smem + L.pts;
        for (int g = 0; g < G; ++g) {
            const TTFdRow row = prog.row[g];
            double result[QPT];
            if (row.m == 0) {
                tt_chain<QPT, MODE>(P, cores, smem, L, pingpong, 0, nullptr, nullptr, result);
            } else {
                double h[TT_MAX_ACTIVE];
                double ctr[TT_MAX_ACTIVE][QPT];
                for (int t = 0; t < row.m; ++t) {
                    const int k = row.dim[t];
                    const double a = P.lo[k], b = P.hi[k];
                    h[t] = (b - a) * 1e-4;  // tensor_train.py:2356-2359
#pragma unroll
                    for (int qq = 0; qq < QPT; ++qq)
                        ctr[t][qq] = tt_nudge(
                            spts[(qq * blockDim.x + threadIdx.x) * P.D + P.perm[k]], a, b, h[t]);
                }
                const int m = row.m;
                double ox[TT_MAX_ACTIVE][QPT];
                if (m == 2 && row.ord[0] == 1 && row.ord[1] == 1) {
                    // tensor_train.py:2405-2426: (f_pp - f_pm - f_mp + f_mm) / (4 h1 h2)
                    double f[4][QPT];
#pragma unroll 1
                    for (int e = 0; e < 4; ++e) {
                        const double s1 = (e & 2) ? -h[0] : h[0];
                        const double s2 = (e & 1) ? -h[1] : h[1];
#pragma unroll
                        for (int qq = 0; qq < QPT; ++qq) {
                            ox[0][qq] = ctr[0][qq] + s1;
                            ox[1][qq] = ctr[1][qq] + s2;
                        }
                        tt_chain<QPT, MODE>(P, cores, smem, L, pingpong, 2, row.dim, ox, f[e]);
                    }
#pragma unroll
                    for (int qq = 0; qq < QPT; ++qq)
                        result[qq] = (f[0][qq] - f[1][qq] - f[2][qq] + f[3][qq]) / (4.0 * h[0] * h[1]);
                } else {
                    // nested stencils (tensor_train.py:2372-2403 for m == 1, 2428-2463 otherwise):
                    // the first active dim is the outermost level.
                    double f0[3][QPT], f1[3][QPT], f2[3][QPT];
#pragma unroll 1
                    for (int e0 = 0; e0 < 3; ++e0) {
                        if (e0 == 1 && row.ord[0] == 1) continue;
#pragma unroll
                        for (int qq = 0; qq < QPT; ++qq)
                            ox[0][qq] = e0 == 0 ? ctr[0][qq] + h[0]
                                                : (e0 == 1 ? ctr[0][qq] : ctr[0][qq] - h[0]);
                        if (m == 1) {
                            tt_chain<QPT, MODE>(P, cores, smem, L, pingpong, 1, row.dim, ox, f0[e0]);
                            continue;
                        }
#pragma unroll 1
                        for (int e1 = 0; e1 < 3; ++e1) {
                            if (e1 == 1 && row.ord[1] == 1) continue;
#pragma unroll
                            for (int qq = 0; qq < QPT; ++qq)
                                ox[1][qq] = e1 == 0 ? ctr[1][qq] + h[1]
                                                    : (e1 == 1 ? ctr[1][qq] : ctr[1][qq] - h[1]);
                            if (m == 2) {
                                tt_chain<QPT, MODE>(P, cores, smem, L, pingpong, 2, row.dim, ox,
                                                    f1[e1]);
                                continue;
                            }
#pragma unroll 1
                            for (int e2 = 0; e2 < 3; ++e2) {
                                if (e2 == 1 && row.ord[2] == 1) continue;
#pragma unroll
                                for (int qq = 0; qq < QPT; ++qq)
                                    ox[2][qq] = e2 == 0 ? ctr[2][qq] + h[2]
                                                        : (e2 == 1 ? ctr[2][qq] : ctr[2][qq] - h[2]);
                                tt_chain<QPT, MODE>(P, cores, smem, L, pingpong, 3, row.dim, ox,
                                                    f2[e2]);
                            }
#pragma unroll
                            for (int qq = 0; qq < QPT; ++qq)
                                f1[e1][qq] = tt_fd_reduce(row.ord[2], f2[0][qq], f2[1][qq],
                                                          f2[2][qq], h[2]);
                        }
#pragma unroll
                        for (int qq = 0; qq < QPT; ++qq)
                            f0[e0][qq] =
                                tt_fd_reduce(row.ord[1], f1[0][qq], f1[1][qq], f1[2][qq], h[1]);
                    }
#pragma unroll
                    for (int qq = 0; qq < QPT; ++qq)
                        result[qq] = tt_fd_reduce(row.ord[0], f0[0][qq], f0[1][qq], f0[2][qq], h[0]);
                }
            }
#pragma unroll
            for (int qq = 0; qq < QPT; ++qq) {
                const int64_t q = q0 + qq * blockDim.x + threadIdx.x;
                if (q < N) out[q * G + g] = result[qq];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// algo 2: shared left/right partial products
//
// With every stencil point differing from the query in ONE coordinate a, the interpolant along
// that coordinate is the degree n_a-1 Chebyshev series  f(x_a) = sum_j y_j T_j(s(x_a)),
//   y_j = L_a . G_a[:, j, :] . R_{a+1},
// L_a = product of the contracted cores left of a, R_{a+1} = product right of a, both at the
// query's own coordinates.  One left sweep (continued from slot to slot), one right sweep per
// differentiated dim and one coefficient pass give every stencil value of that dim for n_a FMAs
// each, instead of a full chain per stencil point.  The finite-difference formulas, step h and
// boundary nudge are the reference's (tensor_train.py:2356-2403).
// ---------------------------------------------------------------------------------------------

struct TTSharedSmem {
    int cores;
    int v;      // [nbuf][rmaxp][QPT][threads]
    int total;
};

__host__ __device__ inline TTSharedSmem tt_shared_smem_layout(const TTParams &P, int mode,
                                                              int pingpong, int qpt, int threads) {
    TTSharedSmem L;
    L.cores = 0;
    const int core_area =
        mode == TT_RESIDENT ? P.total + P.totalT : (mode == TT_STREAM ? P.maxcore : 0);
    L.v = (core_area + 1) & ~1;
    L.total = L.v + (2 + (pingpong ? 1 : 0)) * P.rmaxp * qpt * threads;
    L.total = (L.total + 1) & ~1;
    return L;
}

// Fetch one packed core (forward or transposed) for the whole CTA.
template <int MODE>
__device__ __forceinline__ const double *tt_core_ptr(const double *__restrict__ g_cores, double *smem,
                                                     int cores_off, int off, int count) {
    if (MODE == TT_RESIDENT) return smem + cores_off + off;
    if (MODE == TT_STREAM) {
        __syncthreads();
        const double2 *src = reinterpret_cast<const double2 *>(g_cores + off);
        double2 *dst = reinterpret_cast<double2 *>(smem + cores_off);
        for (int e = threadIdx.x; e < count / 2; e += blockDim.x) dst[e] = src[e];
        __syncthreads();
        return smem + cores_off;
    }
    return g_cores + off;
}

// One chunk of the coefficient pass: y += sum_l (sum_i L[i] g[i][j][l0+l]) * R[l0+l]
template <int W, int QPT>
__device__ __forceinline__ void tt_coeff_chunk(const double *__restrict__ gj, int istride, int r_in,
                                               const double *vL, const double *vR, int vstride,
                                               double (&y)[QPT]) {
    double acc[QPT][W];
#pragma unroll
    for (int qq = 0; qq < QPT; ++qq)
#pragma unroll
        for (int l = 0; l < W; ++l) acc[qq][l] = 0.0;
#pragma unroll 2
    for (int i = 0; i < r_in; ++i) {
        double li[QPT];
#pragma unroll
        for (int qq = 0; qq < QPT; ++qq) li[qq] = vL[(i * QPT + qq) * vstride];
        const double *gi = gj + (size_t)i * istride;
#pragma unroll
        for (int l = 0; l < W; l += 2) {
            const double2 gg = *reinterpret_cast<const double2 *>(gi + l);
#pragma unroll
            for (int qq = 0; qq < QPT; ++qq) {
                acc[qq][l] = fma(li[qq], gg.x, acc[qq][l]);
                acc[qq][l + 1] = fma(li[qq], gg.y, acc[qq][l + 1]);
            }
        }
    }
#pragma unroll
    for (int l = 0; l < W; ++l)
#pragma unroll
        for (int qq = 0; qq < QPT; ++qq) y[qq] = fma(acc[qq][l], vR[(l * QPT + qq) * vstride], y[qq]);
}

template <int QPT>
__device__ __forceinline__ void tt_coeff_row(const double *__restrict__ gj, int rp, int istride,
                                             int r_in, const double *vL, const double *vR,
                                             int vstride, double (&y)[QPT]) {
    for (int l0 = 0; l0 < rp; l0 += TT_LCMAX) {
        const int w = min(TT_LCMAX, rp - l0);
        const double *gc = gj + l0;
        const double *vr = vR + (size_t)l0 * QPT * vstride;
        switch (w) {
            case 2: tt_coeff_chunk<2, QPT>(gc, istride, r_in, vL, vr, vstride, y); break;
            case 4: tt_coeff_chunk<4, QPT>(gc, istride, r_in, vL, vr, vstride, y); break;
            case 6: tt_coeff_chunk<6, QPT>(gc, istride, r_in, vL, vr, vstride, y); break;
            case 8: tt_coeff_chunk<8, QPT>(gc, istride, r_in, vL, vr, vstride, y); break;
            case 10: tt_coeff_chunk<10, QPT>(gc, istride, r_in, vL, vr, vstride, y); break;
            case 12: tt_coeff_chunk<12, QPT>(gc, istride, r_in, vL, vr, vstride, y); break;
            case 14: tt_coeff_chunk<14, QPT>(gc, istride, r_in, vL, vr, vstride, y); break;
            default: tt_coeff_chunk<16, QPT>(gc, istride, r_in, vL, vr, vstride, y); break;
        }
    }
}

template <int QPT, int MODE>
__global__ void __launch_bounds__(TT_THREADS)
tt_fd_shared_kernel(const __grid_constant__ TTParams P, const __grid_constant__ TTSharedProgram prog,
                    const double *__restrict__ cores, const double *__restrict__ pts, int64_t N,
                    double *__restrict__ out, int pingpong) {
    extern __shared__ __align__(16) double smem[];
    const TTSharedSmem L = tt_shared_smem_layout(P, MODE, pingpong, QPT, blockDim.x);
    const int tid = threadIdx.x;
    const int vstride = blockDim.x;
    const int D = P.D, G = prog.G;
    if (MODE == TT_RESIDENT) {
        const double2 *src = reinterpret_cast<const double2 *>(cores);
        double2 *dst = reinterpret_cast<double2 *>(smem + L.cores);
        for (int e = tid; e < (P.total + P.totalT) / 2; e += vstride) dst[e] = src[e];
        __syncthreads();
    }
    const int vsz = P.rmaxp * QPT * vstride;
    const int tile_rows = vstride * QPT;
    const int64_t ntiles = (N + tile_rows - 1) / tile_rows;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t q0 = tile * tile_rows;
        // rows of this thread's query slots (tail: clamp, results are not stored)
        const double *xrow[QPT];
#pragma unroll
        for (int qq = 0; qq < QPT; ++qq) {
            int64_t q = q0 + qq * vstride + tid;
            if (q >= N) q = N - 1;
            xrow[qq] = pts + q * D;
        }
        double *vL = smem + L.v + tid;
        double *vR = vL + vsz;
        double *vT = pingpong ? vR + vsz : nullptr;
#pragma unroll
        for (int qq = 0; qq < QPT; ++qq) vL[qq * vstride] = 1.0;
        int lpos = 0;  // vL holds the left product over dims [0, lpos)
        for (int t = 0; t < prog.n_slots; ++t) {
            const int a = prog.slot_dim[t];
            double s[QPT];
            // ---- right sweep: vR = M_{a+1} ... M_{D-1} (applied right to left) --------------
#pragma unroll
            for (int qq = 0; qq < QPT; ++qq) vR[qq * vstride] = 1.0;
            for (int k = D - 1; k > a; --k) {
#pragma unroll
                for (int qq = 0; qq < QPT; ++qq)
                    s[qq] = tt_scale(__ldg(xrow[qq] + P.perm[k]), P.lo[k], P.hi[k]);
                const double *g = tt_core_ptr<MODE>(cores, smem, L.cores, P.offT[k],
                                                    P.r[k + 1] * P.n[k] * P.rpT[k]);
                double *vo = pingpong ? vT : vR;
                tt_apply_core<QPT>(g, P.rpT[k], P.r[k + 1], P.n[k], vR, vo, vstride, s);
                if (pingpong) {
                    vT = vR;
                    vR = vo;
                }
            }
            // ---- left sweep continues up to a ---------------------------------------------------
            for (int k = lpos; k < a; ++k) {
#pragma unroll
                for (int qq = 0; qq < QPT; ++qq)
                    s[qq] = tt_scale(__ldg(xrow[qq] + P.perm[k]), P.lo[k], P.hi[k]);
                const double *g = tt_core_ptr<MODE>(cores, smem, L.cores, P.off[k],
                                                    P.r[k] * P.n[k] * P.rp[k]);
                double *vo = pingpong ? vT : vL;
                tt_apply_core<QPT>(g, P.rp[k], P.r[k], P.n[k], vL, vo, vstride, s);
                if (pingpong) {
                    vT = vL;
                    vL = vo;
                }
            }
            lpos = a;
            // ---- stencil abscissae of dim a (reference _fd_step / _nudge_point) ---------------
            const double lo = P.lo[a], hi = P.hi[a];
            const double h = (hi - lo) * 1e-4;
            double tc[4][QPT], tn[4][QPT], tw[4][QPT], f[4][QPT];
#pragma unroll
            for (int qq = 0; qq < QPT; ++qq) {
                const double x = __ldg(xrow[qq] + P.perm[a]);
                const double c = tt_nudge(x, lo, hi, h);
                const double sm[4] = {tt_scale(x, lo, hi), tt_scale(c, lo, hi),
                                      tt_scale(c + h, lo, hi), tt_scale(c - h, lo, hi)};
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    tc[m][qq] = 1.0;
                    tn[m][qq] = sm[m];
                    tw[m][qq] = 2.0 * sm[m];
                    f[m][qq] = 0.0;
                }
            }
            // ---- coefficient pass: y_j = L . G_a[:, j, :] . R, f_m += y_j T_j(s_m) -------------
            const double *ga = tt_core_ptr<MODE>(cores, smem, L.cores, P.off[a],
                                                 P.r[a] * P.n[a] * P.rp[a]);
            const int rp = P.rp[a], na = P.n[a];
            for (int j = 0; j < na; ++j) {
                double y[QPT];
#pragma unroll
                for (int qq = 0; qq < QPT; ++qq) y[qq] = 0.0;
                tt_coeff_row<QPT>(ga + (size_t)j * rp, rp, na * rp, P.r[a], vL, vR, vstride, y);
#pragma unroll
                for (int m = 0; m < 4; ++m)
#pragma unroll
                    for (int qq = 0; qq < QPT; ++qq) {
                        f[m][qq] = fma(y[qq], tc[m][qq], f[m][qq]);
                        const double t2 = fma(tw[m][qq], tn[m][qq], -tc[m][qq]);
                        tc[m][qq] = tn[m][qq];
                        tn[m][qq] = t2;
                    }
            }
            // ---- outputs owned by this slot --------------------------------------------------------
            for (int g = 0; g < G; ++g) {
                const int rs = prog.row_slot[g];
                if (!(rs == t || (rs < 0 && t == 0))) continue;
#pragma unroll
                for (int qq = 0; qq < QPT; ++qq) {
                    double res;
                    if (rs < 0)
                        res = f[0][qq];
                    else
                        res = tt_fd_reduce(prog.row_ord[g], f[2][qq], f[1][qq], f[3][qq], h);
                    const int64_t q = q0 + qq * vstride + tid;
                    if (q < N) out[q * G + g] = res;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------

template <typename K>
static int tt_launch_cfg(K kernel, const TTPlan *pl, int qpt, int64_t N, int *grid, size_t *smem_bytes) {
    const TTSmem L = tt_smem_layout(pl->P, pl->mode, pl->pingpong, qpt, TT_THREADS);
    *smem_bytes = (size_t)L.total * sizeof(double);
    if (*smem_bytes > (size_t)pl->smem_optin)
        return fail(PCB_EUNSUPPORTED, "TT plan needs %zu B of shared memory per CTA (> %d)",
                    *smem_bytes, pl->smem_optin);
    PCB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*smem_bytes));
    int per_sm = 0;
    PCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, TT_THREADS, *smem_bytes));
    if (per_sm < 1) return fail(PCB_ECUDA, "TT kernel does not fit on an SM");
    const int64_t ntiles = (N + (int64_t)TT_THREADS * qpt - 1) / ((int64_t)TT_THREADS * qpt);
    const int64_t cap = (int64_t)pl->sm_count * per_sm;
    *grid = (int)(ntiles < cap ? ntiles : cap);
    return PCB_OK;
}

static int tt_pick_qpt(const TTPlan *pl) {
    // two query slots per thread halve the LDS traffic per DFMA; fall back to one when the
    // chain vectors would not fit beside the cores.
    const TTSmem L2 = tt_smem_layout(pl->P, pl->mode, pl->pingpong, 2, TT_THREADS);
    return (size_t)L2.total * sizeof(double) <= (size_t)pl->smem_optin ? 2 : 1;
}

}  // namespace pcb

using namespace pcb;

extern "C" PCB_API int pcb_tt_plan_create(int dev, int D, const int32_t *n, const int32_t *ranks,
                                  const double *lo, const double *hi, const int32_t *dim_order,
                                  const double *cores_cat, void **plan) {
    PCB_REQUIRE(plan && n && ranks && lo && hi && cores_cat, "null argument");
    PCB_REQUIRE(D >= 1 && D <= PCB_MAX_DIMS, "num_dimensions %d outside [1, %d]", D, PCB_MAX_DIMS);
    PCB_REQUIRE(ranks[0] == 1 && ranks[D] == 1, "boundary TT ranks must be 1");
    TTPlan *pl = new TTPlan();
    pl->kind = PLAN_TT;
    pl->dev = dev;
    int cc = 0;
    if (int rc = device_props(dev, &pl->sm_count, &pl->smem_optin, &cc)) {
        delete pl;
        return rc;
    }
    TTParams &P = pl->P;
    memset(&P, 0, sizeof(P));
    P.D = D;
    std::vector<char> seen(D, 0);
    int off = 0, rmaxp = 2, maxcore = 0;
    for (int k = 0; k < D; ++k) {
        const int perm = dim_order ? dim_order[k] : k;
        if (n[k] < 1 || ranks[k] < 1 || ranks[k + 1] < 1 || perm < 0 || perm >= D || seen[perm] ||
            !(lo[k] < hi[k])) {
            delete pl;
            return fail(PCB_EINVAL, "invalid TT description at storage dim %d", k);
        }
        seen[perm] = 1;
        P.n[k] = n[k];
        P.r[k] = ranks[k];
        P.rp[k] = round_up(ranks[k + 1], 2);
        P.off[k] = off;
        P.perm[k] = perm;
        P.lo[k] = lo[k];
        P.hi[k] = hi[k];
        const int sz = ranks[k] * n[k] * P.rp[k];
        off += sz;
        if (sz > maxcore) maxcore = sz;
        if (P.rp[k] > rmaxp) rmaxp = P.rp[k];
        if (round_up(ranks[k], 2) > rmaxp) rmaxp = round_up(ranks[k], 2);
        if (P.rp[k] > TT_LCMAX) pl->pingpong = 1;
    }
    P.r[D] = 1;
    P.total = off;
    P.maxcore = maxcore;
    P.rmaxp = rmaxp;

    // transposed copies for right-to-left sweeps: [l][j][i] with i padded to even by zeros
    int offT = off;
    for (int k = 0; k < D; ++k) {
        P.rpT[k] = round_up(ranks[k], 2);
        P.offT[k] = offT;
        const int sz = ranks[k + 1] * n[k] * P.rpT[k];
        offT += sz;
        if (sz > maxcore) maxcore = sz;
    }
    P.totalT = offT - off;
    P.maxcore = maxcore;

    // pack: forward [i][j][l] with l padded to even by zeros, then transposed [l][j][i]
    std::vector<double> packed((size_t)offT, 0.0);
    size_t src = 0;
    for (int k = 0; k < D; ++k) {
        const int r0 = ranks[k], r1 = ranks[k + 1];
        for (int i = 0; i < r0; ++i)
            for (int j = 0; j < n[k]; ++j) {
                double *dst = &packed[(size_t)P.off[k] + ((size_t)i * n[k] + j) * P.rp[k]];
                for (int l = 0; l < r1; ++l) {
                    const double v = cores_cat[src++];
                    dst[l] = v;
                    packed[(size_t)P.offT[k] + ((size_t)l * n[k] + j) * P.rpT[k] + i] = v;
                }
            }
    }
    // placement of the cores: resident in smem if they fit beside one-slot chain vectors
    pl->mode = TT_RESIDENT;
    {
        TTSmem L = tt_smem_layout(P, TT_RESIDENT, pl->pingpong, 1, TT_THREADS);
        if ((size_t)L.total * 8 > (size_t)pl->smem_optin) {
            pl->mode = TT_STREAM;
            L = tt_smem_layout(P, TT_STREAM, pl->pingpong, 1, TT_THREADS);
            if ((size_t)L.total * 8 > (size_t)pl->smem_optin) pl->mode = TT_GLOBAL;
        }
    }
    pl->mode_shared = TT_RESIDENT;
    {
        TTSharedSmem L = tt_shared_smem_layout(P, TT_RESIDENT, pl->pingpong, 1, TT_THREADS);
        if ((size_t)L.total * 8 > (size_t)pl->smem_optin) {
            pl->mode_shared = TT_STREAM;
            L = tt_shared_smem_layout(P, TT_STREAM, pl->pingpong, 1, TT_THREADS);
            if ((size_t)L.total * 8 > (size_t)pl->smem_optin) pl->mode_shared = TT_GLOBAL;
        }
    }
    DeviceGuard guard(dev);
    if (!guard.ok || cudaMalloc(&pl->d_cores, packed.size() * sizeof(double)) != cudaSuccess) {
        delete pl;
        return fail(PCB_ENOMEM, "cudaMalloc of %zu B for TT cores failed on device %d",
                    packed.size() * sizeof(double), dev);
    }
    if (cudaMemcpy(pl->d_cores, packed.data(), packed.size() * sizeof(double),
                   cudaMemcpyHostToDevice) != cudaSuccess) {
        delete pl;
        return fail(PCB_ECUDA, "upload of TT cores failed");
    }
    *plan = pl;
    return PCB_OK;
}

#define TT_LAUNCH_BY_MODE(KERNEL, QPT, ...)                                                       \
    do {                                                                                          \
        int grid = 0;                                                                             \
        size_t smem = 0;                                                                          \
        int rc;                                                                                   \
        switch (pl->mode) {                                                                       \
            case TT_RESIDENT:                                                                     \
                if ((rc = tt_launch_cfg(KERNEL<QPT, TT_RESIDENT>, pl, QPT, N, &grid, &smem))) return rc; \
                KERNEL<QPT, TT_RESIDENT><<<grid, TT_THREADS, smem, st>>>(__VA_ARGS__);            \
                break;                                                                            \
            case TT_STREAM:                                                                       \
                if ((rc = tt_launch_cfg(KERNEL<QPT, TT_STREAM>, pl, QPT, N, &grid, &smem))) return rc; \
                KERNEL<QPT, TT_STREAM><<<grid, TT_THREADS, smem, st>>>(__VA_ARGS__);              \
                break;                                                                            \
            default:                                                                              \
                if ((rc = tt_launch_cfg(KERNEL<QPT, TT_GLOBAL>, pl, QPT, N, &grid, &smem))) return rc; \
                KERNEL<QPT, TT_GLOBAL><<<grid, TT_THREADS, smem, st>>>(__VA_ARGS__);              \
                break;                                                                            \
        }                                                                                         \
        g_launches.fetch_add(1);                                                                  \
        PCB_CUDA(cudaGetLastError());                                                             \
    } while (0)

extern "C" PCB_API int pcb_tt_eval(void *plan, const double *d_points, int64_t N, double *d_out, void *stream) {
    TTPlan *pl = static_cast<TTPlan *>(plan);
    PCB_REQUIRE(pl && pl->kind == PLAN_TT, "not a TT plan");
    PCB_REQUIRE(N >= 0, "negative N");
    if (N == 0) return PCB_OK;
    PCB_REQUIRE(d_points && d_out, "null device pointer");
    DeviceGuard guard(pl->dev);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (tt_pick_qpt(pl) == 2)
        TT_LAUNCH_BY_MODE(tt_value_kernel, 2, pl->P, pl->d_cores, d_points, N, d_out, pl->pingpong);
    else
        TT_LAUNCH_BY_MODE(tt_value_kernel, 1, pl->P, pl->d_cores, d_points, N, d_out, pl->pingpong);
    return PCB_OK;
}

static int tt_build_program(const TTPlan *pl, int G, const int32_t *orders, TTFdProgram *prog,
                            int *max_active) {
    const TTParams &P = pl->P;
    PCB_REQUIRE(G >= 1 && G <= TT_MAX_G, "number of derivative rows %d outside [1, %d]", G, TT_MAX_G);
    prog->G = G;
    *max_active = 0;
    for (int g = 0; g < G; ++g) {
        TTFdRow &row = prog->row[g];
        row.m = 0;
        for (int k = 0; k < P.D; ++k) {  // storage frame: order of storage dim k is orders[g][perm[k]]
            const int o = orders[(size_t)g * P.D + P.perm[k]];
            PCB_REQUIRE(o >= 0, "negative derivative order");
            if (o == 0) continue;
            // reference: ValueError(f"Derivative order {order} not supported (use 1 or 2)")
            PCB_REQUIRE(o <= 2, "Derivative order %d not supported (use 1 or 2)", o);
            if (row.m == TT_MAX_ACTIVE)
                return fail(PCB_EUNSUPPORTED,
                            "finite-difference rows with more than %d differentiated dims are not "
                            "supported on the device", TT_MAX_ACTIVE);
            row.dim[row.m] = k;
            row.ord[row.m] = o;
            ++row.m;
        }
        if (row.m > *max_active) *max_active = row.m;
    }
    return PCB_OK;
}

// Rows that differentiate at most one dim each can share partial products (algo 2).
static bool tt_build_shared_program(const TTFdProgram &prog, TTSharedProgram *sp) {
    sp->G = prog.G;
    sp->n_slots = 0;
    bool used[PCB_MAX_DIMS] = {false};
    for (int g = 0; g < prog.G; ++g) {
        if (prog.row[g].m > 1) return false;
        if (prog.row[g].m == 1) used[prog.row[g].dim[0]] = true;
    }
    int slot_of[PCB_MAX_DIMS];
    for (int k = 0; k < PCB_MAX_DIMS; ++k) {
        slot_of[k] = -1;
        if (used[k]) {
            slot_of[k] = sp->n_slots;
            sp->slot_dim[sp->n_slots++] = k;
        }
    }
    if (sp->n_slots == 0) return false;  // values only: nothing to share
    for (int g = 0; g < prog.G; ++g) {
        sp->row_slot[g] = prog.row[g].m == 1 ? slot_of[prog.row[g].dim[0]] : -1;
        sp->row_ord[g] = prog.row[g].m == 1 ? prog.row[g].ord[0] : 0;
    }
    return true;
}

template <int QPT, int MODE>
static int tt_launch_shared(const TTPlan *pl, const TTSharedProgram &sp, const double *d_points,
                            int64_t N, double *d_out, cudaStream_t st) {
    const TTSharedSmem L = tt_shared_smem_layout(pl->P, MODE, pl->pingpong, QPT, TT_THREADS);
    const size_t smem = (size_t)L.total * sizeof(double);
    auto kernel = tt_fd_shared_kernel<QPT, MODE>;
    PCB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    PCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, TT_THREADS, smem));
    if (per_sm < 1) return fail(PCB_ECUDA, "TT shared-FD kernel does not fit on an SM");
    const int64_t ntiles = (N + (int64_t)TT_THREADS * QPT - 1) / ((int64_t)TT_THREADS * QPT);
    const int64_t cap = (int64_t)pl->sm_count * per_sm;
    kernel<<<(int)(ntiles < cap ? ntiles : cap), TT_THREADS, smem, st>>>(pl->P, sp, pl->d_cores,
                                                                        d_points, N, d_out,
                                                                        pl->pingpong);
    g_launches.fetch_add(1);
    PCB_CUDA(cudaGetLastError());
    return PCB_OK;
}

static int tt_shared_qpt(const TTPlan *pl) {
    const TTSharedSmem L2 = tt_shared_smem_layout(pl->P, pl->mode_shared, pl->pingpong, 2, TT_THREADS);
    return (size_t)L2.total * sizeof(double) <= (size_t)pl->smem_optin ? 2 : 1;
}

// Which algorithm pcb_tt_eval_fd(algo = 0) runs for these rows: 1 or 2 (negative on error).
extern "C" PCB_API int pcb_tt_fd_algo(void *plan, int G, const int32_t *orders) {
    TTPlan *pl = static_cast<TTPlan *>(plan);
    PCB_REQUIRE(pl && pl->kind == PLAN_TT, "not a TT plan");
    PCB_REQUIRE(orders, "null orders");
    TTFdProgram prog;
    int max_active = 0;
    if (int rc = tt_build_program(pl, G, orders, &prog, &max_active)) return rc;
    TTSharedProgram sp;
    return tt_build_shared_program(prog, &sp) ? 2 : 1;
}

extern "C" PCB_API int pcb_tt_eval_fd(void *plan, const double *d_points, int64_t N, int G,
                                      const int32_t *orders, double *d_out, int algo, void *stream) {
    TTPlan *pl = static_cast<TTPlan *>(plan);
    PCB_REQUIRE(pl && pl->kind == PLAN_TT, "not a TT plan");
    PCB_REQUIRE(orders, "null orders");
    PCB_REQUIRE(N >= 0, "negative N");
    PCB_REQUIRE(algo >= 0 && algo <= 2, "algo %d not available", algo);
    TTFdProgram prog;
    int max_active = 0;
    if (int rc = tt_build_program(pl, G, orders, &prog, &max_active)) return rc;
    TTSharedProgram sp;
    const bool can_share = tt_build_shared_program(prog, &sp);
    if (algo == 2 && !can_share)
        return fail(PCB_EUNSUPPORTED, "algo 2 needs at least one differentiated dim and at most "
                    "one per row");
    if (N == 0) return PCB_OK;
    PCB_REQUIRE(d_points && d_out, "null device pointer");
    DeviceGuard guard(pl->dev);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (algo == 2 || (algo == 0 && can_share)) {
        const int qpt = tt_shared_qpt(pl);
        switch (pl->mode_shared) {
            case TT_RESIDENT:
                return qpt == 2 ? tt_launch_shared<2, TT_RESIDENT>(pl, sp, d_points, N, d_out, st)
                                : tt_launch_shared<1, TT_RESIDENT>(pl, sp, d_points, N, d_out, st);
            case TT_STREAM:
                return qpt == 2 ? tt_launch_shared<2, TT_STREAM>(pl, sp, d_points, N, d_out, st)
                                : tt_launch_shared<1, TT_STREAM>(pl, sp, d_points, N, d_out, st);
            default:
                return qpt == 2 ? tt_launch_shared<2, TT_GLOBAL>(pl, sp, d_points, N, d_out, st)
                                : tt_launch_shared<1, TT_GLOBAL>(pl, sp, d_points, N, d_out, st);
        }
    }
    if (tt_pick_qpt(pl) == 2)
        TT_LAUNCH_BY_MODE(tt_fd_general_kernel, 2, pl->P, prog, pl->d_cores, d_points, N, d_out,
                          pl->pingpong);
    else
        TT_LAUNCH_BY_MODE(tt_fd_general_kernel, 1, pl->P, prog, pl->d_cores, d_points, N, d_out,
                          pl->pingpong);
    return PCB_OK;
}
