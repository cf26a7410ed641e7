// TT kernels that evaluate whole chains: values (pcb_tt_eval) and the one-chain-per-stencil-point
// finite-difference kernel (pcb_tt_eval_fd algo 1).  See pcb_tt.cuh for the design.
#include "pcb_tt.cuh"

namespace pcb {

template <int QPT, int MODE, int LC, int MAXT>
__global__ void __launch_bounds__(MAXT)
tt_value_kernel(const __grid_constant__ TTParams P, const double *__restrict__ cores,
                const double *__restrict__ pts, int64_t N, double *__restrict__ out, int pingpong) {
    extern __shared__ __align__(16) double smem[];
    tt_load_resident<MODE>(cores, smem, P.total);
    const int v_off = (tt_core_area(P, MODE, false) + 1) & ~1;
    const int64_t tile_rows = (int64_t)blockDim.x * QPT;
    const int64_t ntiles = (N + tile_rows - 1) / tile_rows;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t q0 = tile * tile_rows;
        const double *xrow[QPT];
        tt_query_rows<QPT>(pts, N, P.D, q0, q0 + (int64_t)gridDim.x * tile_rows, xrow);
        double res[QPT];
        tt_chain<QPT, MODE, LC>(P, cores, smem, v_off, pingpong, xrow, 0, nullptr, nullptr, res);
#pragma unroll
        for (int qq = 0; qq < QPT; ++qq) {
            const int64_t q = q0 + qq * (int64_t)blockDim.x + threadIdx.x;
            if (q < N) out[q] = res[qq];
        }
    }
}

// algo 1: one full chain per stencil point, the reference's own evaluation count and formulas.
template <int QPT, int MODE, int LC, int MAXT>
__global__ void __launch_bounds__(MAXT)
tt_fd_general_kernel(const __grid_constant__ TTParams P, const __grid_constant__ TTFdProgram prog,
                     const double *__restrict__ cores, const double *__restrict__ pts, int64_t N,
                     double *__restrict__ out, int pingpong) {
    extern __shared__ __align__(16) double smem[];
    tt_load_resident<MODE>(cores, smem, P.total);
    const int v_off = (tt_core_area(P, MODE, false) + 1) & ~1;
    const int64_t tile_rows = (int64_t)blockDim.x * QPT;
    const int64_t ntiles = (N + tile_rows - 1) / tile_rows;
    const int G = prog.G;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t q0 = tile * tile_rows;
        const double *xrow[QPT];
        tt_query_rows<QPT>(pts, N, P.D, q0, q0 + (int64_t)gridDim.x * tile_rows, xrow);
        for (int g = 0; g < G; ++g) {
            const TTFdRow row = prog.row[g];
            double result[QPT];
            if (row.m == 0) {
                tt_chain<QPT, MODE, LC>(P, cores, smem, v_off, pingpong, xrow, 0, nullptr, nullptr,
                                        result);
            } else {
                double h[TT_MAX_ACTIVE];
                double ctr[TT_MAX_ACTIVE][QPT];
                for (int t = 0; t < row.m; ++t) {
                    const int k = row.dim[t];
                    const double a = P.lo[k], b = P.hi[k];
                    h[t] = (b - a) * 1e-4;  // tensor_train.py:2356-2359
#pragma unroll
                    for (int qq = 0; qq < QPT; ++qq)
                        ctr[t][qq] = tt_nudge(__ldg(xrow[qq] + P.perm[k]), a, b, h[t]);
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
                        tt_chain<QPT, MODE, LC>(P, cores, smem, v_off, pingpong, xrow, 2, row.dim, ox,
                                                f[e]);
                    }
#pragma unroll
                    for (int qq = 0; qq < QPT; ++qq)
                        result[qq] = (f[0][qq] - f[1][qq] - f[2][qq] + f[3][qq]) / (4.0 * h[0] * h[1]);
                } else {
                    // nested stencils (tensor_train.py:2372-2403 for m == 1, 2428-2463 otherwise):
                    // the first differentiated dim is the outermost level.
                    double f0[3][QPT], f1[3][QPT], f2[3][QPT];
#pragma unroll 1
                    for (int e0 = 0; e0 < 3; ++e0) {
                        if (e0 == 1 && row.ord[0] == 1) continue;
#pragma unroll
                        for (int qq = 0; qq < QPT; ++qq)
                            ox[0][qq] = e0 == 0 ? ctr[0][qq] + h[0]
                                                : (e0 == 1 ? ctr[0][qq] : ctr[0][qq] - h[0]);
                        if (m == 1) {
                            tt_chain<QPT, MODE, LC>(P, cores, smem, v_off, pingpong, xrow, 1, row.dim,
                                                    ox, f0[e0]);
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
                                tt_chain<QPT, MODE, LC>(P, cores, smem, v_off, pingpong, xrow, 2,
                                                        row.dim, ox, f1[e1]);
                                continue;
                            }
#pragma unroll 1
                            for (int e2 = 0; e2 < 3; ++e2) {
                                if (e2 == 1 && row.ord[2] == 1) continue;
#pragma unroll
                                for (int qq = 0; qq < QPT; ++qq)
                                    ox[2][qq] = e2 == 0 ? ctr[2][qq] + h[2]
                                                        : (e2 == 1 ? ctr[2][qq] : ctr[2][qq] - h[2]);
                                tt_chain<QPT, MODE, LC>(P, cores, smem, v_off, pingpong, xrow, 3,
                                                        row.dim, ox, f2[e2]);
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
                const int64_t q = q0 + qq * (int64_t)blockDim.x + threadIdx.x;
                if (q < N) out[q * G + g] = result[qq];
            }
        }
    }
}

#define TT_TRY_VALUE(Q, L, T)                                                                  \
    if (c.qpt == Q && c.lc == L && c.threads <= T && !done) {                                  \
        done = true;                                                                           \
        rc = tt_launch_kernel(tt_value_kernel<Q, MODE_, L, T>, pl, c, N, st, pl->P, pl->d_cores, \
                              d_points, N, d_out, c.pingpong);                                 \
    }
#define TT_TRY_GENERAL(Q, L, T)                                                                \
    if (c.qpt == Q && c.lc == L && c.threads <= T && !done) {                                  \
        done = true;                                                                           \
        rc = tt_launch_kernel(tt_fd_general_kernel<Q, MODE_, L, T>, pl, c, N, st, pl->P, prog,   \
                              pl->d_cores, d_points, N, d_out, c.pingpong);                    \
    }

int tt_launch_value(const TTPlan *pl, const double *d_points, int64_t N, double *d_out, cudaStream_t st) {
    const TTCfg &c = pl->cfg_chain;
    bool done = false;
    int rc = PCB_OK;
    if (c.mode == TT_RESIDENT) {
        constexpr int MODE_ = TT_RESIDENT;
        TT_RESIDENT_CONFIGS(TT_TRY_VALUE)
    } else if (c.mode == TT_STREAM) {
        constexpr int MODE_ = TT_STREAM;
        TT_GENERIC_CONFIGS(TT_TRY_VALUE)
    } else {
        constexpr int MODE_ = TT_GLOBAL;
        TT_GENERIC_CONFIGS(TT_TRY_VALUE)
    }
    if (!done) return fail(PCB_EUNSUPPORTED, "no TT value kernel for qpt=%d lc=%d threads=%d", c.qpt, c.lc, c.threads);
    return rc;
}

int tt_launch_general(const TTPlan *pl, const TTFdProgram &prog, const double *d_points, int64_t N,
                      double *d_out, cudaStream_t st) {
    const TTCfg &c = pl->cfg_chain;
    bool done = false;
    int rc = PCB_OK;
    if (c.mode == TT_RESIDENT) {
        constexpr int MODE_ = TT_RESIDENT;
        TT_RESIDENT_CONFIGS(TT_TRY_GENERAL)
    } else if (c.mode == TT_STREAM) {
        constexpr int MODE_ = TT_STREAM;
        TT_GENERIC_CONFIGS(TT_TRY_GENERAL)
    } else {
        constexpr int MODE_ = TT_GLOBAL;
        TT_GENERIC_CONFIGS(TT_TRY_GENERAL)
    }
    if (!done) return fail(PCB_EUNSUPPORTED, "no TT finite-difference kernel for qpt=%d lc=%d threads=%d", c.qpt, c.lc, c.threads);
    return rc;
}

}  // namespace pcb
