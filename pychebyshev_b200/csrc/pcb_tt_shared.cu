// pcb_tt_eval_fd algo 2: finite differences with shared left/right partial products.
//
// With every stencil point differing from the query in ONE coordinate a, the interpolant along
// that coordinate is the degree n_a-1 Chebyshev series  f(x_a) = sum_j y_j T_j(s(x_a)),
//   y_j = L_a . G_a[:, j, :] . R_{a+1},
// L_a = product of the contracted cores left of a, R_{a+1} = product right of a, both at the
// query's own coordinates.  One left sweep (continued from slot to slot), one right sweep per
// differentiated dim (over transposed cores) and one coefficient pass give every stencil value of
// that dim for O(n_a) FMAs each, instead of a full chain per stencil point.  The stencil values
// are summed with Clenshaw's recurrence as the coefficients y_j arrive (j descending).
// Step h, boundary nudge and the difference formulas are the reference's
// (tensor_train.py:2356-2403).
#include "pcb_tt.cuh"

namespace pcb {

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
    // software pipeline over i: row i+1 (stride istride) is loaded before the FMAs of row i
    constexpr int H = (W + 1) / 2;
    const double *gp = gj;
    double2 gcur[H];
#pragma unroll
    for (int l = 0; l < H; ++l) gcur[l] = *reinterpret_cast<const double2 *>(gp + 2 * l);
#pragma unroll 2
    for (int i = 0; i < r_in; ++i) {
        double li[QPT];
#pragma unroll
        for (int qq = 0; qq < QPT; ++qq) li[qq] = vL[(i * QPT + qq) * vstride];
        if (i + 1 < r_in) gp += istride;  // the last iteration re-reads its own row
        double2 gnext[H];
#pragma unroll
        for (int l = 0; l < H; ++l) gnext[l] = *reinterpret_cast<const double2 *>(gp + 2 * l);
#pragma unroll
        for (int l = 0; l < H; ++l) {
#pragma unroll
            for (int qq = 0; qq < QPT; ++qq) {
                acc[qq][2 * l] = fma(li[qq], gcur[l].x, acc[qq][2 * l]);
                if (2 * l + 1 < W) acc[qq][2 * l + 1] = fma(li[qq], gcur[l].y, acc[qq][2 * l + 1]);
            }
        }
#pragma unroll
        for (int l = 0; l < H; ++l) gcur[l] = gnext[l];
    }
#pragma unroll
    for (int l = 0; l < W; ++l)
#pragma unroll
        for (int qq = 0; qq < QPT; ++qq) y[qq] = fma(acc[qq][l], vR[(l * QPT + qq) * vstride], y[qq]);
}

template <int QPT, int LC>
__device__ __forceinline__ void tt_coeff_row(const double *__restrict__ gj, int r_out, int istride,
                                             int r_in, const double *vL, const double *vR,
                                             int vstride, double (&y)[QPT]) {
    for (int l0 = 0; l0 < r_out; l0 += LC) {
        const int w = min(LC, r_out - l0);
        const double *gc = gj + l0;
        const double *vr = vR + (size_t)l0 * QPT * vstride;
        TT_CHUNK_SWITCH(tt_coeff_chunk, gc, istride, r_in, vL, vr, vstride, y)
    }
}

template <int QPT, int MODE, int LC, int MAXT>
__global__ void __launch_bounds__(MAXT)
tt_fd_shared_kernel(const __grid_constant__ TTParams P, const __grid_constant__ TTSharedProgram prog,
                    const double *__restrict__ cores, const double *__restrict__ pts, int64_t N,
                    double *__restrict__ out, int pingpong) {
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x;
    const int vstride = blockDim.x;
    const int D = P.D, G = prog.G;
    tt_load_resident<MODE>(cores, smem, P.total + P.totalT);
    const int v_off = (tt_core_area(P, MODE, true) + 1) & ~1;
    const int vsz = P.rmaxp * QPT * vstride;
    const int64_t tile_rows = (int64_t)vstride * QPT;
    const int64_t ntiles = (N + tile_rows - 1) / tile_rows;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t q0 = tile * tile_rows;
        const double *xrow[QPT];
        tt_query_rows<QPT>(pts, N, D, q0, q0 + (int64_t)gridDim.x * tile_rows, xrow);
        double *vL = smem + v_off + tid;
        double *vR = vL + vsz;
        double *vT = pingpong ? vR + vsz : nullptr;
#pragma unroll
        for (int qq = 0; qq < QPT; ++qq) vL[qq * vstride] = 1.0;
        int lpos = 0;  // vL holds the left product over storage dims [0, lpos)
        for (int t = 0; t < prog.n_slots; ++t) {
            const int a = prog.slot_dim[t];
            double s[QPT];
            // ---- right sweep: vR = M_{a+1} ... M_{D-1} . 1, applied right to left ----------------
#pragma unroll
            for (int qq = 0; qq < QPT; ++qq) vR[qq * vstride] = 1.0;
            for (int k = D - 1; k > a; --k) {
#pragma unroll
                for (int qq = 0; qq < QPT; ++qq)
                    s[qq] = tt_scale(__ldg(xrow[qq] + P.perm[k]), P.lo[k], P.hi[k]);
                const double *g =
                    tt_core_ptr<MODE>(cores, smem, P.offT[k], P.r[k + 1] * P.n[k] * P.rpT[k]);
                double *vo = pingpong ? vT : vR;
                tt_apply_core<QPT, LC>(g, P.rpT[k], P.r[k], P.r[k + 1], P.n[k], vR, vo, vstride, s);
                if (pingpong) {
                    vT = vR;
                    vR = vo;
                }
            }
            // ---- left sweep continues up to a ------------------------------------------------------
            for (int k = lpos; k < a; ++k) {
#pragma unroll
                for (int qq = 0; qq < QPT; ++qq)
                    s[qq] = tt_scale(__ldg(xrow[qq] + P.perm[k]), P.lo[k], P.hi[k]);
                const double *g = tt_core_ptr<MODE>(cores, smem, P.off[k], P.r[k] * P.n[k] * P.rp[k]);
                double *vo = pingpong ? vT : vL;
                tt_apply_core<QPT, LC>(g, P.rp[k], P.r[k + 1], P.r[k], P.n[k], vL, vo, vstride, s);
                if (pingpong) {
                    vT = vL;
                    vL = vo;
                }
            }
            lpos = a;
            // ---- stencil abscissae of dim a (reference _fd_step / _nudge_point): ------------------
            //      0: the query itself, 1: nudged centre c, 2: c + h, 3: c - h
            const double lo = P.lo[a], hi = P.hi[a];
            const double h = (hi - lo) * 1e-4;
            double sm[4][QPT], b1[4][QPT], b2[4][QPT];
#pragma unroll
            for (int qq = 0; qq < QPT; ++qq) {
                const double x = __ldg(xrow[qq] + P.perm[a]);
                const double c = tt_nudge(x, lo, hi, h);
                sm[0][qq] = tt_scale(x, lo, hi);
                sm[1][qq] = tt_scale(c, lo, hi);
                sm[2][qq] = tt_scale(c + h, lo, hi);
                sm[3][qq] = tt_scale(c - h, lo, hi);
#pragma unroll
                for (int m = 0; m < 4; ++m) b1[m][qq] = b2[m][qq] = 0.0;
            }
            // ---- coefficient pass, j descending; Clenshaw: b_j = y_j + 2 s b_{j+1} - b_{j+2} -----
            // y_j = sum_{i,l} L_i G[i][j][l] R_l is symmetric in (L, i) <-> (R, l): run it over the
            // layout whose register-accumulated index is the wider one (fewer LDS per DFMA)
            const bool fwd = P.r[a + 1] >= P.r[a];
            const int na = P.n[a];
            const int rp = fwd ? P.rp[a] : P.rpT[a];
            const int r_acc = fwd ? P.r[a + 1] : P.r[a], r_rows = fwd ? P.r[a] : P.r[a + 1];
            const double *ga = tt_core_ptr<MODE>(cores, smem, fwd ? P.off[a] : P.offT[a],
                                                 r_rows * na * rp);
            const double *v_rows = fwd ? vL : vR, *v_acc = fwd ? vR : vL;
            double y0[QPT];
            for (int j = na - 1; j >= 0; --j) {
                double y[QPT];
#pragma unroll
                for (int qq = 0; qq < QPT; ++qq) y[qq] = 0.0;
                tt_coeff_row<QPT, LC>(ga + (size_t)j * rp, r_acc, na * rp, r_rows, v_rows, v_acc,
                                      vstride, y);
                if (j > 0) {
#pragma unroll
                    for (int m = 0; m < 4; ++m)
#pragma unroll
                        for (int qq = 0; qq < QPT; ++qq) {
                            const double bn = fma(2.0 * sm[m][qq], b1[m][qq], y[qq] - b2[m][qq]);
                            b2[m][qq] = b1[m][qq];
                            b1[m][qq] = bn;
                        }
                } else {
#pragma unroll
                    for (int qq = 0; qq < QPT; ++qq) y0[qq] = y[qq];
                }
            }
            // f(s) = y_0 + s b_1 - b_2
            double f[4][QPT];
#pragma unroll
            for (int m = 0; m < 4; ++m)
#pragma unroll
                for (int qq = 0; qq < QPT; ++qq)
                    f[m][qq] = fma(sm[m][qq], b1[m][qq], y0[qq] - b2[m][qq]);
            // ---- outputs owned by this slot (value rows ride on the first slot) -------------------
            for (int g = 0; g < G; ++g) {
                const int rs = prog.row_slot[g];
                if (!(rs == t || (rs < 0 && t == 0))) continue;
#pragma unroll
                for (int qq = 0; qq < QPT; ++qq) {
                    const double res =
                        rs < 0 ? f[0][qq]
                               : tt_fd_reduce(prog.row_ord[g], f[2][qq], f[1][qq], f[3][qq], h);
                    const int64_t q = q0 + qq * (int64_t)vstride + tid;
                    if (q < N) out[q * G + g] = res;
                }
            }
        }
    }
}

#define TT_TRY_SHARED(Q, L, T)                                                                  \
    if (c.qpt == Q && c.lc == L && c.threads <= T && !done) {                                   \
        done = true;                                                                            \
        rc = tt_launch_kernel(tt_fd_shared_kernel<Q, MODE_, L, T>, pl, c, N, st, pl->P, prog,     \
                              pl->d_cores, d_points, N, d_out, c.pingpong);                     \
    }

int tt_launch_shared(const TTPlan *pl, const TTSharedProgram &prog, const double *d_points,
                     int64_t N, double *d_out, cudaStream_t st) {
    const TTCfg &c = pl->cfg_shared;
    bool done = false;
    int rc = PCB_OK;
    if (c.mode == TT_RESIDENT) {
        constexpr int MODE_ = TT_RESIDENT;
        TT_SHARED_RESIDENT_CONFIGS(TT_TRY_SHARED)
    } else if (c.mode == TT_STREAM) {
        constexpr int MODE_ = TT_STREAM;
        TT_GENERIC_CONFIGS(TT_TRY_SHARED)
    } else {
        constexpr int MODE_ = TT_GLOBAL;
        TT_GENERIC_CONFIGS(TT_TRY_SHARED)
    }
    if (!done) return fail(PCB_EUNSUPPORTED, "no TT shared-FD kernel for qpt=%d lc=%d threads=%d", c.qpt, c.lc, c.threads);
    return rc;
}

}  // namespace pcb
