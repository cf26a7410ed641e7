// TT kernels whose coefficient cores live in the 64 KB __constant__ bank and reach the FP64 pipe
// through the UNIFORM DATAPATH: `LDCU.64 URx, c[0x3][URy + imm]` loads a warp-uniform core element
// into a uniform register and `DFMA Rd, Ra, URx, Rd` consumes it directly -- no shared-memory
// wavefront, no LSU slot, no vector register for the operand.
//
// Why: the broadcast-LDS kernels (pcb_tt_chain/shared.cu) are co-limited by the shared-memory
// pipe -- a broadcast LDS.128 costs 2 wavefronts, so with two query slots per thread the inner
// loop needs 12 wavefronts per 24 DFMA, 92 % of the shared-memory bandwidth at full FP64 rate
// (ncu: 60-72 % FP64 pipe).  The same loop fed from the constant bank measures 93 % of the FP64
// peak (tools/probes/ldcu_probe*.cu; flat up to a 63 KB table).
//
// Two things ptxas needs before it emits LDCU instead of a per-lane LDC:
//   * the core element is indexed as c_tt[int], never through a pointer that crossed a call;
//   * NO loop around the chain whose induction depends on blockIdx (a persistent tile loop makes
//     it drop the uniformity proof) -> these kernels run one tile per CTA.  Nothing is staged per
//     CTA, so that costs nothing.
//
// The bank holds the unpadded forward cores G_k[i][j][l] and, for the finite-difference kernel,
// the transposed cores G_k^T[l][j][i] of ONE plan per device at a time; `ttc_make_resident`
// uploads on a plan switch, ordered against kernels of the previous plan on other streams.
// Trains that do not fit (8192 doubles, ranks <= 16) use the shared-memory kernels.
#include <algorithm>

#include "pcb_cbank.cuh"
#include "pcb_tt.cuh"

namespace pcb {

__constant__ double c_tt[TT_CONST_MAX];

// ---- one step of a sweep: chunk of W output columns, rows (i, j) contiguous with stride W --------
//   acc[l] = sum_{i,j} v[i] T_j(s) c_tt[base + (i n + j) W + l]
template <int W, int QPT>
__device__ __forceinline__ void ttc_chunk(int base, int r_in, int n, const double *v_in, int vstride,
                                          const double (&s)[QPT], double *v_out) {
    double acc[QPT][W], twos[QPT];
#pragma unroll
    for (int qq = 0; qq < QPT; ++qq) {
        twos[qq] = 2.0 * s[qq];
#pragma unroll
        for (int l = 0; l < W; ++l) acc[qq][l] = 0.0;
    }
    int gp = base;
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
#pragma unroll
            for (int l = 0; l < W; ++l) {
                const double gv = c_tt[gp + l];  // LDCU.64
#pragma unroll
                for (int qq = 0; qq < QPT; ++qq) acc[qq][l] = fma(c0[qq], gv, acc[qq][l]);
            }
#pragma unroll
            for (int qq = 0; qq < QPT; ++qq) {
                const double c2 = fma(twos[qq], c1[qq], -c0[qq]);  // T_{j+2} = 2 s T_{j+1} - T_j
                c0[qq] = c1[qq];
                c1[qq] = c2;
            }
            gp += W;
        }
    }
#pragma unroll
    for (int l = 0; l < W; ++l)
#pragma unroll
        for (int qq = 0; qq < QPT; ++qq) v_out[(l * QPT + qq) * vstride] = acc[qq][l];
}

// ---- coefficient pass: y_j = sum_{rows i} vrow[i] (sum_l c_tt[...][l] vacc[l]), j descending, with
//      Clenshaw over the four stencil abscissae: b_j = y_j + 2 s b_{j+1} - b_{j+2} -----------------------
template <int W, int QPT>
__device__ __forceinline__ void ttc_coeff(int base, int r_rows, int n, const double *v_rows,
                                          const double *v_acc, int vstride,
                                          const double (&sm)[4][QPT], double (&f)[4][QPT]) {
    // Clenshaw with the uniform recurrence b_j = y_j + 2 s b_{j+1} - b_{j+2} for j = n-1 .. 0 and
    // f = b_0 - s b_1; only tw = 2 s is kept live (s = tw / 2 exactly).
    double b1[4][QPT], b2[4][QPT], tw[4][QPT];
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int qq = 0; qq < QPT; ++qq) {
            b1[m][qq] = b2[m][qq] = 0.0;
            tw[m][qq] = 2.0 * sm[m][qq];
        }
    for (int j = n - 1; j >= 0; --j) {
        double acc[QPT][W];
#pragma unroll
        for (int qq = 0; qq < QPT; ++qq)
#pragma unroll
            for (int l = 0; l < W; ++l) acc[qq][l] = 0.0;
        int gp = base + j * W;
        for (int i = 0; i < r_rows; ++i) {
            double li[QPT];
#pragma unroll
            for (int qq = 0; qq < QPT; ++qq) li[qq] = v_rows[(i * QPT + qq) * vstride];
#pragma unroll
            for (int l = 0; l < W; ++l) {
                const double gv = c_tt[gp + l];
#pragma unroll
                for (int qq = 0; qq < QPT; ++qq) acc[qq][l] = fma(li[qq], gv, acc[qq][l]);
            }
            gp += n * W;
        }
        double y[QPT];
#pragma unroll
        for (int qq = 0; qq < QPT; ++qq) y[qq] = 0.0;
#pragma unroll
        for (int l = 0; l < W; ++l)
#pragma unroll
            for (int qq = 0; qq < QPT; ++qq)
                y[qq] = fma(acc[qq][l], v_acc[(l * QPT + qq) * vstride], y[qq]);
#pragma unroll
        for (int m = 0; m < 4; ++m)
#pragma unroll
            for (int qq = 0; qq < QPT; ++qq) {
                const double bn = fma(tw[m][qq], b1[m][qq], y[qq] - b2[m][qq]);
                b2[m][qq] = b1[m][qq];
                b1[m][qq] = bn;
            }
    }
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int qq = 0; qq < QPT; ++qq) f[m][qq] = fma(-0.5 * tw[m][qq], b2[m][qq], b1[m][qq]);
}

#define TTC_CASE(FN, WW, ...)                             \
    case WW:                                              \
        if constexpr (WW <= RMAX) FN<WW, QPT>(__VA_ARGS__); \
        break;
#define TTC_SWITCH(FN, ...)                                                                       \
    switch (w) {                                                                                  \
        TTC_CASE(FN, 1, __VA_ARGS__) TTC_CASE(FN, 2, __VA_ARGS__) TTC_CASE(FN, 3, __VA_ARGS__)    \
        TTC_CASE(FN, 4, __VA_ARGS__) TTC_CASE(FN, 5, __VA_ARGS__) TTC_CASE(FN, 6, __VA_ARGS__)    \
        TTC_CASE(FN, 7, __VA_ARGS__) TTC_CASE(FN, 8, __VA_ARGS__) TTC_CASE(FN, 9, __VA_ARGS__)    \
        TTC_CASE(FN, 10, __VA_ARGS__) TTC_CASE(FN, 11, __VA_ARGS__) TTC_CASE(FN, 12, __VA_ARGS__) \
        TTC_CASE(FN, 13, __VA_ARGS__) TTC_CASE(FN, 14, __VA_ARGS__) TTC_CASE(FN, 15, __VA_ARGS__) \
        TTC_CASE(FN, 16, __VA_ARGS__)                                                             \
    }

// v_out = step(v_in); ranks <= 16 means a single chunk, so v_out may alias v_in
template <int QPT, int RMAX>
__device__ __forceinline__ void ttc_step(int base, int w, int r_in, int n, const double *v_in,
                                         double *v_out, int vstride, const double (&s)[QPT]) {
    TTC_SWITCH(ttc_chunk, base, r_in, n, v_in, vstride, s, v_out)
}

template <int QPT, int RMAX>
__device__ __forceinline__ void ttc_coeff_pass(int base, int w, int r_rows, int n,
                                               const double *v_rows, const double *v_acc,
                                               int vstride, const double (&sm)[4][QPT],
                                               double (&f)[4][QPT]) {
    TTC_SWITCH(ttc_coeff, base, r_rows, n, v_rows, v_acc, vstride, sm, f)
}

// ---- values ------------------------------------------------------------------------------------------
template <int QPT, int RMAX, int MAXT, int MINB = 0>
__global__ void __launch_bounds__(MAXT, MINB)
ttc_value_kernel(const __grid_constant__ TTParams P, const double *__restrict__ pts, int64_t N,
                 double *__restrict__ out) {
    extern __shared__ __align__(16) double smem[];
    const int vstride = blockDim.x;
    const int64_t q0 = (int64_t)blockIdx.x * vstride * QPT;  // one tile per CTA
    const double *xrow[QPT];
    tt_query_rows<QPT>(pts, N, P.D, q0, N, xrow);
    double *vbuf = smem + threadIdx.x;
#pragma unroll
    for (int qq = 0; qq < QPT; ++qq) vbuf[qq * vstride] = 1.0;
    for (int k = 0; k < P.D; ++k) {
        double s[QPT];
#pragma unroll
        for (int qq = 0; qq < QPT; ++qq)
            s[qq] = tt_scale(__ldg(xrow[qq] + P.perm[k]), P.lo[k], P.hi[k]);
        ttc_step<QPT, RMAX>(P.coff[k], P.r[k + 1], P.r[k], P.n[k], vbuf, vbuf, vstride, s);
    }
#pragma unroll
    for (int qq = 0; qq < QPT; ++qq) {
        const int64_t q = q0 + qq * (int64_t)vstride + threadIdx.x;
        if (q < N) out[q] = vbuf[qq * vstride];
    }
}

// ---- pcb_tt_eval_fd algo 2 (see pcb_tt_shared.cu for the algorithm) -----------------------------------
template <int QPT, int RMAX, int MAXT, int MINB = 0>
__global__ void __launch_bounds__(MAXT, MINB)
ttc_fd_shared_kernel(const __grid_constant__ TTParams P, const __grid_constant__ TTSharedProgram prog,
                     const double *__restrict__ pts, int64_t N, double *__restrict__ out) {
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x;
    const int vstride = blockDim.x;
    const int D = P.D, G = prog.G;
    const int64_t q0 = (int64_t)blockIdx.x * vstride * QPT;  // one tile per CTA
    const double *xrow[QPT];
    tt_query_rows<QPT>(pts, N, D, q0, N, xrow);
    double *vL = smem + tid;
    double *vR = vL + P.rmaxp * QPT * vstride;
#pragma unroll
    for (int qq = 0; qq < QPT; ++qq) vL[qq * vstride] = 1.0;
    int lpos = 0;  // vL holds the left product over storage dims [0, lpos)
    for (int t = 0; t < prog.n_slots; ++t) {
        const int a = prog.slot_dim[t];
        double s[QPT];
        // right sweep over the transposed cores: vR = M_{a+1} ... M_{D-1} . 1
#pragma unroll
        for (int qq = 0; qq < QPT; ++qq) vR[qq * vstride] = 1.0;
        for (int k = D - 1; k > a; --k) {
#pragma unroll
            for (int qq = 0; qq < QPT; ++qq)
                s[qq] = tt_scale(__ldg(xrow[qq] + P.perm[k]), P.lo[k], P.hi[k]);
            ttc_step<QPT, RMAX>(P.coffT[k], P.r[k], P.r[k + 1], P.n[k], vR, vR, vstride, s);
        }
        // left sweep continues up to a
        for (int k = lpos; k < a; ++k) {
#pragma unroll
            for (int qq = 0; qq < QPT; ++qq)
                s[qq] = tt_scale(__ldg(xrow[qq] + P.perm[k]), P.lo[k], P.hi[k]);
            ttc_step<QPT, RMAX>(P.coff[k], P.r[k + 1], P.r[k], P.n[k], vL, vL, vstride, s);
        }
        lpos = a;
        // stencil abscissae (reference _fd_step / _nudge_point): 0 query, 1 centre c, 2 c+h, 3 c-h
        const double lo = P.lo[a], hi = P.hi[a];
        const double h = (hi - lo) * 1e-4;
        double sm[4][QPT], f[4][QPT];
#pragma unroll
        for (int qq = 0; qq < QPT; ++qq) {
            const double x = __ldg(xrow[qq] + P.perm[a]);
            const double c = tt_nudge(x, lo, hi, h);
            sm[0][qq] = tt_scale(x, lo, hi);
            sm[1][qq] = tt_scale(c, lo, hi);
            sm[2][qq] = tt_scale(c + h, lo, hi);
            sm[3][qq] = tt_scale(c - h, lo, hi);
        }
        // coefficient pass over the layout whose register-accumulated index is the wider one
        if (P.r[a + 1] >= P.r[a])
            ttc_coeff_pass<QPT, RMAX>(P.coff[a], P.r[a + 1], P.r[a], P.n[a], vL, vR, vstride, sm, f);
        else
            ttc_coeff_pass<QPT, RMAX>(P.coffT[a], P.r[a], P.r[a + 1], P.n[a], vR, vL, vstride, sm, f);
        for (int g = 0; g < G; ++g) {
            const int rs = prog.row_slot[g];
            if (!(rs == t || (rs < 0 && t == 0))) continue;
#pragma unroll
            for (int qq = 0; qq < QPT; ++qq) {
                const double res = rs < 0 ? f[0][qq]
                                          : tt_fd_reduce(prog.row_ord[g], f[2][qq], f[1][qq], f[3][qq], h);
                const int64_t q = q0 + qq * (int64_t)vstride + tid;
                if (q < N) out[q * G + g] = res;
            }
        }
    }
}

// ---- large trains: one launch per core, chain state in global memory ---------------------------------
// A train whose cores do not fit in the bank together (10-D rank 20: 285 KB) still has cores that
// fit one at a time (rank 20, 11 nodes: 34 KB).  The chain then runs as one launch per core with
// the bank holding that core, and the per-query chain vector (<= 64 doubles) travels through
// global memory between launches as planes S[l][query] (coalesced; 16 r bytes per query per core
// against 2 r^2 n flop, far below the HBM roofline).  Output columns are processed in chunks of
// <= 16 register accumulators; the bank image of a core is chunk-major: chunk c holds
// [(i n + j) W_c + l].
constexpr int TTG_MAX_CHUNKS = 4;  // ranks up to 64
constexpr int TTG_MAX_RANK = 16 * TTG_MAX_CHUNKS;

struct TTGStep {
    int r_in, r_out, n, col, D;
    double lo, hi;
    int nchunks, cbase[TTG_MAX_CHUNKS], cw[TTG_MAX_CHUNKS];
    const double *s_in;  // r_in planes, or nullptr for the vector [1]
    double *s_out;       // r_out planes (or the output column when r_out == 1 ends a value chain)
    int64_t pstride;     // plane stride of s_in / s_out
    int64_t q_begin;     // first query of this tile
};

template <int W, int QPT>
__device__ __forceinline__ void ttc_chunk_g(int base, int r_in, int n, const double *v_in, int vstride,
                                            const double (&s)[QPT], double *g_out, int64_t pstride,
                                            const bool (&live)[QPT]) {
    double acc[QPT][W], twos[QPT];
#pragma unroll
    for (int qq = 0; qq < QPT; ++qq) {
        twos[qq] = 2.0 * s[qq];
#pragma unroll
        for (int l = 0; l < W; ++l) acc[qq][l] = 0.0;
    }
    int gp = base;
    for (int i = 0; i < r_in; ++i) {
        double c0[QPT], c1[QPT];
#pragma unroll
        for (int qq = 0; qq < QPT; ++qq) {
            const double vi = v_in[(i * QPT + qq) * vstride];
            c0[qq] = vi;
            c1[qq] = vi * s[qq];
        }
#pragma unroll 2
        for (int j = 0; j < n; ++j) {
#pragma unroll
            for (int l = 0; l < W; ++l) {
                const double gv = c_tt[gp + l];  // LDCU.64
#pragma unroll
                for (int qq = 0; qq < QPT; ++qq) acc[qq][l] = fma(c0[qq], gv, acc[qq][l]);
            }
#pragma unroll
            for (int qq = 0; qq < QPT; ++qq) {
                const double c2 = fma(twos[qq], c1[qq], -c0[qq]);
                c0[qq] = c1[qq];
                c1[qq] = c2;
            }
            gp += W;
        }
    }
#pragma unroll
    for (int l = 0; l < W; ++l)
#pragma unroll
        for (int qq = 0; qq < QPT; ++qq)
            if (live[qq]) g_out[l * pstride + qq * vstride] = acc[qq][l];
}

template <int QPT, int MAXT>
__global__ void __launch_bounds__(MAXT, 512 / MAXT)
ttc_gstep_kernel(const __grid_constant__ TTGStep a, const double *__restrict__ pts, int64_t N) {
    constexpr int RMAX = 16;
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, vstride = blockDim.x;
    const int64_t ql0 = (int64_t)blockIdx.x * vstride * QPT + tid;  // tile-local index of slot 0
    double s[QPT];
    bool live[QPT];
    double *vbuf = smem + tid;
#pragma unroll
    for (int qq = 0; qq < QPT; ++qq) {
        const int64_t ql = ql0 + qq * (int64_t)vstride;
        int64_t q = a.q_begin + ql;
        live[qq] = q < N;
        if (!live[qq]) q = N - 1;
        s[qq] = tt_scale(__ldg(pts + q * a.D + a.col), a.lo, a.hi);
        if (a.s_in) {
            for (int i = 0; i < a.r_in; ++i) vbuf[(i * QPT + qq) * vstride] = a.s_in[i * a.pstride + ql];
        } else {
            vbuf[qq * vstride] = 1.0;
        }
    }
    // (compile-time chunk indices: a dynamically indexed kernel-parameter array lands in vector
    // registers and takes every bank address with it off the uniform datapath)
    int l0 = 0;
#pragma unroll
    for (int c = 0; c < TTG_MAX_CHUNKS; ++c) {
        if (c < a.nchunks) {
            const int w = a.cw[c];
            double *g_out = a.s_out + l0 * a.pstride + ql0;
            TTC_SWITCH(ttc_chunk_g, a.cbase[c], a.r_in, a.n, vbuf, vstride, s, g_out, a.pstride, live)
            l0 += w;
        }
    }
}

struct TTGCoeff {
    int r_rows, r_acc, n, col, D, G;
    double lo, hi;
    int nchunks, cbase[TTG_MAX_CHUNKS], cw[TTG_MAX_CHUNKS];
    const double *s_rows, *s_acc;  // planes, or nullptr for the vector [1]
    int64_t pstride, q_begin;
    int row_kind[TT_MAX_G];  // 0 skip, -1 value row, 1 / 2 central difference of that order
    double *out;             // (N, G) row-major
};

// y_j += sum_l (sum_i vrow[i] G[i][j][l0 + l]) vacc[l0 + l] for one chunk of the accumulated index
template <int W, int QPT>
__device__ __forceinline__ void ttc_ychunk(int base, int r_rows, int n, const double *v_rows,
                                           const double *v_acc, int vstride, double *ybuf) {
    for (int j = 0; j < n; ++j) {
        double acc[QPT][W];
#pragma unroll
        for (int qq = 0; qq < QPT; ++qq)
#pragma unroll
            for (int l = 0; l < W; ++l) acc[qq][l] = 0.0;
        int gp = base + j * W;
        for (int i = 0; i < r_rows; ++i) {
            double li[QPT];
#pragma unroll
            for (int qq = 0; qq < QPT; ++qq) li[qq] = v_rows[(i * QPT + qq) * vstride];
#pragma unroll
            for (int l = 0; l < W; ++l) {
                const double gv = c_tt[gp + l];
#pragma unroll
                for (int qq = 0; qq < QPT; ++qq) acc[qq][l] = fma(li[qq], gv, acc[qq][l]);
            }
            gp += n * W;
        }
#pragma unroll
        for (int qq = 0; qq < QPT; ++qq) {
            double y0 = 0.0, y1 = 0.0;
#pragma unroll
            for (int l = 0; l < W; ++l) {
                const double va = v_acc[(l * QPT + qq) * vstride];
                if (l & 1)
                    y1 = fma(acc[qq][l], va, y1);
                else
                    y0 = fma(acc[qq][l], va, y0);
            }
            ybuf[(j * QPT + qq) * vstride] += y0 + y1;
        }
    }
}

// (minBlocks = 3 is a register budget, not an occupancy target -- shared memory admits one CTA per
// SM at rank 20: at 80 registers ptxas keeps 87 % of the bank reads on the uniform datapath, at the
// unbounded 127 only 25 %; tools/check_sass.py)
template <int QPT, int MAXT>
__global__ void __launch_bounds__(MAXT, 3)
ttc_gcoeff_kernel(const __grid_constant__ TTGCoeff a, const double *__restrict__ pts, int64_t N) {
    constexpr int RMAX = 16;
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, vstride = blockDim.x;
    const int64_t ql0 = (int64_t)blockIdx.x * vstride * QPT + tid;
    double *v_rows = smem + tid;
    double *v_acc = v_rows + (size_t)a.r_rows * QPT * vstride;
    double *ybuf = v_acc + (size_t)a.r_acc * QPT * vstride;
    const double h = (a.hi - a.lo) * 1e-4;
    double sm[4][QPT];
    bool live[QPT];
    int64_t qg[QPT];
#pragma unroll
    for (int qq = 0; qq < QPT; ++qq) {
        const int64_t ql = ql0 + qq * (int64_t)vstride;
        int64_t q = a.q_begin + ql;
        live[qq] = q < N;
        if (!live[qq]) q = N - 1;
        qg[qq] = q;
        // stencil abscissae (reference _fd_step / _nudge_point): 0 query, 1 centre c, 2 c+h, 3 c-h
        const double x = __ldg(pts + q * a.D + a.col);
        const double c = tt_nudge(x, a.lo, a.hi, h);
        sm[0][qq] = tt_scale(x, a.lo, a.hi);
        sm[1][qq] = tt_scale(c, a.lo, a.hi);
        sm[2][qq] = tt_scale(c + h, a.lo, a.hi);
        sm[3][qq] = tt_scale(c - h, a.lo, a.hi);
        if (a.s_rows) {
            for (int i = 0; i < a.r_rows; ++i) v_rows[(i * QPT + qq) * vstride] = a.s_rows[i * a.pstride + ql];
        } else {
            v_rows[qq * vstride] = 1.0;
        }
        if (a.s_acc) {
            for (int l = 0; l < a.r_acc; ++l) v_acc[(l * QPT + qq) * vstride] = a.s_acc[l * a.pstride + ql];
        } else {
            v_acc[qq * vstride] = 1.0;
        }
        for (int j = 0; j < a.n; ++j) ybuf[(j * QPT + qq) * vstride] = 0.0;
    }
    int l0 = 0;
#pragma unroll
    for (int c = 0; c < TTG_MAX_CHUNKS; ++c) {
        if (c < a.nchunks) {
            const int w = a.cw[c];
            const double *vacc_c = v_acc + (size_t)l0 * QPT * vstride;
            TTC_SWITCH(ttc_ychunk, a.cbase[c], a.r_rows, a.n, v_rows, vacc_c, vstride, ybuf)
            l0 += w;
        }
    }
    // Clenshaw over the four abscissae: b_j = y_j + 2 s b_{j+1} - b_{j+2}, f = b_0 - s b_1
    double b1[4][QPT], b2[4][QPT];
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int qq = 0; qq < QPT; ++qq) b1[m][qq] = b2[m][qq] = 0.0;
    for (int j = a.n - 1; j >= 0; --j) {
#pragma unroll
        for (int qq = 0; qq < QPT; ++qq) {
            const double y = ybuf[(j * QPT + qq) * vstride];
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const double bn = fma(2.0 * sm[m][qq], b1[m][qq], y - b2[m][qq]);
                b2[m][qq] = b1[m][qq];
                b1[m][qq] = bn;
            }
        }
    }
    for (int g = 0; g < a.G; ++g) {
        const int kind = a.row_kind[g];
        if (kind == 0) continue;
#pragma unroll
        for (int qq = 0; qq < QPT; ++qq) {
            double f[4];
#pragma unroll
            for (int m = 0; m < 4; ++m) f[m] = fma(-sm[m][qq], b2[m][qq], b1[m][qq]);
            const double res = kind < 0 ? f[0] : tt_fd_reduce(kind, f[2], f[1], f[3], h);
            if (live[qq]) a.out[qg[qq] * a.G + g] = res;
        }
    }
}

// ---- host side ----------------------------------------------------------------------------------------
static ConstBank g_bank;  // residency of one image per device in c_tt (pcb_cbank.cuh)

void ttc_forget(const TTPlan *pl) {
    for (const auto &kv : pl->images) g_bank.forget(pl->dev, kv.second.id);
    for (const auto &kv : pl->gimages) g_bank.forget(pl->dev, kv.second.id);
}

// The image holding forward cores `need_fwd` and transposed cores `need_T` (bit k = core k), or
// nullptr when they do not fit in the bank.
const TTPlan::ConstImage *TTPlan::const_image(uint64_t need_fwd, uint64_t need_T) {
    if (!const_enabled) return nullptr;
    std::lock_guard<std::mutex> lock(image_mutex);
    const uint64_t key = need_fwd | (need_T << 32);
    auto it = images.find(key);
    if (it != images.end()) return it->second.data.empty() ? nullptr : &it->second;
    ConstImage &img = images[key];
    size_t total = 0;
    for (int k = 0; k < P.D; ++k) {
        const size_t sz = (size_t)(core_off[k + 1] - core_off[k]);
        total += ((need_fwd >> k) & 1 ? sz : 0) + ((need_T >> k) & 1 ? sz : 0);
    }
    if (total > (size_t)TT_CONST_MAX) return nullptr;  // remembered as an empty image
    img.id = next_plan_id();
    img.data.reserve(total);
    for (int k = 0; k < P.D; ++k) {
        img.coff[k] = img.coffT[k] = 0;
        if ((need_fwd >> k) & 1) {
            img.coff[k] = (int)img.data.size();
            img.data.insert(img.data.end(), h_fwd.begin() + core_off[k], h_fwd.begin() + core_off[k + 1]);
        }
    }
    for (int k = 0; k < P.D; ++k)
        if ((need_T >> k) & 1) {
            img.coffT[k] = (int)img.data.size();
            img.data.insert(img.data.end(), h_T.begin() + core_off[k], h_T.begin() + core_off[k + 1]);
        }
    return &img;
}

template <typename K, typename... Args>
static int ttc_launch(K kernel, const TTPlan *pl, const TTPlan::ConstImage *img, int qpt, int threads,
                      int nbuf, int64_t N, cudaStream_t st, Args... args) {
    const size_t smem = (size_t)nbuf * pl->P.rmaxp * qpt * threads * sizeof(double);
    if (smem > (size_t)pl->smem_optin)
        return fail(PCB_EUNSUPPORTED, "TT uniform-path kernel needs %zu B of shared memory", smem);
    PCB_CUDA(allow_dynamic_smem(kernel, smem, pl->smem_optin));
    const int64_t tile = (int64_t)threads * qpt;
    const int64_t ntiles = (N + tile - 1) / tile;
    if (ntiles > 0x7fffffffLL)
        return fail(PCB_EINVAL, "batch of %lld queries is too large for one launch", (long long)N);
    if (int rc = g_bank.acquire(pl->dev, img->id, st, [&](cudaStream_t s) {
            return cudaMemcpyToSymbolAsync(c_tt, img->data.data(), img->data.size() * sizeof(double), 0,
                                           cudaMemcpyHostToDevice, s);
        }))
        return rc;
    kernel<<<(int)ntiles, threads, smem, st>>>(args...);
    const cudaError_t e = cudaGetLastError();
    g_bank.release(pl->dev, st);
    g_launches.fetch_add(1);
    if (e != cudaSuccess) return fail(PCB_ECUDA, "TT kernel launch failed: %s", cudaGetErrorString(e));
    return PCB_OK;
}

// rank class of the plan: 8, 12 or 16 (only chunk widths up to it are compiled in)
static int ttc_rank_class(const TTPlan *pl) {
    int rmax = 1;
    for (int k = 0; k <= pl->P.D; ++k) rmax = pl->P.r[k] > rmax ? pl->P.r[k] : rmax;
    return rmax <= 8 ? 8 : (rmax <= 12 ? 12 : 16);
}

// the kernel's view of the plan: core offsets of THIS image
static TTParams ttc_params(const TTPlan *pl, const TTPlan::ConstImage *img) {
    TTParams P = pl->P;
    for (int k = 0; k < P.D; ++k) {
        P.coff[k] = img->coff[k];
        P.coffT[k] = img->coffT[k];
    }
    return P;
}

// Kernel variants (query slots per thread, rank class, threads per CTA = launch bound).  The plan
// asks for (qpt, threads); an exact match is used when it exists and fits, otherwise the first
// variant of the plan's rank class whose chain-vector buffers fit in shared memory.
template <int Q, int R, int T, int B>
static int ttc_value_go(const TTPlan *pl, const TTPlan::ConstImage *img, const TTParams &P,
                        const double *d_points, int64_t N, double *d_out, cudaStream_t st) {
    return ttc_launch(ttc_value_kernel<Q, R, T, B>, pl, img, Q, T, 1, N, st, P, d_points, N, d_out);
}
template <int Q, int R, int T, int B>
static int ttc_shared_go(const TTPlan *pl, const TTPlan::ConstImage *img, const TTParams &P,
                         const TTSharedProgram &prog, const double *d_points, int64_t N,
                         double *d_out, cudaStream_t st) {
    return ttc_launch(ttc_fd_shared_kernel<Q, R, T, B>, pl, img, Q, T, 2, N, st, P, prog, d_points, N,
                      d_out);
}
struct TTCValueVariant {
    int q, r, t, b;  // query slots per thread, rank class, threads per CTA, CTAs per SM
    int (*go)(const TTPlan *, const TTPlan::ConstImage *, const TTParams &, const double *, int64_t,
              double *, cudaStream_t);
};
struct TTCSharedVariant {
    int q, r, t, b;
    int (*go)(const TTPlan *, const TTPlan::ConstImage *, const TTParams &, const TTSharedProgram &,
              const double *, int64_t, double *, cudaStream_t);
};
#define VV(Q, R, T, B) {Q, R, T, B, ttc_value_go<Q, R, T, B>}
#define SV(Q, R, T, B) {Q, R, T, B, ttc_shared_go<Q, R, T, B>}
// first entry of a rank class = its default; B = 2: two CTAs per SM, so one CTA's coordinate loads,
// chain-vector initialisation and stores overlap the other's arithmetic
static const TTCValueVariant kValueVariants[] = {
    // B = 0: no minBlocks bound (the register budget follows from the thread count alone); B = 2: two
    // CTAs per SM.  ptxas' uniform-datapath decision depends on the register budget it is given
    // (tools/check_sass.py): <2,12,512> keeps its bank reads on LDCU at 64 registers (B = 2) and
    // loses them at 122 (B = 0/1) -- 3.6e9 against 2.05e9 values/s on the 5-D train.
    VV(2, 8, 512, 2), VV(2, 12, 512, 2), VV(2, 16, 512, 0),
    VV(3, 12, 320, 0), VV(2, 16, 512, 2),
};
static const TTCSharedVariant kSharedVariants[] = {
    SV(2, 8, 512, 0), SV(2, 12, 512, 0), SV(2, 16, 384, 0),
    SV(2, 12, 384, 0), SV(2, 12, 256, 2), SV(2, 8, 256, 2),
};
#undef VV
#undef SV

template <typename V, size_t NV>
static const V *ttc_pick(const V (&tab)[NV], const TTPlan *pl, int qpt, int threads, int rc_, int nbuf) {
    auto fits = [&](const V &v) {  // all resident CTAs of an SM share its shared memory
        const size_t ctas = v.b > 1 ? v.b : 1;
        return ctas * nbuf * pl->P.rmaxp * v.q * v.t * sizeof(double) + ctas * 1024 <=
               (size_t)pl->smem_optin + 1024;
    };
    for (const V &v : tab)
        if (v.q == qpt && v.r == rc_ && v.t == threads && fits(v)) return &v;
    for (const V &v : tab)
        if (v.r == rc_ && fits(v)) return &v;
    return nullptr;
}

int ttc_launch_value(TTPlan *pl, const double *d_points, int64_t N, double *d_out, cudaStream_t st,
                     bool *fits) {
    const uint64_t all = pl->P.D >= 32 ? 0xffffffffull : ((1ull << pl->P.D) - 1);
    const TTPlan::ConstImage *img = pl->const_image(all, 0);
    *fits = img != nullptr;
    if (!img) return PCB_OK;
    const TTParams P = ttc_params(pl, img);
    const TTCValueVariant *v = ttc_pick(kValueVariants, pl, pl->const_qpt_value, pl->const_threads_value,
                                        ttc_rank_class(pl), 1);
    if (!v) return fail(PCB_EUNSUPPORTED, "no uniform-path TT value kernel for this configuration");
    return v->go(pl, img, P, d_points, N, d_out, st);
}

// cores ttc_fd_shared_kernel reads for `prog`: left sweep up to the last slot, right sweeps down to
// the first, each slot's own core in the orientation of its coefficient pass
static void ttc_shared_need(const TTPlan *pl, const TTSharedProgram &prog, uint64_t *need_fwd,
                            uint64_t *need_T) {
    *need_fwd = *need_T = 0;
    const int a_min = prog.slot_dim[0], a_max = prog.slot_dim[prog.n_slots - 1];
    for (int k = 0; k < a_max; ++k) *need_fwd |= 1ull << k;
    for (int k = a_min + 1; k < pl->P.D; ++k) *need_T |= 1ull << k;
    for (int t = 0; t < prog.n_slots; ++t) {
        const int a = prog.slot_dim[t];
        if (pl->P.r[a + 1] >= pl->P.r[a])
            *need_fwd |= 1ull << a;
        else
            *need_T |= 1ull << a;
    }
}

static int ttc_launch_shared_one(TTPlan *pl, const TTPlan::ConstImage *img, const TTSharedProgram &prog,
                                 const double *d_points, int64_t N, double *d_out, cudaStream_t st) {
    const TTParams P = ttc_params(pl, img);
    // both chain-vector buffers must fit in shared memory (ttc_pick checks)
    const TTCSharedVariant *v = ttc_pick(kSharedVariants, pl, pl->const_qpt_shared,
                                         pl->const_threads_shared, ttc_rank_class(pl), 2);
    if (!v) return fail(PCB_EUNSUPPORTED, "no uniform-path TT shared-FD kernel for this configuration");
    return v->go(pl, img, P, prog, d_points, N, d_out, st);
}

// Split of a Greek set whose union of cores does not fit the bank (Greeks at both ends of a long
// train need every core in both orientations): one sub-program per differentiated dim, each with
// its own image and writing only its own output rows (value rows go with the first).
static bool ttc_split(TTPlan *pl, const TTSharedProgram &prog, TTSharedProgram *sub,
                      const TTPlan::ConstImage **imgs) {
    constexpr int SKIP = 1 << 20;  // row_slot that matches no slot and is not a value row
    if (prog.n_slots < 2) return false;
    for (int t = 0; t < prog.n_slots; ++t) {
        sub[t].G = prog.G;
        sub[t].n_slots = 1;
        sub[t].slot_dim[0] = prog.slot_dim[t];
        for (int g = 0; g < prog.G; ++g) {
            const int rs = prog.row_slot[g];
            sub[t].row_slot[g] = rs == t ? 0 : ((rs < 0 && t == 0) ? -1 : SKIP);
            sub[t].row_ord[g] = prog.row_ord[g];
        }
        uint64_t need_fwd, need_T;
        ttc_shared_need(pl, sub[t], &need_fwd, &need_T);
        imgs[t] = pl->const_image(need_fwd, need_T);
        if (!imgs[t]) return false;
    }
    return true;
}

// 0: does not fit the bank; 1: one launch; 2: one launch per differentiated dim
int ttc_shared_fits(TTPlan *pl, const TTSharedProgram &prog) {
    if (!pl->const_enabled) return 0;
    uint64_t need_fwd, need_T;
    ttc_shared_need(pl, prog, &need_fwd, &need_T);
    if (pl->const_image(need_fwd, need_T)) return 1;
    TTSharedProgram sub[TT_MAX_G];
    const TTPlan::ConstImage *imgs[TT_MAX_G];
    return ttc_split(pl, prog, sub, imgs) ? 2 : 0;
}

int ttc_launch_shared(TTPlan *pl, const TTSharedProgram &prog, const double *d_points, int64_t N,
                      double *d_out, cudaStream_t st, bool *fits) {
    uint64_t need_fwd, need_T;
    ttc_shared_need(pl, prog, &need_fwd, &need_T);
    const TTPlan::ConstImage *img = pl->const_image(need_fwd, need_T);
    *fits = img != nullptr;
    if (img) return ttc_launch_shared_one(pl, img, prog, d_points, N, d_out, st);
    TTSharedProgram sub[TT_MAX_G];
    const TTPlan::ConstImage *imgs[TT_MAX_G];
    if (!ttc_split(pl, prog, sub, imgs)) return PCB_OK;  // *fits stays false: shared-memory kernel
    *fits = true;
    for (int t = 0; t < prog.n_slots; ++t)  // left sweeps restart per launch
        if (int rc = ttc_launch_shared_one(pl, imgs[t], sub[t], d_points, N, d_out, st)) return rc;
    return PCB_OK;
}

// ---- per-core launches (large trains) -----------------------------------------------------------------
constexpr int TTG_WAVES = 27;  // a tile = this many full waves of step CTAs (2 per SM x 512 queries):
                               // ~4.1M queries on 148 SMs, no partial last wave, 2.6 GB of FD scratch at rank 20
// step kernel: two 256-thread CTAs per SM, so that one CTA's state load / store phases overlap the
// other's arithmetic (one 512-thread CTA: 2.94e8 values/s on the 10-D rank-20 train)
constexpr int TTG_THREADS_STEP = 256, TTG_THREADS_COEFF = 256, TTG_QPT = 2;

// chunk-major image of core k (orientation 0: forward [i][j][l]; 1: transposed [l][j][i])
static const TTPlan::ConstImage *ttg_image(TTPlan *pl, int k, int orient) {
    std::lock_guard<std::mutex> lock(pl->image_mutex);
    const int key = 2 * k + orient;
    auto it = pl->gimages.find(key);
    if (it != pl->gimages.end()) return &it->second;
    TTPlan::ConstImage &img = pl->gimages[key];
    const TTParams &P = pl->P;
    const int rows = orient ? P.r[k + 1] : P.r[k], cols = orient ? P.r[k] : P.r[k + 1], n = P.n[k];
    const double *src = (orient ? pl->h_T.data() : pl->h_fwd.data()) + pl->core_off[k];  // [rows][n][cols]
    const int nchunks = (cols + 15) / 16;
    img.id = next_plan_id();
    img.data.resize((size_t)rows * n * cols);
    int l0 = 0;
    size_t pos = 0;
    for (int c = 0; c < TTG_MAX_CHUNKS; ++c) img.coff[c] = img.coffT[c] = 0;
    for (int c = 0; c < nchunks; ++c) {
        const int w = cols / nchunks + (c < cols % nchunks ? 1 : 0);  // balanced widths, all <= 16
        img.coff[c] = (int)pos;
        img.coffT[c] = w;
        for (int i = 0; i < rows; ++i)
            for (int j = 0; j < n; ++j)
                for (int l = 0; l < w; ++l)
                    img.data[pos++] = src[((size_t)i * n + j) * cols + l0 + l];
        l0 += w;
    }
    img.coff[TTG_MAX_CHUNKS] = nchunks;
    // page-lock the image: an upload from pageable memory synchronises the stream first, which
    // would drain the GPU between the per-core launches
    cudaHostRegister(img.data.data(), img.data.size() * sizeof(double), cudaHostRegisterDefault);
    cudaGetLastError();  // registration is an optimisation only
    return &img;
}

template <typename K, typename A>
static int ttg_launch(K kernel, TTPlan *pl, const TTPlan::ConstImage *img, const A &args, int threads,
                      size_t smem, int64_t tile_n, const double *d_points, int64_t N, cudaStream_t st) {
    if (smem > (size_t)pl->smem_optin)
        return fail(PCB_EUNSUPPORTED, "TT per-core kernel needs %zu B of shared memory", smem);
    PCB_CUDA(allow_dynamic_smem(kernel, smem, pl->smem_optin));
    const int64_t cta = (int64_t)threads * TTG_QPT;
    const int64_t blocks = (tile_n + cta - 1) / cta;
    if (int rc = g_bank.acquire(pl->dev, img->id, st, [&](cudaStream_t s) {
            return cudaMemcpyToSymbolAsync(c_tt, img->data.data(), img->data.size() * sizeof(double), 0,
                                           cudaMemcpyHostToDevice, s);
        }))
        return rc;
    kernel<<<(int)blocks, threads, smem, st>>>(args, d_points, N);
    const cudaError_t e = cudaGetLastError();
    g_bank.release(pl->dev, st);
    g_launches.fetch_add(1);
    if (e != cudaSuccess) return fail(PCB_ECUDA, "TT kernel launch failed: %s", cudaGetErrorString(e));
    return PCB_OK;
}

// one core applied to the chain state: forward core k (left sweep) or transposed core k (right sweep)
static int ttg_step(TTPlan *pl, int k, int orient, const double *s_in, double *s_out, int64_t pstride,
                    int64_t q_begin, int64_t tile_n, const double *d_points, int64_t N, cudaStream_t st) {
    const TTParams &P = pl->P;
    const TTPlan::ConstImage *img = ttg_image(pl, k, orient);
    TTGStep a;
    a.r_in = orient ? P.r[k + 1] : P.r[k];
    a.r_out = orient ? P.r[k] : P.r[k + 1];
    a.n = P.n[k];
    a.col = P.perm[k];
    a.D = P.D;
    a.lo = P.lo[k];
    a.hi = P.hi[k];
    a.nchunks = img->coff[TTG_MAX_CHUNKS];
    for (int c = 0; c < TTG_MAX_CHUNKS; ++c) {
        a.cbase[c] = img->coff[c];
        a.cw[c] = img->coffT[c];
    }
    a.s_in = a.r_in == 1 ? nullptr : s_in;
    a.s_out = s_out;
    a.pstride = pstride;
    a.q_begin = q_begin;
    const size_t smem = (size_t)a.r_in * TTG_QPT * TTG_THREADS_STEP * sizeof(double);
    return ttg_launch(ttc_gstep_kernel<TTG_QPT, TTG_THREADS_STEP>, pl, img, a, TTG_THREADS_STEP, smem,
                      tile_n, d_points, N, st);
}

// Stream-ordered scratch for the chain states of one tile, from a PRIVATE memory pool owned by
// this library (one per device, created on first use).  The pool keeps up to TTG_POOL_KEEP bytes
// across calls (a default-threshold pool hands everything back to the driver at every
// synchronisation, i.e. a fresh allocation per call); the device's default pool and PyTorch's
// caching allocator are left untouched.
constexpr uint64_t TTG_POOL_KEEP = 6ull << 30;
static cudaMemPool_t ttg_pool(int dev) {
    static std::mutex m;
    static cudaMemPool_t pools[ConstBank::MAX_DEV] = {nullptr};
    if (dev < 0 || dev >= ConstBank::MAX_DEV) return nullptr;
    std::lock_guard<std::mutex> lock(m);
    if (!pools[dev]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        cudaMemPool_t pool = nullptr;
        if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        uint64_t keep = TTG_POOL_KEEP;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        pools[dev] = pool;
    }
    return pools[dev];
}

struct TTGScratch {
    double *p = nullptr;
    cudaStream_t st;
    TTGScratch(int dev, size_t doubles, cudaStream_t s) : st(s) {
        cudaMemPool_t pool = ttg_pool(dev);
        const cudaError_t e = pool ? cudaMallocFromPoolAsync(&p, doubles * sizeof(double), pool, s)
                                   : cudaMallocAsync(&p, doubles * sizeof(double), s);
        if (e != cudaSuccess) {
            cudaGetLastError();
            p = nullptr;
        }
    }
    ~TTGScratch() {
        if (p) cudaFreeAsync(p, st);
    }
};

static int ttg_rmax(const TTPlan *pl) {
    int rmax = 1;
    for (int k = 0; k <= pl->P.D; ++k) rmax = std::max(rmax, pl->P.r[k]);
    return rmax;
}

int ttg_launch_value(TTPlan *pl, const double *d_points, int64_t N, double *d_out, cudaStream_t st) {
    const TTParams &P = pl->P;
    const int64_t tile = std::min<int64_t>((int64_t)TTG_WAVES * pl->sm_count * 1024, (N + 1023) / 1024 * 1024);
    const int rmax = ttg_rmax(pl);
    TTGScratch scratch(pl->dev, (size_t)2 * tile * rmax, st);
    if (!scratch.p) return fail(PCB_ENOMEM, "cannot allocate %zu B of TT chain state", (size_t)16 * tile * rmax);
    double *buf[2] = {scratch.p, scratch.p + (size_t)tile * rmax};
    for (int64_t q0 = 0; q0 < N; q0 += tile) {
        const int64_t tn = std::min(tile, N - q0);
        int cur = 0;
        for (int k = 0; k < P.D; ++k) {
            double *dst = k == P.D - 1 ? d_out + q0 : buf[cur ^ 1];
            if (int rc = ttg_step(pl, k, 0, buf[cur], dst, tile, q0, tn, d_points, N, st)) return rc;
            cur ^= 1;
        }
    }
    return PCB_OK;
}

bool ttg_shared_fits(const TTPlan *pl, const TTSharedProgram &prog) {
    if (!pl->gstream_fd_ok) return false;
    for (int t = 0; t < prog.n_slots; ++t) {
        const int a = prog.slot_dim[t];
        const size_t smem = (size_t)(pl->P.r[a] + pl->P.r[a + 1] + pl->P.n[a]) * TTG_QPT *
                            TTG_THREADS_COEFF * sizeof(double);
        if (smem > (size_t)pl->smem_optin) return false;
    }
    return true;
}

int ttg_launch_shared(TTPlan *pl, const TTSharedProgram &prog, const double *d_points, int64_t N,
                      double *d_out, cudaStream_t st, bool *fits) {
    const TTParams &P = pl->P;
    const int D = P.D;
    // the coefficient pass of every differentiated dim keeps r_rows + r_acc + n rows in shared memory
    *fits = ttg_shared_fits(pl, prog);
    if (!*fits) return PCB_OK;  // caller falls back to the shared-memory kernels
    const int64_t tile = std::min<int64_t>((int64_t)TTG_WAVES * pl->sm_count * 1024, (N + 1023) / 1024 * 1024);
    const int rmax = ttg_rmax(pl);
    TTGScratch scratch(pl->dev, (size_t)4 * tile * rmax, st);
    if (!scratch.p) return fail(PCB_ENOMEM, "cannot allocate %zu B of TT chain state", (size_t)32 * tile * rmax);
    double *L[2] = {scratch.p, scratch.p + (size_t)tile * rmax};
    double *R[2] = {L[1] + (size_t)tile * rmax, L[1] + (size_t)2 * tile * rmax};
    for (int64_t q0 = 0; q0 < N; q0 += tile) {
        const int64_t tn = std::min(tile, N - q0);
        int lcur = 0, lpos = 0;  // L[lcur] = left product over storage dims [0, lpos)
        for (int t = 0; t < prog.n_slots; ++t) {
            const int a = prog.slot_dim[t];
            int rcur = 0;
            for (int k = D - 1; k > a; --k) {  // right sweep on transposed cores
                if (int rc = ttg_step(pl, k, 1, R[rcur], R[rcur ^ 1], tile, q0, tn, d_points, N, st)) return rc;
                rcur ^= 1;
            }
            for (int k = lpos; k < a; ++k) {  // left sweep continues
                if (int rc = ttg_step(pl, k, 0, L[lcur], L[lcur ^ 1], tile, q0, tn, d_points, N, st)) return rc;
                lcur ^= 1;
            }
            lpos = a;
            // coefficient pass on core a, register-accumulated index = the wider rank
            const int orient = P.r[a + 1] >= P.r[a] ? 0 : 1;
            const TTPlan::ConstImage *img = ttg_image(pl, a, orient);
            TTGCoeff c;
            c.r_rows = orient ? P.r[a + 1] : P.r[a];
            c.r_acc = orient ? P.r[a] : P.r[a + 1];
            c.n = P.n[a];
            c.col = P.perm[a];
            c.D = D;
            c.G = prog.G;
            c.lo = P.lo[a];
            c.hi = P.hi[a];
            c.nchunks = img->coff[TTG_MAX_CHUNKS];
            for (int i = 0; i < TTG_MAX_CHUNKS; ++i) {
                c.cbase[i] = img->coff[i];
                c.cw[i] = img->coffT[i];
            }
            const double *vl = a == 0 ? nullptr : L[lcur], *vr = a == D - 1 ? nullptr : R[rcur];
            c.s_rows = orient ? vr : vl;
            c.s_acc = orient ? vl : vr;
            c.pstride = tile;
            c.q_begin = q0;
            for (int g = 0; g < prog.G; ++g) {
                const int rs = prog.row_slot[g];
                c.row_kind[g] = rs == t ? prog.row_ord[g] : ((rs < 0 && t == 0) ? -1 : 0);
            }
            c.out = d_out;
            const size_t smem =
                (size_t)(c.r_rows + c.r_acc + c.n) * TTG_QPT * TTG_THREADS_COEFF * sizeof(double);
            if (int rc = ttg_launch(ttc_gcoeff_kernel<TTG_QPT, TTG_THREADS_COEFF>, pl, img, c,
                                    TTG_THREADS_COEFF, smem, tn, d_points, N, st))
                return rc;
        }
    }
    return PCB_OK;
}

}  // namespace pcb
