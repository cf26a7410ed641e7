// Thread-per-query barycentric evaluator on tensor-product Chebyshev grids (device side).
//
// Shared by the full-tensor FMA path (pcb_full_eval algo 1), ChebyshevSpline pieces and
// ChebyshevSlider slides.  Follows reference barycentric.py:1035-1046 per point:
//   diff = x - nodes; |diff| < 1e-14 (first hit) -> take that slice, else w/diff normalised.
//
// * Weight rows: w_hat_i = (w_i / d_i) / sum_k (w_k / d_k) is evaluated in the division-free
//   product form  a_i = w_i * prod_{j != i} d_j,  w_hat_i = a_i / sum_k a_k  (prefix/suffix
//   products, ONE division per dimension instead of n): an IEEE fp64 division costs ~25
//   instructions, and on 15x15 spline pieces the 30 divisions were 2/3 of the work.  Same value
//   up to a few ulp.
// * The G output tensors of a grid are stored INTERLEAVED in blocks of GB outputs,
//   t[block][element][GB], so one 128-bit load feeds two FMAs.
// * The last dimension's weights are held in registers (n_last <= GRID_NL), the outer ones in a
//   per-thread shared-memory column; the tensor is contracted depth-first with one running sum
//   per nesting level and output in registers (template recursion over the dimension).
#pragma once

#include "pcb_common.cuh"

namespace pcb {

constexpr int GRID_MAXD = 8;        // deepest grid the FMA evaluator instantiates
constexpr int GRID_NL = 16;         // last-axis weights kept in registers up to this many nodes
constexpr double NODE_EPS = 1e-14;  // reference barycentric.py:1040

// One tensor-product grid (a full interpolant, one spline piece, or one slider slide).
struct GridDesc {
    int D;
    int n[GRID_MAXD];
    int node_off;          // offset (doubles) of this grid's nodes/weights, dims concatenated
    int sum_n;             // sum of n[]
    long long size;        // prod of n[]
    long long tensor_off;  // offset (doubles) of output block 0; block b at + b * size * GB
};

// Host helper: interleave G tensors (each `size` doubles) into ceil(G/GB) blocks [elem][GB].
inline void grid_interleave(const double *const *tensors, int G, int GB, long long size, double *dst) {
    const int nblk = (G + GB - 1) / GB;
    for (int b = 0; b < nblk; ++b)
        for (long long e = 0; e < size; ++e)
            for (int j = 0; j < GB; ++j) {
                const int g = b * GB + j;
                dst[((long long)b * size + e) * GB + j] = g < G && tensors[g] ? tensors[g][e] : 0.0;
            }
}
inline int grid_pick_gb(int G) { return G >= 3 ? 4 : (G == 2 ? 2 : 1); }

// Does this spline plan (pcb_piecewise.cu) run from the constant bank?  Used by the full-tensor plan,
// which hands small tensors to a one-piece spline plan to get the uniform-datapath evaluator.
bool spline_plan_uses_bank(const void *plan);

#ifdef __CUDACC__

// Normalised barycentric weight row of one dimension into row[i * stride], i < n (also used as
// scratch for the prefix products).  Reference barycentric.py:1038-1045 (node coincidence: first
// |x - node| < 1e-14 -> one-hot row).
__device__ __forceinline__ void grid_weight_row(double x, int n, const double *__restrict__ nodes,
                                                const double *__restrict__ weights, double *row,
                                                int stride) {
    // pass 1: prefix products prod_{j < i} d_j
    int hit = -1;
    double pre = 1.0;
    for (int i = 0; i < n; ++i) {
        const double d = x - __ldg(nodes + i);
        if (hit < 0 && fabs(d) < NODE_EPS) hit = i;
        row[i * stride] = pre;
        pre *= d;
    }
    if (hit >= 0) {
        for (int i = 0; i < n; ++i) row[i * stride] = i == hit ? 1.0 : 0.0;
        return;
    }
    // pass 2 (descending): a_i = w_i * prefix_i * suffix_i
    double suf = 1.0, sum = 0.0;
    for (int i = n - 1; i >= 0; --i) {
        const double a = __ldg(weights + i) * row[i * stride] * suf;
        row[i * stride] = a;
        sum += a;
        suf *= x - __ldg(nodes + i);
    }
    if (fabs(sum) > 1e-280 && fabs(sum) < 1e280) {
        const double inv = 1.0 / sum;
        for (int i = 0; i < n; ++i) row[i * stride] *= inv;
        return;
    }
    // products left the fp64 range (very wide domain / far extrapolation): the reference's own
    // form, one division per node
    sum = 0.0;
    for (int i = 0; i < n; ++i) {
        const double w = __ldg(weights + i) / (x - __ldg(nodes + i));
        row[i * stride] = w;
        sum += w;
    }
    for (int i = 0; i < n; ++i) row[i * stride] /= sum;
}

template <int GB>
__device__ __forceinline__ void grid_load(const double *__restrict__ p, double (&v)[GB]) {
    if constexpr (GB == 1) {
        v[0] = __ldg(p);
    } else if constexpr (GB == 2) {
        const double2 a = __ldg(reinterpret_cast<const double2 *>(p));
        v[0] = a.x;
        v[1] = a.y;
    } else {
        const double2 a = __ldg(reinterpret_cast<const double2 *>(p));
        const double2 b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
        v[0] = a.x;
        v[1] = a.y;
        v[2] = b.x;
        v[3] = b.y;
    }
}

// Depth-first contraction.  ws: weight rows of dims [LEVEL, D-1) in a per-thread smem column;
// wl[]: the last dim's weights in registers (or, when n_last > GRID_NL, also from ws).
template <int LEVEL, int D, int GB>
struct GridContract {
    __device__ __forceinline__ static void run(const double *__restrict__ t, const int *n,
                                               const long long *stride, const double *ws,
                                               int wstride, const double (&wl)[GRID_NL],
                                               bool last_in_regs, double (&out)[GB]) {
#pragma unroll
        for (int j = 0; j < GB; ++j) out[j] = 0.0;
        const int nl = n[LEVEL];
        if constexpr (LEVEL == D - 1) {
            // four independent partial sums per output: the row dot product is otherwise one
            // dependent FMA chain of n_last links and the evaluator is latency-bound
            double part[4][GB];
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int j = 0; j < GB; ++j) part[c][j] = 0.0;
            if (last_in_regs) {
#pragma unroll
                for (int i = 0; i < GRID_NL; ++i) {
                    if (i < nl) {
                        double v[GB];
                        grid_load<GB>(t + (size_t)i * GB, v);
#pragma unroll
                        for (int j = 0; j < GB; ++j) part[i & 3][j] = fma(wl[i], v[j], part[i & 3][j]);
                    }
                }
            } else {
                for (int i0 = 0; i0 < nl; i0 += 4) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (i0 + c < nl) {
                            double v[GB];
                            grid_load<GB>(t + (size_t)(i0 + c) * GB, v);
                            const double w = ws[(i0 + c) * wstride];
#pragma unroll
                            for (int j = 0; j < GB; ++j) part[c][j] = fma(w, v[j], part[c][j]);
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < GB; ++j) out[j] = (part[0][j] + part[1][j]) + (part[2][j] + part[3][j]);
        } else {
            const long long st = stride[LEVEL] * GB;
            const double *wnext = ws + (size_t)nl * wstride;
            for (int i = 0; i < nl; ++i) {
                const double w = ws[i * wstride];
                if (w != 0.0) {  // a one-hot row (node hit) skips the other sub-tensors
                    double sub[GB];
                    GridContract<(LEVEL + 1 < D ? LEVEL + 1 : LEVEL), D, GB>::run(
                        t + i * st, n, stride, wnext, wstride, wl, last_in_regs, sub);
#pragma unroll
                    for (int j = 0; j < GB; ++j) out[j] = fma(w, sub[j], out[j]);
                }
            }
        }
    }
};

// Weight rows of all dims of grid `gd` for coordinates x(d) into the per-thread smem column `ws`
// (dims concatenated); the last row is then copied into registers when it fits.  Returns whether
// it did.
template <typename Coord>
__device__ __forceinline__ bool grid_weights(const GridDesc &gd, const double *__restrict__ nodes,
                                             const double *__restrict__ weights, Coord x, double *ws,
                                             int wstride, double (&wl)[GRID_NL]) {
    int off = 0;
    for (int d = 0; d < gd.D; ++d) {
        grid_weight_row(x(d), gd.n[d], nodes + gd.node_off + off, weights + gd.node_off + off,
                        ws + (size_t)off * wstride, wstride);
        off += gd.n[d];
    }
    const int nl = gd.n[gd.D - 1];
    const bool last_in_regs = nl <= GRID_NL;
    if (last_in_regs) {
        const double *row = ws + (size_t)(off - nl) * wstride;
#pragma unroll
        for (int i = 0; i < GRID_NL; ++i) wl[i] = i < nl ? row[i * wstride] : 0.0;
    }
    return last_in_regs;
}

// Contract one output block with the prepared weights.  DM = deepest grid this instantiation
// handles (keeps the register allocation of low-dimensional plans small).
#define GRID_CASE(K)                                                                              \
    case K:                                                                                       \
        if constexpr (K <= DM)                                                                    \
            GridContract<0, K, GB>::run(t, gd.n, stride, ws, wstride, wl, last_in_regs, out);     \
        break;

template <int GB, int DM>
__device__ __forceinline__ void grid_contract(const GridDesc &gd, const double *__restrict__ t,
                                              const double *ws, int wstride,
                                              const double (&wl)[GRID_NL], bool last_in_regs,
                                              double (&out)[GB]) {
    long long stride[DM];
    long long s = 1;
    for (int d = gd.D - 1; d >= 0; --d) {
        stride[d] = s;
        s *= gd.n[d];
    }
#pragma unroll
    for (int j = 0; j < GB; ++j) out[j] = 0.0;
    switch (gd.D) {
        GRID_CASE(1) GRID_CASE(2) GRID_CASE(3) GRID_CASE(4) GRID_CASE(5) GRID_CASE(6) GRID_CASE(7)
        GRID_CASE(8)
    }
}

// All G outputs of one grid at the prepared weights: out[g * ostride], g < G.
template <int GB, int DM>
__device__ __forceinline__ void grid_eval_outputs(const GridDesc &gd, const double *__restrict__ tensors,
                                                  int G, const double *ws, int wstride,
                                                  const double (&wl)[GRID_NL], bool last_in_regs,
                                                  double *out, int ostride) {
    for (int b = 0; b * GB < G; ++b) {
        double r[GB];
        grid_contract<GB, DM>(gd, tensors + gd.tensor_off + (long long)b * gd.size * GB, ws, wstride,
                              wl, last_in_regs, r);
#pragma unroll
        for (int j = 0; j < GB; ++j)
            if (b * GB + j < G) out[(size_t)(b * GB + j) * ostride] = r[j];
    }
}

// Launch-time choice of the (GB, DM) instantiation of a kernel template K<GB, DM>.
inline int grid_pick_dm(int maxD) { return maxD <= 2 ? 2 : (maxD <= 3 ? 3 : (maxD <= 4 ? 4 : 8)); }
#define GRID_KERNEL_TABLE(K, gb, dm)                                                             \
    ((gb) == 4 ? ((dm) == 2 ? (const void *)K<4, 2> : (dm) == 3 ? (const void *)K<4, 3>           \
                             : (dm) == 4 ? (const void *)K<4, 4> : (const void *)K<4, 8>)         \
     : (gb) == 2 ? ((dm) == 2 ? (const void *)K<2, 2> : (dm) == 3 ? (const void *)K<2, 3>         \
                               : (dm) == 4 ? (const void *)K<2, 4> : (const void *)K<2, 8>)       \
                 : ((dm) == 2 ? (const void *)K<1, 2> : (dm) == 3 ? (const void *)K<1, 3>         \
                               : (dm) == 4 ? (const void *)K<1, 4> : (const void *)K<1, 8>))

// Integer piece lookup, bit-exact with reference spline.py:677-690:
//   idx_d = searchsorted(knots_d, x, side="right") = #{knot <= x}, NaN sorts last -> len(knots);
//   the clip to pieces_d - 1 = len(knots_d) is then a no-op; flat index in C-order.
__device__ __forceinline__ int spline_piece_index(int D, const int *__restrict__ num_knots,
                                                  const int *__restrict__ knot_off,
                                                  const double *__restrict__ knots,
                                                  const double *__restrict__ x /* D coords */) {
    int flat = 0;
    for (int d = 0; d < D; ++d) {
        const int nk = num_knots[d];
        int idx = 0;
        if (nk > 0) {
            const double xv = x[d];
            if (xv != xv) {
                idx = nk;
            } else {
                const double *kn = knots + knot_off[d];
                for (int k = 0; k < nk; ++k) idx += (__ldg(kn + k) <= xv) ? 1 : 0;
            }
        }
        flat = flat * (nk + 1) + idx;
    }
    return flat;
}

#endif  // __CUDACC__

}  // namespace pcb
