// Thread-per-query barycentric evaluator on tensor-product Chebyshev grids (device side).
//
// Shared by the full-tensor FMA path (pcb_full_eval algo 1), ChebyshevSpline pieces and
// ChebyshevSlider slides.  Follows reference barycentric.py:1035-1046 per point:
//   diff = x - nodes; |diff| < 1e-14 (first hit) -> take that slice, else w/diff normalised.
// The per-dimension weight rows are normalised once (w_hat = (w/diff) / sum(w/diff)), parked in a
// per-thread shared-memory column, and the tensor is contracted depth-first with one running
// sum per nesting level held in registers (template recursion over the dimension).
#pragma once

#include "pcb_common.cuh"

namespace pcb {

constexpr int GRID_MAXD = 8;       // deepest grid the FMA evaluator instantiates
constexpr double NODE_EPS = 1e-14;  // reference barycentric.py:1040

// One tensor-product grid (a full interpolant, one spline piece, or one slider slide).
struct GridDesc {
    int D;
    int n[GRID_MAXD];
    int node_off;         // offset (doubles) of this grid's nodes/weights, dims concatenated
    int sum_n;            // sum of n[]
    long long size;       // prod of n[]
    long long tensor_off;  // offset (doubles) of tensor 0; tensor g at tensor_off + g * size
};

// Normalised barycentric weight row of one dimension into the thread's smem column.
// ws[i * stride], i < n.  Reference barycentric.py:1038-1045.
__device__ __forceinline__ void grid_weight_row(double x, int n, const double *__restrict__ nodes,
                                                const double *__restrict__ weights, double *ws,
                                                int stride) {
    int hit = -1;
    double sum = 0.0;
    for (int i = 0; i < n; ++i) {
        const double diff = x - __ldg(nodes + i);
        if (hit < 0 && fabs(diff) < NODE_EPS) hit = i;
        const double w = __ldg(weights + i) / diff;
        ws[i * stride] = w;
        sum += w;
    }
    if (hit >= 0) {
        for (int i = 0; i < n; ++i) ws[i * stride] = (i == hit) ? 1.0 : 0.0;
    } else {
        const double inv = 1.0 / sum;
        for (int i = 0; i < n; ++i) ws[i * stride] *= inv;
    }
}

template <int LEVEL, int D>
struct GridContract {
    __device__ __forceinline__ static double run(const double *__restrict__ t, const int *n,
                                                 const long long *stride, const double *ws,
                                                 int wstride) {
        double acc = 0.0;
        const int nl = n[LEVEL];
        if (LEVEL == D - 1) {
#pragma unroll 4
            for (int i = 0; i < nl; ++i) acc = fma(__ldg(t + i), ws[i * wstride], acc);
        } else {
            const long long st = stride[LEVEL];
            const double *wnext = ws + (size_t)nl * wstride;
            for (int i = 0; i < nl; ++i) {
                const double w = ws[i * wstride];
                // a zero weight row entry (node hit) skips the whole sub-tensor
                if (w != 0.0)
                    acc = fma(w, GridContract<(LEVEL + 1 < D ? LEVEL + 1 : LEVEL), D>::run(
                                     t + i * st, n, stride, wnext, wstride),
                              acc);
            }
        }
        return acc;
    }
};

// Contract one tensor with the weight rows in ws (dims concatenated, each n[d] entries).
__device__ __forceinline__ double grid_contract(const GridDesc &gd, const double *__restrict__ t,
                                                const double *ws, int wstride) {
    long long stride[GRID_MAXD];
    long long s = 1;
    for (int d = gd.D - 1; d >= 0; --d) {
        stride[d] = s;
        s *= gd.n[d];
    }
    switch (gd.D) {
        case 1: return GridContract<0, 1>::run(t, gd.n, stride, ws, wstride);
        case 2: return GridContract<0, 2>::run(t, gd.n, stride, ws, wstride);
        case 3: return GridContract<0, 3>::run(t, gd.n, stride, ws, wstride);
        case 4: return GridContract<0, 4>::run(t, gd.n, stride, ws, wstride);
        case 5: return GridContract<0, 5>::run(t, gd.n, stride, ws, wstride);
        case 6: return GridContract<0, 6>::run(t, gd.n, stride, ws, wstride);
        case 7: return GridContract<0, 7>::run(t, gd.n, stride, ws, wstride);
        default: return GridContract<0, 8>::run(t, gd.n, stride, ws, wstride);
    }
}

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

}  // namespace pcb
