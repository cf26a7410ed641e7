/*
 * oracle.c -- plain-C restatement of the reference's vectorised evaluation algorithms.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/np_oracle.py for the rules): used by tests/ as a fast
 * checker at sizes the NumPy oracle cannot reach, and by bench.py as the timed CPU baseline
 * ("port").  Never linked into, or called from, the product path.
 *
 * Parity status: PINNED -- tests/test_oracle_golden.py checks every function against outputs of
 * the unmodified reference (the .npz files under tests/golden).
 *
 * Each function cites the reference lines it follows (src/pychebyshev/ of the reference).
 * Threads: single-threaded C; callers shard the query range over threads (ctypes releases the
 * GIL), see oracle/c_oracle.py.  The arithmetic per point is sequential.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))
#define ORC_MAXN 256
#define ORC_MAXR 256

ORC_API int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* numpy.polynomial.chebyshev.chebval(x, eye(n)) -- Clenshaw on unit coefficient vectors
 * (tensor_train.py:2257-2259; NumPy's chebval recurrence). q[j] = T_j(x). */
static void cheb_basis(double x, int n, double *q) {
    for (int j = 0; j < n; ++j) {
        double c0, c1;
        if (n == 1) {
            c0 = 1.0;
            c1 = 0.0;
        } else if (n == 2) {
            c0 = j == 0 ? 1.0 : 0.0;
            c1 = j == 1 ? 1.0 : 0.0;
        } else {
            const double x2 = 2.0 * x;
            c0 = (n - 2 == j) ? 1.0 : 0.0;
            c1 = (n - 1 == j) ? 1.0 : 0.0;
            for (int i = 3; i <= n; ++i) {
                const double tmp = c0;
                c0 = ((n - i == j) ? 1.0 : 0.0) - c1;
                c1 = tmp + c1 * x2;
            }
        }
        q[j] = c0 + c1 * x;
    }
}

/* One TT value, storage-frame point p (tensor_train.py:2199-2214 / 2252-2263):
 * result <- result @ (sum_j q_j core[:, j, :]) for each dimension. */
static double tt_value(int D, const int32_t *n, const int32_t *r, const double *lo, const double *hi,
                       const double *cores, const double *p) {
    double res[ORC_MAXR], nxt[ORC_MAXR], q[ORC_MAXN], v[ORC_MAXR];
    res[0] = 1.0;
    const double *core = cores;
    for (int d = 0; d < D; ++d) {
        const int r0 = r[d], r1 = r[d + 1], nd = n[d];
        const double scaled = 2.0 * (p[d] - lo[d]) / (hi[d] - lo[d]) - 1.0;
        cheb_basis(scaled, nd, q);
        for (int k = 0; k < r1; ++k) nxt[k] = 0.0;
        for (int i = 0; i < r0; ++i) {
            for (int k = 0; k < r1; ++k) {
                double acc = 0.0;
                for (int j = 0; j < nd; ++j) acc += q[j] * core[((size_t)i * nd + j) * r1 + k];
                v[k] = acc;
            }
            for (int k = 0; k < r1; ++k) nxt[k] += res[i] * v[k];
        }
        memcpy(res, nxt, sizeof(double) * r1);
        core += (size_t)r0 * nd * r1;
    }
    return res[0];
}

/* ChebyshevTT.eval_batch (tensor_train.py:2217-2265). pts (N, D) user frame. */
ORC_API void orc_tt_eval_batch(int D, const int32_t *n, const int32_t *r, const double *lo,
                               const double *hi, const int32_t *dim_order, const double *cores,
                               const double *pts, int64_t N, double *out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        double p[64];
        for (int k = 0; k < D; ++k) p[k] = pts[i * D + dim_order[k]];
        out[i] = tt_value(D, n, r, lo, hi, cores, p);
    }
}

/* tensor_train.py:2361-2370 */
static void nudge(double *p, int d, double a, double b, double h) {
    const double need = h * 1.5;
    if (p[d] - a < need) p[d] = a + need;
    if (b - p[d] < need) p[d] = b - need;
}

/* tensor_train.py:2428-2463 (and 2372-2403 for a single active dim) */
static int fd_nested(int D, const int32_t *n, const int32_t *r, const double *lo, const double *hi,
                     const double *cores, const double *p, const int *adim, const int *aord, int m,
                     double *result) {
    if (m == 0) {
        *result = tt_value(D, n, r, lo, hi, cores, p);
        return 0;
    }
    const int d = adim[0], order = aord[0];
    const double h = (hi[d] - lo[d]) * 1e-4;
    double pt[64], up[64], dn[64];
    memcpy(pt, p, sizeof(double) * D);
    nudge(pt, d, lo[d], hi[d], h);
    memcpy(up, pt, sizeof(double) * D);
    memcpy(dn, pt, sizeof(double) * D);
    up[d] += h;
    dn[d] -= h;
    double fp, fm, fc;
    if (order == 1) {
        if (fd_nested(D, n, r, lo, hi, cores, up, adim + 1, aord + 1, m - 1, &fp)) return -1;
        if (fd_nested(D, n, r, lo, hi, cores, dn, adim + 1, aord + 1, m - 1, &fm)) return -1;
        *result = (fp - fm) / (2.0 * h);
        return 0;
    }
    if (order == 2) {
        if (fd_nested(D, n, r, lo, hi, cores, up, adim + 1, aord + 1, m - 1, &fp)) return -1;
        if (fd_nested(D, n, r, lo, hi, cores, pt, adim + 1, aord + 1, m - 1, &fc)) return -1;
        if (fd_nested(D, n, r, lo, hi, cores, dn, adim + 1, aord + 1, m - 1, &fm)) return -1;
        *result = (fp - 2.0 * fc + fm) / (h * h);
        return 0;
    }
    return -1; /* "Derivative order k not supported (use 1 or 2)" */
}

/* Loop of ChebyshevTT.eval_multi (tensor_train.py:2267-2463). orders (G, D) user frame.
 * Returns 0, or -1 if an order > 2 was requested. */
ORC_API int orc_tt_eval_multi_batch(int D, const int32_t *n, const int32_t *r, const double *lo,
                                    const double *hi, const int32_t *dim_order, const double *cores,
                                    const double *pts, int64_t N, int G, const int32_t *orders,
                                    double *out) {
    int bad = 0;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        double p[64];
        for (int k = 0; k < D; ++k) p[k] = pts[i * D + dim_order[k]];
        for (int g = 0; g < G; ++g) {
            int adim[64], aord[64], m = 0;
            for (int k = 0; k < D; ++k) {
                const int o = orders[(size_t)g * D + dim_order[k]];
                if (o > 0) {
                    adim[m] = k;
                    aord[m] = o;
                    ++m;
                }
            }
            double res = 0.0;
            if (m == 2 && aord[0] == 1 && aord[1] == 1) {
                /* tensor_train.py:2405-2426 */
                const int d1 = adim[0], d2 = adim[1];
                const double h1 = (hi[d1] - lo[d1]) * 1e-4, h2 = (hi[d2] - lo[d2]) * 1e-4;
                double pt[64], w[64];
                memcpy(pt, p, sizeof(double) * D);
                nudge(pt, d1, lo[d1], hi[d1], h1);
                nudge(pt, d2, lo[d2], hi[d2], h2);
                double f[4];
                for (int e = 0; e < 4; ++e) {
                    memcpy(w, pt, sizeof(double) * D);
                    w[d1] += (e & 2) ? -h1 : h1;
                    w[d2] += (e & 1) ? -h2 : h2;
                    f[e] = tt_value(D, n, r, lo, hi, cores, w);
                }
                res = (f[0] - f[1] - f[2] + f[3]) / (4.0 * h1 * h2);
            } else if (fd_nested(D, n, r, lo, hi, cores, p, adim, aord, m, &res)) {
#pragma omp atomic write
                bad = 1;
            }
            out[i * G + g] = res;
        }
    }
    return bad ? -1 : 0;
}

/* ChebyshevApproximation.vectorized_eval_batch per point (barycentric.py:1035-1046):
 * contraction from the last axis to the first; exact-node slice when |x - node| < 1e-14 (first
 * hit); otherwise (current @ (w/diff)) / sum(w/diff).  `tensor` is already differentiated.
 * scratch: 2 * (size / n[D-1]) doubles. */
static double full_value(int D, const int32_t *n, const double *nodes, const double *weights,
                         const double *tensor, int64_t size, const double *x, double *scratch) {
    const double *cur = tensor;
    int64_t len = size;
    int off = 0;
    for (int d = 0; d < D; ++d) off += n[d];
    double *bufs[2] = {scratch, scratch + size / n[D - 1] + 1};
    int which = 0;
    for (int d = D - 1; d >= 0; --d) {
        const int nd = n[d];
        off -= nd;
        const double *nd_nodes = nodes + off, *nd_w = weights + off;
        const int64_t rows = len / nd;
        double *dst = bufs[which];
        int hit = -1;
        double w[ORC_MAXN];
        for (int i = 0; i < nd; ++i)
            if (hit < 0 && fabs(x[d] - nd_nodes[i]) < 1e-14) hit = i;
        if (hit >= 0) {
            for (int64_t rr = 0; rr < rows; ++rr) dst[rr] = cur[rr * nd + hit];
        } else {
            double sum = 0.0;
            for (int i = 0; i < nd; ++i) {
                w[i] = nd_w[i] / (x[d] - nd_nodes[i]);
                sum += w[i];
            }
            for (int64_t rr = 0; rr < rows; ++rr) {
                double acc = 0.0;
                const double *row = cur + rr * nd;
                for (int i = 0; i < nd; ++i) acc += row[i] * w[i];
                dst[rr] = acc / sum;
            }
        }
        cur = dst;
        len = rows;
        which ^= 1;
    }
    return cur[0];
}

/* G pre-differentiated tensors (concatenated, each `size` doubles) -> out (N, G) */
ORC_API void orc_full_eval_batch(int D, const int32_t *n, const double *nodes, const double *weights,
                                 int G, const double *tensors, const double *pts, int64_t N,
                                 double *out) {
    int64_t size = 1;
    for (int d = 0; d < D; ++d) size *= n[d];
#pragma omp parallel
    {
        double *scratch = (double *)malloc(sizeof(double) * (2 * (size / n[D - 1] + 1)));
#pragma omp for schedule(static)
        for (int64_t i = 0; i < N; ++i)
            for (int g = 0; g < G; ++g)
                out[i * G + g] = full_value(D, n, nodes, weights, tensors + (size_t)g * size, size,
                                            pts + i * D, scratch);
        free(scratch);
    }
}

/* spline.py:677-690: idx_d = searchsorted(knots_d, x, side="right") (NaN sorts last), clipped to
 * pieces_d - 1, C-order ravel. */
static int piece_index(int D, const int32_t *num_knots, const double *knots, const double *x) {
    int flat = 0;
    const double *kn = knots;
    for (int d = 0; d < D; ++d) {
        const int nk = num_knots[d];
        int idx = 0;
        if (nk > 0) {
            if (x[d] != x[d]) {
                idx = nk;
            } else {
                /* binary search for the first knot > x (side="right") */
                int lo = 0, hi = nk;
                while (lo < hi) {
                    const int mid = (lo + hi) / 2;
                    if (kn[mid] <= x[d]) lo = mid + 1; else hi = mid;
                }
                idx = lo;
            }
            if (idx > nk) idx = nk;
        }
        flat = flat * (nk + 1) + idx;
        kn += nk;
    }
    return flat;
}

ORC_API void orc_spline_lookup(int D, const int32_t *num_knots, const double *knots,
                               const double *pts, int64_t N, int32_t *piece) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) piece[i] = piece_index(D, num_knots, knots, pts + i * D);
}

/* ChebyshevSpline.eval_batch (spline.py:633-700). Pieces in C-order; piece p has node counts
 * piece_n[p*D..], nodes/weights at node_off[p], G tensors at tensor_off[p]. */
ORC_API void orc_spline_eval_batch(int D, const int32_t *num_knots, const double *knots, int P,
                                   const int32_t *piece_n, const int64_t *node_off,
                                   const int64_t *tensor_off, const double *nodes,
                                   const double *weights, int G, const double *tensors,
                                   const double *pts, int64_t N, double *out, int32_t *piece_out) {
    int64_t maxsize = 1;
    for (int p = 0; p < P; ++p) {
        int64_t s = 1;
        for (int d = 0; d < D; ++d) s *= piece_n[p * D + d];
        if (s > maxsize) maxsize = s;
    }
#pragma omp parallel
    {
        double *scratch = (double *)malloc(sizeof(double) * (2 * maxsize + 2));
#pragma omp for schedule(static)
        for (int64_t i = 0; i < N; ++i) {
            const int p = piece_index(D, num_knots, knots, pts + i * D);
            if (piece_out) piece_out[i] = p;
            int64_t size = 1;
            for (int d = 0; d < D; ++d) size *= piece_n[p * D + d];
            for (int g = 0; g < G; ++g)
                out[i * G + g] = full_value(D, piece_n + p * D, nodes + node_off[p],
                                            weights + node_off[p],
                                            tensors + tensor_off[p] + (size_t)g * size, size,
                                            pts + i * D, scratch);
        }
        free(scratch);
    }
}

/* Loop of ChebyshevSlider.eval (slider.py:247-318). out_slide[g]: -1 value row, -2 zero row,
 * s derivative owned by slide s; tensor_off[g*S+s] offset of the tensor that row uses. */
ORC_API void orc_slider_eval_batch(int D, int S, const int32_t *group_size, const int32_t *group_dims,
                                   const int32_t *slide_n, const int64_t *node_off,
                                   const double *nodes, const double *weights, double pivot, int G,
                                   const int32_t *out_slide, const int64_t *tensor_off,
                                   const double *tensors, const double *pts, int64_t N, double *out) {
    int64_t maxsize = 1;
    int goff[1025];
    goff[0] = 0;
    for (int s = 0; s < S; ++s) {
        goff[s + 1] = goff[s] + group_size[s];
        int64_t sz = 1;
        for (int d = 0; d < group_size[s]; ++d) sz *= slide_n[goff[s] + d];
        if (sz > maxsize) maxsize = sz;
    }
#pragma omp parallel
    {
        double *scratch = (double *)malloc(sizeof(double) * (2 * maxsize + 2));
#pragma omp for schedule(static)
        for (int64_t i = 0; i < N; ++i) {
            for (int g = 0; g < G; ++g) {
                const int os = out_slide[g];
                double res;
                if (os == -2) {
                    res = 0.0;
                } else {
                    res = os == -1 ? pivot : 0.0;
                    const int s_lo = os == -1 ? 0 : os, s_hi = os == -1 ? S : os + 1;
                    for (int s = s_lo; s < s_hi; ++s) {
                        double x[64];
                        int64_t sz = 1;
                        for (int d = 0; d < group_size[s]; ++d) {
                            x[d] = pts[i * D + group_dims[goff[s] + d]];
                            sz *= slide_n[goff[s] + d];
                        }
                        const double v = full_value(group_size[s], slide_n + goff[s],
                                                    nodes + node_off[s], weights + node_off[s],
                                                    tensors + tensor_off[(size_t)g * S + s], sz, x,
                                                    scratch);
                        res = os == -1 ? res + (v - pivot) : v;
                    }
                }
                out[i * G + g] = res;
            }
        }
        free(scratch);
    }
}
