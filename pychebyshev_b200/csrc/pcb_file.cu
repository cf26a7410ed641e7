// Native `.pcb` v1 loader: file -> device plan without Python.
//
// The native equivalent of the reference's stand-alone readers (examples/binary_reader/reader.c,
// readers/rust, readers/julia): parses the layout of reference _binary.py:157-421, rebuilds the
// grid the way the reference does on load (nodes: _extrude_slice.py:66-70 with numpy's chebpts1
// = sin(pi/(2n) * (-n+1, -n+3, ..)); weights: barycentric.py:43-49) and creates a VALUE plan
// (G = 1, no derivatives: derivative tensors need the reference's BLAS recipe bit for bit, see
// DESIGN.md §2, so Greeks go through the Python host).  Error texts follow _binary.py.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <new>
#include <stdexcept>

#include "pcb_common.cuh"

namespace pcb {

struct PcbFile {
    int kind = 0;  // 1 approximation, 2 spline
    int D = 0;
    std::vector<double> lo, hi;
    std::vector<int32_t> n;
    std::vector<int32_t> num_knots;
    std::vector<double> knots;
    int P = 1;
    std::vector<double> values;  // P tensors of prod(n) doubles, C-order
};

static bool take(const std::vector<unsigned char> &raw, size_t &pos, void *dst, size_t bytes) {
    if (pos + bytes > raw.size()) return false;
    memcpy(dst, raw.data() + pos, bytes);
    pos += bytes;
    return true;
}

static int parse_pcb(const char *path, PcbFile *out) {
    FILE *f = fopen(path, "rb");
    if (!f) return fail(PCB_EINVAL, "cannot open %s", path);
    std::vector<unsigned char> raw;
    unsigned char buf[1 << 16];
    size_t got;
    while ((got = fread(buf, 1, sizeof(buf), f)) > 0) raw.insert(raw.end(), buf, buf + got);
    fclose(f);
    if (raw.size() < 12)
        return fail(PCB_EINVAL, "unexpected EOF reading header (wanted 12 bytes, got %zu)", raw.size());
    if (memcmp(raw.data(), "PCB\0", 4) != 0) return fail(PCB_EINVAL, "not a PyChebyshev binary file (bad magic)");
    if (raw[4] != 1)
        return fail(PCB_EINVAL, "unsupported .pcb major version %d (this build reads major 1)", raw[4]);
    uint16_t tag;
    memcpy(&tag, raw.data() + 6, 2);
    if (raw[8] | raw[9] | raw[10] | raw[11])
        return fail(PCB_EINVAL, "reserved header bytes nonzero — file may be corrupt");
    if (tag != 1 && tag != 2) return fail(PCB_EINVAL, "unknown class_tag %d", (int)tag);
    out->kind = tag;
    size_t pos = 12;
    uint32_t D;
    if (!take(raw, pos, &D, 4)) return fail(PCB_EINVAL, "unexpected EOF reading uint32");
    if (D < 1 || D > 64) return fail(PCB_EINVAL, "num_dimensions must be >= 1, got %u", D);
    out->D = (int)D;
    out->lo.resize(D);
    out->hi.resize(D);
    out->n.resize(D);
    if (!take(raw, pos, out->lo.data(), 8 * D) || !take(raw, pos, out->hi.data(), 8 * D))
        return fail(PCB_EINVAL, "unexpected EOF reading f64 array");
    for (uint32_t d = 0; d < D; ++d)
        if (out->lo[d] >= out->hi[d])  // exactly the reference's test (_binary.py:262-266): NaN bounds pass
            return fail(PCB_EINVAL, "domain[%u]: lo (%g) must be < hi (%g)", d, out->lo[d], out->hi[d]);
    std::vector<uint32_t> nn(D);
    if (!take(raw, pos, nn.data(), 4 * D)) return fail(PCB_EINVAL, "unexpected EOF reading uint32 array");
    // every size below is bounded by the bytes that are actually left in the file before anything
    // is allocated: a crafted header cannot request more memory than the file is long
    const size_t max_doubles = (raw.size() - pos) / 8;
    size_t per = 1;
    for (uint32_t d = 0; d < D; ++d) {
        if (nn[d] < 1) return fail(PCB_EINVAL, "n_nodes[%u] must be >= 1, got %u", d, nn[d]);
        if (nn[d] > 0x7fffffffu || per > max_doubles / nn[d])
            return fail(PCB_EINVAL, "unexpected EOF reading f64 array (tensor of n_nodes[%u]=%u "
                        "exceeds the %zu bytes left)", d, nn[d], raw.size() - pos);
        out->n[d] = (int32_t)nn[d];
        per *= nn[d];
    }
    out->P = 1;
    out->num_knots.assign(D, 0);
    if (tag == 2) {
        std::vector<uint32_t> nk(D);
        if (!take(raw, pos, nk.data(), 4 * D)) return fail(PCB_EINVAL, "unexpected EOF reading uint32 array");
        const size_t left = (raw.size() - pos) / 8;
        size_t total = 0, expect = 1;
        for (uint32_t d = 0; d < D; ++d) {
            if (nk[d] > left || total + nk[d] > left)
                return fail(PCB_EINVAL, "unexpected EOF reading f64 array (num_knots[%u]=%u)", d, nk[d]);
            out->num_knots[d] = (int32_t)nk[d];
            total += nk[d];
            if (expect > left / ((size_t)nk[d] + 1) + 1)  // more pieces than doubles in the file
                return fail(PCB_EINVAL, "unexpected EOF reading f64 array (too many pieces)");
            expect *= (size_t)nk[d] + 1;
        }
        out->knots.resize(total);
        if (total && !take(raw, pos, out->knots.data(), 8 * total))
            return fail(PCB_EINVAL, "unexpected EOF reading f64 array");
        size_t off = 0;
        for (uint32_t d = 0; d < D; ++d) {
            for (uint32_t k = 0; k + 1 < nk[d]; ++k)
                if (!(out->knots[off + k] < out->knots[off + k + 1]))
                    return fail(PCB_EINVAL, "knots in dim %u not strictly ascending", d);
            off += nk[d];
        }
        uint32_t P;
        if (!take(raw, pos, &P, 4)) return fail(PCB_EINVAL, "unexpected EOF reading uint32");
        if ((size_t)P != expect)
            return fail(PCB_EINVAL, "num_pieces=%u does not match prod(num_knots+1)=%zu", P, expect);
        if (P > 0x7fffffffu) return fail(PCB_EINVAL, "num_pieces=%u too large", P);
        out->P = (int)P;
    }
    const size_t avail = (raw.size() - pos) / 8;
    if (per > avail / (size_t)out->P)
        return fail(PCB_EINVAL, "unexpected EOF reading f64 array (wanted %zu x %d doubles, %zu bytes left)",
                    per, out->P, raw.size() - pos);
    out->values.resize(per * (size_t)out->P);
    if (!take(raw, pos, out->values.data(), 8 * per * (size_t)out->P))
        return fail(PCB_EINVAL, "unexpected EOF reading f64 array");
    // what the reference's reader goes on to check by constructing through from_values
    // (spline.py:1272-1291, barycentric.py:1879-1880): knots strictly inside the domain, finite values
    if (out->kind == 2) {
        size_t off = 0;
        for (uint32_t d = 0; d < D; ++d) {
            for (uint32_t k = 0; k < (uint32_t)out->num_knots[d]; ++k) {
                const double v = out->knots[off + k];
                if (!(out->lo[d] < v && v < out->hi[d]))
                    return fail(PCB_EINVAL, "Knot %g for dimension %u is not strictly inside domain [%g, %g]",
                                v, d, out->lo[d], out->hi[d]);
            }
            off += (size_t)out->num_knots[d];
        }
    }
    for (double v : out->values)
        if (!std::isfinite(v)) return fail(PCB_EINVAL, "tensor_values contains NaN or Inf");
    return PCB_OK;
}

// numpy chebpts1 mapped to [lo, hi], ascending (_extrude_slice.py:66-70)
static void make_nodes(double lo, double hi, int n, double *out) {
    const double pi = 3.14159265358979323846;
    for (int k = 0; k < n; ++k) out[k] = 0.5 * (lo + hi) + 0.5 * (hi - lo) * std::sin(0.5 * pi / n * (double)(-n + 1 + 2 * k));
    std::sort(out, out + n);
}

// barycentric.py:43-49: sequential division
static void make_weights(const double *x, int n, double *w) {
    for (int i = 0; i < n; ++i) {
        w[i] = 1.0;
        for (int j = 0; j < n; ++j)
            if (j != i) w[i] /= x[i] - x[j];
    }
}

// numpy's pairwise summation of a contiguous row (np.sum over the last axis): < 8 sequential,
// otherwise 8 interleaved partial sums combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) plus a
// sequential tail; rows longer than 128 split recursively.
static double np_pairwise_sum(const double *a, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res += a[i];
        return res;
    }
    if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        int i = 8;
        for (; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
}

// barycentric.py:52-77: D_ij = w_j / ((x_i - x_j) * w_i), diagonal = -(row sum)
static void make_diff_matrix(const double *x, const double *w, int n, double *dm) {
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < n; ++j) {
            const double c = i == j ? 1.0 : x[i] - x[j];
            dm[(size_t)i * n + j] = i == j ? 0.0 : w[j] / (c * w[i]);
        }
        dm[(size_t)i * n + i] = -np_pairwise_sum(dm + (size_t)i * n, n);
    }
}

// One pass of _apply_derivative_passes on the host (small spline pieces): the same sequential FMA
// chain as the device kernel (pcb_tensor.cu) and OpenBLAS.
static void host_deriv_pass(const double *src, double *dst, long long outer, int n, long long inner,
                            const double *dm) {
    for (long long o = 0; o < outer; ++o)
        for (long long i = 0; i < inner; ++i)
            for (int j = 0; j < n; ++j) {
                double acc = 0.0;
                for (int k = 0; k < n; ++k)
                    acc = std::fma(src[(o * n + k) * inner + i], dm[(size_t)j * n + k], acc);
                dst[(o * n + j) * inner + i] = acc;
            }
}

static void put(std::vector<unsigned char> &out, const void *p, size_t bytes) {
    const unsigned char *b = static_cast<const unsigned char *>(p);
    out.insert(out.end(), b, b + bytes);
}

// _binary.py:157-185 (header) + :208-236 / :289-346 (bodies); little-endian host assumed (x86-64)
static std::vector<unsigned char> serialize_pcb(const PcbFile &pf) {
    std::vector<unsigned char> out;
    const unsigned char head[12] = {'P', 'C', 'B', 0, 1, 0, (unsigned char)pf.kind, 0, 0, 0, 0, 0};
    put(out, head, 12);
    const uint32_t D = (uint32_t)pf.D;
    put(out, &D, 4);
    put(out, pf.lo.data(), 8 * (size_t)pf.D);
    put(out, pf.hi.data(), 8 * (size_t)pf.D);
    for (int d = 0; d < pf.D; ++d) {
        const uint32_t v = (uint32_t)pf.n[d];
        put(out, &v, 4);
    }
    if (pf.kind == 2) {
        for (int d = 0; d < pf.D; ++d) {
            const uint32_t v = (uint32_t)pf.num_knots[d];
            put(out, &v, 4);
        }
        put(out, pf.knots.data(), 8 * pf.knots.size());
        const uint32_t P = (uint32_t)pf.P;
        put(out, &P, 4);
    }
    put(out, pf.values.data(), 8 * pf.values.size());
    return out;
}

static int write_file(const char *path, const std::vector<unsigned char> &raw) {
    FILE *f = fopen(path, "wb");
    if (!f) return fail(PCB_EINVAL, "cannot open %s for writing", path);
    const size_t done = fwrite(raw.data(), 1, raw.size(), f);
    if (fclose(f) != 0 || done != raw.size()) return fail(PCB_EINVAL, "short write to %s", path);
    return PCB_OK;
}

static int check_grid(int D, const double *lo, const double *hi, const int32_t *n) {
    PCB_REQUIRE(D >= 1 && D <= 64, "num_dimensions must be >= 1, got %d", D);
    for (int d = 0; d < D; ++d) {
        PCB_REQUIRE(!(lo[d] >= hi[d]), "domain[%d]: lo (%g) must be < hi (%g)", d, lo[d], hi[d]);
        PCB_REQUIRE(n[d] >= 1, "n_nodes[%d] must be >= 1, got %d", d, n[d]);
    }
    return PCB_OK;
}

}  // namespace pcb

using namespace pcb;

// ---- native writer (reference _binary.py:208-236, 289-346): byte-identical files -----------------------
extern "C" PCB_API int pcb_file_write_approx(const char *path, int D, const double *lo, const double *hi,
                                             const int32_t *n, const double *tensor) {
    PCB_REQUIRE(path && lo && hi && n && tensor, "null argument");
    if (int rc = check_grid(D, lo, hi, n)) return rc;
    try {
        PcbFile pf;
        pf.kind = 1;
        pf.D = D;
        pf.lo.assign(lo, lo + D);
        pf.hi.assign(hi, hi + D);
        pf.n.assign(n, n + D);
        size_t per = 1;
        for (int d = 0; d < D; ++d) per *= (size_t)n[d];
        for (size_t i = 0; i < per; ++i)
            if (!std::isfinite(tensor[i])) return fail(PCB_EINVAL, "tensor_values contains NaN or Inf");
        pf.values.assign(tensor, tensor + per);
        return write_file(path, serialize_pcb(pf));
    } catch (const std::bad_alloc &) {
        return fail(PCB_ENOMEM, "out of host memory while writing %s", path);
    }
}

extern "C" PCB_API int pcb_file_write_spline(const char *path, int D, const double *lo, const double *hi,
                                             const int32_t *n, const int32_t *num_knots,
                                             const double *knots_cat, int P,
                                             const double *const *piece_tensors) {
    PCB_REQUIRE(path && lo && hi && n && num_knots && piece_tensors, "null argument");
    if (int rc = check_grid(D, lo, hi, n)) return rc;
    try {
        PcbFile pf;
        pf.kind = 2;
        pf.D = D;
        pf.lo.assign(lo, lo + D);
        pf.hi.assign(hi, hi + D);
        pf.n.assign(n, n + D);
        pf.num_knots.assign(num_knots, num_knots + D);
        size_t total = 0;
        long long expect = 1;
        for (int d = 0; d < D; ++d) {
            PCB_REQUIRE(num_knots[d] >= 0, "num_knots[%d] must be >= 0", d);
            total += (size_t)num_knots[d];
            expect *= (long long)num_knots[d] + 1;
        }
        PCB_REQUIRE(total == 0 || knots_cat, "null knots");
        PCB_REQUIRE((long long)P == expect, "num_pieces=%d does not match prod(num_knots+1)=%lld", P, expect);
        if (total) pf.knots.assign(knots_cat, knots_cat + total);
        pf.P = P;
        size_t per = 1;
        for (int d = 0; d < D; ++d) per *= (size_t)n[d];
        pf.values.reserve(per * (size_t)P);
        for (int p = 0; p < P; ++p) {
            PCB_REQUIRE(piece_tensors[p], "Cannot save an unbuilt ChebyshevSpline");
            pf.values.insert(pf.values.end(), piece_tensors[p], piece_tensors[p] + per);
        }
        return write_file(path, serialize_pcb(pf));
    } catch (const std::bad_alloc &) {
        return fail(PCB_ENOMEM, "out of host memory while writing %s", path);
    }
}

// bytes -> parsed file -> bytes: the native reader and writer composed (round-trip identity)
extern "C" PCB_API int pcb_file_rewrite(const char *in_path, const char *out_path) {
    PCB_REQUIRE(in_path && out_path, "null argument");
    try {
        PcbFile pf;
        if (int rc = parse_pcb(in_path, &pf)) return rc;
        return write_file(out_path, serialize_pcb(pf));
    } catch (const std::bad_alloc &) {
        return fail(PCB_ENOMEM, "out of host memory while rewriting %s", in_path);
    }
}

// The grid arrays the native loader derives from (lo, hi, n): nodes, barycentric weights and the
// differentiation matrix (n x n), exposed so tests can compare them with the NumPy recipes.
extern "C" PCB_API int pcb_file_grid_arrays(double lo, double hi, int n, double *nodes, double *weights,
                                            double *dmat) {
    PCB_REQUIRE(nodes && weights && n >= 1 && lo < hi, "bad grid description");
    make_nodes(lo, hi, n, nodes);
    make_weights(nodes, n, weights);
    if (dmat) make_diff_matrix(nodes, weights, n, dmat);
    return PCB_OK;
}

// kind: 1 = ChebyshevApproximation, 2 = ChebyshevSpline
static int plan_from_file(int dev, const char *path, int G, const int32_t *orders, void **plan,
                          int *kind, int *D);

extern "C" PCB_API int pcb_plan_from_file(int dev, const char *path, void **plan, int *kind, int *D) {
    return pcb_plan_from_file_orders(dev, path, 0, nullptr, plan, kind, D);
}

// G derivative-order rows (G x D, HOST) -> plan with G outputs per query; G = 0: values only.
extern "C" PCB_API int pcb_plan_from_file_orders(int dev, const char *path, int G, const int32_t *orders,
                                                 void **plan, int *kind, int *D) {
    try {  // no C++ exception may cross the C ABI
        return plan_from_file(dev, path, G, orders, plan, kind, D);
    } catch (const std::bad_alloc &) {
        return fail(PCB_ENOMEM, "out of host memory while loading %s", path ? path : "(null)");
    } catch (const std::exception &e) {
        return fail(PCB_EINVAL, "loading %s failed: %s", path ? path : "(null)", e.what());
    }
}

static int plan_from_file(int dev, const char *path, int G, const int32_t *orders, void **plan,
                          int *kind, int *D) {
    PCB_REQUIRE(path && plan, "null argument");
    PCB_REQUIRE(G >= 0 && G <= 64 && (G == 0 || orders), "bad derivative rows");
    PcbFile pf;
    if (int rc = parse_pcb(path, &pf)) return rc;
    std::vector<int32_t> value_row((size_t)pf.D, 0);
    if (G == 0) {
        G = 1;
        orders = value_row.data();
    }
    for (int i = 0; i < G * pf.D; ++i)
        PCB_REQUIRE(orders[i] >= 0 && orders[i] <= 8, "derivative order %d outside [0, 8]", orders[i]);
    if (kind) *kind = pf.kind;
    if (D) *D = pf.D;
    size_t per = 1;
    int sum_n = 0;
    for (int d = 0; d < pf.D; ++d) {
        per *= (size_t)pf.n[d];
        sum_n += pf.n[d];
    }
    if (pf.kind == 1) {
        std::vector<double> nodes(sum_n), weights(sum_n);
        int off = 0;
        for (int d = 0; d < pf.D; ++d) {
            make_nodes(pf.lo[d], pf.hi[d], pf.n[d], nodes.data() + off);
            make_weights(nodes.data() + off, pf.n[d], weights.data() + off);
            off += pf.n[d];
        }
        // differentiation matrices natively, derivative tensors on the device (N3)
        std::vector<double> dms;
        off = 0;
        for (int d = 0; d < pf.D; ++d) {
            const size_t at = dms.size();
            dms.resize(at + (size_t)pf.n[d] * pf.n[d]);
            make_diff_matrix(nodes.data() + off, weights.data() + off, pf.n[d], dms.data() + at);
            off += pf.n[d];
        }
        bool small_n = true;
        for (int d = 0; d < pf.D; ++d) small_n = small_n && pf.n[d] <= 64;
        PCB_REQUIRE(small_n, "n_nodes above 64 are not supported by the native loader");
        return pcb_full_plan_create_from_values(dev, pf.D, pf.n.data(), nodes.data(), weights.data(),
                                                dms.data(), pf.values.data(), G, orders, plan);
    }
    // spline: pieces in C-order over the per-dimension interval indices, flat n_nodes
    std::vector<int32_t> piece_n((size_t)pf.P * pf.D);
    std::vector<double> nodes((size_t)pf.P * sum_n), weights((size_t)pf.P * sum_n);
    std::vector<const double *> tensors((size_t)pf.P * G);
    std::vector<std::vector<double>> derived;  // derivative tensors of the pieces (host, small)
    derived.reserve((size_t)pf.P * G);
    std::vector<int> idx(pf.D, 0);
    std::vector<size_t> koff(pf.D, 0);
    for (int d = 1; d < pf.D; ++d) koff[d] = koff[d - 1] + pf.num_knots[d - 1];
    for (int p = 0; p < pf.P; ++p) {
        int off = 0;
        for (int d = 0; d < pf.D; ++d) {
            piece_n[(size_t)p * pf.D + d] = pf.n[d];
            const double a = idx[d] == 0 ? pf.lo[d] : pf.knots[koff[d] + idx[d] - 1];
            const double b = idx[d] == pf.num_knots[d] ? pf.hi[d] : pf.knots[koff[d] + idx[d]];
            double *nd = nodes.data() + (size_t)p * sum_n + off;
            make_nodes(a, b, pf.n[d], nd);
            make_weights(nd, pf.n[d], weights.data() + (size_t)p * sum_n + off);
            off += pf.n[d];
        }
        const double *value_t = pf.values.data() + (size_t)p * per;
        for (int g = 0; g < G; ++g) {
            const int32_t *o = orders + (size_t)g * pf.D;
            const double *cur = value_t;
            int poff = sum_n;
            for (int d = pf.D - 1; d >= 0; --d) {  // _apply_derivative_passes: d = D-1 .. 0
                poff -= pf.n[d];
                if (o[d] == 0) continue;
                std::vector<double> dm((size_t)pf.n[d] * pf.n[d]);
                make_diff_matrix(nodes.data() + (size_t)p * sum_n + poff,
                                 weights.data() + (size_t)p * sum_n + poff, pf.n[d], dm.data());
                long long outer = 1, inner = 1;
                for (int e = 0; e < d; ++e) outer *= pf.n[e];
                for (int e = d + 1; e < pf.D; ++e) inner *= pf.n[e];
                for (int rep = 0; rep < o[d]; ++rep) {
                    derived.emplace_back(per);
                    host_deriv_pass(cur, derived.back().data(), outer, pf.n[d], inner, dm.data());
                    cur = derived.back().data();
                }
            }
            tensors[(size_t)p * G + g] = cur;
        }
        for (int d = pf.D - 1; d >= 0; --d) {  // C-order increment
            if (++idx[d] <= pf.num_knots[d]) break;
            idx[d] = 0;
        }
    }
    return pcb_spline_plan_create(dev, pf.D, pf.num_knots.data(), pf.knots.empty() ? nullptr : pf.knots.data(),
                                  pf.P, piece_n.data(), nodes.data(), weights.data(), G, tensors.data(),
                                  plan);
}

// Values of any plan kind at N points (the plan's own number of outputs per point).
extern "C" PCB_API int pcb_plan_eval(void *plan, const double *d_points, int64_t N, double *d_out,
                                     void *stream) {
    PCB_REQUIRE(plan, "null plan");
    switch (static_cast<PlanBase *>(plan)->kind) {
        case PLAN_TT: return pcb_tt_eval(plan, d_points, N, d_out, stream);
        case PLAN_FULL: return pcb_full_eval(plan, d_points, N, d_out, 0, stream);
        case PLAN_SPLINE: return pcb_spline_eval(plan, d_points, N, d_out, nullptr, stream);
        case PLAN_SLIDER: return pcb_slider_eval(plan, d_points, N, d_out, stream);
    }
    return fail(PCB_EINVAL, "not a plan");
}
