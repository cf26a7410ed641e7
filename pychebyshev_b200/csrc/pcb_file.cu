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
        if (!(out->lo[d] < out->hi[d]))
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

}  // namespace pcb

using namespace pcb;

// kind: 1 = ChebyshevApproximation, 2 = ChebyshevSpline
static int plan_from_file(int dev, const char *path, void **plan, int *kind, int *D);

extern "C" PCB_API int pcb_plan_from_file(int dev, const char *path, void **plan, int *kind, int *D) {
    try {  // no C++ exception may cross the C ABI
        return plan_from_file(dev, path, plan, kind, D);
    } catch (const std::bad_alloc &) {
        return fail(PCB_ENOMEM, "out of host memory while loading %s", path ? path : "(null)");
    } catch (const std::exception &e) {
        return fail(PCB_EINVAL, "loading %s failed: %s", path ? path : "(null)", e.what());
    }
}

static int plan_from_file(int dev, const char *path, void **plan, int *kind, int *D) {
    PCB_REQUIRE(path && plan, "null argument");
    PcbFile pf;
    if (int rc = parse_pcb(path, &pf)) return rc;
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
        const double *tensor = pf.values.data();
        return pcb_full_plan_create(dev, pf.D, pf.n.data(), nodes.data(), weights.data(), 1, &tensor, plan);
    }
    // spline: pieces in C-order over the per-dimension interval indices, flat n_nodes
    std::vector<int32_t> piece_n((size_t)pf.P * pf.D);
    std::vector<double> nodes((size_t)pf.P * sum_n), weights((size_t)pf.P * sum_n);
    std::vector<const double *> tensors(pf.P);
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
        tensors[p] = pf.values.data() + (size_t)p * per;
        for (int d = pf.D - 1; d >= 0; --d) {  // C-order increment
            if (++idx[d] <= pf.num_knots[d]) break;
            idx[d] = 0;
        }
    }
    return pcb_spline_plan_create(dev, pf.D, pf.num_knots.data(), pf.knots.empty() ? nullptr : pf.knots.data(),
                                  pf.P, piece_n.data(), nodes.data(), weights.data(), 1, tensors.data(),
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
