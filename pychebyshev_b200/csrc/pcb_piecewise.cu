// ChebyshevSpline (piece lookup + per-piece evaluation) and ChebyshevSlider (additive slides).
//
// Spline: replaces ChebyshevSpline.eval_batch (reference spline.py:633-700).  The routing
//   (spline.py:677-690) is integer work and bit-exact; the per-piece evaluation is the
//   thread-per-query barycentric evaluator of pcb_grid.cuh on the piece's own nodes/weights.
//   The reference groups points by piece and calls the piece evaluator per group; on the device
//   every query simply indexes its piece's descriptor, so no bucketing pass is needed.
// Slider: replaces a loop of ChebyshevSlider.eval (reference slider.py:247-318).
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "pcb_cbank.cuh"
#include "pcb_grid.cuh"

namespace pcb {

constexpr int PW_THREADS = 128;

// descriptor of a grid whose tensor / nodes / weights live in the constant bank (see below)
struct BankGrid {
    int D;
    int n[GRID_MAXD];
    int size;        // prod(n)
    int tensor_off;  // bank index of output block 0
    int node_off;    // bank index of the nodes (dims concatenated)
    int weight_off;  // bank index of the weights
    int scale_off;   // bank index of D prescale factors, then D scaled hit thresholds
    int dims[GRID_MAXD];  // coordinate of the query each grid dimension reads
    int outputs;          // tensors stored for this grid (slider: 0 = nothing to do)
};

struct SplinePlan : PlanBase {
    int D = 0, P = 0, G = 0, GB = 1;
    int max_sum_n = 0;
    int *d_num_knots = nullptr;  // [D] then knot_off [D]
    int *d_knot_off = nullptr;
    double *d_knots = nullptr;
    GridDesc *d_desc = nullptr;  // [P]
    double *d_nodes = nullptr;   // all pieces: nodes then weights (same offsets)
    double *d_weights = nullptr;
    double *d_tensors = nullptr;  // per piece: [block][elem][GB]
    // uniform-datapath path: all piece tensors + descriptors in this module's constant bank.  When
    // the G outputs do not fit together they are split into parts of consecutive outputs, one
    // launch (and one bank image) per part.
    struct BankPart {
        uint64_t id = 0;
        int g0 = 0, G = 0, GB = 1;  // outputs [g0, g0 + G)
        std::vector<double> h_bank;
        std::vector<BankGrid> h_desc;
    };
    bool bank_ok = false;
    std::vector<BankPart> parts;
    size_t bank_smem = 0;
    // 2-D splines on the FP64 tensor cores: piece tensors in MMA fragment order [piece][output]
    bool dmma2d_ok = false;
    int nfix = 0;  // node count shared by every dimension of every piece (0: mixed)
    double *d_frags = nullptr;
    // 3-D splines on the tensor cores (joint-K): fragment images per (piece, output)
    bool dmma3d_ok = false;
    int kb3max = 0, g3_per = 1, g3_warps = 8;  // K blocks of the largest piece; outputs per launch; warps per CTA
    ~SplinePlan() override;
    void free_all() {
        if (d_frags) cudaFree(d_frags);
        if (d_num_knots) cudaFree(d_num_knots);
        if (d_knots) cudaFree(d_knots);
        if (d_desc) cudaFree(d_desc);
        if (d_nodes) cudaFree(d_nodes);
        if (d_tensors) cudaFree(d_tensors);
    }
};

struct SliderPlan : PlanBase {
    int D = 0, S = 0, G = 0, GB = 1;
    int max_sum_n = 0, max_D = 1;
    double pivot = 0.0;
    GridDesc *d_desc = nullptr;  // [S]
    int *d_ints = nullptr;       // group_off [S+1] | group dims | out_slide [G] | row_out [G] | slide_G [S]
    int *d_group_dims = nullptr, *d_out_slide = nullptr, *d_row_out = nullptr, *d_slide_G = nullptr;
    double *d_nodes = nullptr;
    double *d_weights = nullptr;
    double *d_tensors = nullptr;  // per slide: its outputs interleaved [block][elem][GB]
    bool bank_ok = false;
    uint64_t plan_id = 0;
    std::vector<double> h_bank;
    std::vector<BankGrid> h_desc;
    size_t bank_smem = 0;
    // sliders of 2-D slides on the FP64 tensor cores: slide tensors in MMA fragment order
    bool dmma2d_ok = false;
    double *d_frags = nullptr;
    int *d_frag_off = nullptr;
    int nfrag = 0;
    int nfix = 0;  // node count shared by both dimensions of every slide (0: mixed)
    ~SliderPlan() override;
    void free_all() {
        if (d_frags) cudaFree(d_frags);
        if (d_frag_off) cudaFree(d_frag_off);
        if (d_desc) cudaFree(d_desc);
        if (d_ints) cudaFree(d_ints);
        if (d_nodes) cudaFree(d_nodes);
        if (d_tensors) cudaFree(d_tensors);
    }
};

// HBM-bound stream (8 D bytes in, 4 out per query).  Four independent queries per thread per
// iteration, coordinates loaded before any of them is used: with one 16-byte load in flight per
// thread the SM holds ~32 KB in flight, just short of bandwidth x latency.
constexpr int LOOKUP_ILP = 2, LOOKUP_MAXD = 4;

__global__ void __launch_bounds__(256, 8)
spline_lookup_kernel(int D, const int *__restrict__ num_knots, const int *__restrict__ knot_off,
                     const double *__restrict__ knots, const double *__restrict__ pts, int64_t N,
                     int32_t *__restrict__ piece, int aligned16) {
    const int64_t step = (int64_t)gridDim.x * blockDim.x * LOOKUP_ILP;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x * LOOKUP_ILP + threadIdx.x; base < N; base += step) {
        if (D > LOOKUP_MAXD) {  // uniform: generic path straight from memory
#pragma unroll
            for (int k = 0; k < LOOKUP_ILP; ++k) {
                const int64_t q = base + (int64_t)k * blockDim.x;
                if (q < N) piece[q] = spline_piece_index(D, num_knots, knot_off, knots, pts + q * D);
            }
            continue;
        }
        double x[LOOKUP_ILP][LOOKUP_MAXD];
#pragma unroll
        for (int k = 0; k < LOOKUP_ILP; ++k) {
            const int64_t q = base + (int64_t)k * blockDim.x;
            if (D == 2 && aligned16) {  // 128-bit loads need a 16-byte aligned base (uniform branch)
                const double2 v = q < N ? __ldg(reinterpret_cast<const double2 *>(pts) + q) : make_double2(0.0, 0.0);
                x[k][0] = v.x;
                x[k][1] = v.y;
            } else {
#pragma unroll
                for (int d = 0; d < LOOKUP_MAXD; ++d) x[k][d] = (d < D && q < N) ? __ldg(pts + q * D + d) : 0.0;
            }
        }
#pragma unroll
        for (int k = 0; k < LOOKUP_ILP; ++k) {
            const int64_t q = base + (int64_t)k * blockDim.x;
            // idx_d = #{knot <= x} (NaN -> all knots), C-order ravel: spline_piece_index on registers
            int flat = 0;
#pragma unroll
            for (int d = 0; d < LOOKUP_MAXD; ++d) {
                if (d < D) {
                    const int nk = num_knots[d];
                    const double xv = x[k][d];
                    int idx = 0;
                    if (xv != xv) {
                        idx = nk;
                    } else {
                        const double *kn = knots + knot_off[d];
                        for (int t = 0; t < nk; ++t) idx += (__ldg(kn + t) <= xv) ? 1 : 0;
                    }
                    flat = flat * (nk + 1) + idx;
                }
            }
            if (q < N) piece[q] = flat;
        }
    }
}

template <int GB, int DM>
__global__ void __launch_bounds__(PW_THREADS)
spline_eval_kernel(int D, int G, const int *__restrict__ num_knots, const int *__restrict__ knot_off,
                   const double *__restrict__ knots, const GridDesc *__restrict__ desc,
                   const double *__restrict__ nodes, const double *__restrict__ weights,
                   const double *__restrict__ tensors, const double *__restrict__ pts, int64_t N,
                   double *__restrict__ out, int32_t *__restrict__ piece_out) {
    extern __shared__ __align__(16) double smem[];
    double *ws = smem + threadIdx.x;
    const int stride = blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < N;
         q += (int64_t)gridDim.x * blockDim.x) {
        const double *x = pts + q * D;
        const int p = spline_piece_index(D, num_knots, knot_off, knots, x);
        if (piece_out) piece_out[q] = p;
        const GridDesc gd = desc[p];
        double wl[GRID_NL];
        const bool regs = grid_weights(gd, nodes, weights, [&](int d) { return __ldg(x + d); }, ws,
                                       stride, wl);
        grid_eval_outputs<GB, DM>(gd, tensors, G, ws, stride, wl, regs, out + q * G, 1);
    }
}

// Row g of the slider reads output `row_out[g]` of slide `out_slide[g]`; value rows
// (out_slide = -1) accumulate output 0 of every slide; cross-slide rows (-2) are exactly 0.
template <int GB, int DM>
__global__ void __launch_bounds__(PW_THREADS)
slider_eval_kernel(int D, int S, int G, double pivot, const GridDesc *__restrict__ desc,
                   const int *__restrict__ group_off, const int *__restrict__ group_dims,
                   const int *__restrict__ out_slide, const int *__restrict__ row_out,
                   const int *__restrict__ slide_G, const double *__restrict__ nodes,
                   const double *__restrict__ weights, const double *__restrict__ tensors,
                   const double *__restrict__ pts, int64_t N, double *__restrict__ out) {
    extern __shared__ __align__(16) double smem[];
    double *ws = smem + threadIdx.x;
    const int stride = blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < N;
         q += (int64_t)gridDim.x * blockDim.x) {
        const double *x = pts + q * D;
        double *o = out + q * G;
        // value rows start from the pivot (slider.py:310), cross-slide rows are exactly 0 (:297)
        for (int g = 0; g < G; ++g) o[g] = out_slide[g] == -1 ? pivot : 0.0;
        for (int s = 0; s < S; ++s) {
            const int sg = slide_G[s];
            if (sg == 0) continue;
            const GridDesc gd = desc[s];
            const int *dims = group_dims + group_off[s];
            double wl[GRID_NL];
            const bool regs = grid_weights(gd, nodes, weights,
                                           [&](int d) { return __ldg(x + dims[d]); }, ws, stride, wl);
            for (int b = 0; b * GB < sg; ++b) {
                double r[GB];
                grid_contract<GB, DM>(gd, tensors + gd.tensor_off + (long long)b * gd.size * GB, ws,
                                  stride, wl, regs, r);
#pragma unroll
                for (int j = 0; j < GB; ++j) {
                    const int so = b * GB + j;
                    if (so >= sg) continue;
                    for (int g = 0; g < G; ++g) {
                        if (out_slide[g] == s && row_out[g] == so)
                            o[g] = r[j];  // derivative row owned by this slide (:301-307)
                        else if (out_slide[g] == -1 && so == 0 && row_out[g] == 0)
                            o[g] = o[g] + (r[j] - pivot);  // left to right over the slides (:310-318)
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Uniform-datapath variants (small plans).  Tensors, nodes and barycentric weights of EVERY piece /
// slide live in this module's constant bank and are read with warp-uniform `LDCU` feeding
// `DFMA/DADD/DMUL ... UR` directly: the broadcast operands cost no LSU slot and no vector register.
// A weight row is built in registers (unnormalised product form, 6 fp64 ops per node, no memory
// traffic); its 1/sum is folded into the level-0 weights.  Control flow is kept warp-uniform:
//   * one query per thread, no grid-stride loop (ptxas needs that to prove uniformity);
//   * the spline kernel first counting-sorts the CTA's queries by piece (match + one shared atomic
//     per piece per warp), so that all but <= P-1 warps of a CTA are single-piece; a warp then loops
//     over the pieces PRESENT in it (REDUX -> uniform mask) and the lanes that belong to the piece
//     store the result (predicated store, no branch around the math).
// ---------------------------------------------------------------------------------------------
constexpr int BANK_THREADS = 256;
constexpr int BANK_MAXD = 4;  // deeper grids of <= 7808 doubles have <= 4 nodes per dim: not worth the build time
// (GB, DM) instantiation of a bank kernel, DM in {2, 3, 4}
#define BANK_KERNEL_TABLE(K, gb, dm)                                                              \
    ((gb) == 4 ? ((dm) == 2 ? (const void *)K<4, 2> : (dm) == 3 ? (const void *)K<4, 3> : (const void *)K<4, 4>) \
     : (gb) == 2 ? ((dm) == 2 ? (const void *)K<2, 2> : (dm) == 3 ? (const void *)K<2, 3> : (const void *)K<2, 4>) \
                 : ((dm) == 2 ? (const void *)K<1, 2> : (dm) == 3 ? (const void *)K<1, 3> : (const void *)K<1, 4>))
constexpr int BANK_DOUBLES = 7808;  // 61 KB of tensors + nodes + weights (+ 2.75 KB descriptors)
constexpr int BANK_GRIDS = 32;      // also the width of the spline kernel's piece-presence mask
__constant__ double c_grid[BANK_DOUBLES];
__constant__ BankGrid c_bgrid[BANK_GRIDS];
static ConstBank g_grid_bank;

// 1/s without the fp64 division's slow-path CALL: a call inside the kernel's loops makes ptxas give
// up on the uniform datapath for the whole kernel (every LDCU becomes a per-lane LDC).  s is a row
// sum of moderate magnitude here (see the power-of-two prescale in bank_build), so the
// flush-to-zero seed is safe; three Newton steps from the 20-bit seed leave <= 1 ulp.
__device__ __forceinline__ double bank_rcp(double s) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
#pragma unroll
    for (int it = 0; it < 3; ++it) y = fma(y, fma(-s, y, 1.0), y);
    return y;
}

// Unnormalised barycentric row a[i] = w_i * prod_{k != i} (x - x_k), i < N, in registers; returns
// sum_i a[i].  A node hit (|x - x_k| < 1e-14, first k) gives the one-hot row with sum 1
// (barycentric.py:147-151).  x, the nodes (bank index nbase) and the hit threshold `eps_bits` are
// in the grid dimension's PRESCALED units: bank_build multiplies them by one power of two per
// dimension so that the node span is ~4 (every product of distances stays near 1 whatever the
// domain width) -- exact, so each distance is the reference's times that power of two and the
// normalised weights are bit-identical to the unscaled product form.  The weights are scaled by
// another power of two (max |w| in [0.5, 1)), which the normalisation cancels.
// N is a template argument (dispatched by a uniform switch) so that no per-node guard is needed:
// ptxas if-converts such guards into per-lane predicates, which drags the bank reads off the
// uniform path.
template <int N, int STORE>
__device__ __forceinline__ double bank_row_n(double x, int nbase, int wbase, long long eps_bits,
                                             double (&a)[GRID_NL], double *ws, int wstride, double fold) {
    double d[N];
    double pre = 1.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        d[i] = x - c_grid[nbase + i];
        a[i] = pre;
        pre *= d[i];
    }
    double suf = 1.0, sum = 0.0;
    int hit = -1;
#pragma unroll
    for (int i = N - 1; i >= 0; --i) {
        a[i] = c_grid[wbase + i] * a[i] * suf;
        sum += a[i];
        suf *= d[i];
        // |d| < eps on the integer pipe (IEEE order of non-negative doubles = integer order; NaN
        // compares false like the reference's fabs(d) < 1e-14); descending: the lowest index wins
        if ((__double_as_longlong(d[i]) & 0x7fffffffffffffffLL) < eps_bits) hit = i;
    }
    if (hit >= 0) {
#pragma unroll
        for (int i = 0; i < N; ++i) a[i] = i == hit ? 1.0 : 0.0;
        sum = 1.0;
    }
    if constexpr (STORE == 2) {
        // raw (unnormalised) row to the smem column, zero-padded to GRID_NL entries
#pragma unroll
        for (int i = 0; i < GRID_NL; ++i) ws[i * wstride] = i < N ? a[i] : 0.0;
    } else if constexpr (STORE == 1) {
        // normalised row to the smem column; `fold` carries the last row's 1/sum into dim 0
        const double inv = bank_rcp(sum);
#pragma unroll
        for (int i = 0; i < N; ++i) ws[i * wstride] = (a[i] * inv) * fold;
    } else {
#pragma unroll
        for (int i = N; i < GRID_NL; ++i) a[i] = 0.0;
    }
    return sum;
}

#define BANK_ROW_CASE(K) \
    case K: return bank_row_n<K, STORE>(x, nbase, wbase, eps_bits, a, ws, wstride, fold);

// Row of an n-node dimension.  STORE 1: normalised (times `fold`) into the smem column ws; 0:
// unnormalised in a[] (entries >= n zero); 2: unnormalised into the smem column, zero-padded to
// GRID_NL entries.  Returns the row sum.
template <int STORE>
__device__ __forceinline__ double bank_row(double x, int n, int nbase, int wbase, long long eps_bits,
                                           double (&a)[GRID_NL], double *ws, int wstride, double fold) {
    static_assert(GRID_NL == 16, "the bank path dispatches 1..16 nodes");
    switch (n) {
        BANK_ROW_CASE(1) BANK_ROW_CASE(2) BANK_ROW_CASE(3) BANK_ROW_CASE(4) BANK_ROW_CASE(5)
        BANK_ROW_CASE(6) BANK_ROW_CASE(7) BANK_ROW_CASE(8) BANK_ROW_CASE(9) BANK_ROW_CASE(10)
        BANK_ROW_CASE(11) BANK_ROW_CASE(12) BANK_ROW_CASE(13) BANK_ROW_CASE(14) BANK_ROW_CASE(15)
        BANK_ROW_CASE(16)
    }
    return 1.0;
}

// Weights of bank grid `g` at coordinates x(d): last dim unnormalised in wl[] (registers), dims
// 0..D-2 normalised in the smem column ws; returns the factor still to be applied to the result
// (1/sum of the last row for D = 1; folded into dim 0's row otherwise).
template <typename Coord>
__device__ __forceinline__ double bank_weights(const BankGrid &g, Coord x, double *ws, int wstride,
                                               double (&wl)[GRID_NL]) {
    const int D = g.D;
    int off_last = 0;
    for (int d = 0; d + 1 < D; ++d) off_last += g.n[d];
    double fold = bank_rcp(bank_row<false>(x(D - 1) * c_grid[g.scale_off + D - 1], g.n[D - 1],
                                           g.node_off + off_last, g.weight_off + off_last,
                                           __double_as_longlong(c_grid[g.scale_off + D + D - 1]), wl,
                                           nullptr, 0, 1.0));
    int off = 0;
    for (int d = 0; d + 1 < D; ++d) {
        double a[GRID_NL];
        bank_row<true>(x(d) * c_grid[g.scale_off + d], g.n[d], g.node_off + off, g.weight_off + off,
                       __double_as_longlong(c_grid[g.scale_off + D + d]), a, ws + off * wstride, wstride,
                       fold);
        fold = 1.0;
        off += g.n[d];
    }
    return fold;
}

// Innermost contraction: sum_i wl[i] * T[base + i*GB + j], i < N, four partial sums per output.
// N is a template argument for the same reason as in bank_row_n (no per-element guards).
template <int N, int GB>
__device__ __forceinline__ void bank_dot_n(int base, const double (&wl)[GRID_NL], double (&out)[GB]) {
    constexpr int C = N < 4 ? N : 4;
    double part[C][GB];
#pragma unroll
    for (int i = 0; i < N; ++i) {
#pragma unroll
        for (int j = 0; j < GB; ++j) {
            const double t = c_grid[base + i * GB + j];
            part[i % C][j] = i < C ? wl[i] * t : fma(wl[i], t, part[i % C][j]);
        }
    }
#pragma unroll
    for (int j = 0; j < GB; ++j) {
        if constexpr (C == 4)
            out[j] = (part[0][j] + part[1][j]) + (part[2][j] + part[3][j]);
        else if constexpr (C == 3)
            out[j] = (part[0][j] + part[1][j]) + part[2][j];
        else if constexpr (C == 2)
            out[j] = part[0][j] + part[1][j];
        else
            out[j] = part[0][j];
    }
}

#define BANK_DOT_CASE(K) \
    case K: bank_dot_n<K, GB>(base, wl, out); break;

template <int LEVEL, int D, int GB>
struct BankContract {
    __device__ __forceinline__ static void run(int base, const BankGrid &g, const double *ws, int wstride,
                                               const double (&wl)[GRID_NL], double (&out)[GB]) {
        const int nl = g.n[LEVEL];
        if constexpr (LEVEL == D - 1) {
            switch (nl) {
                BANK_DOT_CASE(1) BANK_DOT_CASE(2) BANK_DOT_CASE(3) BANK_DOT_CASE(4) BANK_DOT_CASE(5)
                BANK_DOT_CASE(6) BANK_DOT_CASE(7) BANK_DOT_CASE(8) BANK_DOT_CASE(9) BANK_DOT_CASE(10)
                BANK_DOT_CASE(11) BANK_DOT_CASE(12) BANK_DOT_CASE(13) BANK_DOT_CASE(14)
                BANK_DOT_CASE(15) BANK_DOT_CASE(16)
            }
        } else {
#pragma unroll
            for (int j = 0; j < GB; ++j) out[j] = 0.0;
            // stride of this level from compile-time indices only: a dynamically indexed stride
            // array is if-converted with per-lane predicates and drags every bank address (and
            // with it every bank read) off the uniform datapath
            int st = GB;
#pragma unroll
            for (int d = LEVEL + 1; d < D; ++d) st *= g.n[d];
            const double *wnext = ws + (size_t)nl * wstride;
            for (int i = 0; i < nl; ++i) {
                const double w = ws[i * wstride];  // issued before the row's dot product: latency hidden
                double sub[GB];
                BankContract<(LEVEL + 1 < D ? LEVEL + 1 : LEVEL), D, GB>::run(base + i * st, g, wnext,
                                                                             wstride, wl, sub);
#pragma unroll
                for (int j = 0; j < GB; ++j) out[j] = fma(w, sub[j], out[j]);
            }
        }
    }
};

#define BANK_CASE(K)                                                                    \
    case K:                                                                             \
        if constexpr (K <= DM) BankContract<0, K, GB>::run(base, g, ws, wstride, wl, out); \
        break;

// Contract output block `b` of bank grid `g` (a reference INTO c_bgrid: uniform).
template <int GB, int DM>
__device__ __forceinline__ void bank_contract(const BankGrid &g, int b, const double *ws, int wstride,
                                              const double (&wl)[GRID_NL], double scale,
                                              double (&out)[GB]) {
    const int base = g.tensor_off + b * g.size * GB;
#pragma unroll
    for (int j = 0; j < GB; ++j) out[j] = 0.0;
    switch (g.D) {
        BANK_CASE(1) BANK_CASE(2) BANK_CASE(3) BANK_CASE(4)
    }
    if (g.D == 1) {
#pragma unroll
        for (int j = 0; j < GB; ++j) out[j] *= scale;
    }
}

template <int GB, int DM>
__global__ void __launch_bounds__(BANK_THREADS)
spline_bank_kernel(int D, int G, int g0, int Gtot, int P, const int *__restrict__ num_knots,
                   const int *__restrict__ knot_off, const double *__restrict__ knots,
                   const double *__restrict__ pts, int64_t N, double *__restrict__ out,
                   int32_t *__restrict__ piece_out) {
    extern __shared__ __align__(16) double smem[];
    __shared__ int s_cnt[BANK_GRIDS];
    __shared__ unsigned short s_perm[BANK_THREADS];
    __shared__ unsigned char s_piece[BANK_THREADS];
    const int tid = threadIdx.x, lane = tid & 31;
    const int64_t q0 = (int64_t)blockIdx.x * BANK_THREADS;
    int mine;
    {
        // routing (spline.py:677-690) of the query this thread LOADS; tail lanes take the last one
        const int64_t ql = q0 + tid < N ? q0 + tid : N - 1;
        mine = spline_piece_index(D, num_knots, knot_off, knots, pts + ql * D);
        if (piece_out && q0 + tid < N) piece_out[q0 + tid] = mine;
    }
    // counting sort of the CTA's queries by piece
    if (tid < BANK_GRIDS) s_cnt[tid] = 0;
    __syncthreads();
    const unsigned peers = __match_any_sync(0xffffffffu, mine);
    const int leader = __ffs(peers) - 1;
    int warp_off = 0;
    if (lane == leader) warp_off = atomicAdd(&s_cnt[mine], __popc(peers));
    warp_off = __shfl_sync(0xffffffffu, warp_off, leader);
    __syncthreads();
    int pos = warp_off + __popc(peers & ((1u << lane) - 1u));
    for (int p = 0; p < P; ++p) pos += p < mine ? s_cnt[p] : 0;
    s_perm[pos] = (unsigned short)tid;
    s_piece[pos] = (unsigned char)mine;
    __syncthreads();
    // from here on this thread EVALUATES query s_perm[tid] of the CTA, which lies in piece `mine`
    mine = s_piece[tid];
    const int64_t q = q0 + s_perm[tid];
    const bool live = q < N;
    const double *x = pts + (live ? q : N - 1) * D;
    double *o = out + (live ? q : N - 1) * Gtot + g0;  // this launch writes outputs [g0, g0 + G)
    const unsigned present = __reduce_or_sync(0xffffffffu, 1u << mine);
    for (int p = 0; p < P; ++p) {
        if (!((present >> p) & 1u)) continue;  // uniform: REDUX leaves the mask in a uniform register
        const BankGrid &g = c_bgrid[p];
        const bool keep = live && mine == p;
        double wl[GRID_NL];
        const double inv = bank_weights(g, [&](int d) { return __ldg(x + d); }, smem + tid, BANK_THREADS, wl);
        for (int b = 0; b * GB < G; ++b) {
            double r[GB];
            bank_contract<GB, DM>(g, b, smem + tid, BANK_THREADS, wl, inv, r);
#pragma unroll
            for (int j = 0; j < GB; ++j)
                if (keep && b * GB + j < G) o[b * GB + j] = r[j];
        }
    }
}

constexpr int SLIDER_ACC = 4;  // output rows accumulated in registers; further rows in global memory

template <int GB, int DM>
__global__ void __launch_bounds__(BANK_THREADS)
slider_bank_kernel(int D, int S, int G, double pivot, const int *__restrict__ out_slide,
                   const int *__restrict__ row_out, const double *__restrict__ pts, int64_t N,
                   double *__restrict__ out) {
    extern __shared__ __align__(16) double smem[];
    const int64_t q = (int64_t)blockIdx.x * BANK_THREADS + threadIdx.x;
    const bool live = q < N;
    const int64_t qc = live ? q : N - 1;
    const double *x = pts + qc * D;
    double *o = out + qc * G;
    double acc[SLIDER_ACC];
#pragma unroll
    for (int g = 0; g < SLIDER_ACC; ++g) acc[g] = (g < G && out_slide[g] == -1) ? pivot : 0.0;
    for (int g = SLIDER_ACC; g < G; ++g)
        if (live) o[g] = out_slide[g] == -1 ? pivot : 0.0;
    for (int s = 0; s < S; ++s) {
        // every value that steers control flow comes from the constant bank: a branch on a GLOBAL
        // load that follows per-lane global stores is not provably uniform for ptxas
        const BankGrid &gr = c_bgrid[s];
        const int sg = gr.outputs;
        if (sg == 0) continue;
        double wl[GRID_NL];
        double *ws = smem + threadIdx.x;
        const double inv = bank_weights(gr, [&](int d) { return __ldg(x + gr.dims[d]); }, ws, BANK_THREADS, wl);
        for (int b = 0; b * GB < sg; ++b) {
            double r[GB];
            bank_contract<GB, DM>(gr, b, ws, BANK_THREADS, wl, inv, r);
#pragma unroll
            for (int j = 0; j < GB; ++j) {
                const int so = b * GB + j;
                if (so >= sg) continue;
                // row g takes output row_out[g] of slide out_slide[g]; value rows (-1) add
                // (slide value - pivot) of every slide (slider.py:300-318)
#pragma unroll
                for (int g = 0; g < SLIDER_ACC; ++g) {
                    if (g < G) {
                        const int os = out_slide[g], ro = row_out[g];
                        if (os == s && ro == so)
                            acc[g] = r[j];
                        else if (os == -1 && so == 0 && ro == 0)
                            acc[g] = acc[g] + (r[j] - pivot);
                    }
                }
                for (int g = SLIDER_ACC; g < G; ++g) {
                    const int os = out_slide[g], ro = row_out[g];
                    if (live) {
                        if (os == s && ro == so)
                            o[g] = r[j];
                        else if (os == -1 && so == 0 && ro == 0)
                            o[g] = o[g] + (r[j] - pivot);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int g = 0; g < SLIDER_ACC; ++g)
        if (live && g < G) o[g] = acc[g];
}


// ---------------------------------------------------------------------------------------------
// 2-D grids on the FP64 tensor cores (round 2).  A 2-D barycentric value is the bilinear form
//     p(x, y) = a(x)^T T b(y) / (sum a * sum b)
// so for 8 queries at a time the contraction over the FIRST axis is an 8 x n0 by n0 x n1 GEMM:
//     U[q, j] = sum_i A[q, i] T[i, j]           mma.sync.m8n8k4.f64  (SASS DMMA.8x8x4)
// followed by a per-query dot with b(y): 2 products per lane and a quad shuffle.  One DMMA does the
// work of 8 DFMA warp instructions, which is what the thread-per-query evaluator was short of: ncu
// on spline_bank_kernel<1,2> showed FP64 pipe 25 %, issue slots 41 % busy, 1790 instructions per
// query against ~600 FP64 ones.  Here a warp owns 32 queries:
//   1. every lane builds the two weight rows of ITS query in registers (product form, operands on
//      the uniform datapath from the constant bank, exactly as in the bank kernels) and parks them
//      in a per-warp shared tile, i-major with padded strides (36 / 34 doubles) so that both the
//      stores and the fragment loads below are bank-conflict free;
//   2. for each of the 4 row tiles (8 queries) and each piece present in it: A fragments from the
//      tile, B fragments = the piece's tensor pre-arranged in fragment order (staged in shared
//      memory once per CTA), ceil(n0/4) x ceil(n1/8) DMMAs, the fold with b, a quad reduction;
//   3. the owner lane scales by 1 / (sum a * sum b) and stores.
// Splines first counting-sort the CTA's queries by piece (as spline_bank_kernel does), so all but
// <= P - 1 row tiles per CTA are single-piece.  Sliders run the three steps once per slide.
// ---------------------------------------------------------------------------------------------
constexpr int BL_THREADS = 128;                 // 4 warps; ~10 KB of tiles per warp
constexpr int BL_SA = 36, BL_SB = 34;           // padded strides (doubles) of the weight tiles
constexpr int BL_FRAG = 256;                    // doubles per (grid, output): 4 k-blocks x 2 n-tiles x 32 lanes
constexpr int BL_MAX_FRAGS = 24;                // staged in shared memory (48 KB)
constexpr int BL_OUT = 2;                       // outputs folded per pass (2: five 128-thread CTAs fit in an SM)
constexpr int BL_WARP_DOUBLES = 16 * BL_SA + 16 * BL_SB + BL_OUT * 32 + 32;  // tiles, outputs, scratch row

__device__ __forceinline__ void bl_dmma(double &c0, double &c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

// Weight rows of bank grid `g` for this lane's query into the warp tile (lanes that do not `keep`
// write a scratch row instead); returns 1 / (sum a * sum b).
// The rows are stored from INSIDE bank_row_n's per-length cases (mode 2): merging a register row out
// of the switch and storing it afterwards takes every bank read of the kernel off the uniform
// datapath -- and so does a register cap (`__launch_bounds__(.., minBlocks)`), which is why these
// kernels carry none (they need 72-78 registers).
// Kernel variants for plans whose grids all have the same node count in every dimension.
#define PCB_NFIX_KERNEL(nf, ...)                                                        \
    ((nf) == 8    ? (const void *)__VA_ARGS__<8>                                        \
     : (nf) == 11 ? (const void *)__VA_ARGS__<11>                                       \
     : (nf) == 12 ? (const void *)__VA_ARGS__<12>                                       \
     : (nf) == 15 ? (const void *)__VA_ARGS__<15>                                       \
     : (nf) == 16 ? (const void *)__VA_ARGS__<16>                                       \
                  : (const void *)__VA_ARGS__<0>)

// NFIX > 0: every dimension of every grid of the plan has NFIX nodes, known at compile time -- the row
// builder is then straight-line code; through the 16-way switch ptxas tail-merges the per-length
// cases into one chain of blocks joined by uniform branches (~6 extra instructions per node).
template <int NFIX, typename Coord>
__device__ __forceinline__ double bl_rows(const BankGrid &g, Coord x, bool keep, double *sA, double *sB,
                                          int lane) {
    double row[GRID_NL];
    double *dump = sA + 16 * BL_SA + 16 * BL_SB + BL_OUT * 32;  // 32 doubles per warp
    double sa, sb;
    if constexpr (NFIX > 0) {
        sa = bank_row_n<NFIX, 2>(x(0) * c_grid[g.scale_off + 0], g.node_off, g.weight_off,
                                 __double_as_longlong(c_grid[g.scale_off + 2 + 0]), row,
                                 (keep ? sA : dump) + lane, keep ? BL_SA : 0, 1.0);
        sb = bank_row_n<NFIX, 2>(x(1) * c_grid[g.scale_off + 1], g.node_off + NFIX, g.weight_off + NFIX,
                                 __double_as_longlong(c_grid[g.scale_off + 2 + 1]), row,
                                 (keep ? sB : dump) + lane, keep ? BL_SB : 0, 1.0);
    } else {
        sa = bank_row<2>(x(0) * c_grid[g.scale_off + 0], g.n[0], g.node_off, g.weight_off,
                         __double_as_longlong(c_grid[g.scale_off + 2 + 0]), row,
                         (keep ? sA : dump) + lane, keep ? BL_SA : 0, 1.0);
        sb = bank_row<2>(x(1) * c_grid[g.scale_off + 1], g.n[1], g.node_off + g.n[0],
                         g.weight_off + g.n[0],
                         __double_as_longlong(c_grid[g.scale_off + 2 + 1]), row,
                         (keep ? sB : dump) + lane, keep ? BL_SB : 0, 1.0);
    }
    return bank_rcp(sa * sb);
}

// Row tile t (queries 8t..8t+7 of the warp) against one grid, outputs [o0, o0 + no): the lanes with
// (lane & 3) == 0 whose row passes `take` write sOut[o * 32 + row].  KB = ceil(n0 / 4) k-blocks and
// NT = ceil(n1 / 8) column tiles are template arguments (dispatched by a uniform switch): runtime
// guards cost an ISETP + SEL per fragment element, more than the MMAs themselves.
template <int KB, int NT>
__device__ __forceinline__ void bl_tile_n(const double *frag, int o0, int no, int t, const double *sA,
                                          const double *sB, double *sOut, int lane, bool take) {
    const int r = lane >> 2, c = lane & 3;
    double af[KB];
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) af[kb] = sA[(4 * kb + c) * BL_SA + 8 * t + r];
    double bq[NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) bq[nt][e] = sB[(8 * nt + 2 * c + e) * BL_SB + 8 * t + r];
    for (int o = 0; o < no; ++o) {
        const double *f = frag + (size_t)(o0 + o) * BL_FRAG + lane;
        double acc[NT][2];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = 0.0;
#pragma unroll
        for (int kb = 0; kb < KB; ++kb)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) bl_dmma(acc[nt][0], acc[nt][1], af[kb], f[(kb * 2 + nt) * 32]);
        double part = acc[0][0] * bq[0][0];
        part = fma(acc[0][1], bq[0][1], part);
        if constexpr (NT > 1) {
            part = fma(acc[1][0], bq[1][0], part);
            part = fma(acc[1][1], bq[1][1], part);
        }
        part += __shfl_xor_sync(0xffffffffu, part, 1);
        part += __shfl_xor_sync(0xffffffffu, part, 2);
        if (c == 0 && take) sOut[o * 32 + 8 * t + r] = part;
    }
}

#define BL_TILE_CASE(KBV, NTV) \
    case (KBV) * 2 + (NTV): bl_tile_n<KBV, NTV>(frag, o0, no, t, sA, sB, sOut, lane, take); break;

__device__ __forceinline__ void bl_tile(const BankGrid &g, const double *frag, int o0, int no, int t,
                                        const double *sA, const double *sB, double *sOut, int lane,
                                        bool take) {
    switch (((g.n[0] + 3) >> 2) * 2 + ((g.n[1] + 7) >> 3)) {
        BL_TILE_CASE(1, 1) BL_TILE_CASE(1, 2) BL_TILE_CASE(2, 1) BL_TILE_CASE(2, 2)
        BL_TILE_CASE(3, 1) BL_TILE_CASE(3, 2) BL_TILE_CASE(4, 1) BL_TILE_CASE(4, 2)
    }
}

// All four row tiles of a single-grid warp at once (2-D): each B fragment is loaded once for four
// DMMAs and eight accumulator chains are in flight.
template <int KB, int NT>
__device__ __forceinline__ void bl_tiles2_n(const double *frag, int o0, int no, int t0, const double *sA,
                                            const double *sB, double *sOut, int lane) {
    const int r = lane >> 2, c = lane & 3;
    double af[KB][2];
#pragma unroll
    for (int kb = 0; kb < KB; ++kb)
#pragma unroll
        for (int t = 0; t < 2; ++t) af[kb][t] = sA[(4 * kb + c) * BL_SA + 8 * (t0 + t) + r];
    for (int o = 0; o < no; ++o) {
        const double *f = frag + (size_t)(o0 + o) * BL_FRAG + lane;
        double acc[2][NT][2];
#pragma unroll
        for (int t = 0; t < 2; ++t)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) acc[t][nt][0] = acc[t][nt][1] = 0.0;
#pragma unroll
        for (int kb = 0; kb < KB; ++kb)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const double b = f[(kb * 2 + nt) * 32];
#pragma unroll
                for (int t = 0; t < 2; ++t) bl_dmma(acc[t][nt][0], acc[t][nt][1], af[kb][t], b);
            }
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            double part = 0.0;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                part = fma(acc[t][nt][0], sB[(8 * nt + 2 * c) * BL_SB + 8 * (t0 + t) + r], part);
                part = fma(acc[t][nt][1], sB[(8 * nt + 2 * c + 1) * BL_SB + 8 * (t0 + t) + r], part);
            }
            part += __shfl_xor_sync(0xffffffffu, part, 1);
            part += __shfl_xor_sync(0xffffffffu, part, 2);
            if (c == 0) sOut[o * 32 + 8 * (t0 + t) + r] = part;
        }
    }
}
#define BL_TILES2_CASE(KBV, NTV) \
    case (KBV) * 2 + (NTV): bl_tiles2_n<KBV, NTV>(frag, o0, no, t0, sA, sB, sOut, lane); break;
__device__ __forceinline__ void bl_tiles2(const BankGrid &g, const double *frag, int o0, int no, int t0,
                                          const double *sA, const double *sB, double *sOut, int lane) {
    switch (((g.n[0] + 3) >> 2) * 2 + ((g.n[1] + 7) >> 3)) {
        BL_TILES2_CASE(1, 1) BL_TILES2_CASE(1, 2) BL_TILES2_CASE(2, 1) BL_TILES2_CASE(2, 2)
        BL_TILES2_CASE(3, 1) BL_TILES2_CASE(3, 2) BL_TILES2_CASE(4, 1) BL_TILES2_CASE(4, 2)
    }
}

template <int KB, int NT>
__device__ __forceinline__ void bl_tiles4_n(const double *frag, int o0, int no, const double *sA,
                                            const double *sB, double *sOut, int lane) {
    const int r = lane >> 2, c = lane & 3;
    double af[KB][4];
#pragma unroll
    for (int kb = 0; kb < KB; ++kb)
#pragma unroll
        for (int t = 0; t < 4; ++t) af[kb][t] = sA[(4 * kb + c) * BL_SA + 8 * t + r];
    for (int o = 0; o < no; ++o) {
        const double *f = frag + (size_t)(o0 + o) * BL_FRAG + lane;
        double acc[4][NT][2];
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) acc[t][nt][0] = acc[t][nt][1] = 0.0;
#pragma unroll
        for (int kb = 0; kb < KB; ++kb)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const double b = f[(kb * 2 + nt) * 32];
#pragma unroll
                for (int t = 0; t < 4; ++t) bl_dmma(acc[t][nt][0], acc[t][nt][1], af[kb][t], b);
            }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            double part = 0.0;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                part = fma(acc[t][nt][0], sB[(8 * nt + 2 * c) * BL_SB + 8 * t + r], part);
                part = fma(acc[t][nt][1], sB[(8 * nt + 2 * c + 1) * BL_SB + 8 * t + r], part);
            }
            part += __shfl_xor_sync(0xffffffffu, part, 1);
            part += __shfl_xor_sync(0xffffffffu, part, 2);
            if (c == 0) sOut[o * 32 + 8 * t + r] = part;
        }
    }
}

#define BL_TILES4_CASE(KBV, NTV) \
    case (KBV) * 2 + (NTV): bl_tiles4_n<KBV, NTV>(frag, o0, no, sA, sB, sOut, lane); break;

__device__ __forceinline__ void bl_tiles4(const BankGrid &g, const double *frag, int o0, int no,
                                          const double *sA, const double *sB, double *sOut, int lane) {
    switch (((g.n[0] + 3) >> 2) * 2 + ((g.n[1] + 7) >> 3)) {
        BL_TILES4_CASE(1, 1) BL_TILES4_CASE(1, 2) BL_TILES4_CASE(2, 1) BL_TILES4_CASE(2, 2)
        BL_TILES4_CASE(3, 1) BL_TILES4_CASE(3, 2) BL_TILES4_CASE(4, 1) BL_TILES4_CASE(4, 2)
    }
}

__device__ __forceinline__ void bl_stage_frags(double *sFrag, const double *__restrict__ frags, int nfrag) {
    for (int e = threadIdx.x; e < nfrag * BL_FRAG; e += BL_THREADS) sFrag[e] = __ldg(frags + e);
}

template <int NFIX>
__global__ void __launch_bounds__(BL_THREADS)
spline2d_dmma_kernel(int G, int P, const int *__restrict__ num_knots, const int *__restrict__ knot_off,
                     const double *__restrict__ knots, const double *__restrict__ frags,
                     const double *__restrict__ pts, int64_t N, double *__restrict__ out,
                     int32_t *__restrict__ piece_out) {
    extern __shared__ __align__(16) double smem[];
    __shared__ int s_cnt[BANK_GRIDS];
    __shared__ unsigned short s_perm[BL_THREADS];
    __shared__ unsigned char s_piece[BL_THREADS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double *sFrag = smem;
    double *sA = smem + (size_t)P * G * BL_FRAG + warp * BL_WARP_DOUBLES;
    double *sB = sA + 16 * BL_SA;
    double *sOut = sB + 16 * BL_SB;
    const int64_t q0 = (int64_t)blockIdx.x * BL_THREADS;
    bl_stage_frags(sFrag, frags, P * G);
    int mine;
    {
        const int64_t ql = q0 + tid < N ? q0 + tid : N - 1;
        mine = spline_piece_index(2, num_knots, knot_off, knots, pts + ql * 2);
        if (piece_out && q0 + tid < N) piece_out[q0 + tid] = mine;
    }
    // counting sort of the CTA's queries by piece (see spline_bank_kernel)
    if (tid < BANK_GRIDS) s_cnt[tid] = 0;
    __syncthreads();
    const unsigned peers = __match_any_sync(0xffffffffu, mine);
    const int leader = __ffs(peers) - 1;
    int warp_off = 0;
    if (lane == leader) warp_off = atomicAdd(&s_cnt[mine], __popc(peers));
    warp_off = __shfl_sync(0xffffffffu, warp_off, leader);
    __syncthreads();
    int pos = warp_off + __popc(peers & ((1u << lane) - 1u));
    for (int p = 0; p < P; ++p) pos += p < mine ? s_cnt[p] : 0;
    s_perm[pos] = (unsigned short)tid;
    s_piece[pos] = (unsigned char)mine;
    __syncthreads();  // also: fragments staged
    mine = s_piece[tid];
    const int64_t q = q0 + s_perm[tid];
    const bool live = q < N;
    const double *x = pts + (live ? q : N - 1) * 2;
    double *o = out + (live ? q : N - 1) * G;
    // 1. weight rows, once per piece present in the warp (uniform loop, predicated stores)
    const unsigned present = __reduce_or_sync(0xffffffffu, 1u << mine);
    double inv = 1.0;
    for (int p = 0; p < P; ++p) {
        if (!((present >> p) & 1u)) continue;
        const double v = bl_rows<NFIX>(c_bgrid[p], [&](int d) { return __ldg(x + d); }, mine == p, sA, sB, lane);
        if (mine == p) inv = v;
    }
    __syncwarp();
    // 2. + 3. row tiles x pieces present in the tile, BL_OUT outputs per pass
    for (int o0 = 0; o0 < G; o0 += BL_OUT) {
        const int no = G - o0 < BL_OUT ? G - o0 : BL_OUT;
        if ((present & (present - 1u)) == 0u) {
            const int p = __ffs(present) - 1;
#pragma unroll 1
            for (int t0 = 0; t0 < 4; t0 += 2) {
                if constexpr (NFIX > 0)
                    bl_tiles2_n<(NFIX + 3) / 4, (NFIX + 7) / 8>(sFrag + (size_t)p * G * BL_FRAG, o0, no, t0, sA,
                                                                sB, sOut, lane);
                else
                    bl_tiles2(c_bgrid[p], sFrag + (size_t)p * G * BL_FRAG, o0, no, t0, sA, sB, sOut, lane);
            }
        } else {
#pragma unroll 1
            for (int t = 0; t < 4; ++t) {
                const int prow = __shfl_sync(0xffffffffu, mine, 8 * t + (lane >> 2));
                const unsigned here = __reduce_or_sync(0xffffffffu, 1u << prow);
                for (int p = 0; p < P; ++p) {
                    if (!((here >> p) & 1u)) continue;
                    bl_tile(c_bgrid[p], sFrag + (size_t)p * G * BL_FRAG, o0, no, t, sA, sB, sOut, lane,
                            prow == p);
                }
            }
        }
        __syncwarp();
        for (int k = 0; k < no; ++k) {
            const double v = sOut[k * 32 + lane] * inv;
            if (live) o[o0 + k] = v;
        }
        __syncwarp();
    }
}

// Slider of 2-D slides: value rows accumulate pivot + sum_s (slide_s - pivot) left to right
// (slider.py:310-318), a derivative row takes its slide's output, cross-slide rows are exactly 0.
// All four row tiles of a warp are multiplied at once (bl_tiles4).  (minBlocks = 6 is an 80-register
// budget: with the four-tile routine ptxas keeps the bank reads on the uniform datapath at 80
// registers and drops them at the unbounded 94; the same routine takes spline2d_dmma_kernel off the
// uniform path at any budget, so that kernel multiplies tile by tile.  tools/check_sass.py)
template <int NFIX>
__global__ void __launch_bounds__(BL_THREADS, 6)
slider2d_dmma_kernel(int D, int S, int G, double pivot, const int *__restrict__ out_slide,
                     const int *__restrict__ row_out, const int *__restrict__ frag_off, int nfrag,
                     const double *__restrict__ frags, const double *__restrict__ pts, int64_t N,
                     double *__restrict__ out) {
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double *sFrag = smem;
    double *sA = smem + (size_t)nfrag * BL_FRAG + warp * BL_WARP_DOUBLES;
    double *sB = sA + 16 * BL_SA;
    double *sOut = sB + 16 * BL_SB;
    const int64_t q = (int64_t)blockIdx.x * BL_THREADS + tid;
    const bool live = q < N;
    const double *x = pts + (live ? q : N - 1) * D;
    double *o = out + (live ? q : N - 1) * G;
    bl_stage_frags(sFrag, frags, nfrag);
    double acc[SLIDER_ACC];
#pragma unroll
    for (int g = 0; g < SLIDER_ACC; ++g) acc[g] = (g < G && out_slide[g] == -1) ? pivot : 0.0;
    for (int g = SLIDER_ACC; g < G; ++g)
        if (live) o[g] = out_slide[g] == -1 ? pivot : 0.0;
    __syncthreads();
    for (int s = 0; s < S; ++s) {
        const BankGrid &gr = c_bgrid[s];
        const int sg = gr.outputs;
        if (sg == 0) continue;
        const double inv = bl_rows<NFIX>(gr, [&](int d) { return __ldg(x + gr.dims[d]); }, true, sA, sB, lane);
        __syncwarp();
        for (int o0 = 0; o0 < sg; o0 += BL_OUT) {
            const int no = sg - o0 < BL_OUT ? sg - o0 : BL_OUT;
            if constexpr (NFIX > 0)
                bl_tiles4_n<(NFIX + 3) / 4, (NFIX + 7) / 8>(sFrag + (size_t)frag_off[s] * BL_FRAG, o0, no, sA, sB,
                                                            sOut, lane);
            else
                bl_tiles4(gr, sFrag + (size_t)frag_off[s] * BL_FRAG, o0, no, sA, sB, sOut, lane);
            __syncwarp();
            for (int k = 0; k < no; ++k) {
                const int so = o0 + k;
                const double r = sOut[k * 32 + lane] * inv;
#pragma unroll
                for (int g = 0; g < SLIDER_ACC; ++g) {
                    if (g < G) {
                        const int os = out_slide[g], ro = row_out[g];
                        if (os == s && ro == so)
                            acc[g] = r;
                        else if (os == -1 && so == 0 && ro == 0)
                            acc[g] = acc[g] + (r - pivot);
                    }
                }
                for (int g = SLIDER_ACC; g < G; ++g) {
                    const int os = out_slide[g], ro = row_out[g];
                    if (live) {
                        if (os == s && ro == so)
                            o[g] = r;
                        else if (os == -1 && so == 0 && ro == 0)
                            o[g] = o[g] + (r - pivot);
                    }
                }
            }
            __syncwarp();
        }
    }
#pragma unroll
    for (int g = 0; g < SLIDER_ACC; ++g)
        if (live && g < G) o[g] = acc[g];
}


// ---------------------------------------------------------------------------------------------
// 3-D spline pieces on the FP64 tensor cores: the last two axes jointly on the MMA K dimension.
//     p(x, y, z) = sum_i a_i  sum_{(j,k)} (b_j c_k) T[i][(j,k)]
// For 8 queries: U[q, i] = A[q, (j,k)] B[(j,k), i],  A = b_j(q) c_k(q) built on the fly,
// B = the piece's tensor in fragment order (shared memory), N = n0 padded to 8; then the dot with
// a(x): 4 FMA per lane and a quad shuffle.  The K index is ordered (j, m, c) with k = 4 m + c, the
// last axis padded to M = ceil(n2 / 4) blocks of 4: lane c of a quad only ever needs c_{4m+c},
// m < M -- at most four values, held in registers for the whole tile -- and b_j changes once per M
// k-steps, so a k-step is one DMUL, the B-fragment loads and the DMMAs: no index table, no
// per-step reads of the weight tiles.  (The first version flattened (j, k) densely, 225 -> 228,
// and looked both factors up per k-step through a table: five shared loads per two DMMAs and a
// load -> load -> DMUL -> DMMA dependency chain per step; this order spends 240 instead of 228
// k-rows on 15^3 and runs 17 % faster, 1.73e9 -> 2.02e9 q/s.)  15^3: 120 DMMA per 8 queries
// replace 3,600 DFMA per thread.
// ---------------------------------------------------------------------------------------------
constexpr int BL3_THREADS = 384;                     // up to 12 warps share one staged copy of the fragments
constexpr int BL3_MAX_KB = 64;                       // n1 * ceil(n2 / 4) <= 64
constexpr int BL3_SA = 33;                           // tile stride: rows are stored lane-consecutive, b is read as
                                                     // quad broadcasts; the odd stride spreads the per-tile a / c reads
constexpr int BL3_WARP_DOUBLES = 3 * 16 * BL3_SA + BL_OUT * 32 + 32;  // three weight tiles, outputs, scratch

// TL 8-query row tiles (t0 .. t0 + TL - 1) against one piece: frag = [n1][M][NT = 2][32] per output.
// One B-fragment load feeds TL DMMAs and the warp carries 2 TL NT independent accumulator chains.
template <int NT, int M, int TL>
__device__ __forceinline__ void bl3_tile_nm(int n1, int n2, int fragstride, const double *frag, int o0,
                                            int no, int t0, const double *sA, const double *sB,
                                            const double *sC, double *sOut, int lane, unsigned take) {
    const int r = lane >> 2, c = lane & 3;
    const int q = 8 * t0 + r;  // row of tile t0; tile t0 + tl is 8 tl further
    double aq[TL][NT][2], cz[TL][M];
#pragma unroll
    for (int tl = 0; tl < TL; ++tl) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) aq[tl][nt][e] = sA[(8 * nt + 2 * c + e) * BL3_SA + q + 8 * tl];
#pragma unroll
        for (int m = 0; m < M; ++m) cz[tl][m] = 4 * m + c < n2 ? sC[(4 * m + c) * BL3_SA + q + 8 * tl] : 0.0;
    }
    for (int o = 0; o < no; ++o) {
        const double *f = frag + (size_t)(o0 + o) * fragstride + lane;
        double acc[TL][NT][2], accb[TL][NT][2];
#pragma unroll
        for (int tl = 0; tl < TL; ++tl)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
                acc[tl][nt][0] = acc[tl][nt][1] = accb[tl][nt][0] = accb[tl][nt][1] = 0.0;
#pragma unroll 2
        for (int j = 0; j < n1; ++j) {
            double bj[TL];
#pragma unroll
            for (int tl = 0; tl < TL; ++tl) bj[tl] = sB[j * BL3_SA + q + 8 * tl];
#pragma unroll
            for (int m = 0; m < M; ++m) {
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const double b = f[((j * M + m) * 2 + nt) * 32];
#pragma unroll
                    for (int tl = 0; tl < TL; ++tl) {
                        const double a = bj[tl] * cz[tl][m];
                        if ((M & 1) ? ((j + m) & 1) : (m & 1))
                            bl_dmma(accb[tl][nt][0], accb[tl][nt][1], a, b);
                        else
                            bl_dmma(acc[tl][nt][0], acc[tl][nt][1], a, b);
                    }
                }
            }
        }
#pragma unroll
        for (int tl = 0; tl < TL; ++tl) {
            double part = 0.0;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                part = fma(acc[tl][nt][0] + accb[tl][nt][0], aq[tl][nt][0], part);
                part = fma(acc[tl][nt][1] + accb[tl][nt][1], aq[tl][nt][1], part);
            }
            part += __shfl_xor_sync(0xffffffffu, part, 1);
            part += __shfl_xor_sync(0xffffffffu, part, 2);
            if (c == 0 && ((take >> tl) & 1u)) sOut[o * 32 + q + 8 * tl] = part;
        }
    }
}

template <int NT, int TL>
__device__ __forceinline__ void bl3_tile_n(int n1, int n2, int fragstride, const double *frag, int o0,
                                           int no, int t0, const double *sA, const double *sB,
                                           const double *sC, double *sOut, int lane, unsigned take) {
    switch ((n2 + 3) >> 2) {  // uniform: the piece's shape comes from the constant bank
    case 1: bl3_tile_nm<NT, 1, TL>(n1, n2, fragstride, frag, o0, no, t0, sA, sB, sC, sOut, lane, take); break;
    case 2: bl3_tile_nm<NT, 2, TL>(n1, n2, fragstride, frag, o0, no, t0, sA, sB, sC, sOut, lane, take); break;
    case 3: bl3_tile_nm<NT, 3, TL>(n1, n2, fragstride, frag, o0, no, t0, sA, sB, sC, sOut, lane, take); break;
    default: bl3_tile_nm<NT, 4, TL>(n1, n2, fragstride, frag, o0, no, t0, sA, sB, sC, sOut, lane, take); break;
    }
}

// One launch evaluates outputs [g0, g0 + G) of Gtot (the fragment images of all outputs of a launch
// must fit in shared memory next to the per-warp tiles).
// TL row tiles at once; MINB = 0 leaves the register budget to ptxas.  Which budget keeps the
// weight-row bank reads on the uniform datapath is an empirical matter (tools/check_sass.py):
// <1, 0> does (80 registers), <1, 1> and <2, 0 / 1 / 2> read the whole bank per lane, <2, 3> does.
// Measured on 15^3 (values): <1, 0> 2.03e9 q/s, <2, 3> 1.85e9, <2, 0> 1.92e9, <4, 0> 1.64e9 -- one
// B-fragment load per two or four DMMAs and twice / four times the accumulator chains do not pay;
// only <1, 0> is instantiated.
template <int TL, int MINB, int NFIX>
__global__ void __launch_bounds__(BL3_THREADS, MINB)
spline3d_dmma_kernel(int G, int g0, int Gtot, int P, int KBmax, const int *__restrict__ num_knots,
                     const int *__restrict__ knot_off, const double *__restrict__ knots,
                     const double *__restrict__ frags, const double *__restrict__ pts, int64_t N,
                     double *__restrict__ out,
                     int32_t *__restrict__ piece_out, int qper) {
    extern __shared__ __align__(16) double smem[];
    __shared__ int s_cnt[BANK_GRIDS];
    __shared__ unsigned short s_perm[BL3_THREADS];
    __shared__ unsigned char s_piece[BL3_THREADS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fragstride = KBmax * 64;                       // doubles per (piece, output)
    double *sFrag = smem;
    double *sA = smem + (size_t)P * G * fragstride + warp * BL3_WARP_DOUBLES;
    double *sB = sA + 16 * BL3_SA;
    double *sC = sB + 16 * BL3_SA;
    double *sOut = sC + 16 * BL3_SA;
    // The CTA takes qper <= blockDim.x queries.  With few pieces the launcher leaves 8 (P - 1) slots
    // free so that every piece's run of sorted queries can start on an 8-query tile boundary: no
    // row tile then mixes pieces, none is contracted twice, and all warps finish together (a CTA
    // waits for its slowest warp; with 2 pieces one warp would run 5 tile passes against 4).
    const int64_t q0 = (int64_t)blockIdx.x * qper;
    const bool aligned = qper < (int)blockDim.x;
    // The fragment images (image of (piece p, output g0 + o) -> slot p * G + o) arrive by TMA bulk
    // copies behind an mbarrier while the CTA routes, sorts and builds its weight rows; the first
    // reader is the contraction.
    __shared__ __align__(8) uint64_t s_bar;
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_arrive_expect_tx(&s_bar, (uint32_t)(P * G * fragstride * sizeof(double)));
        for (int pg = 0; pg < P * G; ++pg)
            bulk_g2s(sFrag + (size_t)pg * fragstride,
                     frags + (size_t)((pg / G) * Gtot + g0 + pg % G) * fragstride,
                     (uint32_t)(fragstride * sizeof(double)), &s_bar);
    }
    const bool has = tid < qper;  // slots beyond qper carry no query of their own
    int mine;
    {
        const int64_t ql = has && q0 + tid < N ? q0 + tid : (q0 < N ? q0 : N - 1);
        mine = spline_piece_index(3, num_knots, knot_off, knots, pts + ql * 3);
        if (piece_out && has && q0 + tid < N) piece_out[q0 + tid] = mine;
    }
    if (tid < BANK_GRIDS) s_cnt[tid] = 0;
    s_perm[tid] = 0xffffu;
    __syncthreads();
    const unsigned peers = __match_any_sync(0xffffffffu, has ? mine : -1);
    const int leader = __ffs(peers) - 1;
    int warp_off = 0;
    if (has && lane == leader) warp_off = atomicAdd(&s_cnt[mine], __popc(peers));
    warp_off = __shfl_sync(0xffffffffu, warp_off, leader);
    __syncthreads();
    if (has) {
        int pos = warp_off + __popc(peers & ((1u << lane) - 1u));
        for (int p = 0; p < P; ++p) pos += p < mine ? (aligned ? (s_cnt[p] + 7) & ~7 : s_cnt[p]) : 0;
        s_perm[pos] = (unsigned short)tid;
        s_piece[pos] = (unsigned char)mine;
    }
    __syncthreads();
    int src = s_perm[tid];
    if (src == 0xffff) {  // padding slot: evaluates the CTA's first query on its run's piece, stores nothing
        int p = 0;
        for (int end = (s_cnt[0] + 7) & ~7; tid >= end && p + 1 < P; end += (s_cnt[p] + 7) & ~7) ++p;
        mine = p;
    } else {
        mine = s_piece[tid];
    }
    const int64_t q = src == 0xffff ? N : q0 + src;
    const bool live = q < N;
    const double *x = pts + (live ? q : N - 1) * 3;
    double *o = out + (live ? q : N - 1) * Gtot + g0;
    // 1. weight rows of the three dims (entry 15 of every tile row set stays zero-padded by mode 2)
    const unsigned present = __reduce_or_sync(0xffffffffu, 1u << mine);
    double inv = 1.0;
    for (int p = 0; p < P; ++p) {
        if (!((present >> p) & 1u)) continue;
        const BankGrid &g = c_bgrid[p];
        const bool keep = mine == p;
        double row[GRID_NL];
        double *dump = sOut + BL_OUT * 32;
        double sa, sb, sc;
        if constexpr (NFIX > 0) {
            sa = bank_row_n<NFIX, 2>(__ldg(x + 0) * c_grid[g.scale_off + 0], g.node_off, g.weight_off,
                                     __double_as_longlong(c_grid[g.scale_off + 3 + 0]), row,
                                     (keep ? sA : dump) + lane, keep ? BL3_SA : 0, 1.0);
            sb = bank_row_n<NFIX, 2>(__ldg(x + 1) * c_grid[g.scale_off + 1], g.node_off + NFIX,
                                     g.weight_off + NFIX, __double_as_longlong(c_grid[g.scale_off + 3 + 1]),
                                     row, (keep ? sB : dump) + lane, keep ? BL3_SA : 0, 1.0);
            sc = bank_row_n<NFIX, 2>(__ldg(x + 2) * c_grid[g.scale_off + 2], g.node_off + 2 * NFIX,
                                     g.weight_off + 2 * NFIX, __double_as_longlong(c_grid[g.scale_off + 3 + 2]),
                                     row, (keep ? sC : dump) + lane, keep ? BL3_SA : 0, 1.0);
        } else {
            sa = bank_row<2>(__ldg(x + 0) * c_grid[g.scale_off + 0], g.n[0], g.node_off, g.weight_off,
                             __double_as_longlong(c_grid[g.scale_off + 3 + 0]), row,
                             (keep ? sA : dump) + lane, keep ? BL3_SA : 0, 1.0);
            sb = bank_row<2>(__ldg(x + 1) * c_grid[g.scale_off + 1], g.n[1], g.node_off + g.n[0],
                             g.weight_off + g.n[0], __double_as_longlong(c_grid[g.scale_off + 3 + 1]), row,
                             (keep ? sB : dump) + lane, keep ? BL3_SA : 0, 1.0);
            sc = bank_row<2>(__ldg(x + 2) * c_grid[g.scale_off + 2], g.n[2], g.node_off + g.n[0] + g.n[1],
                             g.weight_off + g.n[0] + g.n[1], __double_as_longlong(c_grid[g.scale_off + 3 + 2]),
                             row, (keep ? sC : dump) + lane, keep ? BL3_SA : 0, 1.0);
        }
        const double v = bank_rcp(sa * sb * sc);
        if (keep) inv = v;
    }
    __syncwarp();
    mbar_wait(&s_bar, 0);  // fragment images have landed (the barrier's init is two __syncthreads old)
    for (int o0 = 0; o0 < G; o0 += BL_OUT) {
        const int no = G - o0 < BL_OUT ? G - o0 : BL_OUT;
#pragma unroll 1
        for (int t0 = 0; t0 < 4; t0 += TL) {
            unsigned here = 0, prow[TL];
#pragma unroll
            for (int tl = 0; tl < TL; ++tl) {
                prow[tl] = __shfl_sync(0xffffffffu, mine, 8 * (t0 + tl) + (lane >> 2));
                here |= __reduce_or_sync(0xffffffffu, 1u << prow[tl]);
            }
            for (int p = 0; p < P; ++p) {
                if (!((here >> p) & 1u)) continue;
                const BankGrid &g = c_bgrid[p];
                const double *frag = sFrag + (size_t)p * G * fragstride;
                unsigned take = 0;
#pragma unroll
                for (int tl = 0; tl < TL; ++tl) take |= (prow[tl] == (unsigned)p ? 1u : 0u) << tl;
                if constexpr (NFIX > 0)
                    bl3_tile_nm<(NFIX + 7) / 8, (NFIX + 3) / 4, TL>(NFIX, NFIX, fragstride, frag, o0, no, t0, sA, sB,
                                                                    sC, sOut, lane, take);
                else if (g.n[0] > 8)
                    bl3_tile_n<2, TL>(g.n[1], g.n[2], fragstride, frag, o0, no, t0, sA, sB, sC, sOut, lane, take);
                else
                    bl3_tile_n<1, TL>(g.n[1], g.n[2], fragstride, frag, o0, no, t0, sA, sB, sC, sOut, lane, take);
            }
        }
        __syncwarp();
        for (int k = 0; k < no; ++k) {
            const double v = sOut[k * 32 + lane] * inv;
            if (live) o[o0 + k] = v;
        }
        __syncwarp();
    }
}

// Fragment image of one n0 x n1 x n2 tensor for the joint-K 3-D kernel, M = ceil(n2 / 4):
//   f[((j * M + m) * 2 + nt) * 32 + lane] = T[i = 8 nt + lane / 4][j][k = 4 m + lane % 4]   (0 outside)
static void bl3_make_frag(const double *t, int n0, int n1, int n2, double *f) {
    const int M = (n2 + 3) / 4;
    for (int j = 0; j < n1; ++j)
        for (int m = 0; m < M; ++m)
            for (int nt = 0; nt < 2; ++nt)
                for (int lane = 0; lane < 32; ++lane) {
                    const int i = 8 * nt + lane / 4, k = 4 * m + lane % 4;
                    f[((j * M + m) * 2 + nt) * 32 + lane] =
                        (i < n0 && k < n2) ? t[((size_t)i * n1 + j) * n2 + k] : 0.0;
                }
}

// Fragment image of one n0 x n1 tensor: f[(kb * 2 + nt) * 32 + lane] = T[4 kb + lane % 4][8 nt + lane / 4]
static void bl_make_frag(const double *t, int n0, int n1, double *f) {
    for (int kb = 0; kb < 4; ++kb)
        for (int nt = 0; nt < 2; ++nt)
            for (int lane = 0; lane < 32; ++lane) {
                const int i = 4 * kb + lane % 4, j = 8 * nt + lane / 4;
                f[(kb * 2 + nt) * 32 + lane] = (i < n0 && j < n1) ? t[(size_t)i * n1 + j] : 0.0;
            }
}

// Bank image of a set of grids: [tensors (already interleaved, GridDesc.tensor_off) | nodes |
// weights].  Returns false when the plan does not qualify for the bank path.
static bool bank_build(const std::vector<GridDesc> &desc, const std::vector<double> &il,
                       long long tensor_total, const double *nodes_cat, const double *weights_cat,
                       long long node_total, const int *dims_cat, const int *outputs, int smem_optin,
                       std::vector<double> *h_bank, std::vector<BankGrid> *h_desc, size_t *smem) {
    if (getenv("PCB_NO_BANK")) return false;
    if (desc.size() > (size_t)BANK_GRIDS) return false;
    int rows = 0;
    for (const GridDesc &gd : desc) {
        if (gd.D > BANK_MAXD) return false;
        for (int d = 0; d < gd.D; ++d)
            if (gd.n[d] > GRID_NL) return false;
        rows = std::max(rows, gd.sum_n - gd.n[gd.D - 1]);
    }
    *smem = (size_t)std::max(rows, 1) * BANK_THREADS * sizeof(double);
    if (*smem + 1024 > (size_t)smem_optin) return false;
    long long dims_total = 0;
    for (const GridDesc &gd : desc) dims_total += gd.D;
    if (tensor_total + 2 * node_total + 2 * dims_total > BANK_DOUBLES) return false;
    h_bank->assign(il.begin(), il.begin() + tensor_total);
    h_bank->insert(h_bank->end(), nodes_cat, nodes_cat + node_total);
    h_bank->insert(h_bank->end(), weights_cat, weights_cat + node_total);
    h_bank->resize(h_bank->size() + 2 * dims_total);
    double *bank_nodes = h_bank->data() + tensor_total;
    double *bank_weights = bank_nodes + node_total;
    double *bank_scale = bank_weights + node_total;
    h_desc->clear();
    int dpos = 0;
    for (const GridDesc &gd : desc) {
        BankGrid g;
        memset(&g, 0, sizeof(g));
        g.D = gd.D;
        for (int d = 0; d < gd.D; ++d) g.n[d] = gd.n[d];
        g.size = (int)gd.size;
        g.tensor_off = (int)gd.tensor_off;
        g.node_off = (int)tensor_total + gd.node_off;
        g.weight_off = (int)(tensor_total + node_total) + gd.node_off;
        g.scale_off = (int)(tensor_total + 2 * node_total) + 2 * dpos;
        int off = gd.node_off;
        for (int d = 0; d < gd.D; ++d) {
            g.dims[d] = dims_cat ? dims_cat[dpos + d] : d;
            // power-of-two prescale (exact): node span -> [2, 4), max |weight| -> [0.5, 1)
            double lo = bank_nodes[off], hi = bank_nodes[off], wmax = 0.0;
            for (int i = 0; i < gd.n[d]; ++i) {
                lo = std::min(lo, bank_nodes[off + i]);
                hi = std::max(hi, bank_nodes[off + i]);
                wmax = std::max(wmax, std::fabs(bank_weights[off + i]));
            }
            int e = 0;
            double sc = 1.0;
            if (hi > lo && std::isfinite(hi - lo)) {
                std::frexp(hi - lo, &e);  // hi - lo = f * 2^e, f in [0.5, 1)
                sc = std::ldexp(1.0, 2 - e);
            }
            int ew = 0;
            if (wmax > 0.0 && std::isfinite(wmax)) std::frexp(wmax, &ew);
            for (int i = 0; i < gd.n[d]; ++i) {
                bank_nodes[off + i] *= sc;
                bank_weights[off + i] = std::ldexp(bank_weights[off + i], -ew);
            }
            bank_scale[2 * dpos + d] = sc;
            bank_scale[2 * dpos + gd.D + d] = NODE_EPS * sc;
            off += gd.n[d];
        }
        dpos += gd.D;
        g.outputs = outputs ? outputs[h_desc->size()] : 1;
        h_desc->push_back(g);
    }
    return true;
}

static int bank_launch(const PlanBase *pl, uint64_t plan_id, const std::vector<double> &h_bank,
                       const std::vector<BankGrid> &h_desc, const void *kernel, void **args,
                       size_t smem, int64_t N, cudaStream_t st, int threads = BANK_THREADS,
                       int per_cta = 0) {  // queries per CTA when not one per thread
    PCB_CUDA(allow_dynamic_smem(kernel, smem, pl->smem_optin));
    if (per_cta <= 0) per_cta = threads;
    const int64_t blocks = (N + per_cta - 1) / per_cta;
    PCB_REQUIRE(blocks <= 0x7fffffffLL, "batch too large for one launch");
    const int rc = g_grid_bank.acquire(pl->dev, plan_id, st, [&](cudaStream_t s) {
        cudaError_t e = cudaMemcpyToSymbolAsync(c_grid, h_bank.data(), h_bank.size() * sizeof(double), 0,
                                                cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) return e;
        return cudaMemcpyToSymbolAsync(c_bgrid, h_desc.data(), h_desc.size() * sizeof(BankGrid), 0,
                                       cudaMemcpyHostToDevice, s);
    });
    if (rc) return rc;
    const cudaError_t e = cudaLaunchKernel(kernel, dim3((unsigned)blocks), dim3(threads), args, smem, st);
    g_grid_bank.release(pl->dev, st);
    g_launches.fetch_add(1);
    PCB_CUDA(e);
    PCB_CUDA(cudaGetLastError());
    return PCB_OK;
}

template <typename T>
static bool upload(T **dptr, const T *src, size_t count) {
    if (count == 0) count = 1;
    if (cudaMalloc(dptr, count * sizeof(T)) != cudaSuccess) return false;
    return !src || cudaMemcpy(*dptr, src, count * sizeof(T), cudaMemcpyHostToDevice) == cudaSuccess;
}

static int grid_launch_dims(const PlanBase *pl, const void *kernel, int threads, size_t smem,
                            int64_t N, int *grid) {
    PCB_CUDA(allow_dynamic_smem(kernel, smem, pl->smem_optin));
    int per_sm = 0;
    PCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    const int64_t want = (N + threads - 1) / threads;
    const int64_t cap = (int64_t)pl->sm_count * (per_sm > 0 ? per_sm : 1);
    *grid = (int)(want < cap ? want : cap);
    return PCB_OK;
}

bool spline_plan_uses_bank(const void *plan) {
    const SplinePlan *pl = static_cast<const SplinePlan *>(plan);
    return pl && pl->kind == PLAN_SPLINE && pl->bank_ok;
}

SplinePlan::~SplinePlan() {
    for (const BankPart &part : parts) g_grid_bank.forget(dev, part.id);
    free_all();
}
SliderPlan::~SliderPlan() {
    g_grid_bank.forget(dev, plan_id);
    free_all();
}

}  // namespace pcb

using namespace pcb;

extern "C" PCB_API int pcb_spline_plan_create(int dev, int D, const int32_t *num_knots,
                                              const double *knots_cat, int P, const int32_t *piece_n,
                                              const double *piece_nodes_cat,
                                              const double *piece_weights_cat, int G,
                                              const double *const *piece_tensors_host, void **plan) {
    PCB_REQUIRE(plan && num_knots && piece_n && piece_nodes_cat && piece_weights_cat &&
                    piece_tensors_host, "null argument");
    PCB_REQUIRE(D >= 1 && D <= GRID_MAXD, "num_dimensions %d outside [1, %d]", D, GRID_MAXD);
    PCB_REQUIRE(G >= 1 && G <= 64, "number of derivative tensors %d outside [1, 64]", G);
    long long expect = 1;
    int total_knots = 0;
    std::vector<int> meta(2 * D);
    for (int d = 0; d < D; ++d) {
        PCB_REQUIRE(num_knots[d] >= 0, "negative knot count");
        meta[d] = num_knots[d];
        meta[D + d] = total_knots;
        total_knots += num_knots[d];
        expect *= num_knots[d] + 1;
    }
    PCB_REQUIRE(expect == P, "num_pieces=%d does not match prod(num_knots+1)=%lld", P, expect);
    PCB_REQUIRE(total_knots == 0 || knots_cat, "null knots");

    SplinePlan *pl = new SplinePlan();
    pl->kind = PLAN_SPLINE;
    pl->dev = dev;
    pl->D = D;
    pl->P = P;
    pl->G = G;
    pl->GB = grid_pick_gb(G);
    int cc = 0;
    if (int rc = device_props(dev, &pl->sm_count, &pl->smem_optin, &cc)) {
        delete pl;
        return rc;
    }
    const int nblk = (G + pl->GB - 1) / pl->GB;
    std::vector<GridDesc> desc(P);
    long long node_total = 0, tensor_total = 0;
    for (int p = 0; p < P; ++p) {
        GridDesc &gd = desc[p];
        memset(&gd, 0, sizeof(gd));
        gd.D = D;
        gd.size = 1;
        gd.node_off = (int)node_total;
        gd.tensor_off = tensor_total;
        for (int d = 0; d < D; ++d) {
            const int nn = piece_n[(size_t)p * D + d];
            if (nn < 1) {
                delete pl;
                return fail(PCB_EINVAL, "piece %d: n_nodes[%d] must be >= 1", p, d);
            }
            gd.n[d] = nn;
            gd.sum_n += nn;
            gd.size *= nn;
        }
        node_total += gd.sum_n;
        tensor_total += gd.size * nblk * pl->GB;
        if (gd.sum_n > pl->max_sum_n) pl->max_sum_n = gd.sum_n;
    }
    std::vector<double> il((size_t)tensor_total);
    for (int p = 0; p < P; ++p)
        grid_interleave(piece_tensors_host + (size_t)p * G, G, pl->GB, desc[p].size,
                        il.data() + desc[p].tensor_off);
    // Bank path (pieces of different shapes are fine; P <= 32 for the kernel's presence mask): all G
    // outputs in one image if they fit, else parts of 4 / 2 / 1 consecutive outputs.
    for (int per : {G, 4, 2, 1}) {
        if (per > G || (per != G && per >= G)) continue;
        std::vector<SplinePlan::BankPart> parts;
        bool fits = true;
        for (int g0 = 0; g0 < G && fits; g0 += per) {
            SplinePlan::BankPart part;
            part.g0 = g0;
            part.G = std::min(per, G - g0);
            part.GB = grid_pick_gb(part.G);
            const int pblk = (part.G + part.GB - 1) / part.GB;
            std::vector<GridDesc> pdesc(desc);
            long long ptotal = 0;
            for (int p = 0; p < P; ++p) {
                pdesc[p].tensor_off = ptotal;
                ptotal += pdesc[p].size * pblk * part.GB;
            }
            std::vector<double> pil((size_t)ptotal);
            for (int p = 0; p < P; ++p)
                grid_interleave(piece_tensors_host + (size_t)p * G + g0, part.G, part.GB, pdesc[p].size,
                                pil.data() + pdesc[p].tensor_off);
            size_t smem_part = 0;
            fits = bank_build(pdesc, pil, ptotal, piece_nodes_cat, piece_weights_cat, node_total, nullptr,
                              nullptr, pl->smem_optin, &part.h_bank, &part.h_desc, &smem_part);
            if (fits) {
                part.id = next_plan_id();
                pl->bank_smem = std::max(pl->bank_smem, smem_part);
                parts.push_back(std::move(part));
            }
        }
        if (fits) {
            pl->parts = std::move(parts);
            pl->bank_ok = true;
            break;
        }
    }
    DeviceGuard guard(dev);
    bool ok = guard.ok && upload(&pl->d_num_knots, meta.data(), meta.size()) &&
              upload(&pl->d_knots, knots_cat, (size_t)total_knots) &&
              upload(&pl->d_desc, desc.data(), desc.size()) &&
              upload<double>(&pl->d_nodes, nullptr, (size_t)(2 * node_total)) &&
              upload(&pl->d_tensors, il.data(), il.size());
    if (ok) {
        pl->d_knot_off = pl->d_num_knots + D;
        pl->d_weights = pl->d_nodes + node_total;
        ok = cudaMemcpy(pl->d_nodes, piece_nodes_cat, node_total * 8, cudaMemcpyHostToDevice) == cudaSuccess &&
             cudaMemcpy(pl->d_weights, piece_weights_cat, node_total * 8, cudaMemcpyHostToDevice) == cudaSuccess;
    }
    if (!ok) {
        delete pl;
        return fail(PCB_ECUDA, "device allocation/upload for the spline plan failed");
    }
    // 2-D pieces of at most 16 x 16 nodes: tensor-core path (nodes / weights from the bank image)
    if (D == 2 && pl->bank_ok && P <= BANK_GRIDS && P * G <= BL_MAX_FRAGS && !getenv("PCB_NO_DMMA2D")) {
        bool fits = true;
        for (int p = 0; p < P; ++p) fits = fits && desc[p].n[0] <= 16 && desc[p].n[1] <= 16;
        if (fits) {
            std::vector<double> fr((size_t)P * G * BL_FRAG);
            for (int p = 0; p < P; ++p)
                for (int g = 0; g < G; ++g)
                    bl_make_frag(piece_tensors_host[(size_t)p * G + g], desc[p].n[0], desc[p].n[1],
                                 fr.data() + ((size_t)p * G + g) * BL_FRAG);
            pl->dmma2d_ok = upload(&pl->d_frags, fr.data(), fr.size());
            if (!pl->dmma2d_ok) cudaGetLastError();
        }
    }
    pl->nfix = desc[0].n[0];
    for (int p = 0; p < P; ++p)
        for (int d = 0; d < D; ++d)
            if (desc[p].n[d] != pl->nfix) pl->nfix = 0;
    // 3-D pieces of at most 16 nodes per dim: joint-K tensor-core path
    if (D == 3 && pl->bank_ok && P <= BANK_GRIDS && !getenv("PCB_NO_DMMA3D")) {
        bool fits = true;
        int kbmax = 1;
        for (int p = 0; p < P; ++p) {
            fits = fits && desc[p].n[0] <= 16 && desc[p].n[1] <= 16 && desc[p].n[2] <= 16;
            kbmax = std::max(kbmax, desc[p].n[1] * ((desc[p].n[2] + 3) / 4));
        }
        const size_t fragstride = (size_t)kbmax * 64;
        // warps per CTA and outputs per launch: 12 warps and as many outputs as fit next to their
        // tiles, else 8 warps.  Measured on 15^3: 12 warps 2.35e9 values/s against 1.93e9 with 8; for
        // value + delta two 12-warp launches (1.18e9 q/s) beat one 8-warp launch of both (1.14e9).
        int per = 0, warps = 0;
        const int force_warps = getenv("PCB_BL3_WARPS") ? atoi(getenv("PCB_BL3_WARPS")) : 0;
        for (int w : {12, 8})
            for (int gl = G; gl >= 1 && !per; --gl) {
                if (force_warps && w != force_warps) continue;
                const size_t bytes = (size_t)P * gl * fragstride * 8 + (size_t)w * BL3_WARP_DOUBLES * 8;
                if (bytes + 2048 <= (size_t)pl->smem_optin) {
                    per = gl;
                    warps = w;
                    break;
                }
            }
        if (fits && kbmax <= BL3_MAX_KB && per >= 1) {
            pl->g3_per = per;
            pl->g3_warps = warps;
            std::vector<double> fr((size_t)P * G * fragstride, 0.0);
            for (int p = 0; p < P; ++p)
                for (int g = 0; g < G; ++g)
                    bl3_make_frag(piece_tensors_host[(size_t)p * G + g], desc[p].n[0], desc[p].n[1],
                                  desc[p].n[2], fr.data() + ((size_t)p * G + g) * fragstride);
            pl->kb3max = kbmax;
            pl->dmma3d_ok = upload(&pl->d_frags, fr.data(), fr.size());
            if (!pl->dmma3d_ok) cudaGetLastError();
        }
    }
    *plan = pl;
    return PCB_OK;
}

extern "C" PCB_API int pcb_spline_lookup(void *plan, const double *d_points, int64_t N,
                                         int32_t *d_piece, void *stream) {
    SplinePlan *pl = static_cast<SplinePlan *>(plan);
    PCB_REQUIRE(pl && pl->kind == PLAN_SPLINE, "not a spline plan");
    PCB_REQUIRE(N >= 0, "negative N");
    if (N == 0) return PCB_OK;
    PCB_REQUIRE(d_points && d_piece, "null device pointer");
    DeviceGuard guard(pl->dev);
    const int64_t want = (N + 256 * LOOKUP_ILP - 1) / (256 * LOOKUP_ILP);
    const int64_t cap = (int64_t)pl->sm_count * 8;
    spline_lookup_kernel<<<(int)(want < cap ? want : cap), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        pl->D, pl->d_num_knots, pl->d_knot_off, pl->d_knots, d_points, N, d_piece,
        (reinterpret_cast<uintptr_t>(d_points) & 15) == 0 ? 1 : 0);
    g_launches.fetch_add(1);
    PCB_CUDA(cudaGetLastError());
    return PCB_OK;
}

extern "C" PCB_API int pcb_spline_eval(void *plan, const double *d_points, int64_t N, double *d_out,
                                       int32_t *d_piece, void *stream) {
    SplinePlan *pl = static_cast<SplinePlan *>(plan);
    PCB_REQUIRE(pl && pl->kind == PLAN_SPLINE, "not a spline plan");
    PCB_REQUIRE(N >= 0, "negative N");
    if (N == 0) return PCB_OK;
    PCB_REQUIRE(d_points && d_out, "null device pointer");
    DeviceGuard guard(pl->dev);
    const size_t smem = (size_t)pl->max_sum_n * PW_THREADS * sizeof(double);
    if (!pl->bank_ok && smem > (size_t)pl->smem_optin)
        return fail(PCB_EUNSUPPORTED, "weight rows of %d nodes do not fit in shared memory", pl->max_sum_n);
    if (pl->dmma2d_ok) {
        const SplinePlan::BankPart &part = pl->parts[0];  // nodes / weights / scales are in every image
        const size_t dsm = ((size_t)pl->P * pl->G * BL_FRAG + (BL_THREADS / 32) * BL_WARP_DOUBLES) * sizeof(double);
        void *dargs[] = {(void *)&pl->G, (void *)&pl->P, (void *)&pl->d_num_knots, (void *)&pl->d_knot_off,
                         (void *)&pl->d_knots, (void *)&pl->d_frags, (void *)&d_points, (void *)&N,
                         (void *)&d_out, (void *)&d_piece};
        const int nf = getenv("PCB_NO_NFIX") ? 0 : pl->nfix;  // every dim of every piece has nf nodes
        const void *k2 = PCB_NFIX_KERNEL(nf, spline2d_dmma_kernel);
        return bank_launch(pl, part.id, part.h_bank, part.h_desc, k2, dargs, dsm, N,
                           static_cast<cudaStream_t>(stream), BL_THREADS);
    }
    if (pl->dmma3d_ok) {
        const SplinePlan::BankPart &part = pl->parts[0];
        const size_t fragstride = (size_t)pl->kb3max * 64;
        for (int g0 = 0; g0 < pl->G; g0 += pl->g3_per) {
            int gl = std::min(pl->g3_per, pl->G - g0);
            const size_t dsm = ((size_t)pl->P * gl * fragstride +
                                (size_t)pl->g3_warps * BL3_WARP_DOUBLES) * sizeof(double);
            int32_t *piece = g0 == 0 ? d_piece : nullptr;
            // few pieces: 8 (P - 1) free slots per CTA let every piece's run start on a tile boundary
            const int threads = 32 * pl->g3_warps;
            const int slack = 8 * (pl->P - 1);
            int qper = slack * 8 <= threads && !getenv("PCB_BL3_UNALIGNED") ? threads - slack : threads;
            void *dargs[] = {(void *)&gl, (void *)&g0, (void *)&pl->G, (void *)&pl->P, (void *)&pl->kb3max,
                             (void *)&pl->d_num_knots, (void *)&pl->d_knot_off, (void *)&pl->d_knots,
                             (void *)&pl->d_frags, (void *)&d_points, (void *)&N,
                             (void *)&d_out, (void *)&piece, (void *)&qper};
            // Fixed-node variants.  Uncapped, ptxas reads their bank operands per lane (LDC); a
            // 56-register budget (minBlocks 3) brings the uniform datapath back but spills.  Measured
            // on 15^3: generic 2.71e9 q/s, fixed + LDC 2.81e9, fixed + capped 2.64e9 -- the straight-line
            // rows matter more than the operand path, so the variants stay uncapped.
            const int nf3 = getenv("PCB_NO_NFIX") ? 0 : pl->nfix;
            const void *k3 = nf3 == 8    ? (const void *)spline3d_dmma_kernel<1, 0, 8>
                             : nf3 == 11 ? (const void *)spline3d_dmma_kernel<1, 0, 11>
                             : nf3 == 12 ? (const void *)spline3d_dmma_kernel<1, 0, 12>
                             : nf3 == 15 ? (const void *)spline3d_dmma_kernel<1, 0, 15>
                             : nf3 == 16 ? (const void *)spline3d_dmma_kernel<1, 0, 16>
                                         : (const void *)spline3d_dmma_kernel<1, 0, 0>;
            if (int rc = bank_launch(pl, part.id, part.h_bank, part.h_desc, k3, dargs, dsm, N,
                                     static_cast<cudaStream_t>(stream), threads, qper))
                return rc;
        }
        return PCB_OK;
    }
    if (pl->bank_ok) {
        for (const SplinePlan::BankPart &part : pl->parts) {
            const void *uk = BANK_KERNEL_TABLE(spline_bank_kernel, part.GB, grid_pick_dm(pl->D));
            int32_t *piece = part.g0 == 0 ? d_piece : nullptr;
            void *uargs[] = {(void *)&pl->D, (void *)&part.G, (void *)&part.g0, (void *)&pl->G, (void *)&pl->P,
                             (void *)&pl->d_num_knots, (void *)&pl->d_knot_off, (void *)&pl->d_knots,
                             (void *)&d_points, (void *)&N, (void *)&d_out, (void *)&piece};
            if (int rc = bank_launch(pl, part.id, part.h_bank, part.h_desc, uk, uargs, pl->bank_smem, N,
                                     static_cast<cudaStream_t>(stream)))
                return rc;
        }
        return PCB_OK;
    }
    const void *kernel = GRID_KERNEL_TABLE(spline_eval_kernel, pl->GB, grid_pick_dm(pl->D));
    int grid = 0;
    if (int rc = grid_launch_dims(pl, kernel, PW_THREADS, smem, N, &grid)) return rc;
    void *args[] = {(void *)&pl->D, (void *)&pl->G, (void *)&pl->d_num_knots, (void *)&pl->d_knot_off,
                    (void *)&pl->d_knots, (void *)&pl->d_desc, (void *)&pl->d_nodes,
                    (void *)&pl->d_weights, (void *)&pl->d_tensors, (void *)&d_points, (void *)&N,
                    (void *)&d_out, (void *)&d_piece};
    PCB_CUDA(cudaLaunchKernel(kernel, dim3(grid), dim3(PW_THREADS), args, smem,
                              static_cast<cudaStream_t>(stream)));
    g_launches.fetch_add(1);
    PCB_CUDA(cudaGetLastError());
    return PCB_OK;
}

extern "C" PCB_API int pcb_slider_plan_create(int dev, int D, int S, const int32_t *group_size,
                                              const int32_t *group_dims_cat, const int32_t *slide_n_cat,
                                              const double *slide_nodes_cat,
                                              const double *slide_weights_cat, double pivot_value, int G,
                                              const int32_t *out_slide,
                                              const double *const *slide_tensors_host, void **plan) {
    PCB_REQUIRE(plan && group_size && group_dims_cat && slide_n_cat && slide_nodes_cat &&
                    slide_weights_cat && out_slide && slide_tensors_host, "null argument");
    PCB_REQUIRE(D >= 1 && D <= 4096 && S >= 1 && S <= D, "invalid slider shape D=%d S=%d", D, S);
    PCB_REQUIRE(G >= 1 && G <= 64, "number of output rows %d outside [1, 64]", G);
    SliderPlan *pl = new SliderPlan();
    pl->kind = PLAN_SLIDER;
    pl->dev = dev;
    pl->D = D;
    pl->S = S;
    pl->G = G;
    pl->pivot = pivot_value;
    int cc = 0;
    if (int rc = device_props(dev, &pl->sm_count, &pl->smem_optin, &cc)) {
        delete pl;
        return rc;
    }
    std::vector<GridDesc> desc(S);
    std::vector<int> group_off(S + 1);
    int gpos = 0;
    long long node_total = 0;
    for (int s = 0; s < S; ++s) {
        GridDesc &gd = desc[s];
        memset(&gd, 0, sizeof(gd));
        const int gs = group_size[s];
        if (gs < 1 || gs > GRID_MAXD) {
            delete pl;
            return fail(PCB_EUNSUPPORTED, "slide %d has %d dims (supported: 1..%d)", s, gs, GRID_MAXD);
        }
        gd.D = gs;
        if (gs > pl->max_D) pl->max_D = gs;
        gd.size = 1;
        gd.node_off = (int)node_total;
        group_off[s] = gpos;
        for (int d = 0; d < gs; ++d) {
            const int dim = group_dims_cat[gpos + d];
            const int nn = slide_n_cat[gpos + d];
            if (dim < 0 || dim >= D || nn < 1) {
                delete pl;
                return fail(PCB_EINVAL, "slide %d: invalid dim/n_nodes", s);
            }
            gd.n[d] = nn;
            gd.sum_n += nn;
            gd.size *= nn;
        }
        gpos += gs;
        node_total += gd.sum_n;
        if (gd.sum_n > pl->max_sum_n) pl->max_sum_n = gd.sum_n;
    }
    group_off[S] = gpos;
    // per-slide output lists: output 0 = value tensor (when a value row exists), then one output
    // per derivative row owned by the slide
    std::vector<std::vector<const double *>> outs(S);
    std::vector<int> row_out(G, 0);
    int first_value_row = -1;
    for (int g = 0; g < G; ++g) {
        if (out_slide[g] < -2 || out_slide[g] >= S) {
            delete pl;
            return fail(PCB_EINVAL, "out_slide[%d]=%d invalid", g, out_slide[g]);
        }
        if (out_slide[g] == -1 && first_value_row < 0) first_value_row = g;
    }
    for (int s = 0; s < S; ++s)
        outs[s].push_back(first_value_row >= 0 ? slide_tensors_host[(size_t)first_value_row * S + s]
                                               : nullptr);
    for (int g = 0; g < G; ++g) {
        const int os = out_slide[g];
        if (os == -1) {
            for (int s = 0; s < S; ++s)
                if (!slide_tensors_host[(size_t)g * S + s]) {
                    delete pl;
                    return fail(PCB_EINVAL, "missing tensor for row %d slide %d", g, s);
                }
        } else if (os >= 0) {
            if (!slide_tensors_host[(size_t)g * S + os]) {
                delete pl;
                return fail(PCB_EINVAL, "missing tensor for row %d slide %d", g, os);
            }
            row_out[g] = (int)outs[os].size();
            outs[os].push_back(slide_tensors_host[(size_t)g * S + os]);
        }
    }
    std::vector<int> slide_G(S);
    int maxg = 1;
    for (int s = 0; s < S; ++s) {
        // a slide with only the (absent) value slot and no derivative rows has nothing to do
        slide_G[s] = (outs[s].size() == 1 && !outs[s][0]) ? 0 : (int)outs[s].size();
        if (slide_G[s] > maxg) maxg = slide_G[s];
    }
    pl->GB = grid_pick_gb(maxg);
    long long tensor_total = 0;
    for (int s = 0; s < S; ++s) {
        desc[s].tensor_off = tensor_total;
        tensor_total += desc[s].size * ((slide_G[s] + pl->GB - 1) / pl->GB) * pl->GB;
    }
    std::vector<double> il((size_t)(tensor_total > 0 ? tensor_total : 1));
    for (int s = 0; s < S; ++s)
        if (slide_G[s] > 0)
            grid_interleave(outs[s].data(), slide_G[s], pl->GB, desc[s].size,
                            il.data() + desc[s].tensor_off);
    pl->plan_id = next_plan_id();
    pl->bank_ok = bank_build(desc, il, tensor_total, slide_nodes_cat, slide_weights_cat, node_total,
                             group_dims_cat, slide_G.data(), pl->smem_optin, &pl->h_bank, &pl->h_desc, &pl->bank_smem);
    std::vector<int> ints(group_off);
    for (int i = 0; i < gpos; ++i) ints.push_back(group_dims_cat[i]);
    for (int g = 0; g < G; ++g) ints.push_back(out_slide[g]);
    for (int g = 0; g < G; ++g) ints.push_back(row_out[g]);
    for (int s = 0; s < S; ++s) ints.push_back(slide_G[s]);
    DeviceGuard guard(dev);
    bool ok = guard.ok && upload(&pl->d_desc, desc.data(), desc.size()) &&
              upload(&pl->d_ints, ints.data(), ints.size()) &&
              upload<double>(&pl->d_nodes, nullptr, (size_t)(2 * node_total)) &&
              upload(&pl->d_tensors, il.data(), il.size());
    if (ok) {
        pl->d_group_dims = pl->d_ints + S + 1;
        pl->d_out_slide = pl->d_group_dims + gpos;
        pl->d_row_out = pl->d_out_slide + G;
        pl->d_slide_G = pl->d_row_out + G;
        pl->d_weights = pl->d_nodes + node_total;
        ok = cudaMemcpy(pl->d_nodes, slide_nodes_cat, node_total * 8, cudaMemcpyHostToDevice) == cudaSuccess &&
             cudaMemcpy(pl->d_weights, slide_weights_cat, node_total * 8, cudaMemcpyHostToDevice) == cudaSuccess;
    }
    if (!ok) {
        delete pl;
        return fail(PCB_ECUDA, "device allocation/upload for the slider plan failed");
    }
    // all slides 2-D with at most 16 x 16 nodes: tensor-core path
    if (pl->bank_ok && S <= BANK_GRIDS && !getenv("PCB_NO_DMMA2D")) {
        bool fits = true;
        int nfrag = 0;
        std::vector<int> foff(S, 0);
        for (int sl = 0; sl < S; ++sl) {
            fits = fits && desc[sl].D == 2 && desc[sl].n[0] <= 16 && desc[sl].n[1] <= 16;
            if (sl == 0) pl->nfix = desc[0].n[0];
            if (desc[sl].n[0] != pl->nfix || desc[sl].n[1] != pl->nfix) pl->nfix = 0;
            foff[sl] = nfrag;
            nfrag += slide_G[sl];
        }
        if (fits && nfrag >= 1 && nfrag <= BL_MAX_FRAGS) {
            std::vector<double> fr((size_t)nfrag * BL_FRAG, 0.0);
            for (int sl = 0; sl < S; ++sl)
                for (int k = 0; k < slide_G[sl]; ++k)
                    if (outs[sl][k])
                        bl_make_frag(outs[sl][k], desc[sl].n[0], desc[sl].n[1],
                                     fr.data() + (size_t)(foff[sl] + k) * BL_FRAG);
            pl->nfrag = nfrag;
            pl->dmma2d_ok = upload(&pl->d_frags, fr.data(), fr.size()) &&
                            upload(&pl->d_frag_off, foff.data(), foff.size());
            if (!pl->dmma2d_ok) cudaGetLastError();
        }
    }
    *plan = pl;
    return PCB_OK;
}

extern "C" PCB_API int pcb_slider_eval(void *plan, const double *d_points, int64_t N, double *d_out,
                                       void *stream) {
    SliderPlan *pl = static_cast<SliderPlan *>(plan);
    PCB_REQUIRE(pl && pl->kind == PLAN_SLIDER, "not a slider plan");
    PCB_REQUIRE(N >= 0, "negative N");
    if (N == 0) return PCB_OK;
    PCB_REQUIRE(d_points && d_out, "null device pointer");
    DeviceGuard guard(pl->dev);
    const size_t smem = (size_t)pl->max_sum_n * PW_THREADS * sizeof(double);
    if (!pl->bank_ok && smem > (size_t)pl->smem_optin)
        return fail(PCB_EUNSUPPORTED, "weight rows of %d nodes do not fit in shared memory", pl->max_sum_n);
    if (pl->dmma2d_ok) {
        const size_t dsm = ((size_t)pl->nfrag * BL_FRAG + (BL_THREADS / 32) * BL_WARP_DOUBLES) * sizeof(double);
        void *dargs[] = {(void *)&pl->D, (void *)&pl->S, (void *)&pl->G, (void *)&pl->pivot,
                         (void *)&pl->d_out_slide, (void *)&pl->d_row_out, (void *)&pl->d_frag_off,
                         (void *)&pl->nfrag, (void *)&pl->d_frags, (void *)&d_points, (void *)&N,
                         (void *)&d_out};
        const int nf = getenv("PCB_NO_NFIX") ? 0 : pl->nfix;  // every dim of every slide has nf nodes
        return bank_launch(pl, pl->plan_id, pl->h_bank, pl->h_desc, PCB_NFIX_KERNEL(nf, slider2d_dmma_kernel),
                           dargs, dsm, N, static_cast<cudaStream_t>(stream), BL_THREADS);
    }
    if (pl->bank_ok) {
        const void *uk = BANK_KERNEL_TABLE(slider_bank_kernel, pl->GB, grid_pick_dm(pl->max_D));
        void *uargs[] = {(void *)&pl->D, (void *)&pl->S, (void *)&pl->G, (void *)&pl->pivot,
                         (void *)&pl->d_out_slide, (void *)&pl->d_row_out, (void *)&d_points, (void *)&N,
                         (void *)&d_out};
        return bank_launch(pl, pl->plan_id, pl->h_bank, pl->h_desc, uk, uargs, pl->bank_smem, N,
                           static_cast<cudaStream_t>(stream));
    }
    const void *kernel = GRID_KERNEL_TABLE(slider_eval_kernel, pl->GB, grid_pick_dm(pl->max_D));
    int grid = 0;
    if (int rc = grid_launch_dims(pl, kernel, PW_THREADS, smem, N, &grid)) return rc;
    void *args[] = {(void *)&pl->D, (void *)&pl->S, (void *)&pl->G, (void *)&pl->pivot,
                    (void *)&pl->d_desc, (void *)&pl->d_ints, (void *)&pl->d_group_dims,
                    (void *)&pl->d_out_slide, (void *)&pl->d_row_out, (void *)&pl->d_slide_G,
                    (void *)&pl->d_nodes, (void *)&pl->d_weights, (void *)&pl->d_tensors,
                    (void *)&d_points, (void *)&N, (void *)&d_out};
    PCB_CUDA(cudaLaunchKernel(kernel, dim3(grid), dim3(PW_THREADS), args, smem,
                              static_cast<cudaStream_t>(stream)));
    g_launches.fetch_add(1);
    PCB_CUDA(cudaGetLastError());
    return PCB_OK;
}
