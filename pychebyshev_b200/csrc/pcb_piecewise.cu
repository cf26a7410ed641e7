// ChebyshevSpline (piece lookup + per-piece evaluation) and ChebyshevSlider (additive slides).
//
// Spline: replaces ChebyshevSpline.eval_batch (reference spline.py:633-700).  The routing
//   (spline.py:677-690) is integer work and bit-exact; the per-piece evaluation is the
//   thread-per-query barycentric evaluator of pcb_grid.cuh on the piece's own nodes/weights.
//   The reference groups points by piece and calls the piece evaluator per group; on the device
//   every query simply indexes its piece's descriptor, so no bucketing pass is needed.
// Slider: replaces a loop of ChebyshevSlider.eval (reference slider.py:247-318).
#include <cstdlib>

#include "pcb_cbank.cuh"
#include "pcb_grid.cuh"

namespace pcb {

constexpr int PW_THREADS = 128;

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
    // uniform-datapath path: all piece tensors + descriptors in this module's constant bank
    bool bank_ok = false;
    uint64_t plan_id = 0;
    std::vector<double> h_bank;
    std::vector<GridDesc> h_desc;
    ~SplinePlan() override;
    void free_all() {
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
    std::vector<GridDesc> h_desc;
    ~SliderPlan() override;
    void free_all() {
        if (d_desc) cudaFree(d_desc);
        if (d_ints) cudaFree(d_ints);
        if (d_nodes) cudaFree(d_nodes);
        if (d_tensors) cudaFree(d_tensors);
    }
};

__global__ void __launch_bounds__(256)
spline_lookup_kernel(int D, const int *__restrict__ num_knots, const int *__restrict__ knot_off,
                     const double *__restrict__ knots, const double *__restrict__ pts, int64_t N,
                     int32_t *__restrict__ piece) {
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < N;
         q += (int64_t)gridDim.x * blockDim.x)
        piece[q] = spline_piece_index(D, num_knots, knot_off, knots, pts + q * D);
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
// Uniform-datapath variants (small plans): every piece / slide tensor lives in this module's
// constant bank and is read with warp-uniform `LDCU` feeding `DFMA ... UR` directly, so the tensor
// costs no LSU slot and no vector register.  Control flow is kept warp-uniform: the spline kernel
// loops over the pieces PRESENT in the warp (vote) and every lane contracts that piece's tensor
// with its own weights; the lanes that belong to the piece keep the result (select, no branch).
// One query per thread and no grid-stride loop (ptxas needs that to prove uniformity).
// ---------------------------------------------------------------------------------------------
constexpr int BANK_DOUBLES = 7680;  // 60 KB of tensors + 64 descriptors of 72 B in the 64 KB bank
constexpr int BANK_GRIDS = 32;  // also the width of the spline kernel's piece-presence mask
__constant__ double c_grid[BANK_DOUBLES];
__constant__ GridDesc c_gdesc[BANK_GRIDS];
static ConstBank g_grid_bank;

template <int LEVEL, int D, int GB>
struct GridContractU {
    __device__ __forceinline__ static void run(int base, const GridDesc &gd, const int (&stride)[GRID_MAXD],
                                               const double *ws, int wstride,
                                               const double (&wl)[GRID_NL], double (&out)[GB]) {
        const int nl = gd.n[LEVEL];
        if constexpr (LEVEL == D - 1) {
            double part[4][GB];
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int j = 0; j < GB; ++j) part[c][j] = 0.0;
#pragma unroll
            for (int i = 0; i < GRID_NL; ++i) {
                if (i < nl) {
#pragma unroll
                    for (int j = 0; j < GB; ++j)
                        part[i & 3][j] = fma(wl[i], c_grid[base + i * GB + j], part[i & 3][j]);
                }
            }
#pragma unroll
            for (int j = 0; j < GB; ++j) out[j] = (part[0][j] + part[1][j]) + (part[2][j] + part[3][j]);
        } else {
#pragma unroll
            for (int j = 0; j < GB; ++j) out[j] = 0.0;
            const int st = stride[LEVEL] * GB;
            const double *wnext = ws + (size_t)nl * wstride;
            for (int i = 0; i < nl; ++i) {
                double sub[GB];
                GridContractU<(LEVEL + 1 < D ? LEVEL + 1 : LEVEL), D, GB>::run(base + i * st, gd, stride,
                                                                              wnext, wstride, wl, sub);
                const double w = ws[i * wstride];
#pragma unroll
                for (int j = 0; j < GB; ++j) out[j] = fma(w, sub[j], out[j]);
            }
        }
    }
};

#define GRIDU_CASE(K)                                                                   \
    case K:                                                                             \
        if constexpr (K <= DM) GridContractU<0, K, GB>::run(base, gd, stride, ws, wstride, wl, out); \
        break;

// Contract output block `b` of bank-resident grid `gd` (a reference INTO c_gdesc: uniform).
template <int GB, int DM>
__device__ __forceinline__ void grid_contract_u(const GridDesc &gd, int b, const double *ws,
                                                int wstride, const double (&wl)[GRID_NL],
                                                double (&out)[GB]) {
    int stride[GRID_MAXD];
    int s = 1;
    for (int d = gd.D - 1; d >= 0; --d) {
        stride[d] = s;
        s *= gd.n[d];
    }
    const int base = (int)gd.tensor_off + b * (int)gd.size * GB;
#pragma unroll
    for (int j = 0; j < GB; ++j) out[j] = 0.0;
    switch (gd.D) {
        GRIDU_CASE(1) GRIDU_CASE(2) GRIDU_CASE(3) GRIDU_CASE(4) GRIDU_CASE(5) GRIDU_CASE(6)
        GRIDU_CASE(7) GRIDU_CASE(8)
    }
}

template <int GB, int DM>
__global__ void __launch_bounds__(PW_THREADS)
spline_uniform_kernel(int D, int G, int P, const int *__restrict__ num_knots,
                      const int *__restrict__ knot_off, const double *__restrict__ knots,
                      const double *__restrict__ nodes, const double *__restrict__ weights,
                      const double *__restrict__ pts, int64_t N, double *__restrict__ out,
                      int32_t *__restrict__ piece_out) {
    extern __shared__ __align__(16) double smem[];
    double *ws = smem + threadIdx.x;
    const int stride = blockDim.x;
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const double *x = pts + (q < N ? q : N - 1) * D;  // tail lanes recompute the last query
    const int mine = spline_piece_index(D, num_knots, knot_off, knots, x);
    if (piece_out && q < N) piece_out[q] = mine;
    double wl[GRID_NL];
    {
        // The bank path requires every piece to have the same node counts (the reference's splines
        // do: spline.py n_nodes is per dimension), so trip counts come from piece 0 -- uniform --
        // and only the node / weight OFFSET is per lane.  Per-lane trip counts (or doing this under
        // `if (mine == p)`) make ptxas abandon the uniform datapath for the whole kernel.
        const long long shift = c_gdesc[mine].node_off - c_gdesc[0].node_off;
        grid_weights(c_gdesc[0], nodes + shift, weights + shift, [&](int d) { return __ldg(x + d); }, ws,
                     stride, wl);
    }
    // pieces present in this warp: REDUX leaves the mask in a uniform register, so the skip below
    // is a uniform branch (P <= 32 on this path)
    const unsigned present = __reduce_or_sync(0xffffffffu, 1u << mine);
    for (int b = 0; b * GB < G; ++b) {
        double r[GB];
#pragma unroll
        for (int j = 0; j < GB; ++j) r[j] = 0.0;
        for (int p = 0; p < P; ++p) {
            if (!((present >> p) & 1u)) continue;
            double sub[GB];
            grid_contract_u<GB, DM>(c_gdesc[p], b, ws, stride, wl, sub);
#pragma unroll
            for (int j = 0; j < GB; ++j) r[j] = mine == p ? sub[j] : r[j];
        }
        if (q < N) {
#pragma unroll
            for (int j = 0; j < GB; ++j)
                if (b * GB + j < G) out[q * G + b * GB + j] = r[j];
        }
    }
}

template <int GB, int DM>
__global__ void __launch_bounds__(PW_THREADS)
slider_uniform_kernel(int D, int S, int G, double pivot, const int *__restrict__ group_off,
                      const int *__restrict__ group_dims, const int *__restrict__ out_slide,
                      const int *__restrict__ row_out, const int *__restrict__ slide_G,
                      const double *__restrict__ nodes, const double *__restrict__ weights,
                      const double *__restrict__ pts, int64_t N, double *__restrict__ out) {
    extern __shared__ __align__(16) double smem[];
    double *ws = smem + threadIdx.x;
    const int stride = blockDim.x;
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t qc = q < N ? q : N - 1;
    const double *x = pts + qc * D;
    double *o = out + qc * G;  // tail lanes redo (and rewrite) the last query: same values
    for (int g = 0; g < G; ++g) o[g] = out_slide[g] == -1 ? pivot : 0.0;
    for (int s = 0; s < S; ++s) {
        const int sg = slide_G[s];
        if (sg == 0) continue;
        const GridDesc &gd = c_gdesc[s];
        const int *dims = group_dims + group_off[s];
        double wl[GRID_NL];
        grid_weights(gd, nodes, weights, [&](int d) { return __ldg(x + dims[d]); }, ws, stride, wl);
        for (int b = 0; b * GB < sg; ++b) {
            double r[GB];
            grid_contract_u<GB, DM>(gd, b, ws, stride, wl, r);
#pragma unroll
            for (int j = 0; j < GB; ++j) {
                const int so = b * GB + j;
                if (so >= sg) continue;
                for (int g = 0; g < G; ++g) {
                    if (out_slide[g] == s && row_out[g] == so)
                        o[g] = r[j];
                    else if (out_slide[g] == -1 && so == 0 && row_out[g] == 0)
                        o[g] = o[g] + (r[j] - pivot);
                }
            }
        }
    }
}

// Can these grids live in the bank?  (tensors incl. all output blocks, descriptors, n_last)
static bool bank_fits(const std::vector<GridDesc> &desc, long long tensor_total) {
    if (tensor_total > BANK_DOUBLES || desc.size() > (size_t)BANK_GRIDS) return false;
    for (const GridDesc &gd : desc)
        if (gd.n[gd.D - 1] > GRID_NL) return false;
    return true;
}

static bool same_shape(const std::vector<GridDesc> &desc) {
    for (const GridDesc &gd : desc) {
        if (gd.D != desc[0].D) return false;
        for (int d = 0; d < gd.D; ++d)
            if (gd.n[d] != desc[0].n[d]) return false;
    }
    return true;
}

static int bank_acquire(int dev, uint64_t plan_id, const std::vector<double> &h_bank,
                        const std::vector<GridDesc> &h_desc, cudaStream_t st) {
    return g_grid_bank.acquire(dev, plan_id, st, [&](cudaStream_t s) {
        cudaError_t e = cudaMemcpyToSymbolAsync(c_grid, h_bank.data(), h_bank.size() * sizeof(double), 0,
                                                cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) return e;
        return cudaMemcpyToSymbolAsync(c_gdesc, h_desc.data(), h_desc.size() * sizeof(GridDesc), 0,
                                       cudaMemcpyHostToDevice, s);
    });
}

template <typename T>
static bool upload(T **dptr, const T *src, size_t count) {
    if (count == 0) count = 1;
    if (cudaMalloc(dptr, count * sizeof(T)) != cudaSuccess) return false;
    return !src || cudaMemcpy(*dptr, src, count * sizeof(T), cudaMemcpyHostToDevice) == cudaSuccess;
}

static int grid_launch_dims(const PlanBase *pl, const void *kernel, int threads, size_t smem,
                            int64_t N, int *grid) {
    PCB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    PCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    const int64_t want = (N + threads - 1) / threads;
    const int64_t cap = (int64_t)pl->sm_count * (per_sm > 0 ? per_sm : 1);
    *grid = (int)(want < cap ? want : cap);
    return PCB_OK;
}

SplinePlan::~SplinePlan() {
    g_grid_bank.forget(dev, plan_id);
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
    pl->plan_id = next_plan_id();
    pl->bank_ok = bank_fits(desc, tensor_total) && same_shape(desc) && !getenv("PCB_NO_BANK");
    if (pl->bank_ok) {
        pl->h_bank = il;
        pl->h_desc = desc;
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
    const int64_t want = (N + 255) / 256;
    const int64_t cap = (int64_t)pl->sm_count * 8;
    spline_lookup_kernel<<<(int)(want < cap ? want : cap), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        pl->D, pl->d_num_knots, pl->d_knot_off, pl->d_knots, d_points, N, d_piece);
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
    if (smem > (size_t)pl->smem_optin)
        return fail(PCB_EUNSUPPORTED, "weight rows of %d nodes do not fit in shared memory", pl->max_sum_n);
    if (pl->bank_ok) {
        const void *uk = GRID_KERNEL_TABLE(spline_uniform_kernel, pl->GB, grid_pick_dm(pl->D));
        PCB_CUDA(cudaFuncSetAttribute(uk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int64_t blocks = (N + PW_THREADS - 1) / PW_THREADS;
        PCB_REQUIRE(blocks <= 0x7fffffffLL, "batch too large for one launch");
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        if (int rc = bank_acquire(pl->dev, pl->plan_id, pl->h_bank, pl->h_desc, st)) return rc;
        void *uargs[] = {(void *)&pl->D, (void *)&pl->G, (void *)&pl->P, (void *)&pl->d_num_knots,
                         (void *)&pl->d_knot_off, (void *)&pl->d_knots, (void *)&pl->d_nodes,
                         (void *)&pl->d_weights, (void *)&d_points, (void *)&N, (void *)&d_out,
                         (void *)&d_piece};
        const cudaError_t e = cudaLaunchKernel(uk, dim3((unsigned)blocks), dim3(PW_THREADS), uargs, smem, st);
        g_grid_bank.release(pl->dev, st);
        g_launches.fetch_add(1);
        PCB_CUDA(e);
        PCB_CUDA(cudaGetLastError());
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
    pl->bank_ok = bank_fits(desc, tensor_total) && !getenv("PCB_NO_BANK");
    if (pl->bank_ok) {
        pl->h_bank = il;
        pl->h_desc = desc;
    }
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
    if (smem > (size_t)pl->smem_optin)
        return fail(PCB_EUNSUPPORTED, "weight rows of %d nodes do not fit in shared memory", pl->max_sum_n);
    if (pl->bank_ok) {
        const void *uk = GRID_KERNEL_TABLE(slider_uniform_kernel, pl->GB, grid_pick_dm(pl->max_D));
        PCB_CUDA(cudaFuncSetAttribute(uk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int64_t blocks = (N + PW_THREADS - 1) / PW_THREADS;
        PCB_REQUIRE(blocks <= 0x7fffffffLL, "batch too large for one launch");
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        if (int rc = bank_acquire(pl->dev, pl->plan_id, pl->h_bank, pl->h_desc, st)) return rc;
        void *uargs[] = {(void *)&pl->D, (void *)&pl->S, (void *)&pl->G, (void *)&pl->pivot,
                         (void *)&pl->d_ints, (void *)&pl->d_group_dims, (void *)&pl->d_out_slide,
                         (void *)&pl->d_row_out, (void *)&pl->d_slide_G, (void *)&pl->d_nodes,
                         (void *)&pl->d_weights, (void *)&d_points, (void *)&N, (void *)&d_out};
        const cudaError_t e = cudaLaunchKernel(uk, dim3((unsigned)blocks), dim3(PW_THREADS), uargs, smem, st);
        g_grid_bank.release(pl->dev, st);
        g_launches.fetch_add(1);
        PCB_CUDA(e);
        PCB_CUDA(cudaGetLastError());
        return PCB_OK;
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
