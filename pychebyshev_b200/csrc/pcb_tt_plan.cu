// ChebyshevTT: plan creation, launch-configuration choice and the C-ABI entry points.
#include <cstdlib>

#include "pcb_cbank.cuh"
#include "pcb_tt.cuh"

namespace pcb {

static int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

static bool tt_try_cfg(const TTPlan *pl, bool shared, int mode, int qpt, int threads, TTCfg *c) {
    const TTParams &P = pl->P;
    const int lc = mode == TT_RESIDENT ? (P.rmaxp <= 8 ? 8 : (P.rmaxp <= 12 ? 12 : 16)) : 16;
    int pingpong = 0;
    for (int k = 0; k < P.D; ++k)
        if (P.rp[k] > lc || P.rpT[k] > lc) pingpong = 1;
    const int nbuf = (shared ? 2 : 1) + pingpong;
    const size_t bytes = (size_t)tt_smem_doubles(P, mode, shared, nbuf, qpt, threads) * sizeof(double);
    if (bytes > (size_t)pl->smem_optin) return false;
    if (qpt >= 3) return false;  // the shared-memory kernels are instantiated for 1 and 2 query slots
    c->qpt = qpt;
    c->threads = threads;
    c->lc = lc;
    c->mode = mode;
    c->pingpong = pingpong;
    c->smem = bytes;
    return true;
}

// Preference: cores resident in shared memory, two query slots per thread (halves the LDS
// traffic per DFMA), as many threads as the register file allows for that kernel.
int tt_pick_cfg(const TTPlan *pl, bool shared, TTCfg *cfg) {
    const int force_q = env_int("PCB_TT_QPT", 0), force_t = env_int("PCB_TT_THREADS", 0);
    // measured on B200 (tools/tt_sweep.py, 5D Black-Scholes TT, gpurun_out/tt_sweep3.log): chain
    // kernels 2 slots x 512 threads 3.06e9 values/s (3 x 256: 2.79e9, 1 x 512: 2.29e9);
    // shared-FD kernel 2 x 384 threads 1.26e9 q/s (2 x 512: 1.02e9, 3 x 256: 1.19e9)
    static const int chain_pref[][2] = {{2, 512}, {2, 256}, {1, 512}, {1, 256}, {1, 128}};
    static const int shared_pref[][2] = {{2, 384}, {2, 256}, {1, 512}, {1, 256}, {1, 128}};
    for (int mode = TT_RESIDENT; mode <= TT_GLOBAL; ++mode) {
        if (force_q || force_t) {
            const int q = force_q ? force_q : 2, t = force_t ? force_t : 256;
            if (tt_try_cfg(pl, shared, mode, q, t, cfg)) return PCB_OK;
            continue;
        }
        for (int i = 0; i < 5; ++i) {
            const int *p = shared ? shared_pref[i] : chain_pref[i];
            if (tt_try_cfg(pl, shared, mode, p[0], p[1], cfg)) return PCB_OK;
        }
    }
    return fail(PCB_EUNSUPPORTED, "TT plan does not fit in shared memory in any configuration");
}

static int tt_build_program(const TTPlan *pl, int G, const int32_t *orders, TTFdProgram *prog) {
    const TTParams &P = pl->P;
    PCB_REQUIRE(G >= 1 && G <= TT_MAX_G, "number of derivative rows %d outside [1, %d]", G, TT_MAX_G);
    prog->G = G;
    for (int g = 0; g < G; ++g) {
        TTFdRow &row = prog->row[g];
        row.m = 0;
        for (int k = 0; k < P.D; ++k) {  // storage frame: order of storage dim k is orders[g][perm[k]]
            const int o = orders[(size_t)g * P.D + P.perm[k]];
            PCB_REQUIRE(o >= 0, "negative derivative order");
            if (o == 0) continue;
            // reference: ValueError(f"Derivative order {order} not supported (use 1 or 2)")
            PCB_REQUIRE(o <= 2, "Derivative order %d not supported (use 1 or 2)", o);
            if (row.m == TT_MAX_ACTIVE)
                return fail(PCB_EUNSUPPORTED,
                            "finite-difference rows with more than %d differentiated dims are not "
                            "supported on the device", TT_MAX_ACTIVE);
            row.dim[row.m] = k;
            row.ord[row.m] = o;
            ++row.m;
        }
    }
    return PCB_OK;
}

// Rows that differentiate at most one dim each can share partial products (algo 2).
static bool tt_build_shared_program(const TTFdProgram &prog, TTSharedProgram *sp) {
    sp->G = prog.G;
    sp->n_slots = 0;
    bool used[PCB_MAX_DIMS] = {false};
    for (int g = 0; g < prog.G; ++g) {
        if (prog.row[g].m > 1) return false;
        if (prog.row[g].m == 1) used[prog.row[g].dim[0]] = true;
    }
    int slot_of[PCB_MAX_DIMS];
    for (int k = 0; k < PCB_MAX_DIMS; ++k) {
        slot_of[k] = -1;
        if (used[k]) {
            slot_of[k] = sp->n_slots;
            sp->slot_dim[sp->n_slots++] = k;
        }
    }
    if (sp->n_slots == 0) return false;  // values only: nothing to share
    for (int g = 0; g < prog.G; ++g) {
        sp->row_slot[g] = prog.row[g].m == 1 ? slot_of[prog.row[g].dim[0]] : -1;
        sp->row_ord[g] = prog.row[g].m == 1 ? prog.row[g].ord[0] : 0;
    }
    return true;
}

TTPlan::~TTPlan() {
    ttc_forget(this);
    for (auto &kv : gimages) {
        cudaHostUnregister(kv.second.data.data());
        cudaGetLastError();
    }
    if (d_cores) cudaFree(d_cores);
}

}  // namespace pcb

using namespace pcb;

extern "C" PCB_API int pcb_tt_plan_create(int dev, int D, const int32_t *n, const int32_t *ranks,
                                          const double *lo, const double *hi,
                                          const int32_t *dim_order, const double *cores_cat,
                                          void **plan) {
    PCB_REQUIRE(plan && n && ranks && lo && hi && cores_cat, "null argument");
    PCB_REQUIRE(D >= 1 && D <= PCB_MAX_DIMS, "num_dimensions %d outside [1, %d]", D, PCB_MAX_DIMS);
    PCB_REQUIRE(ranks[0] == 1 && ranks[D] == 1, "boundary TT ranks must be 1");
    TTPlan *pl = new TTPlan();
    pl->kind = PLAN_TT;
    pl->dev = dev;
    int cc = 0;
    if (int rc = device_props(dev, &pl->sm_count, &pl->smem_optin, &cc)) {
        delete pl;
        return rc;
    }
    TTParams &P = pl->P;
    memset(&P, 0, sizeof(P));
    P.D = D;
    std::vector<char> seen(D, 0);
    int off = 0, rmaxp = 2, maxcore = 0;
    for (int k = 0; k < D; ++k) {
        const int perm = dim_order ? dim_order[k] : k;
        if (n[k] < 1 || ranks[k] < 1 || ranks[k + 1] < 1 || perm < 0 || perm >= D || seen[perm] ||
            !(lo[k] < hi[k])) {
            delete pl;
            return fail(PCB_EINVAL, "invalid TT description at storage dim %d", k);
        }
        seen[perm] = 1;
        P.n[k] = n[k];
        P.r[k] = ranks[k];
        P.rp[k] = round_up(ranks[k + 1], 2);
        P.rpT[k] = round_up(ranks[k], 2);
        P.off[k] = off;
        P.perm[k] = perm;
        P.lo[k] = lo[k];
        P.hi[k] = hi[k];
        const int sz = ranks[k] * n[k] * P.rp[k];
        off += sz;
        if (sz > maxcore) maxcore = sz;
        if (P.rp[k] > rmaxp) rmaxp = P.rp[k];
        if (P.rpT[k] > rmaxp) rmaxp = P.rpT[k];
    }
    P.r[D] = 1;
    P.total = off;
    P.rmaxp = rmaxp;
    // transposed copies for right-to-left sweeps: [l][j][i] with i padded to even by zeros
    int offT = off;
    for (int k = 0; k < D; ++k) {
        P.offT[k] = offT;
        const int sz = ranks[k + 1] * n[k] * P.rpT[k];
        offT += sz;
        if (sz > maxcore) maxcore = sz;
    }
    P.totalT = offT - off;
    P.maxcore = maxcore;

    // + one row of slack: the software-pipelined kernels load (and discard) the row after the last
    std::vector<double> packed((size_t)offT + rmaxp + 32, 0.0);
    size_t src = 0;
    for (int k = 0; k < D; ++k) {
        const int r0 = ranks[k], r1 = ranks[k + 1];
        for (int i = 0; i < r0; ++i)
            for (int j = 0; j < n[k]; ++j) {
                double *dst = &packed[(size_t)P.off[k] + ((size_t)i * n[k] + j) * P.rp[k]];
                for (int l = 0; l < r1; ++l) {
                    const double v = cores_cat[src++];
                    dst[l] = v;
                    packed[(size_t)P.offT[k] + ((size_t)l * n[k] + j) * P.rpT[k] + i] = v;
                }
            }
    }
    // uniform-datapath path: unpadded forward and transposed cores, packed per launch need into
    // constant-bank images (TTPlan::const_image)
    {
        int fwd = 0, rmax = 1;
        for (int k = 0; k < D; ++k) {
            pl->core_off[k] = fwd;
            fwd += ranks[k] * n[k] * ranks[k + 1];
            if (ranks[k + 1] > rmax) rmax = ranks[k + 1];
        }
        pl->core_off[D] = fwd;
        pl->const_enabled = env_int("PCB_TT_CONST", 1) != 0 && rmax <= TT_CONST_MAX_RANK;
        pl->const_value_ok = pl->const_enabled && fwd <= TT_CONST_MAX;
        pl->const_shared_ok = pl->const_enabled && 2 * fwd <= TT_CONST_MAX;  // any Greek set fits
        if (pl->const_enabled) {
            pl->h_fwd.assign(cores_cat, cores_cat + fwd);
            pl->h_T.resize((size_t)fwd);
            size_t s2 = 0;
            for (int k = 0; k < D; ++k)
                for (int i = 0; i < ranks[k]; ++i)
                    for (int j = 0; j < n[k]; ++j)
                        for (int l = 0; l < ranks[k + 1]; ++l)
                            pl->h_T[(size_t)pl->core_off[k] + ((size_t)l * n[k] + j) * ranks[k] + i] =
                                cores_cat[s2++];
            // (qpt, threads) of the uniform-path kernels; ttc_pick falls back to the rank class's
            // default variant when the pair is not compiled in or does not fit in shared memory
            pl->const_qpt_value = env_int("PCB_TT_QPT_VALUE", env_int("PCB_TT_QPT", 2));
            pl->const_qpt_shared = env_int("PCB_TT_QPT_FD", env_int("PCB_TT_QPT", 2));
            pl->const_threads_value = env_int("PCB_TT_THREADS_VALUE", env_int("PCB_TT_THREADS", 512));
            pl->const_threads_shared = env_int("PCB_TT_THREADS_FD", env_int("PCB_TT_THREADS", 512));
        }
        // large trains: every single core fits in the bank -> one launch per core
        bool each_fits = env_int("PCB_TT_CONST", 1) != 0 && env_int("PCB_TT_GSTREAM", 1) != 0 && rmax <= 64;
        for (int k = 0; k < D; ++k)
            if (ranks[k] * n[k] * ranks[k + 1] > TT_CONST_MAX || n[k] > 64) each_fits = false;
        // shared-memory columns of the per-core kernels (256 threads x 2 query slots x 8 B per row):
        // step: r_in rows; coefficient pass: r_rows + r_acc + n rows
        size_t step_rows = 1, coeff_rows = 1;
        for (int k = 0; k < D; ++k) {
            step_rows = std::max<size_t>(step_rows, std::max(ranks[k], ranks[k + 1]));
            coeff_rows = std::max<size_t>(coeff_rows, (size_t)ranks[k] + ranks[k + 1] + n[k]);
        }
        pl->gstream_ok = each_fits && !pl->const_value_ok && step_rows * 4096 <= (size_t)pl->smem_optin;
        pl->gstream_fd_ok = pl->gstream_ok;  // per Greek set: ttg_launch_shared checks its own slots
        (void)coeff_rows;
        if (pl->gstream_ok && !pl->const_enabled) {  // host copies of the cores (ranks > 16)
            pl->h_fwd.assign(cores_cat, cores_cat + fwd);
            pl->h_T.resize((size_t)fwd);
            size_t s2 = 0;
            for (int k = 0; k < D; ++k)
                for (int i = 0; i < ranks[k]; ++i)
                    for (int j = 0; j < n[k]; ++j)
                        for (int l = 0; l < ranks[k + 1]; ++l)
                            pl->h_T[(size_t)pl->core_off[k] + ((size_t)l * n[k] + j) * ranks[k] + i] =
                                cores_cat[s2++];
        }
    }
    // shared-memory kernel configurations; a train that only the per-core path can run keeps
    // threads = 0 there (its launches then report the error instead of plan creation)
    if (int rc = tt_pick_cfg(pl, false, &pl->cfg_chain)) {
        if (!pl->gstream_ok) {
            delete pl;
            return rc;
        }
        pl->cfg_chain = TTCfg();
    }
    if (int rc = tt_pick_cfg(pl, true, &pl->cfg_shared)) {
        if (!pl->gstream_ok) {
            delete pl;
            return rc;
        }
        pl->cfg_shared = TTCfg();
    }
    DeviceGuard guard(dev);
    if (!guard.ok || cudaMalloc(&pl->d_cores, packed.size() * sizeof(double)) != cudaSuccess) {
        delete pl;
        return fail(PCB_ENOMEM, "cudaMalloc of %zu B for TT cores failed on device %d",
                    packed.size() * sizeof(double), dev);
    }
    if (cudaMemcpy(pl->d_cores, packed.data(), packed.size() * sizeof(double),
                   cudaMemcpyHostToDevice) != cudaSuccess) {
        delete pl;
        return fail(PCB_ECUDA, "upload of TT cores failed");
    }
    *plan = pl;
    return PCB_OK;
}

extern "C" PCB_API int pcb_tt_eval(void *plan, const double *d_points, int64_t N, double *d_out,
                                   void *stream) {
    TTPlan *pl = static_cast<TTPlan *>(plan);
    PCB_REQUIRE(pl && pl->kind == PLAN_TT, "not a TT plan");
    PCB_REQUIRE(N >= 0, "negative N");
    if (N == 0) return PCB_OK;
    PCB_REQUIRE(d_points && d_out, "null device pointer");
    DeviceGuard guard(pl->dev);
    if (pl->gstream_ok)
        return ttg_launch_value(pl, d_points, N, d_out, static_cast<cudaStream_t>(stream));
    if (pl->const_value_ok) {
        bool fits = false;
        const int rc = ttc_launch_value(pl, d_points, N, d_out, static_cast<cudaStream_t>(stream), &fits);
        if (rc || fits) return rc;
    }
    return tt_launch_value(pl, d_points, N, d_out, static_cast<cudaStream_t>(stream));
}

extern "C" PCB_API int pcb_tt_plan_info(void *plan, int32_t *out8) {
    TTPlan *pl = static_cast<TTPlan *>(plan);
    PCB_REQUIRE(pl && pl->kind == PLAN_TT && out8, "not a TT plan");
    out8[0] = pl->const_value_ok;
    out8[1] = pl->const_shared_ok;  // every Greek set fits the bank (per set: pcb_tt_fd_path)
    out8[2] = pl->const_qpt_shared;
    out8[3] = pl->const_threads_value;
    out8[4] = pl->const_threads_shared;
    out8[5] = pl->cfg_chain.mode;
    out8[6] = pl->cfg_shared.qpt;
    out8[7] = pl->cfg_shared.threads;
    return PCB_OK;
}

extern "C" PCB_API int pcb_tt_fd_algo(void *plan, int G, const int32_t *orders) {
    TTPlan *pl = static_cast<TTPlan *>(plan);
    PCB_REQUIRE(pl && pl->kind == PLAN_TT, "not a TT plan");
    PCB_REQUIRE(orders, "null orders");
    TTFdProgram prog;
    if (int rc = tt_build_program(pl, G, orders, &prog)) return rc;
    TTSharedProgram sp;
    return tt_build_shared_program(prog, &sp) ? 2 : 1;
}

// Which kernel family pcb_tt_eval_fd runs for these rows (reporting only; evaluation itself keeps
// no state): 1 one chain per stencil point (general rows), 2 constant-bank shared-product kernel
// (one launch), 3 the same with one launch per differentiated dim, 4 per-core launches (large
// trains), 5 shared-memory shared-product kernel.
extern "C" PCB_API int pcb_tt_fd_path(void *plan, int G, const int32_t *orders, int algo) {
    TTPlan *pl = static_cast<TTPlan *>(plan);
    PCB_REQUIRE(pl && pl->kind == PLAN_TT, "not a TT plan");
    PCB_REQUIRE(orders, "null orders");
    TTFdProgram prog;
    if (int rc = tt_build_program(pl, G, orders, &prog)) return rc;
    TTSharedProgram sp;
    const bool can_share = tt_build_shared_program(prog, &sp);
    if (!(algo == 2 || (algo == 0 && can_share)) || !can_share) return 1;
    if (const int f = ttc_shared_fits(pl, sp)) return 1 + f;
    if (ttg_shared_fits(pl, sp)) return 4;
    return 5;
}

extern "C" PCB_API int pcb_tt_eval_fd(void *plan, const double *d_points, int64_t N, int G,
                                      const int32_t *orders, double *d_out, int algo, void *stream) {
    TTPlan *pl = static_cast<TTPlan *>(plan);
    PCB_REQUIRE(pl && pl->kind == PLAN_TT, "not a TT plan");
    PCB_REQUIRE(orders, "null orders");
    PCB_REQUIRE(N >= 0, "negative N");
    PCB_REQUIRE(algo >= 0 && algo <= 2, "algo %d not available", algo);
    TTFdProgram prog;
    if (int rc = tt_build_program(pl, G, orders, &prog)) return rc;
    TTSharedProgram sp;
    const bool can_share = tt_build_shared_program(prog, &sp);
    if (algo == 2 && !can_share)
        return fail(PCB_EUNSUPPORTED, "algo 2 needs at least one differentiated dim and at most "
                    "one per row");
    if (N == 0) return PCB_OK;
    PCB_REQUIRE(d_points && d_out, "null device pointer");
    DeviceGuard guard(pl->dev);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (algo == 2 || (algo == 0 && can_share)) {
        // constant bank when the cores this Greek set reads fit in it, shared memory otherwise
        bool fits = false;
        if (pl->const_enabled) {
            const int rc = ttc_launch_shared(pl, sp, d_points, N, d_out, st, &fits);
            if (rc || fits) return rc;
        }
        if (pl->gstream_fd_ok) {
            const int rc = ttg_launch_shared(pl, sp, d_points, N, d_out, st, &fits);
            if (rc || fits) return rc;
        }
        return tt_launch_shared(pl, sp, d_points, N, d_out, st);
    }
    return tt_launch_general(pl, prog, d_points, N, d_out, st);
}
