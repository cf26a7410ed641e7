#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's configs.

Metric : fp64 queries/sec, price + Greeks
Main   : configs[1] (C2) -- 5D Black-Scholes ChebyshevTT (reference TT-Cross cores, ranks
         [1,11,11,11,7,1]), price+delta+gamma+vega, 1e8 synthetic uniform queries per GPU per step
         (weak scaling over 1-8 B200, no collective on the data path).

A step = one pass of the hot path over one batch of synthetic queries.

  value        device-resident inputs/outputs, CUDA events on the launching stream, max over ranks
  e2e          the public API with HOST (pinned) NumPy buffers: H2D + kernel + D2H inside the timing;
               e2e.ceiling = the same bytes moved by bare concurrent cudaMemcpyAsync (no kernel)
  gather       (N > 1) the same step followed by an all-gather of the result shards over NCCL --
               off the data path, reported separately
  roofline     of the dominant kernel: FP64 pipe (peak measured live with a DFMA probe) or HBM
  cpu_baseline the UNMODIFIED reference (oracle/_ref/site, pychebyshev 0.21.1) on this box's host
               cores, on a bounded sample of the same workload (rank 0, N=1)
  configs      every other BASELINE.json config (C1, C3 lookup/2-D/3-D, C4, C5 basket/rank-20/slider)
               as its own clocked measurement: q/s, kernel, roofline fraction, clocks.  Under
               --gpus N, C4 and C5 shard a fixed total (1e9 queries for C5) across the ranks.

``--workload KEY`` makes any of those the main line.  ``--impl reference`` times the unmodified
reference alone on all host cores and prints the same line shape.
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "fp64 queries/sec (price+Greeks)"
UNIT = "queries/s"


# ==========================================================================================
# workloads
# ==========================================================================================

def tt_value_flops(cores):
    """Algorithmic flop of one chain: sum_k 2 r_{k-1} n_k r_k."""
    return sum(2 * c.shape[0] * c.shape[1] * c.shape[2] for c in cores)


def tt_shared_flops(cores, active):
    """Reduced count when left/right partial products are shared across the stencil: left chain up
    to the last differentiated dim, right chain down to the first, plus one coefficient contraction
    per differentiated dim."""
    step = [2 * c.shape[0] * c.shape[1] * c.shape[2] for c in cores]
    if not active:
        return sum(step)
    left = sum(step[:max(active)])
    right = sum(step[min(active) + 1:])
    coeff = sum(step[a] + 2 * cores[a].shape[1] * cores[a].shape[2] for a in active)
    return left + right + coeff


def full_flops(n):
    """2 sum_k prod_{j<=k} n_j per output (contraction from the last axis)."""
    tot, p = 0, 1
    for k in n:
        p *= k
        tot += p
    return 2 * tot


def weight_row_flops(n):
    """Barycentric weight rows in product form: ~6 fp64 ops per node + 1/sum per dim."""
    return 6 * sum(n) + 8 * len(n)


class Case:
    """One workload: the interpolant, how to evaluate it, its roofline model and its reference."""

    key = label = config = kernel = ""
    bound = "fp64"          # "fp64" | "hbm"
    scaling = "weak"        # "weak": n per GPU fixed; "strong": total fixed, sharded over ranks
    n_default = 1 << 20     # queries per GPU per step (weak) or in total (strong)
    block = None            # strong workloads larger than memory: resident block re-evaluated
    out_dtype = "float64"
    flop_q = bytes_q = 0.0
    flop_note = ""
    cpu_rate = 1000.0       # rough reference q/s per core, sizes the CPU sample
    cpu_procs = None        # cap on worker processes of the reference arm

    # --- ours ---
    def build(self, dev):   # -> device plan (eval_device / eval / lookup)
        raise NotImplementedError

    def launch(self, plan, pts, out):
        plan.eval_device(pts, out)

    def api(self, obj, h_pts, h_out, dev):  # the public host-buffer call
        raise NotImplementedError

    # --- reference ---
    def ref_object(self):
        raise NotImplementedError

    @staticmethod
    def ref_eval(obj, pts):
        raise NotImplementedError


def _golden_tt(name):
    import _golden as G

    g = G.load(name)
    cores, domain, dim_order = G.tt_parts(g)
    return g, cores, domain, dim_order


class TTCase(Case):
    bound = "fp64"

    def __init__(self, key, fixture, orders, label, config, n_default, scaling="weak", block=None,
                 cpu_rate=500.0):
        self.key, self.fixture, self.label, self.config = key, fixture, label, config
        self.n_default, self.scaling, self.block, self.cpu_rate = n_default, scaling, block, cpu_rate
        g, self.cores, self.sdomain, self.dim_order = _golden_tt(fixture)
        D = len(self.cores)
        if orders == "fixture3":
            orders = [list(map(int, o)) for o in g["fd_orders"][:3]]
        self.orders = orders
        self.D, self.G = D, (len(orders) if orders else 1)
        self.domain = [self.sdomain[self.dim_order.index(u)] for u in range(D)]
        if orders:
            active = sorted({self.dim_order.index(u) for o in orders for u, k in enumerate(o) if k})
            self.flop_q = tt_shared_flops(self.cores, active)
            self.flop_note = ("left sweep + right sweep + one coefficient pass per differentiated "
                              "dim (partial products shared across the stencil)")
        else:
            self.flop_q = tt_value_flops(self.cores)
            self.flop_note = "sum_k 2 r_{k-1} n_k r_k"
        self.bytes_q = 8.0 * (D + self.G)

    def obj(self, dev):
        import pychebyshev_b200 as pcb

        return pcb.ChebyshevTT.from_cores(self.cores, self.sdomain, self.dim_order, device=dev)

    def build(self, dev, algo=0):
        self.tt = self.obj(dev)
        plan = self.tt._plan(dev)
        if self.orders:
            plan = plan.with_orders(np.asarray(self.orders), algo)
        info = plan.info()
        if self.orders:
            self.kernel = info["fd_kernel"].split(" (")[0].replace(" + ", "+")
        else:
            self.kernel = "ttc_value_kernel" if info["uniform_path_values"] else "ttc_gstep_kernel"
        self.plan_info = info
        return plan

    def api(self, obj, h_pts, h_out, dev):
        if self.orders:
            obj.eval_multi_batch(h_pts, self.orders, out=h_out, device=dev)
        else:
            obj.eval_batch(h_pts, out=h_out, device=dev)

    def ref_object(self):
        from oracle import ref_objects as RO

        return (RO.tt_from_cores(self.cores, self.sdomain, self.dim_order), self.orders)

    @staticmethod
    def ref_eval(obj, pts):
        tt, orders = obj
        if orders:   # the reference's price+Greeks route: eval_multi per point (FD stencils)
            for p in pts.tolist():
                tt.eval_multi(p, orders)
        else:
            tt.eval_batch(pts)


class FullCase(Case):
    bound = "fp64"

    def __init__(self, key, which, label, config, n_default, scaling="weak", cpu_rate=400.0,
                 cpu_procs=None):
        from pychebyshev_b200 import workloads as wl

        self.key, self.which, self.label, self.config = key, which, label, config
        self.n_default, self.scaling, self.cpu_rate, self.cpu_procs = n_default, scaling, cpu_rate, cpu_procs
        if which == "bs5d":
            self.domain, self.n_nodes, self.orders, self.func = (wl.BS5D_DOMAIN, wl.BS5D_NODES,
                                                                 wl.BS5D_GREEKS, wl.bs_call_price)
        else:
            self.domain, self.n_nodes, self.orders, self.func = (wl.C4_DOMAIN, wl.C4_NODES,
                                                                 wl.C4_GREEKS, wl.bs6d)
        self.D, self.G = len(self.n_nodes), len(self.orders)
        self.flop_q = self.G * full_flops(self.n_nodes)
        self.flop_note = "G x 2 sum_k prod_{j<=k} n_j (weight rows not counted)"
        self.bytes_q = 8.0 * (self.D + self.G)
        self.kernel = ("full_dmma2_kernel (joint-K)" if which == "bs5d" and not os.environ.get("PCB_NO_DMMA2")
                       else "full_dmma_kernel")

    def obj(self, dev):
        import pychebyshev_b200 as pcb
        from pychebyshev_b200 import _grid, workloads as wl

        nodes = [_grid.cheb_nodes(lo, hi, n) for (lo, hi), n in zip(self.domain, self.n_nodes)]
        return pcb.ChebyshevApproximation.from_values(wl.grid_values(self.func, nodes), self.D,
                                                      self.domain, self.n_nodes, device=dev)

    def build(self, dev, algo=0):
        t0 = time.perf_counter()
        self.cheb = self.obj(dev)
        plan = self.cheb._plan(self.orders, dev)
        self.plan_seconds = time.perf_counter() - t0
        return plan

    def api(self, obj, h_pts, h_out, dev):
        obj.eval_batch_multi(h_pts, self.orders, out=h_out, device=dev)

    def ref_object(self):
        from oracle import ref_objects as RO

        return (RO.full_bs5d() if self.which == "bs5d" else RO.full_c4(), self.orders)

    @staticmethod
    def ref_eval(obj, pts):
        cheb, orders = obj  # vectorized_eval_batch per derivative order, summed
        for o in orders:
            cheb.vectorized_eval_batch(pts, list(o))


class SplineCase(Case):
    def __init__(self, key, dim, orders, label, config, n_default, lookup=False, cpu_rate=20000.0):
        from pychebyshev_b200 import workloads as wl

        self.key, self.dim, self.orders, self.label, self.config = key, dim, orders, label, config
        self.n_default, self.lookup, self.cpu_rate = n_default, lookup, cpu_rate
        if dim == 2:
            self.domain, self.n_nodes, self.knots, self.func = (
                wl.SPLINE2D_DOMAIN, wl.SPLINE2D_NODES, wl.SPLINE2D_KNOTS, wl.payoff2d)
        else:
            self.domain, self.n_nodes, self.knots, self.func = (
                wl.SPLINE3D_DOMAIN, wl.SPLINE3D_NODES, wl.SPLINE3D_KNOTS, wl.payoff3d)
        self.D = dim
        if lookup:
            self.G, self.bound, self.out_dtype = 1, "hbm", "int32"
            self.bytes_q = 8.0 * dim + 4.0
            self.kernel = "spline_lookup_kernel"
        else:
            self.G, self.bound = len(orders), "fp64"
            self.flop_q = self.G * full_flops(self.n_nodes) + weight_row_flops(self.n_nodes)
            self.flop_note = ("G x 2 sum_k prod_{j<=k} n_j + product-form weight rows "
                              "(6 ops per node + 1/sum per dim)")
            self.bytes_q = 8.0 * (dim + self.G)
            self.kernel = ("spline2d_dmma_kernel" if dim == 2 and not os.environ.get("PCB_NO_DMMA2D")
                           else ("spline3d_dmma_kernel" if dim == 3 and not os.environ.get("PCB_NO_DMMA3D")
                                 else "spline_bank_kernel"))

    def obj(self, dev):
        import pychebyshev_b200 as pcb
        from pychebyshev_b200 import workloads as wl

        info = pcb.ChebyshevSpline.nodes(self.D, self.domain, self.n_nodes, self.knots)
        vals = [wl.grid_values(self.func, p["nodes_per_dim"]) for p in info["pieces"]]
        return pcb.ChebyshevSpline.from_values(vals, self.D, self.domain, self.n_nodes, self.knots,
                                               device=dev)

    def build(self, dev, algo=0):
        self.sp = self.obj(dev)
        return self.sp._plan(self.orders or [[0] * self.D], dev)

    def launch(self, plan, pts, out):
        if self.lookup:
            plan.lookup_device(pts, out)
        else:
            plan.eval_device(pts, out)

    def api(self, obj, h_pts, h_out, dev):
        if self.lookup:
            obj.find_pieces(h_pts, device=dev)
        else:
            obj.eval_batch_multi(h_pts, self.orders, out=h_out, device=dev)

    def ref_object(self):
        from oracle import ref_objects as RO

        return (RO.spline2d() if self.D == 2 else RO.spline3d(), self.orders, self.lookup)

    @staticmethod
    def ref_eval(obj, pts):
        sp, orders, lookup = obj
        if lookup:
            from oracle import ref_objects as RO

            RO.spline_lookup(sp, pts)
        else:
            for o in orders:
                sp.eval_batch(pts, list(o))


class SliderCase(Case):
    bound = "fp64"

    def __init__(self, key, label, config, n_default, scaling="strong", block=None, cpu_rate=8000.0):
        from pychebyshev_b200 import workloads as wl

        self.key, self.label, self.config = key, label, config
        self.n_default, self.scaling, self.block, self.cpu_rate = n_default, scaling, block, cpu_rate
        self.domain = wl.C5_DOMAIN
        self.D, self.G = 10, 1
        self.orders = [[0] * 10]
        self.flop_q = 5 * (full_flops([11, 11]) + weight_row_flops([11, 11]))
        self.flop_note = "5 slides x (2 (121+11) + product-form weight rows)"
        self.bytes_q = 8.0 * (self.D + self.G)
        self.kernel = "slider_bank_kernel" if os.environ.get("PCB_NO_DMMA2D") else "slider2d_dmma_kernel"

    def obj(self, dev):
        import _golden as G
        import pychebyshev_b200 as pcb
        from pychebyshev_b200 import _grid

        g = G.load("slider10d")
        part, pivot_value, slides = G.slider_parts(g, _grid.diff_matrix)
        dom = [list(map(float, r)) for r in g["domain"]]
        nn = [int(v) for v in g["n_nodes"]]
        return pcb.ChebyshevSlider.from_slides([s[0] for s in slides], 10, dom, nn, part,
                                               list(g["pivot_point"]), pivot_value, device=dev)

    def build(self, dev, algo=0):
        self.sl = self.obj(dev)
        return self.sl._plan(self.orders, dev)

    def api(self, obj, h_pts, h_out, dev):
        obj.eval_batch_multi(h_pts, self.orders, out=h_out, device=dev)

    def ref_object(self):
        from oracle import ref_objects as RO

        return RO.slider10d()

    @staticmethod
    def ref_eval(obj, pts):
        zero = [0] * 10
        for p in pts.tolist():
            obj.eval(p, zero)


def make_case(key):
    from pychebyshev_b200 import workloads as wl

    V5, V10 = None, None
    table = {
        "tt_bs5d": lambda: TTCase(
            "tt_bs5d", "tt_bs5d", wl.BS5D_GREEKS,
            "5D Black-Scholes ChebyshevTT (reference TT-Cross cores, ranks [1,11,11,11,7,1]): "
            "price+delta+gamma+vega per query via eval_multi_batch -> pcb_tt_eval_fd",
            "C2", 100_000_000, cpu_rate=450.0),
        "tt_bs5d_value": lambda: TTCase(
            "tt_bs5d_value", "tt_bs5d", V5, "5D Black-Scholes ChebyshevTT: eval_batch (values)",
            "C2", 100_000_000, cpu_rate=60000.0),
        "full_bs5d": lambda: FullCase(
            "full_bs5d", "bs5d", "5D Black-Scholes ChebyshevApproximation 11^5: price+delta+gamma+"
            "vega per query via eval_batch_multi -> pcb_full_eval (DMMA)", "C1", 148 * 256 * 64,
            cpu_rate=500.0),
        "full_c4": lambda: FullCase(
            "full_c4", "c4", "6D ChebyshevApproximation 16^6 (4 x 134 MB tensors): price+delta+"
            "gamma+vega, queries sharded across the GPUs", "C4", 148 * 256 * 8, scaling="strong",
            cpu_rate=3.0, cpu_procs=2),
        "spline2d_lookup": lambda: SplineCase(
            "spline2d_lookup", 2, None, "2D ChebyshevSpline, knot at the strike: piece lookup "
            "(int32, bit-exact)", "C3", 100_000_000, lookup=True, cpu_rate=2e7),
        "spline2d": lambda: SplineCase(
            "spline2d", 2, [[0, 0]], "2D ChebyshevSpline 2 x (15x15): lookup + per-piece values",
            "C3", 100_000_000, cpu_rate=30000.0),
        "spline2d_greeks": lambda: SplineCase(
            "spline2d_greeks", 2, [[0, 0], [1, 0]], "2D ChebyshevSpline 2 x (15x15): value + d/dS",
            "C3", 100_000_000, cpu_rate=12000.0),
        "spline3d": lambda: SplineCase(
            "spline3d", 3, [[0, 0, 0]], "3D ChebyshevSpline 2 x 15^3: lookup + per-piece values",
            "C3", 20_000_000, cpu_rate=8000.0),
        "spline3d_greeks": lambda: SplineCase(
            "spline3d_greeks", 3, [[0, 0, 0], [1, 0, 0]], "3D ChebyshevSpline 2 x 15^3: value + d/dS",
            "C3", 20_000_000, cpu_rate=3000.0),
        "tt_basket10d": lambda: TTCase(
            "tt_basket10d", "tt_basket10d", "fixture3",
            "10D basket ChebyshevTT ranks <= 10: value + 2 finite-difference Greeks, 1e9 queries per "
            "step sharded across the GPUs", "C5", 1_000_000_000, scaling="strong",
            block=25_000_000, cpu_rate=500.0),
        "tt_basket10d_value": lambda: TTCase(
            "tt_basket10d_value", "tt_basket10d", V10,
            "10D basket ChebyshevTT ranks <= 10: eval_batch, 1e9 queries per step sharded across "
            "the GPUs", "C5", 1_000_000_000, scaling="strong", block=25_000_000, cpu_rate=35000.0),
        "tt_rank20": lambda: TTCase(
            "tt_rank20", "tt_rank20_10d", "fixture3",
            "10D ChebyshevTT uniform rank 20: value + 2 finite-difference Greeks, 1e8 queries per "
            "step sharded across the GPUs", "C5", 100_000_000, scaling="strong",
            block=12_500_000, cpu_rate=300.0),
        "tt_rank20_value": lambda: TTCase(
            "tt_rank20_value", "tt_rank20_10d", V10,
            "10D ChebyshevTT uniform rank 20: eval_batch, 2e8 queries per step sharded across the "
            "GPUs", "C5", 200_000_000, scaling="strong", block=12_500_000, cpu_rate=15000.0),
        "slider10d": lambda: SliderCase(
            "slider10d", "10D ChebyshevSlider 5 x (11x11): values, 1e9 queries per step sharded "
            "across the GPUs", "C5", 1_000_000_000, block=25_000_000),
    }
    if key not in table:
        raise SystemExit(f"unknown workload {key!r}; choose from {sorted(table)}")
    return table[key]()


MAIN = "tt_bs5d"
BLOCK_WORKLOADS = ["tt_bs5d_value", "full_bs5d", "spline2d_lookup", "spline2d", "spline2d_greeks",
                   "spline3d", "spline3d_greeks", "full_c4", "tt_basket10d", "tt_basket10d_value",
                   "tt_rank20", "tt_rank20_value", "slider10d"]


# ==========================================================================================
# clocks
# ==========================================================================================

class ClockSampler:
    """Samples SM clock + throttle reasons of one GPU while the timed region runs."""

    REASONS = {
        0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
        0x4: "sw_power_cap", 0x80: "hw_power_brake", 0x2: "applications_clocks_setting",
    }

    def __init__(self, index):
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def _sample(self):
        try:
            self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            for bit, name in self.REASONS.items():
                if mask & bit:
                    self.reasons.add(name)
        except Exception:  # noqa: BLE001
            pass

    def _run(self):
        while not self._stop.is_set():
            self._sample()
            self._stop.wait(0.02)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        if self.nv is not None:
            self._sample()
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def visible_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:  # noqa: BLE001
            return local_rank
    return local_rank


# ==========================================================================================
# CPU arm: the unmodified reference on the host cores
# ==========================================================================================

_W_OBJ = None
_W_EVAL = None


def _worker_init(payload, evaluator):
    global _W_OBJ, _W_EVAL
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["OMP_NUM_THREADS"] = "1"
    sys.dont_write_bytecode = True
    import pickle

    from oracle import reference as R

    R.load()
    _W_OBJ = pickle.loads(payload)
    _W_EVAL = evaluator


def _worker_run(pts):
    t0 = time.perf_counter()
    _W_EVAL(_W_OBJ, pts)
    return time.perf_counter() - t0


class ReferenceArm:
    """The reference's own implementation of one workload on ``procs`` host processes (its loops are
    GIL-bound, so processes -- BASELINE.md §3), each holding its own unpickled interpolant."""

    def __init__(self, case, procs=None):
        import pickle
        from concurrent.futures import ProcessPoolExecutor

        from oracle import reference as R

        self.case = case
        self.ref = R.load()
        ncores = os.cpu_count() or 1
        self.procs = max(1, min(procs or ncores, case.cpu_procs or ncores, ncores))
        obj = case.ref_object()
        self.pool = ProcessPoolExecutor(self.procs, initializer=_worker_init,
                                        initargs=(pickle.dumps(obj), type(case).ref_eval))
        self.where = R.where()

    def sample(self, n, seed):
        from pychebyshev_b200 import workloads as wl

        return wl.uniform_queries(self.case.domain, n, seed)

    def run(self, n, seed):
        """Evaluate n fresh queries, sharded over the workers; returns wall seconds."""
        shards = [s for s in np.array_split(self.sample(n, seed), self.procs) if len(s)]
        t0 = time.perf_counter()
        list(self.pool.map(_worker_run, shards))
        return time.perf_counter() - t0

    def close(self):
        self.pool.shutdown()

    def describe(self, n):
        c = self.case
        call = {TTCase: "ChebyshevTT.eval_multi per point" if getattr(c, "orders", None)
                else "ChebyshevTT.eval_batch",
                FullCase: "ChebyshevApproximation.vectorized_eval_batch per derivative order",
                SplineCase: "ChebyshevSpline.eval_batch per derivative order"
                if not getattr(c, "lookup", False) else "searchsorted/clip/ravel_multi_index routing",
                SliderCase: "ChebyshevSlider.eval per point"}[type(c)]
        return (f"{n} uniform queries x {c.G} outputs, unmodified pychebyshev "
                f"{getattr(self.ref, '__version__', '?')} ({call}), {self.procs} process(es)")


def reference_rate(case, seconds=2.0, procs=None, seed=77):
    """q/s of the reference on a bounded sample (~`seconds` of wall after a warm-up)."""
    arm = ReferenceArm(case, procs)
    try:
        n = max(arm.procs * 2, int(case.cpu_rate * arm.procs * seconds))
        arm.run(max(arm.procs, n // 8), seed)            # warm-up: imports, page-in, BLAS init
        wall = arm.run(n, seed + 1)
        if wall < 0.4 * seconds:                          # rate guess was low: one larger sample
            n = int(n * min(8.0, seconds / max(wall, 1e-3)))
            wall = arm.run(n, seed + 2)
        return {"value": n / wall, "unit": UNIT, "cores": arm.procs, "kind": "reference",
                "sample": arm.describe(n), "seconds": wall,
                "reference_path": os.path.relpath(arm.where, ROOT)
                if arm.where.startswith(ROOT) else arm.where}
    finally:
        arm.close()


def tt_stencil_batch_rate(case, n=40000, seed=78):
    """Context: the reference's fastest own route to price+Greeks of a TT -- the same central
    stencils evaluated as 5 batches through ``eval_batch`` (tensor_train.py:2217-2265) instead of
    the per-point ``eval_multi`` loop.  One process."""
    from oracle import reference as R
    from pychebyshev_b200 import workloads as wl

    R.load()
    tt, orders = case.ref_object()
    pts = wl.uniform_queries(case.domain, n, seed)
    active = sorted({u for o in orders for u, k in enumerate(o) if k})

    def go():
        tt.eval_batch(pts)
        for u in active:
            h = (case.domain[u][1] - case.domain[u][0]) * 1e-4
            for s in (+h, -h):
                q = pts.copy()
                q[:, u] += s
                tt.eval_batch(q)
    go()
    t0 = time.perf_counter()
    go()
    wall = time.perf_counter() - t0
    return {"value": n / wall, "unit": UNIT, "cores": 1, "kind": "reference",
            "sample": f"{n} queries, {1 + 2 * len(active)} stencil batches through "
                      "ChebyshevTT.eval_batch (no boundary nudge), 1 process"}


def config_dict(case, n, world):
    """Identical in both arms (the driver compares them)."""
    total = n * world if case.scaling == "weak" else n
    cfg = {"workload": case.label, "key": case.key, "baseline_config": case.config,
           "outputs_per_query": case.G, "scaling": case.scaling,
           "queries_per_step_total": int(total)}
    if case.scaling == "weak":
        cfg["queries_per_gpu_per_step"] = int(n)
    return cfg


def run_reference(args, rank, world):
    if rank != 0:
        return
    case = make_case(args.workload)
    n = args.queries or case.n_default
    arm = ReferenceArm(case)
    per_step = max(arm.procs * 2, int(case.cpu_rate * arm.procs * args.cpu_step_seconds))
    try:
        arm.run(max(arm.procs, per_step // 4), 99)
        for i in range(args.warmup):
            arm.run(per_step, 100 + i)
        walls = [arm.run(per_step, 1000 + i) for i in range(args.steps)]
    finally:
        arm.close()
    total_s = float(sum(walls))
    value = per_step * args.steps / total_s
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_s / args.steps,
        "higher_is_better": True, "scaling": case.scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config_dict(case, n, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.procs, "kind": "reference",
                         "sample": arm.describe(per_step) + " per step",
                         "reference_path": os.path.relpath(arm.where, ROOT)
                         if arm.where.startswith(ROOT) else arm.where},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


# ==========================================================================================
# our arm
# ==========================================================================================

def device_queries(domain, n, dev, seed):
    import torch

    gen = torch.Generator(device=dev).manual_seed(seed)
    D = len(domain)
    lo = torch.tensor([d[0] for d in domain], device=dev, dtype=torch.float64)
    hi = torch.tensor([d[1] for d in domain], device=dev, dtype=torch.float64)
    pts = torch.empty((n, D), dtype=torch.float64, device=dev)
    step = 1 << 22
    for s in range(0, n, step):
        e = min(n, s + step)
        pts[s:e] = lo + (hi - lo) * torch.rand((e - s, D), generator=gen, device=dev,
                                               dtype=torch.float64)
    return pts


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except Exception:  # noqa: BLE001
        return {}


class Timed:
    """Device-timed run of one case on this rank: warm-up, K steps, clocks, launches."""

    def __init__(self, case, dev, local_rank, rank, world, n_req, algo=0):
        import torch

        from pychebyshev_b200 import sharding

        self.case, self.dev, self.world = case, dev, world
        self.plan = case.build(local_rank, algo) if algo else case.build(local_rank)
        n_total = n_req or case.n_default
        if case.scaling == "strong":
            lo, hi = sharding.shard_range(n_total, rank, world)
            self.n_local, self.n_total = hi - lo, n_total
        else:
            self.n_local, self.n_total = n_total, n_total * world
        blk = case.block or self.n_local
        self.block = max(1, min(blk, self.n_local))
        self.passes = [self.block] * (self.n_local // self.block)
        if self.n_local % self.block:
            self.passes.append(self.n_local % self.block)
        self.pts = device_queries(case.domain, self.block, dev, 1234 + rank)
        dt = torch.int32 if case.out_dtype == "int32" else torch.float64
        self.out = torch.empty((self.block, case.G), dtype=dt, device=dev)

    def step(self):
        for m in self.passes:
            self.case.launch(self.plan, self.pts[:m], self.out[:m])

    def run(self, steps, warmup, barrier, gpu_index):
        import torch

        from pychebyshev_b200 import _engine

        for _ in range(warmup):
            self.step()
        barrier()
        l0 = _engine.launch_count()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
               for _ in range(steps)]
        with ClockSampler(gpu_index) as clocks:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for e0, e1 in evs:
                e0.record()
                self.step()
                e1.record()
            b.record()
            barrier()
        self.launches = _engine.launch_count() - l0
        self.total_ms = a.elapsed_time(b)
        self.step_ms = float(np.mean([e0.elapsed_time(e1) for e0, e1 in evs]))
        self.clocks = clocks.summary()
        return self


def roofline_of(case, n_local, step_ms, peak_fp64, peaks, peak_note):
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_ach = case.bytes_q * n_local / (step_ms * 1e-3) / 1e9
    if case.bound == "hbm":
        return {"bound": "hbm", "achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s",
                "frac": hbm_ach / hbm_peak, "bytes_per_query": case.bytes_q, "kernel": case.kernel,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy bandwidth)"
                if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"}
    ach = case.flop_q * n_local / (step_ms * 1e-3) / 1e12
    return {"bound": "fp64", "bound_class": "fp64-fma", "achieved": ach, "peak": peak_fp64,
            "unit": "TFLOP/s", "frac": ach / peak_fp64, "flop_per_query": case.flop_q,
            "flop_model": case.flop_note, "kernel": case.kernel, "peak_source": peak_note,
            "hbm": {"achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s",
                    "bytes_per_query": case.bytes_q, "frac": hbm_ach / hbm_peak}}


def copy_ceiling(dev, h_in, h_out, d_in, d_out, reps=3, barrier=None):
    """The same host buffers moved by bare cudaMemcpyAsync, H2D and D2H concurrently on two
    streams, no kernel: the end-to-end ceiling of this box for these byte counts.  Each scheme
    (whole buffers, 64 MB and 256 MB chunks) is timed like the end-to-end region itself -- one
    barrier, then `reps` rounds back to back, ranks free to drift apart -- and the best scheme
    wins.  Returns seconds per round."""
    import torch

    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    t_in, t_out = torch.from_numpy(h_in).view(-1), torch.from_numpy(h_out).view(-1)
    f_in, f_out = d_in.view(-1), d_out.view(-1)
    best, how = 1e30, ""
    for chunk_mb in (0, 64, 256):
        ci = t_in.numel() if not chunk_mb else (chunk_mb << 20) // 8
        co = t_out.numel() if not chunk_mb else max(1, ci * t_out.numel() // max(1, t_in.numel()))
        for timed in (False, True):          # one untimed round per scheme, then the timed ones
            if barrier:
                barrier()
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for _ in range(reps if timed else 1):
                with torch.cuda.stream(s1):
                    for o in range(0, t_in.numel(), ci):
                        f_in[o:o + ci].copy_(t_in[o:o + ci], non_blocking=True)
                with torch.cuda.stream(s2):
                    for o in range(0, t_out.numel(), co):
                        t_out[o:o + co].copy_(f_out[o:o + co], non_blocking=True)
            torch.cuda.synchronize(dev)
            dt = (time.perf_counter() - t0) / max(1, reps)
            if timed and dt < best:
                best, how = dt, ("whole buffers" if not chunk_mb else f"{chunk_mb} MB chunks")
    return best, how


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    from pychebyshev_b200 import _engine

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    gpu_index = visible_gpu_index(local_rank)
    peaks = load_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    peak_tf, _ = _engine.probe_fp64_peak(0, local_rank)
    peak_dmma, _ = _engine.probe_fp64_peak(1, local_rank)
    peak_note = ("measured live: pcb_probe_fp64_peak DFMA register-chain kernel (DMMA m8n8k4 "
                 f"probe: {peak_dmma:.1f} TFLOP/s); MEASURED_PEAKS.json has no FP64 entry")

    # ---- main workload: device-resident value ---------------------------------------------------
    case = make_case(args.workload)
    n = args.queries or case.n_default
    t = Timed(case, dev, local_rank, rank, world, n, args.algo)
    t.run(args.steps, args.warmup, barrier, gpu_index)
    checksum = float(t.out[:: max(1, t.block // 1000)].double().sum().item())
    total_ms, step_ms = max_over_ranks([t.total_ms, t.step_ms])
    value = t.n_total * args.steps / (total_ms * 1e-3)

    # ---- end to end through the host-buffer API -------------------------------------------------
    D, G = case.D, case.G
    e2e = None
    if not getattr(case, "lookup", False):
        e2e_n = min(t.n_local, args.e2e_queries)
        try:  # all ranks pin their host buffers on one box: stay well inside its free memory
            import psutil

            cap = int(0.4 * psutil.virtual_memory().available / max(1, world) / (8 * (D + G)))
            e2e_n = max(min(e2e_n, 1_000_000), min(e2e_n, cap))
        except Exception:  # noqa: BLE001
            pass
        obj = case.obj(local_rank) if not hasattr(case, "tt") else case.tt
        h_pts = _engine.pinned_empty((e2e_n, D), device=local_rank)
        h_out = _engine.pinned_empty((e2e_n, G), device=local_rank)
        src = t.pts if t.block >= e2e_n else device_queries(case.domain, e2e_n, dev, 4321 + rank)
        torch.from_numpy(h_pts).copy_(src[:e2e_n])
        torch.cuda.synchronize(dev)
        for _ in range(max(1, args.warmup)):
            case.api(obj, h_pts, h_out, local_rank)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            case.api(obj, h_pts, h_out, local_rank)
        torch.cuda.synchronize(dev)
        e2e_s = time.perf_counter() - t0
        # the host path must agree with the device-resident path bit for bit
        chk = torch.empty((4096, G), dtype=torch.float64, device=dev)
        case.launch(t.plan, src[:4096].contiguous(), chk)
        same = bool(np.array_equal(h_out[:4096], chk.cpu().numpy()))
        # bare-copy ceiling for the same buffers (all ranks at once, like the timed loop)
        d_in = torch.empty((e2e_n, D), dtype=torch.float64, device=dev)
        d_o = torch.empty((e2e_n, G), dtype=torch.float64, device=dev)
        barrier()
        ceil_s, ceil_how = copy_ceiling(dev, h_pts, h_out, d_in, d_o, barrier=barrier if world > 1 else None)
        del d_in, d_o
        e2e_s, ceil_s = max_over_ranks([e2e_s, ceil_s])
        e2e_total = e2e_n * world
        bytes_step = e2e_n * 8 * (D + G)
        e2e = {"value": e2e_total * args.steps / e2e_s, "unit": UNIT,
               "h2d_bytes_per_step": e2e_n * D * 8, "d2h_bytes_per_step": e2e_n * G * 8,
               "queries_per_gpu_per_step": e2e_n, "host_path_matches_device_path": same,
               "ceiling": {"value": e2e_total / ceil_s, "unit": UNIT,
                           "gbs_per_gpu": bytes_step / ceil_s / 1e9,
                           "gbs_total": world * bytes_step / ceil_s / 1e9,
                           "how": "same pinned buffers, H2D and D2H cudaMemcpyAsync concurrently on "
                                  "two streams, no kernel, all ranks at once; per scheme one "
                                  "barrier then 3 rounds back to back (like the timed loop), best "
                                  f"of 3 schemes (this rank's best: {ceil_how})"},
               "frac_of_ceiling": (e2e_total * args.steps / e2e_s) / (e2e_total / ceil_s),
               "host_pipeline": _engine.host_pipeline_info(local_rank)}
        del h_pts, h_out

    # ---- optional result gather: off the data path, reported separately --------
    gather = None
    if world > 1 and case.scaling == "weak" and len(t.passes) == 1 and not getattr(case, "lookup", False):
        from pychebyshev_b200 import sharding

        g_steps = max(1, min(args.steps, 5))
        recv = (world - 1) * t.block * G * 8
        # (1) the library recipe: kernel, then NCCL all-gather, serialised
        full = torch.empty((world * t.block, G), dtype=t.out.dtype, device=dev)
        dist.all_gather_into_tensor(full, t.out)          # channel set-up, untimed
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        barrier()
        ev[0].record()
        for _ in range(g_steps):
            dist.all_gather_into_tensor(full, t.out)
        ev[1].record()
        for _ in range(g_steps):
            t.step()
            dist.all_gather_into_tensor(full, t.out)
        ev[2].record()
        barrier()
        alone_ms, both_ms = max_over_ranks([ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])])
        nccl_full = full[rank * t.block: rank * t.block + 4096].clone()
        del full
        torch.cuda.empty_cache()
        gather = {"unit": UNIT, "steps": g_steps, "bytes_received_per_rank_per_step": recv,
                  "nccl": {"value": t.n_total * g_steps / (both_ms * 1e-3),
                           "what": "kernel, then all_gather_into_tensor (NCCL) of every rank's (n x G) "
                                   "shard into one replicated tensor, serialised",
                           "gather_ms": alone_ms / g_steps,
                           "gather_gbs_per_rank": recv / (alone_ms / g_steps * 1e-3) / 1e9}}
        # (2) ours: the kernel writes its slice of a replicated tensor, the copy engines push each
        #     chunk to the peers over NVLink underneath the next chunk's kernel (csrc/pcb_peer.cu)
        try:
            # quarter-size chunks, then halving ones: the only push nothing hides is the last
            # chunk's, so it is the smallest (1/32 of the shard)
            frac = [0.0, 0.25, 0.5, 0.75, 0.875, 0.9375, 0.96875, 1.0]
            chunks = len(frac) - 1
            pg = sharding.PeerGather(t.block, G, local_rank)
            cut = [int(t.block * f) for f in frac]

            def peer_step():
                pg.join()                                  # last step's pushes read these rows
                for a, b in zip(cut, cut[1:]):
                    case.launch(t.plan, t.pts[a:b], pg.local[a:b])
                    pg.push(a, b)

            peer_step()
            pg.publish()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record()
            for _ in range(g_steps):
                peer_step()
            pg.join()
            e1.record()
            barrier()
            (peer_ms,) = max_over_ranks([e0.elapsed_time(e1)])
            got = pg.publish()
            same = True
            for r in range(world):                         # every replica holds every rank's rows
                probe = got[r, :: max(1, t.block // 1000)].double().sum().reshape(1)
                allp = [torch.empty_like(probe) for _ in range(world)]
                dist.all_gather(allp, probe)
                same = same and all(bool(torch.equal(allp[0], x)) for x in allp)
            same = same and bool(torch.equal(got[rank, :4096], nccl_full))
            gather["peer"] = {"value": t.n_total * g_steps / (peer_ms * 1e-3), "chunks": chunks,
                              "what": "the kernel writes its slice of a replicated (world, n, G) tensor; "
                                      "copy engines push each chunk into every peer's tensor over NVLink "
                                      "(CUDA IPC mappings, one stream per peer) under the next chunk's "
                                      "kernel; no NCCL on this path",
                              "of_device_resident_value": (t.n_total * g_steps / (peer_ms * 1e-3)) / value,
                              "egress_gbs_per_rank": recv * g_steps / (peer_ms * 1e-3) / 1e9,
                              "replicas_agree": same}
            pg.close()
            del pg
        except Exception as exc:  # noqa: BLE001  (IPC unavailable on this box: NCCL figure stands)
            gather["peer"] = {"unavailable": repr(exc)}
        gather["value"] = max(gather["nccl"]["value"], gather.get("peer", {}).get("value", 0.0))

    # ---- the other configs, each its own clocked measurement ------------------------------------
    del t.pts, t.out
    torch.cuda.empty_cache()
    rows = []
    if not args.no_configs and args.workload == MAIN:
        for key in BLOCK_WORKLOADS:
            c = make_case(key)
            try:
                tc = Timed(c, dev, local_rank, rank, world, 0)
                tc.run(args.config_steps, max(3, args.config_warmup), barrier, gpu_index)
                tot_ms, st_ms = max_over_ranks([tc.total_ms, tc.step_ms])
                row = {"config": c.config, "key": c.key, "workload": c.label, "n_gpus": world,
                       "scaling": c.scaling, "queries_per_step_total": int(tc.n_total),
                       "value": tc.n_total * args.config_steps / (tot_ms * 1e-3), "unit": UNIT,
                       "ms_per_step": tot_ms / args.config_steps, "steps": args.config_steps,
                       "warmup": max(3, args.config_warmup), "gpu_launches": int(tc.launches),
                       "clocks": tc.clocks,
                       "roofline": roofline_of(c, tc.n_local, st_ms, peak_tf, peaks, "as main line")}
                if c.block and c.block < tc.n_local:
                    row["resident_block"] = (f"{len(tc.passes)} launches per step over a resident "
                                             f"{tc.block}-query block ({tc.block * c.D * 8 >> 20} "
                                             "MiB of inputs, far larger than L2)")
                if hasattr(c, "plan_seconds"):
                    row["plan_create_s"] = c.plan_seconds
                if hasattr(c, "plan_info"):
                    row["plan"] = c.plan_info
                rows.append((c, row))
                del tc
            except Exception as exc:  # noqa: BLE001
                rows.append((c, {"config": c.config, "key": c.key, "error": repr(exc)}))
            torch.cuda.empty_cache()

    if rank != 0:
        return

    # DRAM traffic of the dominant kernel per launch, from the committed `ncu --set full` capture
    traffic, traffic_note = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            cap = json.load(f).get(case.kernel.split("+")[0])
        if cap:
            traffic = (cap["dram_bytes_read"] + cap["dram_bytes_write"]) / cap["queries"] * t.block
            traffic_note = (f"bytes per launch = ncu dram read+write per query ({cap['source']}: "
                            f"{cap['bytes_per_query']} B vs {case.bytes_q:.0f} B algorithmic) x "
                            "queries per launch")
    except Exception:  # noqa: BLE001
        pass
    cfg = config_dict(case, n, world)
    cfg["l2"] = (f"inputs {t.block * D * 8 / 2**20:.0f} MiB + outputs {t.block * G * 8 / 2**20:.0f} "
                 "MiB per launch, far larger than the 126 MB L2 (no flush needed)")
    if args.algo:
        cfg["fd_algo"] = args.algo
    roof = roofline_of(case, t.n_local, step_ms, peak_tf, peaks, peak_note)
    roof.update(traffic=traffic, traffic_note=traffic_note, kernel_ms=step_ms / len(t.passes),
                launches_per_step=len(t.passes))
    if hasattr(case, "plan_info"):
        roof["plan"] = case.plan_info
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": case.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg, "e2e": e2e, "gpu_launches": int(t.launches), "clocks": t.clocks,
        "roofline": roof, "checksum": checksum,
    }
    if gather:
        line["gather"] = gather
    if world == 1 and not args.no_cpu:
        try:
            line["cpu_baseline"] = reference_rate(case, seconds=args.cpu_seconds)
            if isinstance(case, TTCase) and case.orders:
                line["cpu_baseline"]["also"] = {
                    "stencil_batches": tt_stencil_batch_rate(case),
                    "note": "BASELINE.md §3 names the per-point eval_multi loop as the reference's "
                            "price+Greeks route; the stencil-batch figure is its fastest own route",
                }
        except Exception as exc:  # noqa: BLE001
            line["cpu_baseline"] = {"unavailable": repr(exc)}
    out_rows = []
    for c, row in rows:
        if world == 1 and not args.no_cpu and "error" not in row:
            try:
                r = reference_rate(c, seconds=args.config_cpu_seconds, procs=1)
                row["cpu_reference"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
            except Exception as exc:  # noqa: BLE001
                row["cpu_reference"] = {"unavailable": repr(exc)}
        out_rows.append(row)
    if out_rows:
        line["configs"] = out_rows
    _emit(line)


_RESULT_FD = None


def _claim_stdout():
    """Keep stdout for the ONE JSON line: everything else that writes to file descriptor 1 (NCCL's
    version banner, library chatter of child processes) is sent to stderr."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def _emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=MAIN)
    ap.add_argument("--queries", type=int, default=0,
                    help="queries per GPU per step (weak) or in total (strong); 0 = workload default")
    ap.add_argument("--e2e-queries", type=int, default=100_000_000)
    ap.add_argument("--algo", type=int, default=0, help="pcb_tt_eval_fd algo (0 auto)")
    ap.add_argument("--cpu-seconds", type=float, default=4.0)
    ap.add_argument("--cpu-step-seconds", type=float, default=0.5,
                    help="--impl reference: wall seconds of reference work per step (sizes the sample)")
    ap.add_argument("--config-steps", type=int, default=3)
    ap.add_argument("--config-warmup", type=int, default=3)
    ap.add_argument("--config-cpu-seconds", type=float, default=1.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        print("note: timing rules ask for >= 3 warm-up steps", file=sys.stderr)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, local_rank, world)
    finally:
        if world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()


if __name__ == "__main__":
    main()
