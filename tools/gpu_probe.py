#!/usr/bin/env python
"""One-shot GPU exploration: FP64 pipe peaks + device-resident kernel timings (not the bench)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import _golden as G  # noqa: E402
import pychebyshev_b200 as pcb  # noqa: E402
from pychebyshev_b200 import _engine, workloads as wl  # noqa: E402


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def rand_points(domain, n, seed=1):
    gen = torch.Generator(device="cuda").manual_seed(seed)
    lo = torch.tensor([d[0] for d in domain], device="cuda", dtype=torch.float64)
    hi = torch.tensor([d[1] for d in domain], device="cuda", dtype=torch.float64)
    return lo + (hi - lo) * torch.rand((n, len(domain)), generator=gen, device="cuda", dtype=torch.float64)


def main():
    out = {"device": torch.cuda.get_device_name(0), "host_cpus": os.cpu_count()}
    for kind, name in ((0, "dfma"), (1, "dmma")):
        tf, ms = _engine.probe_fp64_peak(kind)
        out[f"peak_{name}_tflops"] = tf
        print(f"peak {name}: {tf:.2f} TFLOP/s ({ms:.3f} ms)", flush=True)

    which = sys.argv[1:] or ["tt", "full", "small"]
    if "tt" in which:
        for name in ("tt_bs5d", "tt_basket10d", "tt_rank20_10d"):
            g = G.load(name)
            cores, domain, dim_order = G.tt_parts(g)
            tt = pcb.ChebyshevTT.from_cores(cores, domain, dim_order)
            n = 20_000_000 if name == "tt_bs5d" else 4_000_000
            udom = [domain[dim_order.index(u)] for u in range(len(domain))]
            pts = rand_points(udom, n)
            res = torch.empty((n, 1), dtype=torch.float64, device="cuda")
            ms = timeit(lambda: tt._plan().eval_device(pts, res))
            fval = sum(2 * c.shape[0] * c.shape[1] * c.shape[2] for c in cores)
            out[f"{name}_value_qps"] = n / ms * 1e3
            out[f"{name}_value_tflops"] = fval * n / ms * 1e3 / 1e12
            print(f"{name} value: {n / ms * 1e3:.3e} q/s  {fval * n / ms / 1e9:.2f} TFLOP/s algorithmic "
                  f"({ms:.2f} ms for {n})", flush=True)
            # value + single-dim derivative rows (price/delta/gamma/vega for the 5D case)
            orders = g["fd_orders"][:4] if name == "tt_bs5d" else g["fd_orders"][:3]
            for algo in (1, 2):
                try:
                    plan = tt._plan().with_orders(orders, algo)
                    res4 = torch.empty((n, len(orders)), dtype=torch.float64, device="cuda")
                    ms = timeit(lambda: plan.eval_device(pts, res4), reps=3, warm=1)
                    out[f"{name}_fd_algo{algo}_qps"] = n / ms * 1e3
                    print(f"{name} price+3 Greeks algo {algo}: {n / ms * 1e3:.3e} q/s ({ms:.2f} ms)", flush=True)
                except Exception as e:  # noqa: BLE001
                    print(f"{name} fd algo {algo}: {type(e).__name__}: {e}", flush=True)
    if "full" in which:
        g = G.load("full_bs5d")
        nodes = G.split(g["nodes_cat"], [int(v) for v in g["n_nodes"]])
        tensor = wl.grid_values(wl.bs_call_price, nodes)
        cheb = pcb.ChebyshevApproximation.from_values(tensor, 5, wl.BS5D_DOMAIN, wl.BS5D_NODES)
        for algo, n in ((2, 148 * 256 * 4), (1, 200_000)):
            pts = rand_points(wl.BS5D_DOMAIN, n)
            plan = cheb._plan(wl.BS5D_GREEKS, algo=algo)
            res4 = torch.empty((n, 4), dtype=torch.float64, device="cuda")
            ms = timeit(lambda: plan.eval_device(pts, res4), reps=3, warm=1)
            flop = 4 * 354310.0
            out[f"full_bs5d_algo{algo}_qps"] = n / ms * 1e3
            out[f"full_bs5d_algo{algo}_tflops"] = flop * n / ms / 1e9
            print(f"full 11^5 price+3 Greeks algo {algo}: {n / ms * 1e3:.3e} q/s "
                  f"{flop * n / ms / 1e9:.2f} TFLOP/s ({ms:.2f} ms for {n})", flush=True)
    if "c4" in which:
        nodes = [np.asarray(x) for x in G.split(G.load("full_c4_16p6")["nodes_cat"], [16] * 6)]
        tensor = wl.grid_values(wl.bs6d, nodes)
        cheb = pcb.ChebyshevApproximation.from_values(tensor, 6, wl.C4_DOMAIN, wl.C4_NODES)
        n = 148 * 256
        pts = rand_points(wl.C4_DOMAIN, n)
        plan = cheb._plan(wl.C4_GREEKS, algo=2)
        res4 = torch.empty((n, 4), dtype=torch.float64, device="cuda")
        ms = timeit(lambda: plan.eval_device(pts, res4), reps=2, warm=1)
        flop = 4 * 35791392.0
        out["full_c4_qps"] = n / ms * 1e3
        out["full_c4_tflops"] = flop * n / ms / 1e9
        print(f"full 16^6 price+3 Greeks dmma: {n / ms * 1e3:.3e} q/s {flop * n / ms / 1e9:.2f} TFLOP/s "
              f"({ms:.1f} ms for {n})", flush=True)
    if "small" in which:
        g = G.load("spline_bs2d")
        from oracle import np_oracle as O
        knots, shape, pieces = G.spline_parts(g, O.diff_matrix)
        sp = pcb.ChebyshevSpline.from_values([p[0] for p in pieces], 2, wl.SPLINE2D_DOMAIN,
                                             wl.SPLINE2D_NODES, knots)
        n = 50_000_000
        pts = rand_points(wl.SPLINE2D_DOMAIN, n)
        plan = sp._plan([[0, 0]])
        res = torch.empty((n, 1), dtype=torch.float64, device="cuda")
        ms = timeit(lambda: plan.eval_device(pts, res), reps=3, warm=1)
        out["spline2d_value_qps"] = n / ms * 1e3
        print(f"spline 2D value: {n / ms * 1e3:.3e} q/s ({ms:.2f} ms)", flush=True)
        t0 = time.perf_counter()
        idx = plan.lookup(pts)
        torch.cuda.synchronize()
        ms = timeit(lambda: plan.lookup(pts), reps=3, warm=1)
        out["spline2d_lookup_qps"] = n / ms * 1e3
        out["spline2d_lookup_gbs"] = n * 20 / ms / 1e6
        print(f"spline 2D lookup: {n / ms * 1e3:.3e} q/s {n * 20 / ms / 1e6:.0f} GB/s", flush=True)
        del t0, idx
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
