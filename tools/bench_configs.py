#!/usr/bin/env python
"""All five BASELINE.json configs on one B200: device-resident throughput, algorithmic
FLOP/s (or GB/s) and the fraction of the measured roofline, written to
gpurun_out/configs.json and profiles-ready markdown (gpurun_out/configs.md).

This is the companion of bench.py (which reports the headline config only, in the driver's JSON
contract).  Timing: CUDA events on the launching stream, best of `reps` after warm-up, inputs far
larger than L2 except where noted.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import _golden as G  # noqa: E402
import pychebyshev_b200 as pcb  # noqa: E402
from oracle import np_oracle as O  # noqa: E402
from pychebyshev_b200 import _engine, workloads as wl  # noqa: E402


def timeit(fn, reps=4, warm=2):
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
    out = torch.empty((n, len(domain)), dtype=torch.float64, device="cuda")
    step = 1 << 22
    for s in range(0, n, step):
        e = min(n, s + step)
        out[s:e] = lo + (hi - lo) * torch.rand((e - s, len(domain)), generator=gen, device="cuda",
                                               dtype=torch.float64)
    return out


def tt_flop(cores):
    return sum(2 * c.shape[0] * c.shape[1] * c.shape[2] for c in cores)


def tt_shared_flop(cores, active):
    step = [2 * c.shape[0] * c.shape[1] * c.shape[2] for c in cores]
    left, right = sum(step[:max(active)]), sum(step[min(active) + 1:])
    return left + right + sum(step[a] + 2 * cores[a].shape[1] * cores[a].shape[2] for a in active)


def full_flop(n):
    tot, p = 0, 1
    for k in n:  # sum_k prod_{j<=k} n_j, contraction from the last axis
        p *= k
        tot += p
    return 2 * tot


def main():
    rows = []
    peak_fp64, _ = _engine.probe_fp64_peak(0)
    peak_dmma, _ = _engine.probe_fp64_peak(1)
    hbm = 6544.7
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:  # noqa: BLE001
        pass

    def add(config, call, n, ms, flop_q=None, bytes_q=None, bound="fp64", note=""):
        qps = n / ms * 1e3
        row = {"config": config, "call": call, "queries": n, "ms": ms, "queries_per_s": qps,
               "bound": bound, "note": note}
        if bound == "fp64":
            row.update(achieved=flop_q * qps / 1e12, peak=peak_fp64, unit="TFLOP/s",
                       flop_per_query=flop_q)
        else:
            row.update(achieved=bytes_q * qps / 1e9, peak=hbm, unit="GB/s", bytes_per_query=bytes_q)
        row["frac"] = row["achieved"] / row["peak"]
        rows.append(row)
        print(f"{config:28s} {call:34s} {qps:10.3e} q/s  {row['achieved']:8.2f} {row['unit']:8s} "
              f"{100 * row['frac']:5.1f}% of {bound} roofline", flush=True)

    # ---- C2: 5D Black-Scholes TT --------------------------------------------------------------
    g = G.load("tt_bs5d")
    cores, domain, order = G.tt_parts(g)
    tt = pcb.ChebyshevTT.from_cores(cores, domain, order)
    n = 50_000_000
    pts = rand_points(domain, n)
    o1 = torch.empty((n, 1), dtype=torch.float64, device="cuda")
    o4 = torch.empty((n, 4), dtype=torch.float64, device="cuda")
    ms = timeit(lambda: tt._plan().eval_device(pts, o1))
    add("C2 TT 5D BS r<=11", "eval_batch (values)", n, ms, tt_flop(cores))
    plan = tt._plan().with_orders(np.asarray(wl.BS5D_GREEKS), 0)
    ms = timeit(lambda: plan.eval_device(pts, o4))
    add("C2 TT 5D BS r<=11", "price+delta+gamma+vega (algo 2)", n, ms, tt_shared_flop(cores, [0, 3]))
    plan1 = tt._plan().with_orders(np.asarray(wl.BS5D_GREEKS), 1)
    ms = timeit(lambda: plan1.eval_device(pts[:10_000_000], o4[:10_000_000]), reps=2, warm=1)
    add("C2 TT 5D BS r<=11", "price+delta+gamma+vega (algo 1)", 10_000_000, ms, 8 * tt_flop(cores),
        note="8 chain evaluations per query, the reference's own count")
    del pts, o1, o4

    # ---- C5: 10D TT -------------------------------------------------------------------------------
    for name, label in (("tt_basket10d", "C5 TT 10D basket r<=10"), ("tt_rank20_10d", "C5 TT 10D rank 20")):
        g = G.load(name)
        cores, domain, order = G.tt_parts(g)
        tt = pcb.ChebyshevTT.from_cores(cores, domain, order)
        n = 20_000_000 if name == "tt_basket10d" else 8_388_608
        pts = rand_points(domain, n)
        o1 = torch.empty((n, 1), dtype=torch.float64, device="cuda")
        ms = timeit(lambda: tt._plan().eval_device(pts, o1), reps=3)
        add(label, "eval_batch (values)", n, ms, tt_flop(cores))
        orders = g["fd_orders"][:3]  # value, d/dx0, d2/dx9
        active = sorted({order.index(u) for o in orders for u, k in enumerate(o) if k > 0})
        o3 = torch.empty((n, 3), dtype=torch.float64, device="cuda")
        p3 = tt._plan().with_orders(orders, 0)
        ms = timeit(lambda: p3.eval_device(pts, o3), reps=3)
        add(label, "value + 2 FD Greeks (algo 2)", n, ms, tt_shared_flop(cores, active))
        del pts, o1, o3

    # ---- C5: slider ---------------------------------------------------------------------------------
    g = G.load("slider10d")
    part, pivot_value, slides = G.slider_parts(g, O.diff_matrix)
    dom = [list(map(float, r)) for r in g["domain"]]
    nn = [int(v) for v in g["n_nodes"]]
    sl = pcb.ChebyshevSlider.from_slides([s[0] for s in slides], 10, dom, nn, part,
                                         list(g["pivot_point"]), pivot_value)
    n = 50_000_000
    pts = rand_points(dom, n)
    o1 = torch.empty((n, 1), dtype=torch.float64, device="cuda")
    ps = sl._plan([[0] * 10])
    ms = timeit(lambda: ps.eval_device(pts, o1), reps=3)
    add("C5 slider 10D 5x(11x11)", "eval_batch (values)", n, ms, 5 * full_flop([11, 11]),
        note="+ 110 weight-row terms per query (not counted)")
    del pts, o1

    # ---- C1: 11^5 full tensor ---------------------------------------------------------------------
    gg = G.load("full_bs5d")
    nodes = G.split(gg["nodes_cat"], [int(v) for v in gg["n_nodes"]])
    cheb = pcb.ChebyshevApproximation.from_values(wl.grid_values(wl.bs_call_price, nodes), 5,
                                                  wl.BS5D_DOMAIN, wl.BS5D_NODES)
    n = 148 * 256 * 8
    pts = rand_points(wl.BS5D_DOMAIN, n)
    o4 = torch.empty((n, 4), dtype=torch.float64, device="cuda")
    pf = cheb._plan(wl.BS5D_GREEKS, algo=2)
    ms = timeit(lambda: pf.eval_device(pts, o4), reps=3, warm=1)
    add("C1 full 11^5", "price+delta+gamma+vega (DMMA)", n, ms, 4 * full_flop(wl.BS5D_NODES),
        note="tensors L2-resident (4 x 1.5 MB prepared)")
    del pts, o4

    # ---- C3: splines -----------------------------------------------------------------------------------
    for name, label, dom, nnodes in (("spline_bs2d", "C3 spline 2D 2x(15x15)", wl.SPLINE2D_DOMAIN, wl.SPLINE2D_NODES),
                                     ("spline_bs3d", "C3 spline 3D 2x(15^3)", wl.SPLINE3D_DOMAIN, wl.SPLINE3D_NODES)):
        g = G.load(name)
        knots, shape, pieces = G.spline_parts(g, O.diff_matrix)
        sp = pcb.ChebyshevSpline.from_values([p[0] for p in pieces], len(dom), dom, nnodes, knots)
        n = 100_000_000 if len(dom) == 2 else 20_000_000
        pts = rand_points(dom, n)
        D = len(dom)
        if D == 2:
            idx = torch.empty(n, dtype=torch.int32, device="cuda")
            ms = timeit(lambda: sp._plan([[0] * D]).lookup(pts))
            add(label, "piece lookup (int32)", n, ms, bytes_q=8 * D + 4, bound="hbm")
            del idx
        o1 = torch.empty((n, 1), dtype=torch.float64, device="cuda")
        ps = sp._plan([[0] * D])
        ms = timeit(lambda: ps.eval_device(pts, o1), reps=3)
        add(label, "eval_batch (values)", n, ms, full_flop(nnodes),
            note=f"+ {sum(nnodes)} weight-row terms per query (not counted)")
        orders = [[0] * D, [1] + [0] * (D - 1)]
        o2 = torch.empty((n, 2), dtype=torch.float64, device="cuda")
        p2 = sp._plan(orders)
        ms = timeit(lambda: p2.eval_device(pts, o2), reps=3)
        add(label, "value + d/dS", n, ms, 2 * full_flop(nnodes))
        del pts, o1, o2

    # ---- C4: 16^6 full tensor (134 MB per output, HBM/L2-streamed) -----------------------------------
    if "--no-c4" not in sys.argv:
        nodes = [np.asarray(x) for x in G.split(G.load("full_c4_16p6")["nodes_cat"], [16] * 6)]
        cheb = pcb.ChebyshevApproximation.from_values(wl.grid_values(wl.bs6d, nodes), 6, wl.C4_DOMAIN,
                                                      wl.C4_NODES)
        n = 148 * 256
        pts = rand_points(wl.C4_DOMAIN, n)
        o4 = torch.empty((n, 4), dtype=torch.float64, device="cuda")
        pf = cheb._plan(wl.C4_GREEKS, algo=2)
        ms = timeit(lambda: pf.eval_device(pts, o4), reps=2, warm=1)
        add("C4 full 16^6", "price+delta+gamma+vega (DMMA)", n, ms, 4 * full_flop(wl.C4_NODES),
            note="4 x 134 MB prepared tensors streamed per 256-query tile")

    out = {"device": torch.cuda.get_device_name(0), "peak_fp64_dfma_tflops": peak_fp64,
           "peak_fp64_dmma_tflops": peak_dmma, "hbm_gbs": hbm, "rows": rows}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w") as f:
        json.dump(out, f, indent=1)
    with open(os.path.join(ROOT, "gpurun_out", "configs.md"), "w") as f:
        f.write(f"device: {out['device']}; measured FP64 peak {peak_fp64:.1f} TFLOP/s (DFMA), "
                f"{peak_dmma:.1f} (DMMA); HBM {hbm:.0f} GB/s\n\n")
        f.write("| config | call | queries/s | achieved | of roofline | note |\n|---|---|---|---|---|---|\n")
        for r in rows:
            f.write(f"| {r['config']} | {r['call']} | {r['queries_per_s']:.3e} | {r['achieved']:.2f} "
                    f"{r['unit']} | {100 * r['frac']:.1f} % ({r['bound']}) | {r['note']} |\n")


if __name__ == "__main__":
    main()
