#!/usr/bin/env python
"""Small, deterministic launch sequence of one kernel for ncu captures.

usage: ncu_target.py {tt_value|tt_fd1|tt_fd2|full_dmma|full_fma|spline|spline_value|lookup|slider} [N]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import _golden as G  # noqa: E402
import pychebyshev_b200 as pcb  # noqa: E402
from pychebyshev_b200 import workloads as wl  # noqa: E402


def rand_points(domain, n, seed=1):
    gen = torch.Generator(device="cuda").manual_seed(seed)
    lo = torch.tensor([d[0] for d in domain], device="cuda", dtype=torch.float64)
    hi = torch.tensor([d[1] for d in domain], device="cuda", dtype=torch.float64)
    return lo + (hi - lo) * torch.rand((n, len(domain)), generator=gen, device="cuda", dtype=torch.float64)


def main():
    which = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    reps = 4
    if which.startswith("tt"):
        name = os.environ.get("TT_CASE", "tt_bs5d")
        g = G.load(name)
        cores, domain, dim_order = G.tt_parts(g)
        tt = pcb.ChebyshevTT.from_cores(cores, domain, dim_order)
        n = n or 8_000_000
        udom = [domain[dim_order.index(u)] for u in range(len(domain))]
        pts = rand_points(udom, n)
        if which == "tt_value":
            fn = lambda: tt.eval_batch(pts)  # noqa: E731
        else:
            orders = wl.BS5D_GREEKS if name == "tt_bs5d" else g["fd_orders"][:3]
            algo = 1 if which == "tt_fd1" else 2
            fn = lambda: tt.eval_multi_batch(pts, orders, algo=algo)  # noqa: E731
    elif which in ("full_dmma", "full_fma", "full_dmma2"):
        g = G.load("full_bs5d")
        nodes = G.split(g["nodes_cat"], [int(v) for v in g["n_nodes"]])
        tensor = wl.grid_values(wl.bs_call_price, nodes)
        cheb = pcb.ChebyshevApproximation.from_values(tensor, 5, wl.BS5D_DOMAIN, wl.BS5D_NODES)
        algo = {"full_dmma": 2, "full_dmma2": 3}.get(which, 1)
        n = n or (148 * 256 if algo >= 2 else 20000)
        pts = rand_points(wl.BS5D_DOMAIN, n)
        fn = lambda: cheb.eval_batch_multi(pts, wl.BS5D_GREEKS, algo=algo)  # noqa: E731
    elif which == "slider":
        from oracle import np_oracle as O

        g = G.load("slider10d")
        part, pivot_value, slides = G.slider_parts(g, O.diff_matrix)
        dom = [list(map(float, r)) for r in g["domain"]]
        sl = pcb.ChebyshevSlider.from_slides([s[0] for s in slides], 10, dom, [int(v) for v in g["n_nodes"]],
                                             part, list(g["pivot_point"]), pivot_value)
        n = n or 10_000_000
        pts = rand_points(dom, n)
        fn = lambda: sl.eval_batch(pts, [0] * 10)  # noqa: E731
    elif which == "spline3d_value":
        from oracle import np_oracle as O

        g = G.load("spline_bs3d")
        knots, shape, pieces = G.spline_parts(g, O.diff_matrix)
        sp = pcb.ChebyshevSpline.from_values([p[0] for p in pieces], 3, wl.SPLINE3D_DOMAIN,
                                             wl.SPLINE3D_NODES, knots)
        n = n or 8_000_000
        pts = rand_points(wl.SPLINE3D_DOMAIN, n)
        fn = lambda: sp.eval_batch(pts, [0, 0, 0])  # noqa: E731
    else:
        from oracle import np_oracle as O

        g = G.load("spline_bs2d")
        knots, shape, pieces = G.spline_parts(g, O.diff_matrix)
        sp = pcb.ChebyshevSpline.from_values([p[0] for p in pieces], 2, wl.SPLINE2D_DOMAIN,
                                             wl.SPLINE2D_NODES, knots)
        n = n or 20_000_000
        pts = rand_points(wl.SPLINE2D_DOMAIN, n)
        if which == "lookup":
            fn = lambda: sp.find_pieces(pts)  # noqa: E731
        elif which == "spline_value":
            fn = lambda: sp.eval_batch(pts, [0, 0])  # noqa: E731
        else:
            fn = lambda: sp.eval_batch_multi(pts, [[0, 0], [1, 0]])  # noqa: E731
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    print("done", which, n)


if __name__ == "__main__":
    main()
