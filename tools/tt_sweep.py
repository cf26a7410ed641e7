#!/usr/bin/env python
"""Launch-configuration sweep of the TT kernels on the 5D Black-Scholes TT (exploration tool).

Each configuration runs in its own process because PCB_TT_QPT / PCB_TT_THREADS are read at plan
creation."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import os, sys, json
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np, torch
import _golden as G
import pychebyshev_b200 as pcb
from pychebyshev_b200 import workloads as wl
name = os.environ.get("TT_CASE", "tt_bs5d")
g = G.load(name); cores, domain, dim_order = G.tt_parts(g)
tt = pcb.ChebyshevTT.from_cores(cores, domain, dim_order)
n = int(os.environ.get("TT_N", "16000000"))
gen = torch.Generator(device="cuda").manual_seed(1)
udom = [domain[dim_order.index(u)] for u in range(len(domain))]
lo = torch.tensor([d[0] for d in udom], device="cuda", dtype=torch.float64)
hi = torch.tensor([d[1] for d in udom], device="cuda", dtype=torch.float64)
pts = lo + (hi - lo) * torch.rand((n, len(domain)), generator=gen, device="cuda", dtype=torch.float64)
orders = wl.BS5D_GREEKS if name == "tt_bs5d" else g["fd_orders"][:3]
def timeit(fn, reps=4):
    fn(); fn(); torch.cuda.synchronize(); best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
res = {}
out1 = torch.empty((n, 1), dtype=torch.float64, device="cuda")
outg = torch.empty((n, len(orders)), dtype=torch.float64, device="cuda")
for key, fn in (("value", lambda: tt._plan().eval_device(pts, out1)),
                ("fd2", lambda: tt._plan().with_orders(orders, 2).eval_device(pts, outg)),
                ("fd1", lambda: tt._plan().with_orders(orders, 1).eval_device(pts, outg))):
    if key == "fd1" and os.environ.get("TT_SKIP_FD1"): continue
    try:
        res[key] = n / timeit(fn) * 1e3
    except Exception as e:
        res[key] = f"{type(e).__name__}: {e}"
print(json.dumps(res))
""" % (ROOT, ROOT)


def main():
    configs = [tuple(int(v) for v in c.split('x')) for c in os.environ.get('TT_CONFIGS', '0x0,2x512,2x384,3x320,3x256,4x256,4x192').split(',')]
    for q, t in configs:
        env = dict(os.environ)
        if q:
            # uniform-path (constant-bank) kernels only; the shared-memory kernels keep their own pick
            for suffix in ("_FD", "_VALUE"):
                env["PCB_TT_QPT" + suffix], env["PCB_TT_THREADS" + suffix] = str(q), str(t)
        env["TT_SKIP_FD1"] = "1"
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
        line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr.strip()[-300:]
        try:
            d = json.loads(line)
            fmt = {k: (f"{v:.3e}" if isinstance(v, float) else v) for k, v in d.items()}
        except Exception:  # noqa: BLE001
            fmt = line
        print(f"qpt={q} threads={t}: {fmt}", flush=True)


if __name__ == "__main__":
    main()
