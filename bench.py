#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

Metric : fp64 queries/sec, price + Greeks (price, delta, gamma, vega)
Config : configs[1] -- 5D Black-Scholes ChebyshevTT (TT-Cross cores, ranks [1,11,11,11,7,1]),
         1e8 synthetic uniformly sampled queries per GPU per step (weak scaling over 1-8 B200).

A step = one pass of the hot path (``ChebyshevTT.eval_multi_batch`` -> ``pcb_tt_eval_fd``) over one
batch of synthetic queries.

  value        device-resident inputs/outputs, CUDA events on the launching stream, max over ranks
  e2e          the same call with HOST (pinned) NumPy buffers: H2D + kernel + D2H inside the timing
  roofline     FP64-pipe roofline of the dominant kernel, peak measured live with the DFMA probe
  cpu_baseline the oracle port of the reference's NumPy path on this box's host cores (rank 0, N=1)

``--impl reference`` times the CPU port alone (all host cores) and prints the same line shape.

Other workloads for exploration: ``--workload full_bs5d | full_c4 | spline2d | tt_value``.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "fp64 queries/sec (price+Greeks)"
UNIT = "queries/s"


# ------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------

def load_tt_bs5d():
    """Reference-built cores of the 5D Black-Scholes TT (tests/golden/tt_bs5d.npz)."""
    import _golden as G

    g = G.load("tt_bs5d")
    cores, domain, dim_order = G.tt_parts(g)
    return cores, domain, dim_order


def tt_flops_per_query(cores, n_stencil):
    """Algorithmic flop: E * sum_k 2 r_{k-1} n_k r_k."""
    return n_stencil * sum(2 * c.shape[0] * c.shape[1] * c.shape[2] for c in cores)


def tt_shared_flops_per_query(cores, active):
    """Reduced count when left/right partial products are shared across the stencil:
    left chain up to the last active dim, right chain down to the first, plus one coefficient
    contraction per active dim."""
    step = [2 * c.shape[0] * c.shape[1] * c.shape[2] for c in cores]
    if not active:
        return sum(step)
    left = sum(step[:max(active)])
    right = sum(step[min(active) + 1:])
    coeff = sum(step[a] + 2 * cores[a].shape[1] * cores[a].shape[2] for a in active)
    return left + right + coeff


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------

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

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def visible_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:  # noqa: BLE001
            return local_rank
    return local_rank


# ------------------------------------------------------------------------------------------
# CPU baseline (oracle port of the reference's NumPy path)
# ------------------------------------------------------------------------------------------

def _cpu_worker(args):
    cores, domain, dim_order, pts, orders = args
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    from oracle import np_oracle as O

    t0 = time.perf_counter()
    O.tt_eval_multi_batch(cores, domain, dim_order, pts, orders)
    return time.perf_counter() - t0


def cpu_baseline_tt(cores, domain, dim_order, orders, per_core=600, seed=99):
    """Reference algorithm (per-point eval_multi, tensor_train.py:2267-2463) as restated in
    oracle/np_oracle.py, one process per host core, each on its own shard of the sample."""
    from concurrent.futures import ProcessPoolExecutor

    from pychebyshev_b200 import workloads as wl

    ncores = os.cpu_count() or 1
    udom = [domain[dim_order.index(u)] for u in range(len(domain))]
    pts = wl.uniform_queries(udom, per_core * ncores, seed)
    shards = np.array_split(pts, ncores)
    jobs = [(cores, domain, dim_order, s, orders) for s in shards]
    with ProcessPoolExecutor(ncores) as pool:
        list(pool.map(_cpu_worker, jobs[:ncores]))  # warm-up: imports, page-in
        t0 = time.perf_counter()
        list(pool.map(_cpu_worker, jobs))
        wall = time.perf_counter() - t0
    return {
        "value": len(pts) / wall, "unit": UNIT, "cores": ncores, "kind": "port",
        "sample": f"{len(pts)} uniform queries x {len(orders)} outputs, oracle/np_oracle.py "
                  f"tt_eval_multi_batch (reference eval_multi per point), {ncores} processes",
        "seconds": wall,
    }


def cpu_baseline_c(cores, domain, dim_order, orders, n=200_000, seed=98):
    """Context only: the plain-C restatement (oracle/c/oracle.c) on all host threads."""
    from oracle import c_oracle as C
    from pychebyshev_b200 import workloads as wl

    ncores = os.cpu_count() or 1
    udom = [domain[dim_order.index(u)] for u in range(len(domain))]
    pts = wl.uniform_queries(udom, n, seed)
    tt = C.TT(cores, domain, dim_order)
    tt.eval_multi_batch(pts[:1000], orders, threads=ncores)
    t0 = time.perf_counter()
    tt.eval_multi_batch(pts, orders, threads=ncores)
    wall = time.perf_counter() - t0
    return {"value": n / wall, "unit": UNIT, "cores": ncores, "kind": "port-c",
            "sample": f"{n} uniform queries, oracle/c/oracle.c, {ncores} threads"}


# ------------------------------------------------------------------------------------------
# arms
# ------------------------------------------------------------------------------------------

def run_reference(args, rank, world):
    if rank != 0:
        return
    from pychebyshev_b200 import workloads as wl

    cores, domain, dim_order = load_tt_bs5d()
    orders = wl.BS5D_GREEKS
    per_step = []
    for i in range(args.warmup + args.steps):
        res = cpu_baseline_tt(cores, domain, dim_order, orders, per_core=args.cpu_per_core,
                              seed=100 + i)
        if i >= args.warmup:
            per_step.append(res)
    total_q = sum(r["value"] * r["seconds"] for r in per_step)
    total_s = sum(r["seconds"] for r in per_step)
    value = total_q / total_s
    last = per_step[-1]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_s / len(per_step),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "5D Black-Scholes ChebyshevTT ranks [1,11,11,11,7,1], "
                               "price+delta+gamma+vega per query (reference eval_multi)",
                   "queries_per_step": int(last["value"] * last["seconds"] + 0.5)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": last["cores"], "kind": "port",
                         "sample": last["sample"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    import pychebyshev_b200 as pcb
    from pychebyshev_b200 import _engine, workloads as wl

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cores, domain, dim_order = load_tt_bs5d()
    orders = wl.BS5D_GREEKS
    G = len(orders)
    D = len(domain)
    tt = pcb.ChebyshevTT.from_cores(cores, domain, dim_order, device=local_rank)
    n = args.queries
    # synthetic uniform queries, generated on the device from a per-rank seed
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    udom = [domain[dim_order.index(u)] for u in range(D)]
    lo = torch.tensor([d[0] for d in udom], device=dev, dtype=torch.float64)
    hi = torch.tensor([d[1] for d in udom], device=dev, dtype=torch.float64)
    pts = torch.empty((n, D), dtype=torch.float64, device=dev)
    step_rows = 1 << 22
    for s in range(0, n, step_rows):
        e = min(n, s + step_rows)
        pts[s:e] = lo + (hi - lo) * torch.rand((e - s, D), generator=gen, device=dev,
                                               dtype=torch.float64)
    out = torch.empty((n, G), dtype=torch.float64, device=dev)
    plan = tt._plan(local_rank).with_orders(np.asarray(orders), args.algo)
    args.algo = plan.resolved_algo()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident value ------------------------------------------------------------
    for _ in range(args.warmup):
        plan.eval_device(pts, out)
    barrier()
    launches0 = _engine.launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(args.steps)]
    with ClockSampler(visible_gpu_index(local_rank)) as clocks:
        t_all0, t_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_all0.record()
        for e0, e1 in evs:
            e0.record()
            plan.eval_device(pts, out)
            e1.record()
        t_all1.record()
        barrier()
    launches = _engine.launch_count() - launches0
    total_ms = t_all0.elapsed_time(t_all1)
    kernel_ms = float(np.mean([e0.elapsed_time(e1) for e0, e1 in evs]))
    checksum = float(out[:: max(1, n // 1000)].sum().item())

    # ---- end to end through the host-buffer API ---------------------------------------------
    e2e_n = min(n, args.e2e_queries)
    try:  # all ranks pin their host buffers on one box: stay well inside its free memory
        import psutil

        cap = int(0.4 * psutil.virtual_memory().available / max(1, world) / (8 * (D + G)))
        e2e_n = max(1_000_000, min(e2e_n, cap))
    except Exception:  # noqa: BLE001
        pass
    h_pts = _engine.pinned_empty((e2e_n, D))
    h_out = _engine.pinned_empty((e2e_n, G))
    torch.from_numpy(h_pts).copy_(pts[:e2e_n])
    torch.cuda.synchronize(dev)
    for _ in range(max(1, args.warmup)):
        tt.eval_multi_batch(h_pts, orders, out=h_out, device=local_rank, algo=args.algo)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        tt.eval_multi_batch(h_pts, orders, out=h_out, device=local_rank, algo=args.algo)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    # the host path must agree with the device-resident path bit for bit
    same = bool(np.array_equal(h_out[:4096], out[:4096].cpu().numpy()))

    # ---- max over ranks --------------------------------------------------------------------
    stats = torch.tensor([total_ms, e2e_s, kernel_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    total_ms, e2e_s, kernel_ms = (float(v) for v in stats.tolist())
    if rank != 0:
        return

    value = world * n * args.steps / (total_ms * 1e-3)
    e2e_value = world * e2e_n * args.steps / e2e_s
    peak_tf, _ = _engine.probe_fp64_peak(0, local_rank)
    peak_dmma, _ = _engine.probe_fp64_peak(1, local_rank)
    active = sorted({dim_order.index(u) for o in orders for u, k in enumerate(o) if k > 0})
    if args.algo == 1:
        n_evals = sum(1 if not any(o) else (2 if max(o) == 1 else 3) for o in orders)
        flop_q = tt_flops_per_query(cores, n_evals)
        flop_note = f"{n_evals} chain evaluations x {tt_flops_per_query(cores, 1)} flop"
    else:
        flop_q = tt_shared_flops_per_query(cores, active)
        flop_note = "shared left/right partial products (reduced count)"
    achieved_tf = flop_q * n / (kernel_ms * 1e-3) / 1e12
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:  # noqa: BLE001
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_bytes = 8.0 * (D + G)
    # DRAM traffic of the dominant kernel per launch, from the committed `ncu --set full` capture
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            cap = json.load(f)["ttc_fd_shared_kernel"]
        if args.algo == 2 and plan.info()["uniform_path_fd"]:
            traffic = (cap["dram_bytes_read"] + cap["dram_bytes_write"]) / cap["queries"] * n
    except Exception:  # noqa: BLE001
        pass
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": "5D Black-Scholes ChebyshevTT (reference TT-Cross cores, ranks "
                        "[1,11,11,11,7,1]): price+delta+gamma+vega per query via "
                        "eval_multi_batch -> pcb_tt_eval_fd",
            "queries_per_gpu_per_step": n, "outputs_per_query": G, "fd_algo": args.algo,
            "l2": f"inputs {n * D * 8 / 2**20:.0f} MiB + outputs {n * G * 8 / 2**20:.0f} MiB per "
                  "step, far larger than the 126 MB L2 (no flush needed)",
        },
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_n * D * 8,
                "d2h_bytes_per_step": e2e_n * G * 8, "queries_per_gpu_per_step": e2e_n,
                "host_path_matches_device_path": same},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
        "roofline": {
            "bound": "fp64", "bound_class": "tensor",  # compute-bound: the FP64 pipe, which DFMA and the
            # FP64 tensor-core instruction (DMMA) share -- same measured peak, SURVEY.md 8(d)
            "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
            "frac": achieved_tf / peak_tf, "traffic": traffic,
            "traffic_note": "bytes per launch = ncu dram read+write per query (profiles/r1_traffic.json, "
                            "67.8 B vs 72 B algorithmic) x queries per launch",
            "kernel": ("tt_fd_general_kernel" if args.algo == 1 else
                       ("ttc_fd_shared_kernel (cores in the constant bank, LDCU -> DFMA)"
                        if plan.info()["uniform_path_fd"] else "tt_fd_shared_kernel")),
            "plan": plan.info(),
            "kernel_ms": kernel_ms, "flop_per_query": flop_q, "flop_model": flop_note,
            "peak_source": "measured live: pcb_probe_fp64_peak DFMA register-chain kernel "
                           f"(DMMA m8n8k4 probe: {peak_dmma:.1f} TFLOP/s); MEASURED_PEAKS.json "
                           "has no FP64 entry",
            "hbm": {"achieved": hbm_bytes * n / (kernel_ms * 1e-3) / 1e9, "peak": hbm_peak,
                    "unit": "GB/s", "bytes_per_query": hbm_bytes,
                    "frac": hbm_bytes * n / (kernel_ms * 1e-3) / 1e9 / hbm_peak},
        },
        "checksum": checksum,
    }
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline_tt(cores, domain, dim_order, orders,
                                               per_core=args.cpu_per_core)
        line["cpu_baseline"].pop("seconds", None)
        try:
            line["cpu_baseline_c"] = cpu_baseline_c(cores, domain, dim_order, orders)
        except Exception as exc:  # noqa: BLE001
            line["cpu_baseline_c"] = {"unavailable": str(exc)}
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
    ap.add_argument("--queries", type=int, default=100_000_000, help="queries per GPU per step")
    ap.add_argument("--e2e-queries", type=int, default=100_000_000)
    ap.add_argument("--algo", type=int, default=0, help="pcb_tt_eval_fd algo (0 auto)")
    ap.add_argument("--cpu-per-core", type=int, default=600)
    ap.add_argument("--no-cpu", action="store_true")
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
