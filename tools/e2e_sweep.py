#!/usr/bin/env python
"""End-to-end (host buffers in, host buffers out) throughput of the headline workload for several
host-pipeline settings, on 1..8 GPUs at once (run under torchrun), next to the bare-copy ceiling
for the same buffers.  Rank 0 prints one JSON document and writes gpurun_out/e2e_sweep_N<world>.json.

    python -m torch.distributed.run --nproc-per-node 8 tools/e2e_sweep.py
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import bench as B  # noqa: E402
from pychebyshev_b200 import _engine  # noqa: E402


def main():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def vmax(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n = int(os.environ.get("E2E_QUERIES", "100000000"))
    case = B.make_case("tt_bs5d")
    case.build(local)
    tt = case.tt
    D, G = case.D, case.G
    h_pts = _engine.pinned_empty((n, D), device=local)
    h_out = _engine.pinned_empty((n, G), device=local)
    src = B.device_queries(case.domain, n, dev, 4321 + rank)
    torch.from_numpy(h_pts).copy_(src)
    torch.cuda.synchronize(dev)
    out = {"n_gpus": world, "queries_per_gpu": n, "bytes_per_query": 8 * (D + G), "rows": []}
    settings = [(64, 3), (16, 4), (32, 4), (128, 3), (256, 2), (32, 6), (64, 2)]
    for chunk_mb, depth in settings:
        _engine.configure_host_pipeline(chunk_mb, depth)
        case.api(tt, h_pts, h_out, local)  # warm-up (allocates the ring)
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            case.api(tt, h_pts, h_out, local)
        torch.cuda.synchronize(dev)
        dt = vmax(time.perf_counter() - t0)
        qps = world * n * 3 / dt
        out["rows"].append({"chunk_mb": chunk_mb, "ring_depth": depth, "queries_per_s": qps,
                            "gbs_total": qps * 8 * (D + G) / 1e9})
        if rank == 0:
            print(f"chunk {chunk_mb} MB x {depth}: {qps:.3e} q/s = {qps * 72 / 1e9:.1f} GB/s", flush=True)
    d_in = torch.empty((n, D), dtype=torch.float64, device=dev)
    d_o = torch.empty((n, G), dtype=torch.float64, device=dev)
    barrier()
    ceil_s, how = B.copy_ceiling(dev, h_pts, h_out, d_in, d_o, barrier=barrier if world > 1 else None)
    ceil_s = vmax(ceil_s)
    out["ceiling"] = {"queries_per_s": world * n / ceil_s, "gbs_total": world * n * 8 * (D + G) / ceil_s / 1e9,
                      "scheme_rank0": how}
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"e2e_sweep_N{world}.json"), "w") as f:
            json.dump(out, f, indent=1)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
