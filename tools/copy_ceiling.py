#!/usr/bin/env python
"""Host<->device copy ceiling of this box, per GPU and in aggregate (N = 1/2/4/8 processes).

What the end-to-end path (host NumPy buffers in, host buffers out) can reach at best: every rank
moves the byte counts of the headline workload (40 B/query in, 32 B/query out) between pinned host
memory and its GPU with bare ``cudaMemcpyAsync`` calls -- no kernel -- in chunks, H2D and D2H
concurrently on two streams.  Reported: H2D alone, D2H alone, both at once, per rank and summed,
with the pinned buffers (a) allocated wherever the process happens to run and (b) first-touched on
the GPU's NUMA node with the driving thread bound there (pychebyshev_b200._numa).

    python tools/copy_ceiling.py                       # one GPU
    torchrun --nproc-per-node 8 tools/copy_ceiling.py  # all GPUs at once

Rank 0 prints one JSON document (also written to gpurun_out/copy_ceiling_N<world>.json) and, at
N = 1, the box topology (nvidia-smi topo -m, lscpu, NUMA nodes, each GPU's PCIe root).
"""
import json
import os
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from pychebyshev_b200 import _numa  # noqa: E402


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=30).stdout.strip()
    except Exception as exc:  # noqa: BLE001
        return f"<{exc}>"


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

    n = int(os.environ.get("COPY_QUERIES", "50000000"))
    b_in, b_out = 40 * n, 32 * n
    d_in = torch.empty(b_in, dtype=torch.uint8, device=dev)
    d_out = torch.empty(b_out, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    results = {}

    def measure(h_in, h_out, chunk, mode):
        def go():
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    for o in range(0, b_in, chunk):
                        d_in[o:o + chunk].copy_(h_in[o:o + chunk], non_blocking=True)
            if mode in ("d2h", "both"):
                oc = chunk * 4 // 5
                with torch.cuda.stream(s2):
                    for o in range(0, b_out, oc):
                        h_out[o:o + oc].copy_(d_out[o:o + oc], non_blocking=True)
        go()
        best = 1e30
        for _ in range(3):
            barrier()
            t0 = time.perf_counter()
            go()
            torch.cuda.synchronize(dev)
            best = min(best, time.perf_counter() - t0)
        moved = (b_in if mode != "d2h" else 0) + (b_out if mode != "h2d" else 0)
        return moved / best / 1e9

    quick = os.environ.get("COPY_QUICK") == "1"   # only: default placement, 64 MB chunks, both directions
    for placement in (("default",) if quick else ("default", "numa_local")):
        if placement == "numa_local":
            bound = _numa.bind_thread(local)
            if not bound and world > 1 and rank != 0:
                pass
        with _numa.bound_to_device(local if placement == "numa_local" else None):
            h_in = torch.empty(b_in, dtype=torch.uint8, pin_memory=True)
            h_out = torch.empty(b_out, dtype=torch.uint8, pin_memory=True)
            h_in[::4096] = 1
            h_out[::4096] = 1
        for chunk_mb in ((64,) if quick else (16, 64, 256)):
            for mode in (("both",) if quick else ("h2d", "d2h", "both")):
                if chunk_mb != 64 and mode != "both":
                    continue
                results[f"{placement}/{chunk_mb}MB/{mode}"] = measure(h_in, h_out, chunk_mb << 20, mode)
        del h_in, h_out
        torch._C._host_emptyCache() if hasattr(torch._C, "_host_emptyCache") else None

    keys = sorted(results)
    mine = torch.tensor([results[k] for k in keys], dtype=torch.float64, device=dev)
    if world > 1:
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
    else:
        allr = [mine]
    node = torch.tensor([float(_numa.gpu_numa_node(local))], dtype=torch.float64, device=dev)
    nodes = [torch.empty_like(node) for _ in range(world)]
    if world > 1:
        dist.all_gather(nodes, node)
    else:
        nodes = [node]
    if rank == 0:
        doc = {"n_gpus": world, "queries_per_rank": n, "bytes_in": b_in, "bytes_out": b_out,
               "host_cpus": os.cpu_count(), "numa_nodes": _numa.n_nodes(),
               "gpu_numa_node_per_rank": [int(v.item()) for v in nodes],
               "unit": "GB/s (H2D+D2H bytes / wall)",
               "per_rank": {k: [float(r[i]) for r in allr] for i, k in enumerate(keys)},
               "total": {k: float(sum(r[i] for r in allr)) for i, k in enumerate(keys)}}
        best = max((v, k) for k, v in doc["total"].items() if k.endswith("/both"))
        doc["ceiling_both_gbs_total"] = best[0]
        doc["ceiling_config"] = best[1]
        doc["ceiling_queries_per_s"] = best[0] * 1e9 / 72.0
        if world == 1:
            doc["topology"] = {
                "nvidia_smi_topo": sh("nvidia-smi topo -m"),
                "lscpu": sh("lscpu | head -40"),
                "numa": sh("cat /sys/devices/system/node/node*/cpulist; "
                           "grep MemTotal /sys/devices/system/node/node*/meminfo"),
                "gpu_pci": sh("nvidia-smi --query-gpu=index,pci.bus_id,pcie.link.gen.current,"
                              "pcie.link.width.current,pcie.link.gen.max --format=csv"),
                "gpu_numa": sh("for d in /sys/bus/pci/devices/*; do if grep -qi 0x10de $d/vendor "
                               "2>/dev/null && grep -q 0x0302 $d/class 2>/dev/null; then echo "
                               "$(basename $d) numa=$(cat $d/numa_node); fi; done"),
                "mem": sh("free -g | head -2"),
                "affinity": sorted(os.sched_getaffinity(0)),
            }
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        tag = f"_q{n // 1000000}M" if quick else ""
        with open(os.path.join(ROOT, "gpurun_out", f"copy_ceiling_N{world}{tag}.json"), "w") as f:
            json.dump(doc, f, indent=1)
        print(json.dumps({k: doc[k] for k in doc if k != "topology"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
