"""The optional result gather without a collective: replicated result tensors mapped through CUDA
IPC, written by the evaluators and replicated by the copy engines (sharding.PeerGather,
csrc/pcb_peer.cu).  Two processes share the one GPU of the test box (IPC works between processes on
the same device exactly as between devices); torch.distributed/gloo only carries the handles."""

import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import _golden as G

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, n, chunks, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        import pychebyshev_b200 as pcb
        from pychebyshev_b200.sharding import PeerGather, shard_range

        dev = rank % torch.cuda.device_count()
        torch.cuda.set_device(dev)
        g = G.load("tt_bs5d")
        cores, domain, dim_order = G.tt_parts(g)
        orders = np.asarray(g["fd_orders"][:4], dtype=np.int32)
        tt = pcb.ChebyshevTT.from_cores(cores, domain, dim_order, device=dev)
        plan = tt._plan(dev).with_orders(orders, 0)
        reps = -(-n // len(g["points"]))
        pts_all = torch.from_numpy(np.tile(g["points"], (reps, 1))[:n]).to(f"cuda:{dev}")
        whole = torch.empty((n, 4), dtype=torch.float64, device=pts_all.device)
        plan.eval_device(pts_all, whole)                      # single-process answer

        rows = -(-n // world)
        pg = PeerGather(rows, 4, dev)
        assert pg.full.shape == (world, rows, 4) and pg.local.data_ptr() == pg.full[rank].data_ptr()
        pg.local.fill_(float("nan"))
        lo, hi = shard_range(n, rank, world)
        mine = pts_all[lo:hi].contiguous()
        step = -(-(hi - lo) // chunks)
        for c in range(0, hi - lo, step):
            e = min(hi - lo, c + step)
            plan.eval_device(mine[c:e], pg.local[c:e])        # the kernel writes the replica itself
            pg.push(c, e)                                     # ... and the copy engines spread it
        full = pg.publish()
        for r in range(world):
            rlo, rhi = shard_range(n, r, world)
            assert torch.equal(full[r, : rhi - rlo], whole[rlo:rhi]), f"rank {rank}: slice of rank {r}"
        with pytest.raises(ValueError):
            pg.push(0, rows + 1)
        # a second round over the same tensors (join before overwriting, then push again)
        pg.join()
        pg.local.zero_()
        pg.push()
        full = pg.publish()
        assert float(full.abs().sum()) == 0.0
        pg.close()
        np.save(os.path.join(tmp, f"ok{rank}.npy"), np.array([1]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,chunks", [(4096, 4), (1001, 3)])
def test_two_process_peer_gather(tmp_path, n, chunks):
    port = 29700 + (os.getpid() % 2000) + n % 7
    mp.spawn(_worker, args=(2, port, n, chunks, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0.npy") and os.path.exists(tmp_path / "ok1.npy")
