"""The N>1 path on CPU: world_size-2 gloo, query sharding with the optional result gather.
The evaluator here is the oracle (this test checks the plumbing, not the CUDA kernels)."""

import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import _golden as G
from pychebyshev_b200.sharding import eval_sharded, shard_range


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 100, 1_000_003):
        for world in (1, 2, 3, 8):
            cuts = [shard_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            for (a, b), (c, d) in zip(cuts, cuts[1:]):
                assert b == c and b >= a
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, n, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from oracle import np_oracle as O

        g = G.load("tt_4d_perm")
        cores, domain, dim_order = G.tt_parts(g)
        pts = torch.from_numpy(g["points"][:n])

        def evaluate(shard):
            v = O.tt_eval_batch(cores, domain, dim_order, shard.numpy())
            return torch.from_numpy(v).reshape(-1, 1)

        local, full = eval_sharded(evaluate, pts, 1)
        lo, hi = shard_range(n, rank, world)
        assert local.shape == (hi - lo, 1)
        ref = torch.from_numpy(g["values"][:n]).reshape(-1, 1)
        assert torch.equal(full, ref)  # every rank holds the whole, correctly ordered result
        local2, none = eval_sharded(evaluate, pts, 1, gather=False)
        assert none is None and torch.equal(local2, ref[lo:hi])
        np.save(os.path.join(tmp, f"ok{rank}.npy"), np.array([1]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [1001, 64])
def test_two_rank_gloo_sharding(tmp_path, n):
    port = 29500 + (os.getpid() % 2000) + n % 7
    mp.spawn(_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0.npy") and os.path.exists(tmp_path / "ok1.npy")
