"""Query sharding across the GPUs of one box (one process per GPU, ``torch.distributed``).

Every query is independent and the interpolant is replicated on each GPU, so
the data path has **no collective**: rank ``r`` evaluates the contiguous row range
``shard_range(N, r, world)`` of the batch with its own device plan.  The only optional message is
the gather of the per-rank result shards (``all_gather_into_tensor`` over NCCL on GPUs -- NVLink 5
/ NVSwitch -- or gloo in the CPU tests), which callers keep off the timed path.
"""

from __future__ import annotations

from typing import Callable, Tuple


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of ``range(n)``: the first ``n % world`` ranks get one extra row."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of size {world}")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def eval_sharded(evaluate: Callable, points, n_outputs: int, group=None, gather: bool = True):
    """Evaluate this rank's shard of ``points`` (an (N, D) tensor every rank holds, or can
    generate) with ``evaluate(shard) -> (n_shard, n_outputs)`` tensor and optionally gather.

    Returns ``(local, full)``; ``full`` is ``None`` when ``gather`` is false.  The gather pads the
    shards to equal length (``all_gather_into_tensor`` needs uniform sizes) and trims afterwards.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = points.shape[0]
    lo, hi = shard_range(n, rank, world)
    local = evaluate(points[lo:hi])
    if local.shape != (hi - lo, n_outputs):
        raise ValueError(f"evaluate returned {tuple(local.shape)}, expected {(hi - lo, n_outputs)}")
    if not gather or world == 1:
        return local, (local if gather else None)
    width = -(-n // world)  # ceil
    padded = torch.zeros((width, n_outputs), dtype=local.dtype, device=local.device)
    padded[: hi - lo] = local
    gathered = torch.empty((world * width, n_outputs), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, padded, group=group)
    parts = []
    for r in range(world):
        rlo, rhi = shard_range(n, r, world)
        parts.append(gathered[r * width: r * width + (rhi - rlo)])
    return local, torch.cat(parts, dim=0)
