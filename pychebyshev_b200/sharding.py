"""Query sharding across the GPUs of one box (one process per GPU, ``torch.distributed``).

Every query is independent and the interpolant is replicated on each GPU, so
the data path has **no collective**: rank ``r`` evaluates the contiguous row range
``shard_range(N, r, world)`` of the batch with its own device plan.  The only optional message is
the gather of the per-rank result shards, which callers keep off the timed path: either
``all_gather_into_tensor`` (NCCL on GPUs, gloo in the CPU tests; ``eval_sharded``), or
``PeerGather`` -- replicated result tensors mapped through CUDA IPC that the evaluators write
directly and the copy engines replicate over NVLink while the next chunk is being evaluated.
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


class _DeviceArray:
    """``__cuda_array_interface__`` holder for memory this library allocated itself."""

    def __init__(self, ptr: int, shape):
        self.__cuda_array_interface__ = {"shape": tuple(int(s) for s in shape), "typestr": "<f8",
                                         "data": (int(ptr), False), "version": 2, "strides": None}


class PeerGather:
    """Every rank's result shard replicated on every GPU of the box WITHOUT a collective.

    Each rank owns ``full`` -- a ``(world, rows, n_outputs)`` float64 tensor in its own HBM -- and
    has the peers' tensors mapped through CUDA IPC (``pcb_peer_alloc`` / ``pcb_peer_open``).  The
    evaluators write straight into ``local = full[rank]``; ``push(lo, hi)`` then copies rows
    ``[lo, hi)`` of that slice into the same place of every peer's tensor with the copy engines
    over NVLink, underneath whatever the SMs do next.  ``publish()`` waits for the pushes and for
    the other ranks: after it, ``full`` is complete on every rank, rank-major like the
    concatenation ``eval_sharded`` returns.

    One process per GPU on one box; ``torch.distributed`` (any backend) only carries the 64-byte
    handles at construction and the final barrier.
    """

    def __init__(self, rows: int, n_outputs: int, device: int, group=None):
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import _lib

        if not dist.is_initialized():
            raise RuntimeError("PeerGather needs an initialised torch.distributed process group")
        self._lib, self._group, self.device = _lib.load(), group, int(device)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.rows, self.n_outputs = int(rows), int(n_outputs)
        if self.rows < 1 or self.n_outputs < 1:
            raise ValueError("rows and n_outputs must be >= 1")
        self._slice_bytes = self.rows * self.n_outputs * 8
        # Every step that can fail on one rank only (allocation, IPC export, mapping a peer) is
        # followed by an exchange of the outcome, so that all ranks raise together instead of one
        # leaving the others in a barrier.
        ptr, handle, err = C.c_void_p(), C.create_string_buffer(64), None
        try:
            _lib.check(self._lib.pcb_peer_alloc(self.device, self.world * self._slice_bytes,
                                                C.byref(ptr), handle))
        except Exception as exc:  # noqa: BLE001
            err = repr(exc)
        self._ptr, self._peers = ptr.value, []
        handles = [None] * self.world
        dist.all_gather_object(handles, (err, handle.raw, self.rows, self.n_outputs), group=group)
        if err is None:
            for r, (err_r, raw, rows_r, g_r) in enumerate(handles):
                if err_r is not None:
                    err = f"rank {r}: {err_r}"
                elif (rows_r, g_r) != (self.rows, self.n_outputs):
                    err = f"rank {r} gathers {(rows_r, g_r)}, this rank {(self.rows, self.n_outputs)}"
                if err is not None:
                    break
                if r == self.rank:
                    continue
                p = C.c_void_p()
                try:
                    _lib.check(self._lib.pcb_peer_open(self.device, raw, C.byref(p)))
                except Exception as exc:  # noqa: BLE001
                    err = f"mapping rank {r}'s tensor: {exc!r}"
                    break
                self._peers.append(p.value)
        outcomes = [None] * self.world
        dist.all_gather_object(outcomes, err, group=group)   # also: every mapping exists from here on
        bad = [f"[rank {r}] {e}" for r, e in enumerate(outcomes) if e is not None]
        if bad:
            for p in self._peers:
                self._lib.pcb_peer_close(self.device, p)
            self._peers = []
            dist.barrier(group=group)
            if self._ptr:
                self._lib.pcb_peer_free(self.device, self._ptr)
            self._ptr = None
            raise RuntimeError("PeerGather unavailable: " + "; ".join(bad))
        self._peer_array = (C.c_void_p * max(1, len(self._peers)))(*self._peers)
        self.full = torch.as_tensor(_DeviceArray(self._ptr, (self.world, self.rows, self.n_outputs)),
                                    device=torch.device("cuda", self.device))
        self.local = self.full[self.rank]

    def _stream(self, stream):
        import torch

        return (stream or torch.cuda.current_stream(self.device)).cuda_stream

    def push(self, lo: int = 0, hi: int | None = None, stream=None) -> None:
        """Copy rows ``[lo, hi)`` of ``local`` to every peer, ordered after what is already enqueued
        on ``stream`` (default: the current stream); the stream itself does not wait."""
        hi = self.rows if hi is None else hi
        if not (0 <= lo <= hi <= self.rows):
            raise ValueError(f"rows [{lo}, {hi}) outside [0, {self.rows})")
        from . import _lib

        off = self.rank * self._slice_bytes + lo * self.n_outputs * 8
        _lib.check(self._lib.pcb_peer_push(self.device, len(self._peers), self._peer_array, off,
                                           self._ptr + off, (hi - lo) * self.n_outputs * 8,
                                           self._stream(stream)))

    def join(self, stream=None) -> None:
        """Make ``stream`` wait for every push issued so far (before ``local`` is overwritten)."""
        from . import _lib

        _lib.check(self._lib.pcb_peer_join(self.device, self._stream(stream)))

    def publish(self, stream=None):
        """Wait for this rank's pushes and for every other rank's: ``full`` is then complete."""
        import torch
        import torch.distributed as dist

        self.join(stream)
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self._group)
        return self.full

    def close(self) -> None:
        import torch
        import torch.distributed as dist

        if self._ptr is None:
            return
        torch.cuda.synchronize(self.device)
        for p in self._peers:
            self._lib.pcb_peer_close(self.device, p)
        self._peers = []
        if dist.is_initialized():
            dist.barrier(group=self._group)   # every mapping is gone before the memory is freed
        self.full = self.local = None
        self._lib.pcb_peer_free(self.device, self._ptr)
        self._ptr = None
