"""Device plans and the host<->device plumbing around the C ABI.

PyTorch owns every device buffer, stream and event used here; the CUDA library
(``libpcb_b200.so``) receives raw pointers and never allocates per call.

Two ways in:

* device-resident: ``points`` is a CUDA ``torch.Tensor`` (N, D) float64 -> one kernel launch on
  the current stream, result is a CUDA tensor (N, G);
* host buffers (the reference's calling convention: NumPy in, NumPy out): the batch is cut into
  chunks that are copied in, evaluated and copied out on a small ring of streams so H2D, kernel
  and D2H of neighbouring chunks overlap.  Pinned inputs/outputs are copied directly; pageable
  ones go through pinned staging buffers.
"""

from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

from . import _lib

# host pipeline: per-chunk input size and ring depth (PCB_HOST_CHUNK_MB / PCB_HOST_NBUF override)
_CHUNK_BYTES = int(os.environ.get("PCB_HOST_CHUNK_MB", "64")) << 20
_NBUF = max(2, int(os.environ.get("PCB_HOST_NBUF", "3")))


def _torch():
    import torch

    return torch


def configure_host_pipeline(chunk_mb=None, ring_depth=None) -> dict:
    """Chunk size (MB of input per chunk) and ring depth of the host-buffer pipeline; existing
    per-device rings are dropped so the next call rebuilds them.  Returns the active settings."""
    global _CHUNK_BYTES, _NBUF
    if chunk_mb is not None:
        _CHUNK_BYTES = max(1, int(chunk_mb)) << 20
    if ring_depth is not None:
        _NBUF = max(2, int(ring_depth))
    with _HostPipe._guard:
        _HostPipe._pipes.clear()
    return {"chunk_mb": _CHUNK_BYTES >> 20, "ring_depth": _NBUF}


def require_device(device=None) -> int:
    """Resolve ``device`` to a CUDA ordinal; raise loudly when there is none."""
    torch = _torch()
    lib = _lib.load()
    if not torch.cuda.is_available() or lib.pcb_device_count() < 1:
        raise _lib.BackendUnavailable(
            "no CUDA device visible: pychebyshev_b200 evaluates on B200 (sm_100a) only and has "
            "no CPU fallback"
        )
    if device is None:
        return torch.cuda.current_device()
    if isinstance(device, int):
        return device
    dev = torch.device(device)
    if dev.type != "cuda":
        raise ValueError(f"device must be a CUDA device, got {device!r}")
    return torch.cuda.current_device() if dev.index is None else dev.index


class _HostPipe:
    """Per-device ring of streams, device buffers and pinned staging buffers (grow-only)."""

    _pipes: dict = {}
    _guard = threading.Lock()

    def __init__(self, dev: int):
        torch = _torch()
        self.dev = dev
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(_NBUF)]
        self.events = [torch.cuda.Event() for _ in range(_NBUF)]
        self.d_in = [None] * _NBUF
        self.d_out = [None] * _NBUF
        self.h_in = [None] * _NBUF
        self.h_out = [None] * _NBUF
        self.lock = threading.Lock()

    @classmethod
    def get(cls, dev: int) -> "_HostPipe":
        with cls._guard:
            if dev not in cls._pipes:
                cls._pipes[dev] = _HostPipe(dev)
            return cls._pipes[dev]

    def _grow(self, bufs, b, numel, **kw):
        """Grow-only buffer of ring slot ``b``.  Device buffers are allocated with the slot's OWN
        stream current, so the caching allocator ties the block to the stream that uses it (a block
        recycled from the caller's stream could otherwise be overwritten by the ring's H2D copy
        while queued work of the caller still reads it)."""
        torch = _torch()
        if bufs[b] is None or bufs[b].numel() < numel:
            if "device" in kw:
                with torch.cuda.stream(self.streams[b]):
                    bufs[b] = None  # free the old block on its own stream first
                    bufs[b] = torch.empty(numel, dtype=torch.float64, **kw)
            else:
                from . import _numa

                with _numa.bound_to_device(self.dev):
                    bufs[b] = torch.empty(numel, dtype=torch.float64, **kw)
        return bufs[b]

    def dev_in(self, b, numel):
        return self._grow(self.d_in, b, numel, device=f"cuda:{self.dev}")

    def dev_out(self, b, numel):
        return self._grow(self.d_out, b, numel, device=f"cuda:{self.dev}")

    def host_in(self, b, numel):
        return self._grow(self.h_in, b, numel, pin_memory=True)

    def host_out(self, b, numel):
        return self._grow(self.h_out, b, numel, pin_memory=True)


def fingerprint(a: np.ndarray) -> tuple:
    """Cheap identity + content token of a host array, used to invalidate cached device plans.

    Rebinding is caught by the data pointer (the caches hold the arrays, so an address cannot be
    reused while cached); in-place writes (``tensor_values[...] = ...``, ``core *= 2``) by a hash
    of the content -- the whole array up to 8192 elements, otherwise ~4096 evenly strided samples
    plus the last 8 elements (a write confined to unsampled elements of a large tensor is not
    seen: call ``_reset_plans()`` / rebind after such surgery).
    """
    n = a.size
    flat = a.reshape(-1) if a.flags.c_contiguous else a.ravel()
    if n <= 8192:
        sample = flat
    else:
        sample = np.concatenate([flat[:: n // 4096], flat[-8:]])
    return (a.__array_interface__["data"][0], a.shape, a.strides, hash(sample.tobytes()))


def pinned_empty(shape, dtype=np.float64, device=None):
    """A NumPy array backed by pinned (page-locked) host memory; keeps its tensor alive.

    With ``device`` given the pages are first-touched by a thread bound to the CPUs of that GPU's
    NUMA node (see :mod:`._numa`), so DMA to/from that GPU does not cross the socket interconnect.
    """
    torch = _torch()
    tdtype = {np.dtype(np.float64): torch.float64, np.dtype(np.int32): torch.int32}[np.dtype(dtype)]
    from . import _numa

    with _numa.bound_to_device(device):
        t = torch.empty(tuple(shape), dtype=tdtype, pin_memory=True)
        if device is not None and t.numel():
            t.view(-1)[:: max(1, 4096 // t.element_size())] = 0  # first touch on this node
    return t.numpy()


def host_pipeline_info(device=None) -> dict:
    """Configuration of the host-buffer pipeline (reporting)."""
    from . import _numa

    return {"chunk_mb": _CHUNK_BYTES >> 20, "ring_depth": _NBUF,
            "numa": _numa.describe(device)}


class DevicePlan:
    """Immutable device-resident copy of one interpolant on one GPU."""

    #: number of int32 side outputs per query (spline piece index)
    def __init__(self, handle, dev: int, ndim: int, G: int):
        self._handle = handle
        self.dev = dev
        self.ndim = ndim
        self.G = G

    def __del__(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h:
            try:
                _lib.load().pcb_plan_destroy(h)
            except Exception:
                pass

    # subclasses: launch one kernel over device pointers on `stream`
    def _launch(self, d_points: int, n: int, d_out: int, stream: int) -> None:
        raise NotImplementedError

    # -- device-resident ---------------------------------------------------------------------
    def eval_device(self, points, out=None):
        torch = _torch()
        if points.dtype != torch.float64 or points.dim() != 2 or points.shape[1] != self.ndim:
            raise ValueError(
                f"points must be a float64 tensor of shape (N, {self.ndim}), got "
                f"{tuple(points.shape)} {points.dtype}"
            )
        if points.device.index != self.dev:
            raise ValueError(f"points live on {points.device}, plan on cuda:{self.dev}")
        points = points.contiguous()
        n = points.shape[0]
        if out is None:
            out = torch.empty((n, self.G), dtype=torch.float64, device=points.device)
        elif (not out.is_contiguous() or out.dtype != torch.float64 or out.numel() != n * self.G
              or out.device != points.device):
            raise ValueError("out must be a contiguous float64 CUDA tensor with N*G elements")
        stream = torch.cuda.current_stream(points.device).cuda_stream
        self._launch(points.data_ptr(), n, out.data_ptr(), stream)
        return out

    # -- host buffers -------------------------------------------------------------------------
    def eval_host(self, points, out=None):
        torch = _torch()
        pts = np.asarray(points)
        if pts.dtype != np.float64:
            pts = pts.astype(np.float64)
        if pts.ndim != 2 or pts.shape[1] != self.ndim:
            raise ValueError(f"points must have shape (N, {self.ndim}), got {pts.shape}")
        pts = np.ascontiguousarray(pts)
        n, D, G = pts.shape[0], self.ndim, self.G
        if out is None:
            out = np.empty((n, G), dtype=np.float64)
        elif out.dtype != np.float64 or out.size != n * G or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float64 array with N*G elements")
        if n == 0:
            return out.reshape(n, G)
        pts_t = torch.from_numpy(pts)
        out_t = torch.from_numpy(out.reshape(n, G))
        in_pinned = pts_t.is_pinned()
        out_pinned = out_t.is_pinned()
        rows = max(1, min(n, _CHUNK_BYTES // (8 * max(D, G))))
        pipe = _HostPipe.get(self.dev)
        with pipe.lock, torch.cuda.device(self.dev):
            pending = [None] * _NBUF  # (lo, hi) whose staged output still has to be copied out
            for ci, lo in enumerate(range(0, n, rows)):
                hi = min(n, lo + rows)
                m = hi - lo
                b = ci % _NBUF
                stream = pipe.streams[b]
                if pending[b] is not None:  # buffer b is being reused: drain its previous chunk
                    pipe.events[b].synchronize()
                    plo, phi = pending[b]
                    out_t[plo:phi].copy_(pipe.h_out[b][: (phi - plo) * G].view(phi - plo, G))
                    pending[b] = None
                elif ci >= _NBUF:
                    pipe.events[b].synchronize()
                d_in = pipe.dev_in(b, rows * D)[: m * D].view(m, D)
                d_out = pipe.dev_out(b, rows * G)[: m * G].view(m, G)
                with torch.cuda.stream(stream):
                    if in_pinned:
                        d_in.copy_(pts_t[lo:hi], non_blocking=True)
                    else:
                        h_in = pipe.host_in(b, rows * D)[: m * D].view(m, D)
                        h_in.copy_(pts_t[lo:hi])
                        d_in.copy_(h_in, non_blocking=True)
                    self._launch(d_in.data_ptr(), m, d_out.data_ptr(), stream.cuda_stream)
                    if out_pinned:
                        out_t[lo:hi].copy_(d_out, non_blocking=True)
                    else:
                        h_out = pipe.host_out(b, rows * G)[: m * G].view(m, G)
                        h_out.copy_(d_out, non_blocking=True)
                        pending[b] = (lo, hi)
                    pipe.events[b].record(stream)
            for b in range(_NBUF):
                pipe.events[b].synchronize()
                if pending[b] is not None:
                    plo, phi = pending[b]
                    out_t[plo:phi].copy_(pipe.h_out[b][: (phi - plo) * G].view(phi - plo, G))
        return out.reshape(n, G)

    def eval(self, points, out=None):
        torch = _torch()
        if isinstance(points, torch.Tensor) and points.is_cuda:
            return self.eval_device(points, out)
        if isinstance(points, torch.Tensor):
            points = points.numpy()
        return self.eval_host(points, out)


# ------------------------------------------------------------------------------------------
# concrete plans
# ------------------------------------------------------------------------------------------

class TTPlan(DevicePlan):
    """``pcb_tt_plan_create`` + ``pcb_tt_eval`` / ``pcb_tt_eval_fd``."""

    def __init__(self, cores, domain, n_nodes, ranks, dim_order, device=None):
        lib = _lib.load()
        dev = require_device(device)
        D = len(cores)
        _, n_p = _lib.as_i32(n_nodes)
        _, r_p = _lib.as_i32(ranks)
        lo, lo_p = _lib.as_f64([float(d[0]) for d in domain])
        hi, hi_p = _lib.as_f64([float(d[1]) for d in domain])
        _, do_p = _lib.as_i32(dim_order)
        cat = np.ascontiguousarray(
            np.concatenate([np.ascontiguousarray(c, dtype=np.float64).ravel() for c in cores]))
        handle = C.c_void_p()
        _lib.check(lib.pcb_tt_plan_create(dev, D, n_p, r_p, lo_p, hi_p, do_p,
                                          cat.ctypes.data_as(_lib._f64p), C.byref(handle)))
        super().__init__(handle, dev, D, 1)
        self._orders = None
        self._orders_keep = None
        self.algo = 0

    def with_orders(self, orders, algo=0):
        """A view of this plan that evaluates G finite-difference rows (shares the device data)."""
        view = object.__new__(TTPlan)
        view._handle = None  # not owning
        view._owner = self
        view.dev, view.ndim = self.dev, self.ndim
        arr = np.ascontiguousarray(np.asarray(orders, dtype=np.int32).reshape(-1, self.ndim))
        view._orders_keep = arr
        view._orders = arr.ctypes.data_as(_lib._i32p)
        view.G = arr.shape[0]
        view.algo = algo
        return view

    FD_PATHS = {1: "tt_fd_general_kernel (one chain per stencil point)",
                2: "ttc_fd_shared_kernel (constant bank, one launch)",
                3: "ttc_fd_shared_kernel (constant bank, one launch per differentiated dim)",
                4: "ttc_gstep_kernel + ttc_gcoeff_kernel (one launch per core)",
                5: "tt_fd_shared_kernel (shared memory)"}

    def info(self) -> dict:
        """Which kernels evaluate this plan (reporting only; the plan keeps no per-call state)."""
        handle = self._handle if self._handle else self._owner._handle
        lib = _lib.load()
        out = (C.c_int32 * 8)()
        _lib.check(lib.pcb_tt_plan_info(handle, out))
        keys = ("uniform_path_values", "uniform_path_any_fd", "uniform_qpt", "uniform_threads_values",
                "uniform_threads_fd", "smem_core_placement", "smem_fd_qpt", "smem_fd_threads")
        info = dict(zip(keys, [int(v) for v in out]))
        if self._orders is not None:
            path = lib.pcb_tt_fd_path(handle, self.G, self._orders, self.algo)
            if path < 0:
                _lib.check(path)
            info["fd_path"] = int(path)
            info["fd_kernel"] = self.FD_PATHS.get(int(path), "?")
            info["uniform_path_fd"] = int(path in (2, 3, 4))
        return info

    def resolved_algo(self) -> int:
        """The algorithm ``pcb_tt_eval_fd`` runs for this view's rows (1 or 2)."""
        if self._orders is None:
            return 0
        if self.algo:
            return self.algo
        handle = self._handle if self._handle else self._owner._handle
        rc = _lib.load().pcb_tt_fd_algo(handle, self.G, self._orders)
        if rc < 0:
            _lib.check(rc)
        return rc

    def _launch(self, d_points, n, d_out, stream):
        lib = _lib.load()
        handle = self._handle if self._handle else self._owner._handle
        if self._orders is None:
            _lib.check(lib.pcb_tt_eval(handle, d_points, n, d_out, stream))
        else:
            _lib.check(lib.pcb_tt_eval_fd(handle, d_points, n, self.G, self._orders, d_out,
                                          self.algo, stream))


class FullPlan(DevicePlan):
    """``pcb_full_plan_create`` + ``pcb_full_eval`` for G pre-differentiated tensors."""

    def __init__(self, n_nodes, nodes, weights, tensors, device=None, algo=0):
        lib = _lib.load()
        dev = require_device(device)
        D = len(n_nodes)
        _, n_p = _lib.as_i32(n_nodes)
        _, nodes_p = _lib.as_f64(np.concatenate([np.asarray(a, dtype=np.float64) for a in nodes]))
        _, w_p = _lib.as_f64(np.concatenate([np.asarray(a, dtype=np.float64) for a in weights]))
        keep = [np.ascontiguousarray(t, dtype=np.float64) for t in tensors]
        for t in keep:
            if t.shape != tuple(n_nodes):
                raise ValueError(f"tensor shape {t.shape} does not match n_nodes {tuple(n_nodes)}")
        ptrs = _lib.ptr_array(keep)
        handle = C.c_void_p()
        _lib.check(lib.pcb_full_plan_create(dev, D, n_p, nodes_p, w_p, len(keep), ptrs,
                                            C.byref(handle)))
        super().__init__(handle, dev, D, len(keep))
        self.algo = algo

    @classmethod
    def from_values(cls, n_nodes, nodes, weights, diff_matrices, values, orders, device=None, algo=0):
        """``pcb_full_plan_create_from_values``: upload the value tensor once; the derivative tensors
        of ``orders`` are made on the device (bit-identical to the host recipe)."""
        lib = _lib.load()
        dev = require_device(device)
        D = len(n_nodes)
        _, n_p = _lib.as_i32(n_nodes)
        _, nodes_p = _lib.as_f64(np.concatenate([np.asarray(a, dtype=np.float64) for a in nodes]))
        _, w_p = _lib.as_f64(np.concatenate([np.asarray(a, dtype=np.float64) for a in weights]))
        _, dm_p = _lib.as_f64(np.concatenate([np.ascontiguousarray(m, dtype=np.float64).ravel()
                                              for m in diff_matrices]))
        vals = np.ascontiguousarray(values, dtype=np.float64)
        if vals.shape != tuple(n_nodes):
            raise ValueError(f"tensor shape {vals.shape} does not match n_nodes {tuple(n_nodes)}")
        ords = np.ascontiguousarray(np.asarray(orders, dtype=np.int32).reshape(-1, D))
        handle = C.c_void_p()
        _lib.check(lib.pcb_full_plan_create_from_values(
            dev, D, n_p, nodes_p, w_p, dm_p, vals.ctypes.data_as(_lib._f64p), ords.shape[0],
            ords.ctypes.data_as(_lib._i32p), C.byref(handle)))
        self = object.__new__(cls)
        DevicePlan.__init__(self, handle, dev, D, ords.shape[0])
        self.algo = algo
        return self

    def _launch(self, d_points, n, d_out, stream):
        _lib.check(_lib.load().pcb_full_eval(self._handle, d_points, n, d_out, self.algo, stream))


class SplinePlan(DevicePlan):
    """``pcb_spline_plan_create`` + ``pcb_spline_eval`` / ``pcb_spline_lookup``."""

    def __init__(self, knots, pieces, device=None):
        """``pieces``: C-order list of ``(n_nodes, nodes, weights, [tensor_g ...])``."""
        lib = _lib.load()
        dev = require_device(device)
        D = len(knots)
        _, nk_p = _lib.as_i32([len(k) for k in knots])
        kcat = [float(v) for k in knots for v in k]
        _, k_p = _lib.as_f64(kcat if kcat else [0.0])
        P = len(pieces)
        G = len(pieces[0][3])
        _, pn_p = _lib.as_i32([list(p[0]) for p in pieces])
        _, nodes_p = _lib.as_f64(np.concatenate([np.asarray(a, dtype=np.float64)
                                                 for p in pieces for a in p[1]]))
        _, w_p = _lib.as_f64(np.concatenate([np.asarray(a, dtype=np.float64)
                                             for p in pieces for a in p[2]]))
        keep = [np.ascontiguousarray(t, dtype=np.float64) for p in pieces for t in p[3]]
        ptrs = _lib.ptr_array(keep)
        handle = C.c_void_p()
        _lib.check(lib.pcb_spline_plan_create(dev, D, nk_p, k_p, P, pn_p, nodes_p, w_p, G, ptrs,
                                              C.byref(handle)))
        super().__init__(handle, dev, D, G)

    def _launch(self, d_points, n, d_out, stream):
        _lib.check(_lib.load().pcb_spline_eval(self._handle, d_points, n, d_out, None, stream))

    def lookup_device(self, points, out=None):
        """Piece indices (int32) of a CUDA (N, D) float64 tensor, on the current stream."""
        torch = _torch()
        if (not isinstance(points, torch.Tensor) or not points.is_cuda
                or points.dtype != torch.float64 or points.dim() != 2
                or points.shape[1] != self.ndim):
            raise ValueError(
                f"points must be a CUDA float64 tensor of shape (N, {self.ndim}), got "
                f"{tuple(getattr(points, 'shape', ()))} {getattr(points, 'dtype', type(points))}")
        if points.device.index != self.dev:
            raise ValueError(f"points live on {points.device}, plan on cuda:{self.dev}")
        pts = points.contiguous()
        n = pts.shape[0]
        if out is None:
            out = torch.empty(n, dtype=torch.int32, device=pts.device)
        elif (out.dtype != torch.int32 or out.numel() != n or not out.is_contiguous()
              or out.device != pts.device):
            raise ValueError("out must be a contiguous int32 CUDA tensor with N elements")
        stream = torch.cuda.current_stream(pts.device).cuda_stream
        _lib.check(_lib.load().pcb_spline_lookup(self._handle, pts.data_ptr(), n, out.data_ptr(),
                                                 stream))
        return out

    def lookup(self, points):
        """Piece indices (int32) for device or host points."""
        torch = _torch()
        lib = _lib.load()
        if isinstance(points, torch.Tensor) and points.is_cuda:
            return self.lookup_device(points)
        pts = np.ascontiguousarray(np.asarray(points, dtype=np.float64))
        if pts.ndim != 2 or pts.shape[1] != self.ndim:
            raise ValueError(f"points must have shape (N, {self.ndim}), got {pts.shape}")
        with torch.cuda.device(self.dev):
            d = torch.from_numpy(pts).to(f"cuda:{self.dev}")
            out = torch.empty(pts.shape[0], dtype=torch.int32, device=d.device)
            stream = torch.cuda.current_stream(d.device).cuda_stream
            _lib.check(lib.pcb_spline_lookup(self._handle, d.data_ptr(), pts.shape[0],
                                             out.data_ptr(), stream))
            return out.cpu().numpy()


class SliderPlan(DevicePlan):
    """``pcb_slider_plan_create`` + ``pcb_slider_eval``."""

    def __init__(self, ndim, partition, slides, pivot_value, out_slide, row_tensors, device=None):
        """``slides``: list of ``(n_nodes, nodes, weights)``; ``row_tensors[g][s]`` tensor or None."""
        lib = _lib.load()
        dev = require_device(device)
        S = len(slides)
        G = len(out_slide)
        _, gs_p = _lib.as_i32([len(g) for g in partition])
        _, gd_p = _lib.as_i32([d for g in partition for d in g])
        _, sn_p = _lib.as_i32([n for s in slides for n in s[0]])
        _, nodes_p = _lib.as_f64(np.concatenate([np.asarray(a, dtype=np.float64)
                                                 for s in slides for a in s[1]]))
        _, w_p = _lib.as_f64(np.concatenate([np.asarray(a, dtype=np.float64)
                                             for s in slides for a in s[2]]))
        _, os_p = _lib.as_i32(out_slide)
        keep = [None if t is None else np.ascontiguousarray(t, dtype=np.float64)
                for row in row_tensors for t in row]
        ptrs = _lib.ptr_array(keep)
        handle = C.c_void_p()
        _lib.check(lib.pcb_slider_plan_create(dev, ndim, S, gs_p, gd_p, sn_p, nodes_p, w_p,
                                              float(pivot_value), G, os_p, ptrs, C.byref(handle)))
        super().__init__(handle, dev, ndim, G)

    def _launch(self, d_points, n, d_out, stream):
        _lib.check(_lib.load().pcb_slider_eval(self._handle, d_points, n, d_out, stream))


class FilePlan(DevicePlan):
    """A plan built natively from a ``.pcb`` file (``pcb_plan_from_file[_orders]``): no Python
    object, no NumPy grid arithmetic -- the C++ twin of the reference's stand-alone readers.
    ``orders``: optional G derivative-order rows -> G outputs per point (values only otherwise)."""

    def __init__(self, path, device=None, orders=None):
        import os

        lib = _lib.load()
        dev = require_device(device)
        handle, kind, ndim = C.c_void_p(), C.c_int32(), C.c_int32()
        fs = os.fsencode(os.fspath(path))
        if orders is None:
            _lib.check(lib.pcb_plan_from_file(dev, fs, C.byref(handle), C.byref(kind), C.byref(ndim)))
            G = 1
        else:
            arr = np.ascontiguousarray(np.asarray(orders, dtype=np.int32))
            if arr.ndim != 2:
                raise ValueError("orders must be a list of derivative-order rows")
            G = arr.shape[0]
            with open(os.fspath(path), "rb") as f:
                head = f.read(16)
            ndim_file = int.from_bytes(head[12:16], "little") if len(head) == 16 else -1
            if ndim_file != arr.shape[1]:
                raise ValueError(
                    f"derivative rows have {arr.shape[1]} entries, file has {ndim_file} dimensions")
            _lib.check(lib.pcb_plan_from_file_orders(dev, fs, G, arr.ctypes.data_as(_lib._i32p),
                                                     C.byref(handle), C.byref(kind), C.byref(ndim)))
        super().__init__(handle, dev, int(ndim.value), G)
        self.kind = {1: "approx", 2: "spline"}[int(kind.value)]

    def _launch(self, d_points, n, d_out, stream):
        _lib.check(_lib.load().pcb_plan_eval(self._handle, d_points, n, d_out, stream))


def write_pcb_native(path, domain, n_nodes, tensors, knots=None) -> None:
    """``.pcb`` v1 through the NATIVE writer (``pcb_file_write_approx`` / ``_spline``): one tensor
    for an approximation, the C-order piece tensors (+ ``knots``) for a spline."""
    import os

    lib = _lib.load()
    D = len(n_nodes)
    _, lo_p = _lib.as_f64([float(d[0]) for d in domain])
    _, hi_p = _lib.as_f64([float(d[1]) for d in domain])
    _, n_p = _lib.as_i32(n_nodes)
    fs = os.fsencode(os.fspath(path))
    if knots is None:
        t = np.ascontiguousarray(tensors, dtype=np.float64)
        _lib.check(lib.pcb_file_write_approx(fs, D, lo_p, hi_p, n_p, t.ctypes.data_as(_lib._f64p)))
        return
    keep = [np.ascontiguousarray(t, dtype=np.float64) for t in tensors]
    _, nk_p = _lib.as_i32([len(k) for k in knots])
    kcat = [float(v) for k in knots for v in k]
    _, k_p = _lib.as_f64(kcat if kcat else [0.0])
    _lib.check(lib.pcb_file_write_spline(fs, D, lo_p, hi_p, n_p, nk_p, k_p, len(keep),
                                         _lib.ptr_array(keep)))


def probe_fp64_peak(kind: int, device=None):
    """Measured FP64 peak (TFLOP/s) of the DFMA (0) or DMMA (1) pipe on this device."""
    lib = _lib.load()
    dev = require_device(device)
    tf, ms = C.c_double(), C.c_double()
    _lib.check(lib.pcb_probe_fp64_peak(dev, kind, C.byref(tf), C.byref(ms)))
    return tf.value, ms.value


def launch_count() -> int:
    return int(_lib.load().pcb_launch_count())
