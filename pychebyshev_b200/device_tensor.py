"""Node-value tensors on the device: derivative passes, slice and extrude.

Device twins of the reference's tensor preparations, for tensors too large to push through NumPy on
every plan creation (16^6: 134 MB per derivative order) and for GPU-resident portfolio algebra:

* :func:`differentiate` -- ``_apply_derivative_passes`` (``barycentric.py:951-990``): for
  ``d = D-1 .. 0``, ``order[d]`` times ``T <- T x_d D_d^T``.  **Bit-identical** to the host recipe
  (``_grid.differentiate_tensor``): every output element is one sequential FMA chain over the
  contracted index, which is what OpenBLAS computes for the reference's ``arr @ D_T``
  (``csrc/pcb_tensor.cu``; ``tests/test_device_tensor.py``).
* :func:`slice_axis` -- ``_slice_tensor`` (``_extrude_slice.py:79-92``): node hit -> ``take``, else
  tensordot with the normalised barycentric weights (weights computed on the host with the
  reference's NumPy expressions; the contraction agrees with NumPy's dgemv to <= 1e-15 of the
  tensor's scale).
* :func:`extrude_axis` -- ``_extrude_tensor`` (``_extrude_slice.py:73-76``): exact replication.

All functions take and return contiguous float64 CUDA ``torch.Tensor`` objects (PyTorch owns the
memory) and run on the current stream.
"""

from __future__ import annotations

import numpy as np

from . import _lib
from ._engine import require_device


def _torch():
    import torch

    return torch


def to_device(values, device=None):
    """Host array (or CUDA tensor) -> contiguous float64 CUDA tensor."""
    torch = _torch()
    dev = require_device(device)
    if isinstance(values, torch.Tensor):
        return values.to(device=f"cuda:{dev}", dtype=torch.float64).contiguous()
    arr = np.ascontiguousarray(values, dtype=np.float64)
    return torch.from_numpy(arr).to(f"cuda:{dev}")


def _check(t):
    torch = _torch()
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64):
        raise ValueError("expected a float64 CUDA tensor")
    return t.contiguous()


def _shape_args(t):
    return _lib.as_i32(list(t.shape))


def deriv_pass(t, axis: int, dmat):
    """``moveaxis(moveaxis(t, axis, -1) @ dmat.T, -1, axis)`` (one pass, out of place)."""
    torch = _torch()
    t = _check(t)
    n = t.shape[axis]
    dm, dm_p = _lib.as_f64(np.ascontiguousarray(dmat, dtype=np.float64).ravel())
    if dm.size != n * n:
        raise ValueError(f"differentiation matrix must be {n} x {n}")
    out = torch.empty_like(t)
    _, shp = _shape_args(t)
    stream = torch.cuda.current_stream(t.device).cuda_stream
    _lib.check(_lib.load().pcb_tensor_deriv(t.device.index, t.dim(), shp, int(axis), dm_p,
                                            t.data_ptr(), out.data_ptr(), stream))
    return out


def differentiate(values, diff_matrices, order, device=None):
    """Device twin of ``_grid.differentiate_tensor`` / the reference's
    ``_apply_derivative_passes``; returns a CUDA tensor of the same shape."""
    t = to_device(values, device)
    if order is None:
        return t
    for d in range(t.dim() - 1, -1, -1):
        for _ in range(int(order[d])):
            t = deriv_pass(t, d, diff_matrices[d])
    return t


def contract_axis(t, axis: int, vec):
    """``np.tensordot(t, vec, axes=([axis], [0]))`` on the device."""
    torch = _torch()
    t = _check(t)
    v, v_p = _lib.as_f64(np.ascontiguousarray(vec, dtype=np.float64).ravel())
    if v.size != t.shape[axis]:
        raise ValueError(f"vector must have {t.shape[axis]} entries")
    shape = list(t.shape)
    del shape[axis]
    out = torch.empty(shape, dtype=torch.float64, device=t.device)
    _, shp = _shape_args(t)
    stream = torch.cuda.current_stream(t.device).cuda_stream
    _lib.check(_lib.load().pcb_tensor_contract(t.device.index, t.dim(), shp, int(axis), v_p,
                                               t.data_ptr(), out.data_ptr(), stream))
    return out


def slice_axis(t, axis: int, nodes, weights, value: float):
    """``_slice_tensor`` (``_extrude_slice.py:79-92``): fix ``axis`` at ``value``."""
    nodes = np.asarray(nodes, dtype=np.float64)
    weights = np.asarray(weights, dtype=np.float64)
    diff = value - nodes
    exact_idx = int(np.argmin(np.abs(diff)))
    if np.abs(diff[exact_idx]) < 1e-14:
        return _check(t).select(axis, exact_idx).contiguous()   # np.take: exact
    w_over_diff = weights / diff
    w_norm = w_over_diff / np.sum(w_over_diff)
    return contract_axis(t, axis, w_norm)


def extrude_axis(t, axis: int, n_new: int):
    """``np.repeat(np.expand_dims(t, axis), n_new, axis)`` on the device."""
    torch = _torch()
    t = _check(t)
    shape = list(t.shape)
    shape.insert(axis, int(n_new))
    out = torch.empty(shape, dtype=torch.float64, device=t.device)
    _, shp = _shape_args(t)
    stream = torch.cuda.current_stream(t.device).cuda_stream
    _lib.check(_lib.load().pcb_tensor_extrude(t.device.index, t.dim(), shp, int(axis), int(n_new),
                                              t.data_ptr(), out.data_ptr(), stream))
    return out
