"""``ChebyshevApproximation``: full-tensor Chebyshev interpolant evaluated on B200.

Same constructor, factories and evaluation entry points as the reference class
(``barycentric.py:271-523, 717-1112, 1173-1243, 1576-1664, 1690-1934``); only the bodies of the
evaluation methods differ: they hand the batch to the CUDA engine (``pcb_full_eval``).  There is
no CPU evaluation path in this package.
"""

from __future__ import annotations

import os
import pickle
import time
from typing import Callable, List, Sequence

import numpy as np

from . import _grid, pcbfile
from ._engine import FullPlan, fingerprint


def _unwrap(value, attr):
    """Accept the typed helpers Domain / Ns / SpecialPoints (reference ``__init__.py:35-66``)."""
    inner = getattr(value, attr, None)
    return value if inner is None else [list(v) if hasattr(v, "__iter__") else v for v in inner]


class _DerivativeIds:
    """Per-object registry ``derivative_order tuple -> sequential int`` (barycentric.py:1173-1243)."""

    def _init_derivative_ids(self):
        self._derivative_id_registry = {}
        self._derivative_id_to_orders = []

    def get_derivative_id(self, derivative_order: Sequence[int]) -> int:
        if len(derivative_order) != self.num_dimensions:
            raise ValueError(
                f"derivative_order length {len(derivative_order)} does not match "
                f"num_dimensions {self.num_dimensions}"
            )
        for d, o in enumerate(derivative_order):
            if not isinstance(o, (int, np.integer)):
                raise ValueError(f"derivative_order[{d}] must be int, got {type(o).__name__}")
            if o < 0 or o > self.max_derivative_order:
                raise ValueError(
                    f"derivative_order[{d}]={o} out of range [0, {self.max_derivative_order}]")
        key = tuple(int(o) for o in derivative_order)
        if key not in self._derivative_id_registry:
            self._derivative_id_registry[key] = len(self._derivative_id_to_orders)
            self._derivative_id_to_orders.append(key)
        return self._derivative_id_registry[key]

    def _resolve_derivative_args(self, derivative_order, derivative_id):
        if derivative_order is not None and derivative_id is not None:
            raise ValueError("provide exactly one of derivative_order or derivative_id, not both")
        if derivative_order is None and derivative_id is None:
            raise ValueError("must provide derivative_order or derivative_id")
        if derivative_id is not None:
            if derivative_id < 0 or derivative_id >= len(self._derivative_id_to_orders):
                raise KeyError(
                    f"unknown derivative_id {derivative_id}; register via get_derivative_id() first")
            return list(self._derivative_id_to_orders[derivative_id])
        return derivative_order


class ChebyshevApproximation(_DerivativeIds):
    """Multi-dimensional Chebyshev interpolant on a full tensor grid.

    ``ChebyshevApproximation(..., special_points=[[100.0], []])`` returns a
    :class:`~pychebyshev_b200.ChebyshevSpline` with knots at the special points, exactly like the
    reference's ``__new__`` dispatch (``barycentric.py:271-339``).
    """

    def __new__(cls, function=None, num_dimensions=None, domain=None, n_nodes=None,
                max_derivative_order=2, error_threshold=None, max_n=64, special_points=None,
                additional_data=None, *, defer_build=False, n_workers=None, device=None):
        special_points = _unwrap(special_points, "knots_per_dim")
        if special_points is not None:
            if num_dimensions is not None and len(special_points) != num_dimensions:
                raise ValueError(
                    f"special_points must have {num_dimensions} entries, got {len(special_points)}")
            for d, sp in enumerate(special_points):
                if not isinstance(sp, (list, tuple)):
                    raise ValueError(
                        f"special_points[{d}] must be a list/tuple of floats, got "
                        f"{type(sp).__name__}: {sp!r}")
            if any(len(sp) > 0 for sp in special_points):
                from .spline import ChebyshevSpline, validate_special_points_shape

                domain = _unwrap(domain, "bounds")
                n_nodes = _unwrap(n_nodes, "counts")
                validate_special_points_shape(special_points, n_nodes, num_dimensions, domain)
                return ChebyshevSpline(
                    function, num_dimensions, domain, n_nodes=n_nodes, knots=special_points,
                    max_derivative_order=max_derivative_order, error_threshold=error_threshold,
                    max_n=max_n, additional_data=additional_data, defer_build=defer_build,
                    n_workers=n_workers, device=device)
        return super().__new__(cls)

    def __init__(self, function: Callable, num_dimensions: int, domain, n_nodes=None,
                 max_derivative_order: int = 2, error_threshold=None, max_n: int = 64,
                 special_points=None, additional_data=None, *, defer_build: bool = False,
                 n_workers=None, device=None):
        domain = _unwrap(domain, "bounds")
        n_nodes = _unwrap(n_nodes, "counts")
        if max_n < 3:
            raise ValueError(f"max_n must be at least 3, got max_n={max_n}.")
        if error_threshold is not None or n_nodes is None or any(n is None for n in n_nodes):
            if n_nodes is None and error_threshold is None and not defer_build:
                raise ValueError(
                    "Must provide either n_nodes (explicit) or error_threshold (auto-N). Got neither.")
            raise NotImplementedError(
                "error_threshold auto-N calibration is a build-time feature outside the "
                "evaluation path this package accelerates; pass explicit n_nodes")
        if len(domain) != num_dimensions or len(n_nodes) != num_dimensions:
            raise ValueError(
                f"len(domain)={len(domain)} and len(n_nodes)={len(n_nodes)} must both equal "
                f"num_dimensions={num_dimensions}")
        self.function = function
        self.num_dimensions = int(num_dimensions)
        self.domain = [list(b) for b in domain]
        self.n_nodes = [int(n) for n in n_nodes]
        self.max_derivative_order = max_derivative_order
        self.error_threshold = error_threshold
        self.max_n = max_n
        self.special_points = _unwrap(special_points, "knots_per_dim")
        self.additional_data = additional_data
        self.n_workers = n_workers
        self.descriptor = ""
        self.device = device
        self.build_time = 0.0
        self.n_evaluations = 0
        self.tensor_values = None
        self._init_derivative_ids()
        self._reset_plans()
        self.nodes, self.weights, self.diff_matrices = _grid.grid_arrays(self.domain, self.n_nodes)
        if defer_build and function is not None:
            raise ValueError("defer_build=True requires function=None")

    # ------------------------------------------------------------------ construction
    def build(self, verbose: bool | int = True) -> None:
        """Evaluate ``function(point, additional_data)`` on the grid (barycentric.py:643-715)."""
        if self.function is None:
            raise RuntimeError(
                "Cannot build: no function assigned. This object was created via from_values() "
                "or load().")
        total = int(np.prod(self.n_nodes))
        if verbose:
            print(f"Building {self.num_dimensions}D Chebyshev approximation ({total:,} evaluations)...")
        t0 = time.time()
        pts = _grid.full_grid_points(self.nodes)
        if self.n_workers is not None and self.n_workers > 1:
            from concurrent.futures import ProcessPoolExecutor

            chunk = max(1, total // (4 * self.n_workers))
            with ProcessPoolExecutor(self.n_workers) as pool:
                vals = list(pool.map(_call_function, [self.function] * total, pts.tolist(),
                                     [self.additional_data] * total, chunksize=chunk))
        else:
            f, data = self.function, self.additional_data
            vals = [float(f(p, data)) for p in pts.tolist()]
        tensor = np.asarray(vals, dtype=np.float64).reshape(self.n_nodes)
        if not np.isfinite(tensor).all():
            bad = int(np.sum(~np.isfinite(tensor)))
            raise ValueError(
                f"function returned non-finite values at {bad} grid point(s); build cannot "
                "proceed with NaN/Inf in tensor_values")
        self.tensor_values = tensor
        self.n_evaluations = total
        self.build_time = time.time() - t0
        self._reset_plans()
        if verbose:
            print(f"  Built in {self.build_time:.3f}s")

    @classmethod
    def from_values(cls, tensor_values, num_dimensions, domain, n_nodes, max_derivative_order=2,
                    *, device=None) -> "ChebyshevApproximation":
        """Interpolant from pre-computed grid values (barycentric.py:1813-1934)."""
        domain = _unwrap(domain, "bounds")
        n_nodes = _unwrap(n_nodes, "counts")
        arr = np.asarray(tensor_values, dtype=float)
        if len(domain) != num_dimensions or len(n_nodes) != num_dimensions:
            raise ValueError(
                f"len(domain)={len(domain)} and len(n_nodes)={len(n_nodes)} must both equal "
                f"num_dimensions={num_dimensions}")
        if arr.shape != tuple(n_nodes):
            raise ValueError(
                f"tensor_values.shape={arr.shape} does not match n_nodes={tuple(n_nodes)}")
        if not np.isfinite(arr).all():
            raise ValueError("tensor_values contains NaN or Inf")
        for d, (lo, hi) in enumerate(domain):
            if lo >= hi:
                raise ValueError(f"domain[{d}]: lo={lo} must be strictly less than hi={hi}")
        obj = cls(None, num_dimensions, domain, list(n_nodes), max_derivative_order, device=device)
        obj.tensor_values = np.array(arr, dtype=np.float64, order="C", copy=True)
        return obj

    def set_original_function_values(self, values) -> None:
        """Fill a ``defer_build=True`` object in place (barycentric.py:480-521)."""
        if self.tensor_values is not None:
            raise RuntimeError(
                "interpolant is already constructed; set_original_function_values() is for "
                "defer_build=True objects")
        arr = np.asarray(values, dtype=np.float64)
        if arr.shape != tuple(self.n_nodes):
            raise ValueError(
                f"values shape {arr.shape} does not match expected {tuple(self.n_nodes)}")
        if not np.isfinite(arr).all():
            raise ValueError("values contains NaN or Inf (must be finite)")
        self.tensor_values = arr.copy()
        self.function = None
        self._reset_plans()

    @staticmethod
    def nodes(num_dimensions, domain, n_nodes) -> dict:  # noqa: F811 (class attr shadowed per instance)
        """Grid description for :meth:`from_values` (barycentric.py:1690-1761)."""
        domain = _unwrap(domain, "bounds")
        n_nodes = _unwrap(n_nodes, "counts")
        if len(domain) != num_dimensions or len(n_nodes) != num_dimensions:
            raise ValueError(
                f"len(domain)={len(domain)} and len(n_nodes)={len(n_nodes)} must both equal "
                f"num_dimensions={num_dimensions}")
        per_dim = [_grid.cheb_nodes(lo, hi, n) for (lo, hi), n in zip(domain, n_nodes)]
        return {"nodes_per_dim": per_dim, "full_grid": _grid.full_grid_points(per_dim),
                "shape": tuple(n_nodes)}

    def is_construction_finished(self) -> bool:
        return self.tensor_values is not None

    # ------------------------------------------------------------------ device plans
    def _reset_plans(self):
        self._plans = {}
        self._plan_tensor_id = None
        self._deriv_cache = {}

    def _validate_caches(self):
        """Drop cached plans / derivative tensors when ``tensor_values`` was rebound or edited."""
        token = fingerprint(self.tensor_values)
        if self._plan_tensor_id != token:
            self._plans, self._deriv_cache = {}, {}
            self._plan_tensor_id = token

    def derivative_tensor(self, order) -> np.ndarray:
        """Pre-differentiated tensor for one derivative multi-index (host, cached)."""
        key = tuple(int(o) for o in order)
        self._validate_caches()
        if key not in self._deriv_cache:
            self._deriv_cache[key] = _grid.differentiate_tensor(
                self.tensor_values, self.diff_matrices, key)
        return self._deriv_cache[key]

    def _plan(self, orders, device=None, algo=0) -> FullPlan:
        if self.tensor_values is None:
            raise RuntimeError("Call build() first")
        orders = _grid.normalize_orders(orders, self.num_dimensions)
        self._validate_caches()
        from ._engine import require_device

        dev = require_device(self.device if device is None else device)
        key = (dev, orders, algo)
        if key not in self._plans:
            if len(self._plans) >= 8:  # bound device memory held by stale order sets
                self._plans.pop(next(iter(self._plans)))
            # (1-D tensors stay on the host recipe: NumPy runs `vector @ D^T` through dgemv, whose
            # lane-split partial sums differ from the chain in the last bit -- and a 1-D tensor is
            # tiny anyway)
            if (os.environ.get("PCB_DEVICE_DERIV", "1") != "0" and max(self.n_nodes) <= 64
                    and self.num_dimensions >= 2):
                # N3: one upload of the value tensor, derivative tensors made on the device
                self._plans[key] = FullPlan.from_values(
                    self.n_nodes, self.nodes, self.weights, self.diff_matrices, self.tensor_values,
                    orders, dev, algo)
            else:
                tensors = [self.derivative_tensor(o) for o in orders]
                self._plans[key] = FullPlan(self.n_nodes, self.nodes, self.weights, tensors, dev,
                                            algo)
        return self._plans[key]

    # ------------------------------------------------------------------ slice / extrude (N3)
    def _derived(self, tensor, nodes, weights, diff_matrices, domain, n_nodes):
        obj = object.__new__(ChebyshevApproximation)
        obj.function = None
        obj.num_dimensions = len(n_nodes)
        obj.domain = domain
        obj.n_nodes = n_nodes
        obj.max_derivative_order = self.max_derivative_order
        obj.error_threshold = None
        obj.max_n = self.max_n
        obj.special_points = None
        obj.additional_data = None
        obj.n_workers = None
        obj.descriptor = ""
        obj.device = self.device
        obj.build_time = 0.0
        obj.n_evaluations = 0
        obj.nodes, obj.weights, obj.diff_matrices = nodes, weights, diff_matrices
        obj.tensor_values = tensor
        obj._init_derivative_ids()
        obj._reset_plans()
        return obj

    def slice(self, params, *, device=None) -> "ChebyshevApproximation":
        """Fix dimensions at given values (reference ``barycentric.py:2067-2154``); the mode
        contractions run on the device (``device_tensor.slice_axis``)."""
        if self.tensor_values is None:
            raise RuntimeError("Call build() first")
        from . import device_tensor as DT

        ndim = self.num_dimensions
        if isinstance(params, tuple) and len(params) == 2 and isinstance(params[0], (int, np.integer)):
            params = [params]
        params = [tuple(p) for p in params]
        if len(params) >= ndim:
            raise ValueError(f"Cannot slice all {ndim} dimensions (would produce 0D result)")
        seen = set()
        for dim_idx, value in params:
            if not isinstance(dim_idx, (int, np.integer)):
                raise TypeError(f"dim_index must be int, got {type(dim_idx).__name__}")
            if dim_idx < 0 or dim_idx >= ndim:
                raise ValueError(f"dim_index {dim_idx} out of range [0, {ndim - 1}]")
            if dim_idx in seen:
                raise ValueError(f"Duplicate dim_index {dim_idx}")
            seen.add(dim_idx)
            lo, hi = self.domain[dim_idx]
            if value < lo or value > hi:
                raise ValueError(
                    f"Slice value {value} for dim {dim_idx} is outside domain [{lo}, {hi}]")
        nodes, weights, dms = list(self.nodes), list(self.weights), list(self.diff_matrices)
        domain = [list(b) for b in self.domain]
        n_nodes = list(self.n_nodes)
        t = DT.to_device(self.tensor_values, self.device if device is None else device)
        for dim_idx, value in sorted(params, key=lambda p: p[0], reverse=True):
            t = DT.slice_axis(t, dim_idx, nodes[dim_idx], weights[dim_idx], value)
            for lst in (nodes, weights, dms, domain, n_nodes):
                del lst[dim_idx]
        return self._derived(t.cpu().numpy(), nodes, weights, dms, domain, n_nodes)

    def extrude(self, params, *, device=None) -> "ChebyshevApproximation":
        """Add dimensions along which the function is constant (reference
        ``barycentric.py:1977-2065``); replication runs on the device."""
        if self.tensor_values is None:
            raise RuntimeError("Call build() first")
        from . import device_tensor as DT

        if isinstance(params, tuple) and len(params) == 3 and isinstance(params[0], (int, np.integer)):
            params = [params]
        params = [tuple(p) for p in params]
        new_ndim = self.num_dimensions + len(params)
        seen = set()
        for dim_idx, bounds, n in params:
            if not isinstance(dim_idx, (int, np.integer)):
                raise TypeError(f"dim_index must be int, got {type(dim_idx).__name__}")
            if dim_idx < 0 or dim_idx >= new_ndim:
                raise ValueError(f"dim_index {dim_idx} out of range [0, {new_ndim - 1}]")
            if dim_idx in seen:
                raise ValueError(f"Duplicate dim_index {dim_idx}")
            seen.add(dim_idx)
            lo, hi = bounds
            if lo >= hi:
                raise ValueError(f"Domain bounds must satisfy lo < hi, got [{lo}, {hi}]")
            if not isinstance(n, (int, np.integer)) or n < 2:
                raise ValueError(f"n_nodes must be int >= 2, got {n}")
        nodes, weights, dms = list(self.nodes), list(self.weights), list(self.diff_matrices)
        domain = [list(b) for b in self.domain]
        n_nodes = list(self.n_nodes)
        t = DT.to_device(self.tensor_values, self.device if device is None else device)
        for dim_idx, (lo, hi), n in sorted(params, key=lambda p: p[0]):
            t = DT.extrude_axis(t, dim_idx, n)
            x = _grid.cheb_nodes(float(lo), float(hi), int(n))
            w = _grid.bary_weights(x)
            nodes.insert(dim_idx, x)
            weights.insert(dim_idx, w)
            dms.insert(dim_idx, _grid.diff_matrix(x, w))
            domain.insert(dim_idx, [lo, hi])
            n_nodes.insert(dim_idx, int(n))
        return self._derived(t.cpu().numpy(), nodes, weights, dms, domain, n_nodes)

    # ------------------------------------------------------------------ evaluation
    def eval_batch_multi(self, points, derivative_orders, *, out=None, device=None, algo=0):
        """Extension: N points x G derivative orders in one launch -> (N, G).

        Oracle: a loop of the reference's ``vectorized_eval_batch`` over ``derivative_orders``.
        ``points`` may be a NumPy array (host path, returns NumPy) or a CUDA ``torch.Tensor``.
        """
        return self._plan(derivative_orders, device, algo).eval(points, out)

    def vectorized_eval_batch(self, points, derivative_order=None, *, derivative_id=None,
                              out=None, device=None):
        """Evaluate at N points (reference ``barycentric.py:992-1047``) -> (N,)."""
        order = self._resolve_derivative_args(derivative_order, derivative_id)
        if self.tensor_values is None:
            raise RuntimeError("Call build() first")
        res = self._plan([order], device).eval(points, out)
        return res.reshape(res.shape[0])

    def vectorized_eval(self, point, derivative_order=None, *, derivative_id=None) -> float:
        """Single point (reference ``barycentric.py:885-949``)."""
        order = self._resolve_derivative_args(derivative_order, derivative_id)
        if self.tensor_values is None:
            raise RuntimeError("Call build() first")
        pts = _grid.point_row(point, self.num_dimensions)
        return float(self._plan([order]).eval(pts)[0, 0])

    def vectorized_eval_multi(self, point, derivative_orders) -> List[float]:
        """Single point, several derivative orders (reference ``barycentric.py:1049-1112``)."""
        if self.tensor_values is None:
            raise RuntimeError("Call build() first")
        pts = _grid.point_row(point, self.num_dimensions)
        return [float(v) for v in self._plan(derivative_orders).eval(pts)[0]]

    # the reference's scalar-loop variants return the same interpolant value
    def eval(self, point, derivative_order=None, *, derivative_id=None) -> float:
        return self.vectorized_eval(point, derivative_order, derivative_id=derivative_id)

    def fast_eval(self, point, derivative_order=None, *, derivative_id=None) -> float:
        return self.vectorized_eval(point, derivative_order, derivative_id=derivative_id)

    # ------------------------------------------------------------------ persistence
    def save(self, path, format: str = "binary") -> None:
        """``format='binary'`` writes ``.pcb`` v1; ``'pickle'`` the Python object."""
        if format == "binary":
            if self.additional_data is not None:
                raise NotImplementedError(
                    "binary format cannot store additional_data; pass format='pickle' or set "
                    "additional_data=None before saving")
            if self.tensor_values is None:
                raise RuntimeError("Cannot save an unbuilt ChebyshevApproximation")
            raw = pcbfile.approx_bytes(self.domain, self.n_nodes, self.tensor_values)
            with open(os.fspath(path), "wb") as f:
                f.write(raw)
        elif format == "pickle":
            with open(os.fspath(path), "wb") as f:
                pickle.dump(self, f)
        else:
            raise ValueError(f"format must be 'binary' or 'pickle', got {format!r}")

    @classmethod
    def load(cls, path, *, device=None) -> "ChebyshevApproximation":
        if pcbfile.is_pcb(path):
            rec = pcbfile.read(path)
            if rec["kind"] != "approx":
                raise ValueError(
                    f"file contains class_tag {pcbfile.TAG_SPLINE}, expected "
                    f"{pcbfile.TAG_APPROX} (ChebyshevApproximation)")
            return cls.from_values(rec["tensor"], rec["num_dimensions"], rec["domain"],
                                   rec["n_nodes"], device=device)
        with open(os.fspath(path), "rb") as f:
            obj = pickle.load(f)
        if not isinstance(obj, cls):
            raise TypeError(f"Expected a {cls.__name__} instance, got {type(obj).__name__}")
        return obj

    def __getstate__(self):
        state = self.__dict__.copy()
        state["function"] = None
        for k in ("_plans", "_deriv_cache", "_plan_tensor_id"):
            state.pop(k, None)
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self._reset_plans()

    def __repr__(self):
        built = self.tensor_values is not None
        return (f"ChebyshevApproximation(dims={self.num_dimensions}, nodes={self.n_nodes}, "
                f"built={built}, backend='b200')")


def _call_function(f, p, data):
    return float(f(p, data))
