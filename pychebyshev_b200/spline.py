"""``ChebyshevSpline``: piecewise Chebyshev interpolant with knots at kinks, evaluated on B200.

Mirrors the reference class (``spline.py:106-267, 325-445, 519-700, 1020-1357``): same
constructor, ``build`` / ``from_values`` / ``nodes`` / ``save`` / ``load``, and the evaluation
entry points ``eval`` / ``eval_multi`` / ``eval_batch``.  Piece routing is the bit-exact integer
kernel ``pcb_spline_lookup``; evaluation is ``pcb_spline_eval`` on the pieces' own grids.
"""

from __future__ import annotations

import itertools
import os
import pickle
import time
from typing import List

import numpy as np

from . import _grid, pcbfile
from ._engine import SplinePlan, fingerprint, require_device
from .approximation import ChebyshevApproximation, _DerivativeIds, _unwrap

KNOT_EPS = 1e-14  # reference spline.py:545


def _nested(n_nodes) -> bool:
    return n_nodes is not None and any(isinstance(x, (list, tuple)) for x in n_nodes)


def validate_special_points_shape(special_points, n_nodes, num_dimensions, domain) -> None:
    """Checks done by the reference before the special-points dispatch (barycentric.py:149-206)."""
    for d in range(num_dimensions):
        lo, hi = domain[d]
        pts = list(special_points[d])
        for k in pts:
            if not (lo < k < hi):
                raise ValueError(
                    f"Special point {k} for dimension {d} is not strictly inside domain [{lo}, {hi}]")
        if pts != sorted(pts):
            raise ValueError(f"special_points for dimension {d} must be sorted")
        if len(set(pts)) != len(pts):
            raise ValueError(f"Coinciding special points in dimension {d}")
    if n_nodes is None:
        return
    flags = [isinstance(x, (list, tuple)) for x in n_nodes]
    if any(flags) and not all(flags):
        raise ValueError(
            f"n_nodes must be fully nested (all dims as lists) when any dim is nested; got mixed "
            f"form {n_nodes!r}")
    if not all(flags):
        raise ValueError(
            f"n_nodes must be nested as List[List[int]] when special_points is present; got "
            f"{n_nodes!r}")
    for d in range(num_dimensions):
        want, got = len(special_points[d]) + 1, len(n_nodes[d])
        if got != want:
            raise ValueError(f"n_nodes[{d}] must have {want} entries (one per sub-interval); got {got}")


def _check_knots(num_dimensions, domain, knots, dup_check=True):
    for d in range(num_dimensions):
        lo, hi = domain[d]
        for k in knots[d]:
            if not (lo < k < hi):
                raise ValueError(
                    f"Knot {k} for dimension {d} is not strictly inside domain [{lo}, {hi}]")
        if list(knots[d]) != sorted(knots[d]):
            raise ValueError(f"Knots for dimension {d} must be sorted")
        if dup_check and len(knots[d]) != len(set(knots[d])):
            raise ValueError(f"Knots for dimension {d} contain duplicates")


def _intervals(domain, knots):
    out = []
    for (lo, hi), ks in zip(domain, knots):
        edges = [lo] + list(ks) + [hi]
        out.append([(edges[i], edges[i + 1]) for i in range(len(edges) - 1)])
    return out


class ChebyshevSpline(_DerivativeIds):
    """Piecewise Chebyshev interpolation with user-specified knots."""

    def __init__(self, function, num_dimensions, domain, n_nodes=None, knots=None,
                 max_derivative_order=2, error_threshold=None, max_n=64, additional_data=None, *,
                 defer_build=False, n_workers=None, device=None):
        domain = _unwrap(domain, "bounds")
        n_nodes = _unwrap(n_nodes, "counts")
        if max_n < 3:
            raise ValueError(f"max_n must be at least 3, got max_n={max_n}.")
        if n_nodes is None:
            if error_threshold is None:
                raise ValueError(
                    "Must provide either n_nodes (explicit) or error_threshold (auto-N). Got neither.")
            raise NotImplementedError("auto-N (error_threshold) builds are outside this package")
        n_nodes = list(n_nodes)
        self._n_nodes_nested = _nested(n_nodes)
        if self._n_nodes_nested and not all(isinstance(x, (list, tuple)) for x in n_nodes):
            raise ValueError(
                "n_nodes must be fully nested (all dims as lists) when any dim is nested; got "
                "mixed form")
        flat = [x for row in n_nodes for x in row] if self._n_nodes_nested else n_nodes
        if any(x is None for x in flat):
            if error_threshold is None:
                raise ValueError(
                    "None entries in n_nodes require error_threshold to be set (auto-N mode).")
            raise NotImplementedError("auto-N (error_threshold) builds are outside this package")
        if knots is None:
            knots = [[] for _ in range(num_dimensions)]
        self.function = function
        self.num_dimensions = int(num_dimensions)
        self.domain = domain
        self.n_nodes = n_nodes
        self.knots = knots
        self.max_derivative_order = max_derivative_order
        self.error_threshold = error_threshold
        self.max_n = max_n
        self.additional_data = additional_data
        self.n_workers = n_workers
        self.descriptor = ""
        self.device = device
        self._init_derivative_ids()
        _check_knots(num_dimensions, domain, knots, dup_check=False)
        self._intervals = _intervals(domain, knots)
        self._shape = tuple(len(iv) for iv in self._intervals)
        if self._n_nodes_nested:
            for d in range(num_dimensions):
                want, got = len(knots[d]) + 1, len(n_nodes[d])
                if got != want:
                    raise ValueError(
                        f"n_nodes[{d}] must have {want} entries (one per sub-interval); got {got}")
                n_nodes[d] = list(n_nodes[d])
        self._pieces: List[ChebyshevApproximation | None] = [None] * int(np.prod(self._shape))
        self._built = False
        self._build_time = 0.0
        self._plans = {}
        if defer_build:
            if function is not None:
                raise ValueError("defer_build=True requires function=None")
            for flat_idx, mi in enumerate(itertools.product(*[range(s) for s in self._shape])):
                self._pieces[flat_idx] = ChebyshevApproximation(
                    None, self.num_dimensions, self._sub_domain(mi), self._piece_n_nodes(mi),
                    max_derivative_order=max_derivative_order, defer_build=True)

    # ------------------------------------------------------------------ construction
    def _sub_domain(self, mi):
        return [list(self._intervals[d][mi[d]]) for d in range(self.num_dimensions)]

    def _piece_n_nodes(self, mi):
        if self._n_nodes_nested:
            return [self.n_nodes[d][mi[d]] for d in range(self.num_dimensions)]
        return list(self.n_nodes)

    @property
    def num_pieces(self) -> int:
        return len(self._pieces)

    def build(self, verbose: bool | int = True) -> None:
        """Build every piece on its own sub-domain (reference ``spline.py:325-412``)."""
        if self.function is None:
            raise RuntimeError(
                "Cannot build: no function assigned. This object was created via from_values() "
                "or load().")
        t0 = time.time()
        if verbose:
            print(f"Building {self.num_dimensions}D Chebyshev Spline ({len(self._pieces)} pieces)...")
        for flat_idx, mi in enumerate(itertools.product(*[range(s) for s in self._shape])):
            piece = ChebyshevApproximation(
                self.function, self.num_dimensions, self._sub_domain(mi), self._piece_n_nodes(mi),
                max_derivative_order=self.max_derivative_order,
                additional_data=self.additional_data, n_workers=self.n_workers)
            piece.build(verbose=False)
            self._pieces[flat_idx] = piece
        self._build_time = time.time() - t0
        self._built = True
        self._plans = {}
        if verbose:
            print(f"Build complete in {self._build_time:.3f}s")

    @staticmethod
    def nodes(num_dimensions, domain, n_nodes, knots) -> dict:
        """Per-piece grids for :meth:`from_values` (reference ``spline.py:1104-1215``)."""
        if _nested(n_nodes):
            raise NotImplementedError(
                "ChebyshevSpline.nodes() accepts only flat n_nodes (one int per dim, shared "
                "across pieces).")
        for d in range(num_dimensions):
            if domain[d][0] >= domain[d][1]:
                raise ValueError(
                    f"domain[{d}]: lo={domain[d][0]} must be strictly less than hi={domain[d][1]}")
        _check_knots(num_dimensions, domain, knots)
        ivs = _intervals(domain, knots)
        shape = tuple(len(iv) for iv in ivs)
        pieces = []
        for mi in np.ndindex(*shape):
            sub = [ivs[d][mi[d]] for d in range(num_dimensions)]
            info = ChebyshevApproximation.nodes(num_dimensions, [list(s) for s in sub], n_nodes)
            pieces.append({"piece_index": mi, "sub_domain": sub,
                           "nodes_per_dim": info["nodes_per_dim"], "full_grid": info["full_grid"],
                           "shape": info["shape"]})
        return {"pieces": pieces, "num_pieces": int(np.prod(shape)), "piece_shape": shape}

    @classmethod
    def from_values(cls, piece_values, num_dimensions, domain, n_nodes, knots,
                    max_derivative_order=2, *, device=None) -> "ChebyshevSpline":
        """Spline from per-piece value tensors in C-order (reference ``spline.py:1217-1357``)."""
        if _nested(n_nodes):
            raise NotImplementedError("from_values() accepts only flat n_nodes")
        if len(domain) != num_dimensions or len(n_nodes) != num_dimensions or \
                len(knots) != num_dimensions:
            raise ValueError("domain, n_nodes and knots must all have num_dimensions entries")
        for d in range(num_dimensions):
            if domain[d][0] >= domain[d][1]:
                raise ValueError(
                    f"domain[{d}]: lo={domain[d][0]} must be strictly less than hi={domain[d][1]}")
        _check_knots(num_dimensions, domain, knots)
        obj = cls(None, num_dimensions, [list(b) for b in domain], list(n_nodes),
                  [list(k) for k in knots], max_derivative_order, device=device)
        if len(piece_values) != len(obj._pieces):
            raise ValueError(f"Expected {len(obj._pieces)} piece_values, got {len(piece_values)}")
        for i, pv in enumerate(piece_values):
            if np.asarray(pv).shape != tuple(n_nodes):
                raise ValueError(
                    f"piece_values[{i}] has shape {np.asarray(pv).shape}, expected {tuple(n_nodes)}")
        for flat_idx, mi in enumerate(np.ndindex(*obj._shape)):
            obj._pieces[flat_idx] = ChebyshevApproximation.from_values(
                piece_values[flat_idx], num_dimensions, obj._sub_domain(mi), list(n_nodes),
                max_derivative_order=max_derivative_order)
        obj._built = True
        return obj

    # ------------------------------------------------------------------ device plans
    def _plan(self, orders, device=None) -> SplinePlan:
        orders = _grid.normalize_orders(orders, self.num_dimensions)
        dev = require_device(self.device if device is None else device)
        token = tuple(fingerprint(p.tensor_values) for p in self._pieces)
        key = (dev, orders)
        hit = self._plans.get(key)
        if hit is None or hit[0] != token:
            if len(self._plans) >= 8:
                self._plans.pop(next(iter(self._plans)))
            pieces = [(p.n_nodes, p.nodes, p.weights, [p.derivative_tensor(o) for o in orders])
                      for p in self._pieces]
            hit = (token, SplinePlan(self.knots, pieces, dev))
            self._plans[key] = hit
        return hit[1]

    # ------------------------------------------------------------------ evaluation
    def find_pieces(self, points, *, device=None):
        """C-order flat piece index per point (int32); the device twin of ``_find_piece``."""
        if not self._built:
            raise RuntimeError("Call build() before eval_batch().")
        return self._plan([[0] * self.num_dimensions], device).lookup(points)

    def _find_piece(self, point):
        flat = int(self.find_pieces(_grid.point_row(point, self.num_dimensions))[0])
        return flat, self._pieces[flat]

    def _check_knot_boundary(self, point, derivative_order) -> None:
        """Derivatives are undefined on a knot (reference ``spline.py:519-550``)."""
        if all(o == 0 for o in derivative_order):
            return
        for d in range(self.num_dimensions):
            if derivative_order[d] > 0:
                for k in self.knots[d]:
                    if abs(point[d] - k) < KNOT_EPS:
                        raise ValueError(
                            f"Derivative w.r.t. dimension {d} is not defined at knot x[{d}]={k}. "
                            "The left and right derivatives may differ at this point.")

    def eval_batch_multi(self, points, derivative_orders, *, out=None, device=None):
        """Extension: N points x G derivative orders in one launch -> (N, G)."""
        if not self._built:
            raise RuntimeError("Call build() before eval_batch().")
        return self._plan(derivative_orders, device).eval(points, out)

    def eval_batch(self, points, derivative_order=None, *, derivative_id=None, out=None,
                   device=None):
        """Evaluate at N points (reference ``spline.py:633-700``) -> (N,).

        A point exactly on a knot is routed to the right-hand piece; no knot check is made for
        derivatives (the reference's batch method does not make one either).
        """
        if not self._built:
            raise RuntimeError("Call build() before eval_batch().")
        order = self._resolve_derivative_args(derivative_order, derivative_id)
        res = self._plan([order], device).eval(points, out)
        return res.reshape(res.shape[0])

    def eval(self, point, derivative_order=None, *, derivative_id=None) -> float:
        """Single point (reference ``spline.py:552-595``)."""
        if not self._built:
            raise RuntimeError("Call build() before eval().")
        order = self._resolve_derivative_args(derivative_order, derivative_id)
        self._check_knot_boundary(point, order)
        pts = _grid.point_row(point, self.num_dimensions)
        return float(self._plan([order]).eval(pts)[0, 0])

    def eval_multi(self, point, derivative_orders) -> List[float]:
        """Single point, several derivative orders (reference ``spline.py:597-631``)."""
        if not self._built:
            raise RuntimeError("Call build() before eval_multi().")
        for o in derivative_orders:
            self._check_knot_boundary(point, o)
        pts = _grid.point_row(point, self.num_dimensions)
        return [float(v) for v in self._plan(derivative_orders).eval(pts)[0]]

    # ------------------------------------------------------------------ persistence
    def save(self, path, format: str = "binary") -> None:
        if format == "binary":
            if any(p is None or p.tensor_values is None for p in self._pieces):
                raise RuntimeError("Cannot save an unbuilt ChebyshevSpline")
            if self.additional_data is not None:
                raise NotImplementedError(
                    "binary format cannot store additional_data; pass format='pickle' or set "
                    "additional_data=None before saving")
            if self._n_nodes_nested:
                raise NotImplementedError(
                    "binary format requires flat n_nodes (shared across pieces); use "
                    "format='pickle' for nested-n_nodes splines")
            raw = pcbfile.spline_bytes(self.domain, self.n_nodes, self.knots,
                                       [p.tensor_values for p in self._pieces])
            with open(os.fspath(path), "wb") as f:
                f.write(raw)
        elif format == "pickle":
            with open(os.fspath(path), "wb") as f:
                pickle.dump(self, f)
        else:
            raise ValueError(f"format must be 'binary' or 'pickle', got {format!r}")

    @classmethod
    def load(cls, path, *, device=None) -> "ChebyshevSpline":
        if pcbfile.is_pcb(path):
            rec = pcbfile.read(path)
            if rec["kind"] != "spline":
                raise ValueError(
                    f"file contains class_tag {pcbfile.TAG_APPROX}, expected "
                    f"{pcbfile.TAG_SPLINE} (ChebyshevSpline)")
            return cls.from_values(rec["pieces"], rec["num_dimensions"], rec["domain"],
                                   rec["n_nodes"], rec["knots"], device=device)
        with open(os.fspath(path), "rb") as f:
            obj = pickle.load(f)
        if not isinstance(obj, cls):
            raise TypeError(f"Expected a {cls.__name__} instance, got {type(obj).__name__}")
        return obj

    def __getstate__(self):
        state = self.__dict__.copy()
        state["function"] = None
        state.pop("_plans", None)
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self._plans = {}

    def __repr__(self):
        return (f"ChebyshevSpline(dims={self.num_dimensions}, pieces={self._shape}, "
                f"built={self._built}, backend='b200')")
