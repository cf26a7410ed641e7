"""``ChebyshevSlider``: additive sliding approximation ``f(z) + sum_i [s_i(x_i) - f(z)]`` on B200.

Mirrors the reference class (``slider.py:80-337``): same constructor, ``build``, ``eval`` and
``eval_multi``.  The reference has no batch method; :meth:`eval_batch` /
:meth:`eval_batch_multi` are the device-side extension whose oracle is a loop of the
reference's ``eval``.  Evaluation runs in ``pcb_slider_eval``.
"""

from __future__ import annotations

import os
import pickle
import time
from typing import List

import numpy as np

from . import _grid
from ._engine import SliderPlan, fingerprint, require_device
from .approximation import ChebyshevApproximation, _DerivativeIds, _unwrap


class ChebyshevSlider(_DerivativeIds):
    """Sum of low-dimensional Chebyshev slides around a pivot point."""

    def __init__(self, function, num_dimensions, domain, n_nodes, partition, pivot_point,
                 max_derivative_order=2, additional_data=None, *, device=None):
        domain = _unwrap(domain, "bounds")
        n_nodes = _unwrap(n_nodes, "counts")
        self.function = function
        self.num_dimensions = num_dimensions
        self.domain = domain
        self.n_nodes = n_nodes
        self.partition = partition
        self.pivot_point = list(pivot_point)
        self.max_derivative_order = max_derivative_order
        self.additional_data = additional_data
        self.descriptor = ""
        self.device = device
        covered = sorted(d for group in partition for d in group)
        if covered != list(range(num_dimensions)):
            raise ValueError(
                f"Partition must cover all dimensions 0..{num_dimensions-1} exactly once. "
                f"Got dimensions: {covered}")
        self._dim_to_slide = {d: s for s, group in enumerate(partition) for d in group}
        self.slides: List[ChebyshevApproximation] = []
        self.pivot_value = 0.0
        self._built = False
        self._init_derivative_ids()
        self._plans = {}

    def build(self, verbose: bool | int = True) -> None:
        """Build every slide with the other coordinates frozen at the pivot (slider.py:128-199)."""
        t0 = time.time()
        f, data = self.function, self.additional_data
        self.pivot_value = f(self.pivot_point, data)
        self.slides = []
        for group in self.partition:
            def on_slide(sub_point, d, _grp=tuple(group)):
                full = list(self.pivot_point)
                for local, dim in enumerate(_grp):
                    full[dim] = sub_point[local]
                return f(full, d)

            slide = ChebyshevApproximation(
                on_slide, len(group), [self.domain[d] for d in group],
                [self.n_nodes[d] for d in group],
                max_derivative_order=self.max_derivative_order, additional_data=data)
            slide.build(verbose=False)
            self.slides.append(slide)
        self._built = True
        self._plans = {}
        if verbose:
            print(f"Built {len(self.slides)} slides in {time.time() - t0:.3f}s")

    @classmethod
    def from_slides(cls, slide_tensors, num_dimensions, domain, n_nodes, partition, pivot_point,
                    pivot_value, max_derivative_order=2, *, device=None) -> "ChebyshevSlider":
        """Assemble a slider from per-slide value tensors (one per partition group)."""
        obj = cls(None, num_dimensions, domain, n_nodes, partition, pivot_point,
                  max_derivative_order, device=device)
        if len(slide_tensors) != len(partition):
            raise ValueError(f"Expected {len(partition)} slide tensors, got {len(slide_tensors)}")
        for group, t in zip(partition, slide_tensors):
            obj.slides.append(ChebyshevApproximation.from_values(
                t, len(group), [list(domain[d]) for d in group], [n_nodes[d] for d in group],
                max_derivative_order))
        obj.pivot_value = float(pivot_value)
        obj._built = True
        return obj

    # ------------------------------------------------------------------ device plan
    def _plan(self, orders, device=None) -> SliderPlan:
        orders = _grid.normalize_orders(orders, self.num_dimensions)
        dev = require_device(self.device if device is None else device)
        token = tuple(fingerprint(s.tensor_values) for s in self.slides) + (float(self.pivot_value),)
        key = (dev, orders)
        hit = self._plans.get(key)
        if hit is None or hit[0] != token:
            out_slide, rows = [], []
            for o in orders:
                active = {self._dim_to_slide[d] for d, k in enumerate(o) if k > 0}
                row = [None] * len(self.slides)
                if not active:      # value: pivot + sum (slide - pivot)
                    out_slide.append(-1)
                    row = [s.derivative_tensor([0] * s.num_dimensions) for s in self.slides]
                elif len(active) > 1:  # cross-slide mixed partial is exactly zero
                    out_slide.append(-2)
                else:
                    s = active.pop()
                    out_slide.append(s)
                    row[s] = self.slides[s].derivative_tensor([o[d] for d in self.partition[s]])
                rows.append(row)
            slides = [(s.n_nodes, s.nodes, s.weights) for s in self.slides]
            if len(self._plans) >= 8:
                self._plans.pop(next(iter(self._plans)))
            hit = (token, SliderPlan(self.num_dimensions, self.partition, slides, self.pivot_value,
                                     out_slide, rows, dev))
            self._plans[key] = hit
        return hit[1]

    # ------------------------------------------------------------------ evaluation
    def eval_batch_multi(self, points, derivative_orders, *, out=None, device=None):
        """Extension: N points x G derivative orders -> (N, G)."""
        if not self._built:
            raise RuntimeError("Call build() before eval().")
        return self._plan(derivative_orders, device).eval(points, out)

    def eval_batch(self, points, derivative_order=None, *, derivative_id=None, out=None,
                   device=None):
        """Extension: N points, one derivative order -> (N,)."""
        if not self._built:
            raise RuntimeError("Call build() before eval().")
        order = self._resolve_derivative_args(derivative_order, derivative_id)
        res = self._plan([order], device).eval(points, out)
        return res.reshape(res.shape[0])

    def eval(self, point, derivative_order=None, *, derivative_id=None) -> float:
        """Single point (reference ``slider.py:247-318``)."""
        if not self._built:
            raise RuntimeError("Call build() before eval().")
        order = self._resolve_derivative_args(derivative_order, derivative_id)
        pts = _grid.point_row(point, self.num_dimensions)
        return float(self._plan([order]).eval(pts)[0, 0])

    def eval_multi(self, point, derivative_orders) -> List[float]:
        """Single point, several derivative orders (reference ``slider.py:320-337``)."""
        if not self._built:
            raise RuntimeError("Call build() before eval().")
        pts = _grid.point_row(point, self.num_dimensions)
        return [float(v) for v in self._plan(derivative_orders).eval(pts)[0]]

    # ------------------------------------------------------------------ persistence
    def save(self, path) -> None:
        with open(os.fspath(path), "wb") as f:
            pickle.dump(self, f)

    @classmethod
    def load(cls, path) -> "ChebyshevSlider":
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
        return (f"ChebyshevSlider(dims={self.num_dimensions}, partition={self.partition}, "
                f"built={self._built}, backend='b200')")
