"""Drop-in acceleration of *reference* objects: ``pychebyshev`` interpolants evaluated on B200.

The reference has no FFI seam: its boundary is the Python method surface of the
four interpolant classes.  This module is the binding a reference maintainer would add
(INTEGRATION.md): ``install(pychebyshev)`` rebinds the *bodies* of the vectorised evaluation
methods

* ``ChebyshevApproximation.vectorized_eval / vectorized_eval_batch / vectorized_eval_multi``
  (``barycentric.py:885-1112``),
* ``ChebyshevSpline.eval / eval_multi / eval_batch`` (``spline.py:552-700``),
* ``ChebyshevTT.eval / eval_batch / eval_multi`` (``tensor_train.py:2127-2320``),
* ``ChebyshevSlider.eval / eval_multi`` (``slider.py:247-337``)

to the CUDA engine; constructors, build, algebra, calculus, serialisation stay the reference's own
code and keep producing the arrays the engine consumes.  Signatures, return types and exceptions
are the reference's (argument resolution is done by the reference's own
``_resolve_derivative_args``; the "not built" / knot / order checks are reproduced by the mirror
classes of this package).

``adopt(ref_obj)`` returns the mirror object of this package that *shares* the reference object's
arrays (nodes, weights, differentiation matrices, tensors, cores) -- nothing is recomputed.  A
mirror is cached per reference object (weakly) and rebuilt when the object's arrays are rebound or
written in place (``_engine.fingerprint``).

There is no CPU fallback: with ``install()`` active and no B200 visible the patched methods raise
``BackendUnavailable``.
"""

from __future__ import annotations

import threading
import weakref

import numpy as np

from . import _engine
from .approximation import ChebyshevApproximation
from .slider import ChebyshevSlider
from .spline import ChebyshevSpline
from .tt import ChebyshevTT

_lock = threading.Lock()
_mirrors: dict = {}       # id(ref_obj) -> (weakref, token, mirror)
_patched: dict = {}       # (class, name) -> original function
_device = None


# ------------------------------------------------------------------------------------------
# adoption: reference object -> mirror sharing its arrays
# ------------------------------------------------------------------------------------------

def _adopt_approx(ref) -> ChebyshevApproximation:
    obj = object.__new__(ChebyshevApproximation)
    obj.function = None
    obj.num_dimensions = int(ref.num_dimensions)
    obj.domain = [list(b) for b in ref.domain]
    obj.n_nodes = [int(n) for n in ref.n_nodes]
    obj.max_derivative_order = ref.max_derivative_order
    obj.error_threshold = None
    obj.max_n = getattr(ref, "max_n", 64)
    obj.special_points = None
    obj.additional_data = None
    obj.n_workers = None
    obj.descriptor = getattr(ref, "descriptor", "")
    obj.device = _device
    obj.build_time = 0.0
    obj.n_evaluations = 0
    obj.tensor_values = ref.tensor_values
    obj._init_derivative_ids()
    obj._reset_plans()
    obj.nodes, obj.weights, obj.diff_matrices = ref.nodes, ref.weights, ref.diff_matrices
    return obj


def _adopt_spline(ref) -> ChebyshevSpline:
    obj = object.__new__(ChebyshevSpline)
    obj.function = None
    obj.num_dimensions = int(ref.num_dimensions)
    obj.domain = ref.domain
    obj.n_nodes = ref.n_nodes
    obj.knots = ref.knots
    obj.max_derivative_order = ref.max_derivative_order
    obj.error_threshold = None
    obj.max_n = getattr(ref, "max_n", 64)
    obj.additional_data = None
    obj.n_workers = None
    obj.descriptor = getattr(ref, "descriptor", "")
    obj.device = _device
    obj._init_derivative_ids()
    obj._n_nodes_nested = bool(getattr(ref, "_n_nodes_nested", False))
    obj._intervals = ref._intervals
    obj._shape = tuple(ref._shape)
    built = bool(getattr(ref, "_built", False)) and all(
        p is not None and p.tensor_values is not None for p in ref._pieces)
    obj._pieces = [_adopt_approx(p) for p in ref._pieces] if built else [None] * len(ref._pieces)
    obj._built = built
    obj._build_time = 0.0
    obj._plans = {}
    return obj


def _adopt_tt(ref) -> ChebyshevTT:
    if not getattr(ref, "_built", False) and getattr(ref, "_coeff_cores", None) is None:
        obj = ChebyshevTT(None, ref.num_dimensions, ref.domain, ref.n_nodes, device=_device)
        return obj
    return ChebyshevTT.from_cores(ref._coeff_cores, ref.domain, getattr(ref, "_dim_order", None),
                                  max_derivative_order=ref.max_derivative_order, device=_device)


def _adopt_slider(ref) -> ChebyshevSlider:
    obj = ChebyshevSlider(None, ref.num_dimensions, ref.domain, ref.n_nodes, ref.partition,
                          ref.pivot_point, ref.max_derivative_order, device=_device)
    if getattr(ref, "_built", False):
        obj.slides = [_adopt_approx(s) for s in ref.slides]
        obj.pivot_value = ref.pivot_value
        obj._built = True
    return obj


def _arrays_of(ref):
    """The arrays whose identity/content define the interpolant (for cache validation)."""
    name = type(ref).__name__
    if name == "ChebyshevApproximation":
        return [ref.tensor_values]
    if name == "ChebyshevSpline":
        return [None if p is None else p.tensor_values for p in ref._pieces]
    if name == "ChebyshevTT":
        return list(ref._coeff_cores or []) + [tuple(getattr(ref, "_dim_order", ()))]
    if name == "ChebyshevSlider":
        return [s.tensor_values for s in ref.slides] + [float(ref.pivot_value)]
    raise TypeError(f"not a PyChebyshev interpolant: {type(ref).__name__}")


_ADOPT = {"ChebyshevApproximation": _adopt_approx, "ChebyshevSpline": _adopt_spline,
          "ChebyshevTT": _adopt_tt, "ChebyshevSlider": _adopt_slider}


def adopt(ref, *, cached: bool = True):
    """Mirror of ``ref`` on the B200 engine, sharing its arrays."""
    make = _ADOPT.get(type(ref).__name__)
    if make is None:
        raise TypeError(f"not a PyChebyshev interpolant: {type(ref).__name__}")
    if not cached:
        return make(ref)
    arrays = _arrays_of(ref)
    token = tuple(_engine.fingerprint(a) if isinstance(a, np.ndarray) else a for a in arrays)
    key = id(ref)
    with _lock:
        hit = _mirrors.get(key)
        if hit is not None and hit[0]() is ref and hit[1] == token:
            return hit[3]
        mirror = make(ref)
        # `arrays` is kept so a freed-and-reallocated array cannot alias a cached identity
        _mirrors[key] = (weakref.ref(ref, lambda _r, k=key: _mirrors.pop(k, None)), token,
                         arrays, mirror)
        return mirror


# ------------------------------------------------------------------------------------------
# patched method bodies
# ------------------------------------------------------------------------------------------

def _approx_vectorized_eval(self, point, derivative_order=None, *, derivative_id=None):
    order = self._resolve_derivative_args(derivative_order, derivative_id)
    return adopt(self).vectorized_eval(point, order)


def _approx_vectorized_eval_batch(self, points, derivative_order=None, *, derivative_id=None):
    order = self._resolve_derivative_args(derivative_order, derivative_id)
    return adopt(self).vectorized_eval_batch(points, order)


def _approx_vectorized_eval_multi(self, point, derivative_orders):
    return adopt(self).vectorized_eval_multi(point, derivative_orders)


def _spline_eval(self, point, derivative_order=None, *, derivative_id=None):
    if not self._built:
        raise RuntimeError("Call build() before eval().")
    order = self._resolve_derivative_args(derivative_order, derivative_id)
    return adopt(self).eval(point, order)


def _spline_eval_multi(self, point, derivative_orders):
    return adopt(self).eval_multi(point, derivative_orders)


def _spline_eval_batch(self, points, derivative_order=None, *, derivative_id=None):
    if not self._built:
        raise RuntimeError("Call build() before eval_batch().")
    order = self._resolve_derivative_args(derivative_order, derivative_id)
    return adopt(self).eval_batch(np.asarray(points), order)


def _tt_eval(self, point):
    self._check_built()
    return adopt(self).eval(point)


def _tt_eval_batch(self, points):
    self._check_built()
    return adopt(self).eval_batch(np.asarray(points))


def _tt_eval_multi(self, point, derivative_orders):
    self._check_built()
    return adopt(self).eval_multi(point, derivative_orders)


def _slider_eval(self, point, derivative_order=None, *, derivative_id=None):
    if not self._built:
        raise RuntimeError("Call build() before eval().")
    order = self._resolve_derivative_args(derivative_order, derivative_id)
    return adopt(self).eval(point, order)


def _slider_eval_multi(self, point, derivative_orders):
    if not self._built:
        raise RuntimeError("Call build() before eval().")
    return adopt(self).eval_multi(point, derivative_orders)


_TABLE = {
    "ChebyshevApproximation": {
        "vectorized_eval": _approx_vectorized_eval,
        "vectorized_eval_batch": _approx_vectorized_eval_batch,
        "vectorized_eval_multi": _approx_vectorized_eval_multi,
    },
    "ChebyshevSpline": {"eval": _spline_eval, "eval_multi": _spline_eval_multi,
                        "eval_batch": _spline_eval_batch},
    "ChebyshevTT": {"eval": _tt_eval, "eval_batch": _tt_eval_batch, "eval_multi": _tt_eval_multi},
    "ChebyshevSlider": {"eval": _slider_eval, "eval_multi": _slider_eval_multi},
}


def install(module=None, *, device=None) -> None:
    """Rebind the reference's vectorised evaluation methods to the B200 engine.

    ``module``: the imported reference package (default: ``import pychebyshev``)."""
    global _device
    if module is None:
        import pychebyshev as module  # noqa: PLC0415
    _engine.require_device(device)  # fail loudly: no CPU fallback behind the patched methods
    _device = device
    with _lock:
        for cname, methods in _TABLE.items():
            cls = getattr(module, cname)
            for mname, fn in methods.items():
                if (cls, mname) in _patched:
                    continue
                orig = cls.__dict__[mname]
                fn.__doc__ = fn.__doc__ or getattr(orig, "__doc__", None)
                _patched[(cls, mname)] = orig
                setattr(cls, mname, fn)


def uninstall() -> None:
    """Restore the reference's own method bodies."""
    with _lock:
        for (cls, mname), orig in _patched.items():
            setattr(cls, mname, orig)
        _patched.clear()
        _mirrors.clear()


def installed() -> bool:
    return bool(_patched)
