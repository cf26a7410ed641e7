"""The interpolants of BASELINE.json's configs as objects OF THE UNMODIFIED REFERENCE.

TEST INFRASTRUCTURE ONLY (tests/, bench.py's CPU legs, smoke).  Every object is made by the
reference's own constructors / factories (``from_values``, ``build``, or -- for cores produced
elsewhere -- the ``__new__`` + attribute injection its factories use, ``tensor_train.py:2946-2964``)
from the synthetic workload definitions in ``pychebyshev_b200/workloads.py`` (input generation
only).  Used for same-box side-by-side parity and as the timed CPU baseline.
"""

from __future__ import annotations

import numpy as np

from pychebyshev_b200 import workloads as wl

from . import reference as R


def _nodes_1d(lo, hi, n):
    """The reference's node recipe, through its own factory (barycentric.py:448-452)."""
    ref = R.load()
    return ref.ChebyshevApproximation.from_values(np.zeros((n,)), 1, [[lo, hi]], [n]).nodes[0]


def full_from_func(func_vec, domain, n_nodes):
    """ChebyshevApproximation.from_values on the analytic function sampled at the grid."""
    ref = R.load()
    nodes = [_nodes_1d(d[0], d[1], n) for d, n in zip(domain, n_nodes)]
    tensor = wl.grid_values(func_vec, nodes)
    return ref.ChebyshevApproximation.from_values(tensor, len(n_nodes), [list(d) for d in domain],
                                                  list(n_nodes))


def full_bs5d():
    """C1: 5D Black-Scholes, 11^5 nodes (reference tests/conftest.py:86-100)."""
    return full_from_func(wl.bs_call_price, wl.BS5D_DOMAIN, wl.BS5D_NODES)


def full_c4():
    """C4: 6D, 16^6 nodes (134 MB)."""
    return full_from_func(wl.bs6d, wl.C4_DOMAIN, wl.C4_NODES)


def tt_from_cores(cores, domain, dim_order=None, max_rank=None):
    """Inject coefficient cores the way the reference's factories do (tensor_train.py:2946-2964)."""
    ref = R.load()
    D = len(cores)
    obj = ref.ChebyshevTT.__new__(ref.ChebyshevTT)
    obj.function = None
    obj.num_dimensions = D
    obj.domain = [list(map(float, b)) for b in domain]
    obj.n_nodes = [int(c.shape[1]) for c in cores]
    obj.max_rank = max_rank or max(int(c.shape[2]) for c in cores)
    obj.tolerance = 1e-6
    obj.max_sweeps = 10
    obj.max_derivative_order = 2
    obj.additional_data = None
    obj.descriptor = ""
    obj.method = "svd"
    obj._coeff_cores = [np.ascontiguousarray(c, dtype=np.float64) for c in cores]
    obj._tt_ranks = [int(c.shape[0]) for c in cores] + [int(cores[-1].shape[2])]
    obj._built = True
    obj._build_time = 0.0
    obj._total_build_evals = 0
    obj._cached_error_estimate = None
    obj._dim_order = list(range(D)) if dim_order is None else [int(v) for v in dim_order]
    return obj


def tt_bs5d_build(seed=42, max_rank=15, max_sweeps=5):
    """C2 built by the reference itself (TT-Cross, tests/conftest.py:127-135): ~2 s."""
    ref = R.load()

    def bs5(x, _):
        return wl.bs5d_scalar(x)

    tt = ref.ChebyshevTT(bs5, 5, wl.BS5D_DOMAIN, wl.BS5D_NODES, max_rank=max_rank,
                         max_sweeps=max_sweeps)
    tt.build(verbose=False, seed=seed)
    return tt


def spline_from_func(func_vec, domain, n_nodes, knots):
    ref = R.load()
    info = ref.ChebyshevSpline.nodes(len(domain), domain, n_nodes, knots)
    vals = [wl.grid_values(func_vec, p["nodes_per_dim"]) for p in info["pieces"]]
    return ref.ChebyshevSpline.from_values(vals, len(domain), domain, n_nodes, knots)


def spline2d():
    return spline_from_func(wl.payoff2d, wl.SPLINE2D_DOMAIN, wl.SPLINE2D_NODES, wl.SPLINE2D_KNOTS)


def spline3d():
    return spline_from_func(wl.payoff3d, wl.SPLINE3D_DOMAIN, wl.SPLINE3D_NODES, wl.SPLINE3D_KNOTS)


def slider10d():
    ref = R.load()
    sl = ref.ChebyshevSlider(wl.basket10d_scalar, wl.C5_DIM, wl.C5_DOMAIN, wl.C5_NODES,
                             wl.C5_PARTITION, wl.C5_PIVOT)
    sl.build(verbose=False)
    return sl


def spline_lookup(sp, pts):
    """The reference's vectorised routing lines (spline.py:677-690), executed on ``sp``'s own
    knots/shape with the same NumPy calls."""
    N = pts.shape[0]
    mi = np.zeros((N, sp.num_dimensions), dtype=int)
    for d in range(sp.num_dimensions):
        if len(sp.knots[d]) > 0:
            idx = np.searchsorted(sp.knots[d], pts[:, d], side="right")
            mi[:, d] = np.clip(idx, 0, sp._shape[d] - 1)
    return np.ravel_multi_index(mi.T, sp._shape).astype(np.int32)
