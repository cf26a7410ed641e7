"""Host-side grid numerics: Chebyshev nodes, barycentric weights, differentiation matrices and
the derivative passes that turn a value tensor into a pre-differentiated tensor.

These arrays are *inputs* of the device engine.  They are produced with the reference's own
NumPy recipes because parity at 1e-12 on Greeks requires bit-identical
derivative tensors (SURVEY.md finding 2):

* nodes      reference ``_extrude_slice.py:66-70`` / ``barycentric.py:448-452``
* weights    reference ``barycentric.py:30-49``
* D matrix   reference ``barycentric.py:52-77``
* passes     reference ``barycentric.py:951-990``
"""

from __future__ import annotations

from typing import List, Sequence

import numpy as np
from numpy.polynomial.chebyshev import chebpts1


def cheb_nodes(lo: float, hi: float, n: int) -> np.ndarray:
    """Ascending Chebyshev points of the first kind mapped to ``[lo, hi]``."""
    return np.sort(0.5 * (lo + hi) + 0.5 * (hi - lo) * chebpts1(n))


def bary_weights(x: np.ndarray) -> np.ndarray:
    """``w_i = 1 / prod_{j != i} (x_i - x_j)`` by sequential division in ``j`` order.

    The reference divides one factor at a time (so intermediate values never overflow for the
    node counts in use); doing the same division sequence for every ``i`` at once keeps the
    result bit-identical while avoiding the double Python loop.
    """
    n = len(x)
    w = np.ones(n)
    rows = np.arange(n)
    for j in range(n):
        keep = rows != j
        w[keep] /= x[keep] - x[j]
    return w


def diff_matrix(x: np.ndarray, w: np.ndarray) -> np.ndarray:
    """Spectral differentiation matrix on nodes ``x`` (Berrut & Trefethen 2004, §9.3):
    ``D_ij = (w_j / w_i) / (x_i - x_j)``, diagonal by the negative-sum trick."""
    gap = x[:, None] - x
    np.fill_diagonal(gap, 1.0)
    dm = w / (gap * w[:, None])
    np.fill_diagonal(dm, 0.0)
    np.fill_diagonal(dm, -dm.sum(axis=1))
    return dm


def grid_arrays(domain: Sequence[Sequence[float]], n_nodes: Sequence[int]):
    """(nodes, weights, diff_matrices) lists for a tensor grid."""
    nodes = [cheb_nodes(float(lo), float(hi), int(n)) for (lo, hi), n in zip(domain, n_nodes)]
    weights = [bary_weights(x) for x in nodes]
    dms = [diff_matrix(x, w) for x, w in zip(nodes, weights)]
    return nodes, weights, dms


def differentiate_tensor(tensor: np.ndarray, dms: List[np.ndarray], order) -> np.ndarray:
    """Apply ``order[d]`` passes of ``D_d^T`` along axis ``d`` for ``d = D-1 .. 0``; C-contiguous.

    Same matmul shapes and axis order as the reference so OpenBLAS sums in the same order.
    """
    out = tensor
    if order is not None:
        for d in range(tensor.ndim - 1, -1, -1):
            for _ in range(int(order[d])):
                out = np.moveaxis(np.moveaxis(out, d, -1) @ dms[d].T, -1, d)
    return np.ascontiguousarray(out, dtype=np.float64)


def full_grid_points(nodes: List[np.ndarray]) -> np.ndarray:
    """All grid points in C-order, shape (prod n, D)."""
    mesh = np.meshgrid(*nodes, indexing="ij")
    return np.stack([m.ravel() for m in mesh], axis=1)


def normalize_orders(orders, ndim: int) -> tuple:
    """Tuple-of-tuples of ints, validated for length."""
    out = []
    for o in orders:
        o = tuple(int(v) for v in o)
        if len(o) != ndim:
            raise ValueError(f"derivative_order must have {ndim} entries, got {len(o)}")
        if any(v < 0 for v in o):
            raise ValueError("derivative orders must be non-negative")
        out.append(o)
    return tuple(out)


def point_row(point, ndim: int) -> np.ndarray:
    """One query as a ``(1, ndim)`` float64 array, read the way the reference's single-point methods
    read it: ``point[d]`` for ``d < ndim`` only (longer sequences are tolerated), and a length-1
    sequence counts as its element (NumPy broadcasting makes ``[0.1] - nodes`` equal ``0.1 - nodes``
    there; the reference's own tests pass such points, tests/test_from_values.py:246-254)."""
    row = np.empty((1, ndim), dtype=np.float64)
    for d in range(ndim):
        v = np.asarray(point[d], dtype=np.float64)
        if v.size != 1:
            raise ValueError(f"point[{d}] must be a scalar, got shape {v.shape}")
        row[0, d] = v.reshape(-1)[0]
    return row

