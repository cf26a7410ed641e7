"""CPU oracle (NumPy) for the vectorised evaluation path of PyChebyshev.

TEST INFRASTRUCTURE ONLY.  This module is a restatement, in plain NumPy, of
the reference's evaluation algorithms.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import it, and only as the checker or the timed CPU
baseline -- never as part of the product path (``pychebyshev_b200`` raises
if its CUDA library is missing; it has no CPU fallback).

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks every function
below against outputs of the unmodified reference (v0.21.1) run in the build
container and stored by ``tests/golden/make_golden.py``.  The reference holds
no golden *numeric* evaluation vectors of its own; its
``.pcb`` golden bytes are pinned in ``tests/test_pcb_format.py``.

Each function cites the reference lines it follows (paths relative to
``/root/reference/src/pychebyshev/``).
"""

from __future__ import annotations

import numpy as np

NODE_EPS = 1e-14  # barycentric.py:1040 -- absolute, physical units


# ------------------------------------------------------------------------------------------
# full tensor (ChebyshevApproximation)
# ------------------------------------------------------------------------------------------

def apply_derivative_passes(tensor, diff_matrices, order):
    """barycentric.py:951-990: ``T <- T x_d (D_d^T)^order[d]`` for d = D-1..0."""
    result = tensor
    if order is None:
        return result
    for d in range(tensor.ndim - 1, -1, -1):
        for _ in range(int(order[d])):
            moved = np.moveaxis(result, d, -1) @ diff_matrices[d].T
            result = np.moveaxis(moved, -1, d)
    return result


def _contract_last(current, vec):
    """barycentric.py:871-883: contract the last axis with a vector."""
    if current.ndim > 2:
        lead = current.shape[:-1]
        return (current.reshape(-1, current.shape[-1]) @ vec).reshape(lead)
    return current @ vec


def full_eval_batch(tensor, nodes, weights, diff_matrices, points, order):
    """barycentric.py:992-1047: hoisted derivative passes, then per point the
    barycentric reduction from the last axis to the first."""
    points = np.asarray(points, dtype=np.float64)
    t = apply_derivative_passes(np.asarray(tensor, dtype=np.float64), diff_matrices, order)
    D = t.ndim
    out = np.empty(points.shape[0])
    for i in range(points.shape[0]):
        cur = t
        for d in range(D - 1, -1, -1):
            diff = points[i, d] - nodes[d]
            hit = np.where(np.abs(diff) < NODE_EPS)[0]
            if len(hit) > 0:
                cur = cur[..., hit[0]]
            else:
                w = weights[d] / diff
                cur = _contract_last(cur, w) / np.sum(w)
        out[i] = float(cur)
    return out


def full_eval_multi_batch(tensor, nodes, weights, diff_matrices, points, orders):
    """N x G table: a loop of :func:`full_eval_batch` over derivative orders
    (the oracle of the batched-multi extension)."""
    return np.stack(
        [full_eval_batch(tensor, nodes, weights, diff_matrices, points, o) for o in orders], axis=1
    )


# ------------------------------------------------------------------------------------------
# tensor train (ChebyshevTT)
# ------------------------------------------------------------------------------------------

def cheb_basis(scaled, n):
    """``chebval(scaled, eye(n)).T`` (tensor_train.py:2257-2259): Clenshaw on the
    unit coefficient vectors, NumPy's chebval recurrence restated.  Returns (N, n)."""
    x = np.asarray(scaled, dtype=np.float64)
    eye = np.eye(n)
    if n == 1:
        return np.ones(x.shape + (1,))
    if n == 2:
        c0 = eye[0][:, None] + 0.0 * x
        c1 = eye[1][:, None] + 0.0 * x
    else:
        x2 = 2.0 * x
        c0 = eye[-2][:, None] + 0.0 * x
        c1 = eye[-1][:, None] + 0.0 * x
        for i in range(3, n + 1):
            tmp = c0
            c0 = eye[-i][:, None] - c1
            c1 = tmp + c1 * x2
    return (c0 + c1 * x).T


def tt_eval_batch(cores, domain, dim_order, points):
    """tensor_train.py:2217-2265.  ``domain`` and ``cores`` are in storage frame;
    ``points`` columns are in the user frame and are permuted by ``dim_order``."""
    pts = np.asarray(points, dtype=np.float64)
    D = len(cores)
    if list(dim_order) != list(range(D)):
        pts = pts[:, list(dim_order)]
    result = np.ones((pts.shape[0], 1, 1))
    for d in range(D):
        a, b = domain[d]
        scaled = 2.0 * (pts[:, d] - a) / (b - a) - 1.0
        q = cheb_basis(scaled, cores[d].shape[1])
        v = np.einsum("nj,ijk->nik", q, cores[d])
        result = np.einsum("nij,njk->nik", result, v)
    return result[:, 0, 0]


def _tt_value_storage(cores, domain, p):
    """tensor_train.py:2199-2214: single point, storage frame."""
    result = np.ones((1, 1))
    for d in range(len(cores)):
        a, b = domain[d]
        scaled = 2.0 * (p[d] - a) / (b - a) - 1.0
        q = cheb_basis(np.array([scaled]), cores[d].shape[1])[0]
        result = result @ np.einsum("j,ijk->ik", q, cores[d])
    return float(result[0, 0])


def _nudge(domain, p, d, h):
    """tensor_train.py:2361-2370 (strict ``<``; two independent ifs)."""
    p = list(p)
    a, b = domain[d]
    need = h * 1.5
    if p[d] - a < need:
        p[d] = a + need
    if b - p[d] < need:
        p[d] = b - need
    return p


def _fd_nested(cores, domain, p, active):
    """tensor_train.py:2428-2463."""
    if not active:
        return _tt_value_storage(cores, domain, p)
    d, order = active[0]
    rest = active[1:]
    a, b = domain[d]
    h = (b - a) * 1e-4
    pt = _nudge(domain, p, d, h)
    up, dn = list(pt), list(pt)
    up[d] += h
    dn[d] -= h
    if order == 1:
        return (_fd_nested(cores, domain, up, rest) - _fd_nested(cores, domain, dn, rest)) / (2.0 * h)
    if order == 2:
        return (
            _fd_nested(cores, domain, up, rest)
            - 2.0 * _fd_nested(cores, domain, pt, rest)
            + _fd_nested(cores, domain, dn, rest)
        ) / (h * h)
    raise ValueError(f"Derivative order {order} not supported (use 1 or 2)")


def tt_eval_multi_point(cores, domain, dim_order, point, orders):
    """tensor_train.py:2267-2463: value or central finite differences, one point."""
    D = len(cores)
    ident = list(dim_order) == list(range(D))
    p = list(point) if ident else [point[dim_order[k]] for k in range(D)]
    out = []
    for o in orders:
        o = list(o) if ident else [o[dim_order[k]] for k in range(D)]
        active = [(d, k) for d, k in enumerate(o) if k > 0]
        if not active:
            out.append(_tt_value_storage(cores, domain, p))
        elif len(active) == 1:
            d, k = active[0]
            if k not in (1, 2):
                raise ValueError(f"Derivative order {k} not supported (use 1 or 2)")
            out.append(_fd_nested(cores, domain, p, active))  # == _fd_single_dim, :2372-2403
        elif len(active) == 2 and active[0][1] == 1 and active[1][1] == 1:
            (d1, _), (d2, _) = active  # tensor_train.py:2405-2426
            h1 = (domain[d1][1] - domain[d1][0]) * 1e-4
            h2 = (domain[d2][1] - domain[d2][0]) * 1e-4
            pt = _nudge(domain, _nudge(domain, p, d1, h1), d2, h2)

            def at(s1, s2):
                q = list(pt)
                q[d1] += s1
                q[d2] += s2
                return _tt_value_storage(cores, domain, q)

            out.append((at(+h1, +h2) - at(+h1, -h2) - at(-h1, +h2) + at(-h1, -h2)) / (4.0 * h1 * h2))
        else:
            out.append(_fd_nested(cores, domain, p, active))
    return out


def tt_eval_multi_batch(cores, domain, dim_order, points, orders):
    """N x G table of :func:`tt_eval_multi_point` (oracle of the batched FD extension)."""
    return np.array([tt_eval_multi_point(cores, domain, dim_order, list(map(float, p)), orders)
                     for p in np.asarray(points, dtype=np.float64)])


# ------------------------------------------------------------------------------------------
# spline (ChebyshevSpline)
# ------------------------------------------------------------------------------------------

def spline_lookup(knots, shape, points):
    """spline.py:677-690: integer piece index, C-order over ``shape``."""
    pts = np.asarray(points, dtype=np.float64)
    mi = np.zeros((pts.shape[0], len(shape)), dtype=int)
    for d in range(len(shape)):
        if len(knots[d]) > 0:
            mi[:, d] = np.searchsorted(knots[d], pts[:, d], side="right")
            np.clip(mi[:, d], 0, shape[d] - 1, out=mi[:, d])
    return np.ravel_multi_index(mi.T, tuple(shape)).astype(np.int32)


def spline_eval_batch(knots, shape, pieces, points, order):
    """spline.py:633-700.  ``pieces`` is the flat C-order list of
    ``(tensor, nodes, weights, diff_matrices)`` tuples."""
    pts = np.asarray(points, dtype=np.float64)
    flat = spline_lookup(knots, shape, pts)
    out = np.empty(pts.shape[0])
    for idx in np.unique(flat):
        mask = flat == idx
        t, nodes, weights, dms = pieces[idx]
        out[mask] = full_eval_batch(t, nodes, weights, dms, pts[mask], order)
    return out


# ------------------------------------------------------------------------------------------
# slider (ChebyshevSlider)
# ------------------------------------------------------------------------------------------

def slider_eval_batch(partition, pivot_value, slides, points, order):
    """slider.py:247-318 applied to each row of ``points``.  ``slides`` is a
    list of ``(tensor, nodes, weights, diff_matrices)``."""
    pts = np.asarray(points, dtype=np.float64)
    dim_to_slide = {d: s for s, grp in enumerate(partition) for d in grp}
    active = {dim_to_slide[d] for d, k in enumerate(order) if k > 0}
    if len(active) > 1:
        return np.zeros(pts.shape[0])
    if len(active) == 1:
        s = active.pop()
        grp = list(partition[s])
        t, nodes, weights, dms = slides[s]
        return full_eval_batch(t, nodes, weights, dms, pts[:, grp], [order[d] for d in grp])
    result = np.full(pts.shape[0], float(pivot_value))
    for s, grp in enumerate(partition):
        t, nodes, weights, dms = slides[s]
        val = full_eval_batch(t, nodes, weights, dms, pts[:, list(grp)], [0] * len(grp))
        result = result + (val - pivot_value)
    return result


# ------------------------------------------------------------------------------------------
# grid recipes (inputs of the engine)
# ------------------------------------------------------------------------------------------

def make_nodes(lo, hi, n):
    """_extrude_slice.py:66-70 / barycentric.py:448-452."""
    std = np.sin(0.5 * np.pi / n * np.arange(-n + 1, n + 1, 2))  # numpy chebpts1
    return np.sort(0.5 * (lo + hi) + 0.5 * (hi - lo) * std)


def barycentric_weights(nodes):
    """barycentric.py:30-49 (sequential division)."""
    n = len(nodes)
    w = np.ones(n)
    for i in range(n):
        for j in range(n):
            if j != i:
                w[i] /= nodes[i] - nodes[j]
    return w


def diff_matrix(nodes, weights):
    """barycentric.py:52-77."""
    c = nodes[:, None] - nodes
    np.fill_diagonal(c, 1.0)
    c = weights / (c * weights[:, None])
    np.fill_diagonal(c, 0.0)
    np.fill_diagonal(c, -c.sum(axis=1))
    return c
