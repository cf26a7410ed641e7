"""Loaders for tests/golden/*.npz (written by tests/golden/make_golden.py from the reference)."""

import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

REL, ABS = 1e-12, 1e-14  # BASELINE.json north_star: 1e-12 relative + 1e-14 absolute


def load(name):
    with np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def split(cat, sizes):
    out, off = [], 0
    for s in sizes:
        out.append(np.ascontiguousarray(cat[off:off + s]))
        off += s
    assert off == len(cat), (off, len(cat))
    return out


def full_parts(g):
    """nodes, weights, diff matrices (lists per dim) of a full_* fixture."""
    n = [int(v) for v in g["n_nodes"]]
    nodes = split(g["nodes_cat"], n)
    weights = split(g["weights_cat"], n)
    dms = [m.reshape(k, k) for m, k in zip(split(g["diff_cat"], [k * k for k in n]), n)]
    return n, nodes, weights, dms


def tt_parts(g):
    n = [int(v) for v in g["n_nodes"]]
    r = [int(v) for v in g["ranks"]]
    sizes = [r[k] * n[k] * r[k + 1] for k in range(len(n))]
    cores = [c.reshape(r[k], n[k], r[k + 1]) for k, c in enumerate(split(g["cores_cat"], sizes))]
    domain = [list(map(float, row)) for row in g["domain"]]
    dim_order = [int(v) for v in g["dim_order"]]
    return cores, domain, dim_order


def spline_parts(g, make_dm):
    """knots (list per dim), shape, pieces [(tensor, nodes, weights, dms)] in C-order."""
    nk = [int(v) for v in g["num_knots"]]
    knots = [list(map(float, k)) for k in split(g["knots_cat"], nk)]
    shape = [int(v) for v in g["shape"]]
    pn = g["piece_n_nodes"]
    tens = split(g["piece_tensors_cat"], [int(np.prod(p)) for p in pn])
    nod = split(g["piece_nodes_cat"], [int(np.sum(p)) for p in pn])
    wts = split(g["piece_weights_cat"], [int(np.sum(p)) for p in pn])
    pieces = []
    for i, p in enumerate(pn):
        p = [int(v) for v in p]
        nodes = split(nod[i], p)
        weights = split(wts[i], p)
        dms = [make_dm(a, b) for a, b in zip(nodes, weights)]
        pieces.append((tens[i].reshape(p), nodes, weights, dms))
    return knots, shape, pieces


def slider_parts(g, make_dm):
    part = [list(map(int, grp)) for grp in g["partition"]]
    n = [int(v) for v in g["n_nodes"]]
    shapes = [[n[d] for d in grp] for grp in part]
    tens = split(g["slide_tensors_cat"], [int(np.prod(s)) for s in shapes])
    nod = split(g["slide_nodes_cat"], [int(np.sum(s)) for s in shapes])
    wts = split(g["slide_weights_cat"], [int(np.sum(s)) for s in shapes])
    slides = []
    for i, s in enumerate(shapes):
        nodes = split(nod[i], s)
        weights = split(wts[i], s)
        dms = [make_dm(a, b) for a, b in zip(nodes, weights)]
        slides.append((tens[i].reshape(s), nodes, weights, dms))
    return part, float(g["pivot_value"]), slides


def assert_close(got, ref, rel=REL, abs_=ABS, what=""):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    err = np.abs(got - ref)
    tol = rel * np.abs(ref) + abs_
    bad = ~(err <= tol)
    if bad.any():
        i = int(np.argmax(np.where(bad, err / np.maximum(tol, 1e-300), 0.0)))
        raise AssertionError(
            f"{what}: {int(bad.sum())}/{bad.size} outside {rel:g}*|ref|+{abs_:g}; worst at flat "
            f"index {i}: got {got.flat[i]!r} ref {ref.flat[i]!r} err {err.flat[i]:.3e}"
        )


def extrapolation_factor(domain, nodes, weights, points):
    """Per-point tolerance multiplier: 1 inside the domain, the Lebesgue function
    prod_d sum_i |l_i(x_d)| outside it.

    Barycentric *extrapolation* is ill-conditioned: any two summation orders (the reference's own
    evaluation paths included) differ by ~eps * Lebesgue(x) * max|f| there, so the flat 1e-12
    bound is only meaningful inside the domain, where the interpolant is meant to be used.
    """
    pts = np.asarray(points, dtype=np.float64)
    fac = np.ones(pts.shape[0])
    for d in range(pts.shape[1]):
        lo, hi = domain[d]
        out = (pts[:, d] < lo) | (pts[:, d] > hi)
        if not out.any():
            continue
        diff = pts[out, d][:, None] - nodes[d][None, :]
        w = weights[d][None, :] / diff
        lam = np.sum(np.abs(w), axis=1) / np.abs(np.sum(w, axis=1))
        fac[out] *= np.maximum(lam, 1.0)
    return fac


def assert_close_scaled(got, ref, factor, what, rel=REL, floor=ABS):
    """|got - ref| <= factor * (rel * |ref| + floor * max(1, max|ref|))."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    finite = np.isfinite(ref)
    scale = max(1.0, float(np.max(np.abs(ref[finite])))) if finite.any() else 1.0
    tol = factor * (rel * np.abs(ref) + floor * scale)
    err = np.abs(got - ref)
    bad = ~(err <= tol) & finite
    if bad.any():
        i = int(np.argmax(np.where(bad, err / np.maximum(tol, 1e-300), 0.0)))
        raise AssertionError(
            f"{what}: {int(bad.sum())}/{bad.size} outside tolerance; worst at {i}: got "
            f"{got.flat[i]!r} ref {ref.flat[i]!r} err {err.flat[i]:.3e} tol {tol.flat[i]:.3e}")


def fd_tolerance(g, domain, dim_order):
    """Propagated finite-difference tolerance per output row:
    ``c * (1e-12 * max|f| + 1e-14) / h^p`` with (c, p) = (1, 1) per first-order dim and
    (4, 2) per second-order dim (nested stencils multiply)."""
    orders = g["fd_orders"]
    scale = float(np.max(np.abs(g["fd_single_values"])))
    base = 1e-12 * scale + 1e-14
    tol = np.empty(orders.shape[0])
    for r, o in enumerate(orders):
        t = base
        for user_dim, k in enumerate(o):
            if k == 0:
                continue
            s = dim_order.index(user_dim)
            h = (domain[s][1] - domain[s][0]) * 1e-4
            t = t * (1.0 / h if k == 1 else 4.0 / (h * h))
        tol[r] = t
    return tol


def spline_factor(g, knots, pieces):
    """Extrapolation factor w.r.t. the piece each point is routed to."""
    pts = g["points"]
    fac = np.ones(len(pts))
    for p in np.unique(g["piece"]):
        t, nodes, w, dm = pieces[p]
        # the piece's sub-domain: nodes are strictly inside it, so recover it from the fixture
        dom = piece_domain(g, knots, p)
        m = g["piece"] == p
        fac[m] = extrapolation_factor(dom, nodes, w, pts[m])
    return fac


def piece_domain(g, knots, flat):
    shape = [int(v) for v in g["shape"]]
    idx = np.unravel_index(int(flat), shape)
    dom = []
    for d, (lo, hi) in enumerate(g["domain"]):
        edges = [float(lo)] + list(knots[d]) + [float(hi)]
        dom.append((edges[idx[d]], edges[idx[d] + 1]))
    return dom
