"""GPU parity: the CUDA path (through the Python classes -> ctypes -> C ABI) against the golden
fixtures written by the reference and against the oracle on fresh seeded inputs.

Tolerances (BASELINE.json north_star): piece indices bit-exact; values and full-tensor /
spline / slider derivatives ``|d| <= 1e-12 |ref| + 1e-14``.  Two documented refinements:

* values whose magnitude is far below the interpolant's scale (deep out-of-the-money prices that
  are the sum of O(10) terms cancelling to 1e-5) cannot agree to 1e-12 *relative* between any two
  summation orders -- the reference's own ``eval`` and ``eval_batch`` differ by 1.6e-13 there
  (SURVEY.md App. B.2).  The absolute floor is therefore taken relative to the interpolant's
  scale: ``1e-14 * max|ref|``.
* TT Greeks are finite differences: the value tolerance is propagated through the stencil,
  ``|d| <= c (1e-12 max|f| + 1e-14) / h^p``.
"""

import numpy as np
import pytest

import _golden as G
from oracle import np_oracle as O

pytestmark = pytest.mark.gpu


def scale_close(got, ref, what, rel=1e-12, floor=1e-14, factor=1.0):
    G.assert_close_scaled(got, ref, factor, what, rel=rel, floor=floor)


# ------------------------------------------------------------------------------------------
# tensor train
# ------------------------------------------------------------------------------------------

TT_CASES = ["tt_bs5d", "tt_4d", "tt_4d_perm", "tt_basket10d", "tt_rank20_10d"]


def _tt(name):
    import pychebyshev_b200 as pcb

    g = G.load(name)
    cores, domain, dim_order = G.tt_parts(g)
    return g, pcb.ChebyshevTT.from_cores(cores, domain, dim_order)


@pytest.mark.parametrize("name", TT_CASES)
def test_tt_eval_batch_matches_reference(name):
    g, tt = _tt(name)
    got = tt.eval_batch(g["points"])
    assert got.shape == g["values"].shape
    scale_close(got, g["values"], f"{name} eval_batch")


@pytest.mark.parametrize("name", TT_CASES)
def test_tt_device_resident_matches_host_path(name):
    import torch

    g, tt = _tt(name)
    d_pts = torch.from_numpy(g["points"]).cuda()
    got = tt.eval_batch(d_pts)
    assert got.is_cuda and got.dtype == torch.float64
    assert np.array_equal(got.cpu().numpy(), tt.eval_batch(g["points"]))


@pytest.mark.parametrize("name", TT_CASES)
@pytest.mark.parametrize("algo", [1, 2, 0])
def test_tt_fd_greeks_match_reference(name, algo):
    """algo 1: one chain per stencil point (any rows); algo 2: shared partial products (rows that
    differentiate at most one dim); algo 0: the library's own choice."""
    g, tt = _tt(name)
    cores, domain, dim_order = G.tt_parts(g)
    if algo == 2:
        keep = np.array([int((o > 0).sum()) <= 1 for o in g["fd_orders"]])
        g = dict(g, fd_orders=g["fd_orders"][keep], fd_values=g["fd_values"][:, keep])
        with pytest.raises(NotImplementedError):  # algo 2 shares products: <= 1 differentiated dim per row
            tt.eval_multi_batch(g["fd_points"][:4], [[1, 1] + [0] * (tt.num_dimensions - 2)], algo=2)
    got = tt.eval_multi_batch(g["fd_points"], g["fd_orders"], algo=algo)
    ref = g["fd_values"]
    assert got.shape == ref.shape
    tol = G.fd_tolerance(g, domain, dim_order)
    err = np.abs(got - ref)
    bad = err > tol[None, :]
    assert not bad.any(), (
        f"{name}: {int(bad.sum())} FD entries outside the propagated tolerance; worst ratio "
        f"{float(np.max(err / tol[None, :])):.3g}")
    # value rows (all-zero orders) are plain values: the flat tolerance applies
    for r, o in enumerate(g["fd_orders"]):
        if not o.any():
            scale_close(got[:, r], ref[:, r], f"{name} eval_multi value row")


def test_tt_order3_raises_like_reference():
    g, tt = _tt("tt_4d")
    with pytest.raises(ValueError, match="not supported"):
        tt.eval_multi([0.1, 0.5, 1.0, -1.0], [[3, 0, 0, 0]])


def test_tt_single_point_api():
    g, tt = _tt("tt_bs5d")
    p = g["fd_points"][3]
    v = tt.eval(list(p))
    scale_close(np.array([v]), g["fd_single_values"][3:4], "eval single", floor=1e-13)
    multi = tt.eval_multi(list(p), [list(o) for o in g["fd_orders"][:4]])
    assert len(multi) == 4 and abs(multi[0] - v) <= 1e-12 * abs(v) + 1e-13


def test_tt_oracle_fresh_seed_large_batch():
    """Fresh inputs (not in the fixture), ragged size, against the oracle."""
    g, tt = _tt("tt_bs5d")
    cores, domain, dim_order = G.tt_parts(g)
    rng = np.random.default_rng(2024)
    dom = np.array(domain)
    n = 100_003  # not a multiple of the tile size
    pts = rng.uniform(dom[:, 0], dom[:, 1], size=(n, 5))
    scale_close(tt.eval_batch(pts), O.tt_eval_batch(cores, domain, dim_order, pts), "fresh batch")


@pytest.mark.parametrize("rank,n_nodes,D", [(28, 10, 4), (33, 7, 5), (17, 16, 3), (12, 11, 10)])
def test_tt_large_rank_per_core_path(rank, n_nodes, D, monkeypatch):
    """Trains whose cores fit the constant bank only one at a time: one launch per core, chain state
    in global memory, output columns in 2-3 register chunks.  Checked against the oracle, against the
    shared-memory kernels (PCB_TT_GSTREAM=0), over several tiles and with a ragged tail."""
    import pychebyshev_b200 as pcb

    rng = np.random.default_rng(rank)
    ranks = [1] + [rank] * (D - 1) + [1]
    ranks[1] = min(rank, 9)  # uneven ranks: the coefficient pass runs in both orientations
    cores = [rng.standard_normal((ranks[k], n_nodes, ranks[k + 1])) / np.sqrt(ranks[k] * n_nodes)
             for k in range(D)]
    domain = [[-1.0 + 0.1 * k, 2.0 + 0.3 * k] for k in range(D)]
    dim_order = list(range(D))[::-1]
    # `domain` is in storage order; user column u is storage dim dim_order.index(u)
    udom = np.array([domain[dim_order.index(u)] for u in range(D)])
    tile = 27 * 148 * 1024  # queries per pass of the per-core executor on a 148-SM B200
    n = tile + 1000 + 37  # two tiles + a ragged tail
    pts = rng.uniform(udom[:, 0], udom[:, 1], size=(n, D))
    orders = [[0] * D, [1] + [0] * (D - 1), [0] * (D - 1) + [2], [0] * D]
    tt = pcb.ChebyshevTT.from_cores(cores, domain, dim_order)
    info = tt._plan().info()
    assert info["uniform_path_values"] == 0  # too large for the single-launch bank kernels
    vals = tt.eval_batch(pts)
    greeks = tt.eval_multi_batch(pts, orders)
    monkeypatch.setenv("PCB_TT_GSTREAM", "0")
    tt2 = pcb.ChebyshevTT.from_cores(cores, domain, dim_order)
    vals2 = tt2.eval_batch(pts)
    greeks2 = tt2.eval_multi_batch(pts, orders)
    scale = float(np.max(np.abs(vals2)))
    assert np.max(np.abs(vals - vals2)) <= 1e-12 * scale + 1e-14
    assert np.array_equal(greeks[:, 0], greeks[:, 3])
    for r, o in enumerate(orders):
        tol = 1e-12 * scale + 1e-14  # propagated through the stencil
        for u, k in enumerate(o):
            h = (udom[u, 1] - udom[u, 0]) * 1e-4
            tol *= 1.0 if k == 0 else (1.0 / h if k == 1 else 4.0 / (h * h))
        assert np.max(np.abs(greeks[:, r] - greeks2[:, r])) <= tol, (r, o)
    sub = np.r_[0:700, tile - 350:tile + 350, n - 700:n]
    ref = O.tt_eval_batch(cores, domain, dim_order, pts[sub])
    scale_close(vals[sub], ref, f"rank {rank} per-core values")


def test_tt_empty_and_single_row():
    g, tt = _tt("tt_4d")
    assert tt.eval_batch(np.zeros((0, 4))).shape == (0,)
    one = tt.eval_batch(g["points"][:1])
    scale_close(one, g["values"][:1], "N=1")


# ------------------------------------------------------------------------------------------
# full tensor
# ------------------------------------------------------------------------------------------

FULL_SMALL = ["full_1d", "full_2d", "full_3d", "full_4d"]


def _full(name, tensor=None):
    import pychebyshev_b200 as pcb

    g = G.load(name)
    if tensor is None:
        tensor = g["tensor"]
    n = [int(v) for v in g["n_nodes"]]
    dom = [list(map(float, r)) for r in g["domain"]]
    return g, pcb.ChebyshevApproximation.from_values(tensor, len(n), dom, n)


def _full_factor(g, cheb):
    """1 inside the domain, the Lebesgue function outside (see _golden.extrapolation_factor)."""
    return G.extrapolation_factor(cheb.domain, cheb.nodes, cheb.weights, g["points"])


@pytest.mark.parametrize("name", FULL_SMALL)
@pytest.mark.parametrize("algo", [1, 2])
def test_full_small_matches_reference(name, algo):
    g, cheb = _full(name)
    if algo == 2 and cheb.num_dimensions < 2:
        pytest.skip("tensor-core path needs D >= 2")
    got = cheb.eval_batch_multi(g["points"], g["orders"], algo=algo)
    assert got.shape == g["values"].shape
    fac = _full_factor(g, cheb)
    for r in range(got.shape[1]):
        scale_close(got[:, r], g["values"][:, r], f"{name} algo={algo} order={g['orders'][r]}",
                    factor=fac)


@pytest.mark.parametrize("name", FULL_SMALL)
def test_full_small_auto_path_uses_bank_evaluator(name):
    """algo 0 hands small tensors to the constant-bank evaluator (a one-piece spline plan): same
    values as the reference and as the global-memory FMA evaluator, also for several outputs."""
    g, cheb = _full(name)
    fac = _full_factor(g, cheb)
    orders = [list(map(int, o)) for o in g["orders"]]
    auto = cheb.eval_batch_multi(g["points"], orders, algo=0)
    fma = cheb.eval_batch_multi(g["points"], orders, algo=1)
    for r, o in enumerate(orders):
        scale_close(auto[:, r], g["values"][:, r], f"{name} auto order={o}", factor=fac)
        scale_close(auto[:, r], fma[:, r], f"{name} auto vs fma order={o}", factor=fac)


def test_full_per_order_api_matches_multi():
    g, cheb = _full("full_3d")
    multi = cheb.eval_batch_multi(g["points"], g["orders"])
    for r, o in enumerate(g["orders"]):
        single = cheb.vectorized_eval_batch(g["points"], [int(v) for v in o])
        assert single.shape == (len(g["points"]),)
        scale_close(single, g["values"][:, r], f"vectorized_eval_batch {o}")
        assert np.array_equal(single, multi[:, r]) or np.allclose(single, multi[:, r], rtol=1e-13)
    did = cheb.get_derivative_id([1, 0, 0])
    by_id = cheb.vectorized_eval_batch(g["points"], derivative_id=did)
    scale_close(by_id, g["values"][:, 1], "derivative_id")


@pytest.mark.parametrize("algo", [1, 2])
def test_full_bs5d_price_and_greeks(algo):
    """BASELINE config 1: 11^5 Black-Scholes, price + delta/gamma/vega (+rho, +cross)."""
    from pychebyshev_b200 import workloads as wl

    g = G.load("full_bs5d")
    nodes = G.split(g["nodes_cat"], [int(v) for v in g["n_nodes"]])
    tensor = wl.grid_values(wl.bs_call_price, nodes)
    g2, cheb = _full("full_bs5d", tensor)
    for d in range(5):
        assert np.array_equal(cheb.nodes[d], nodes[d])
    got = cheb.eval_batch_multi(g["points"], g["orders"], algo=algo)
    fac = _full_factor(g, cheb)
    for r in range(got.shape[1]):
        scale_close(got[:, r], g["values"][:, r], f"bs5d algo={algo} order={g['orders'][r]}",
                    factor=fac)


def test_full_c4_16p6_tensor_core_path():
    """BASELINE config 4: 16^6 (134 MB per tensor), DMMA path, parity on the stored points."""
    from pychebyshev_b200 import workloads as wl

    g = G.load("full_c4_16p6")
    nodes = G.split(g["nodes_cat"], [int(v) for v in g["n_nodes"]])
    tensor = wl.grid_values(wl.bs6d, nodes)
    g2, cheb = _full("full_c4_16p6", tensor)
    got = cheb.eval_batch_multi(g["points"], g["orders"], algo=2)
    fac = _full_factor(g, cheb)
    for r in range(got.shape[1]):
        scale_close(got[:, r], g["values"][:, r], f"c4 order={g['orders'][r]}", factor=fac)


def test_full_single_point_and_errors():
    g, cheb = _full("full_2d")
    p = [float(v) for v in g["points"][0]]
    v = cheb.vectorized_eval(p, [0, 0])
    scale_close(np.array([v]), g["values"][0:1, 0], "vectorized_eval")
    multi = cheb.vectorized_eval_multi(p, [[0, 0], [1, 0]])
    scale_close(np.array(multi), g["values"][0, :2], "vectorized_eval_multi")
    with pytest.raises(ValueError):
        cheb.vectorized_eval_batch(g["points"])
    with pytest.raises(ValueError):
        cheb.vectorized_eval_batch(g["points"], [0, 0], derivative_id=0)
    with pytest.raises(KeyError):
        cheb.vectorized_eval_batch(g["points"], derivative_id=99)


# ------------------------------------------------------------------------------------------
# spline
# ------------------------------------------------------------------------------------------

SPLINES = ["spline_abs1d", "spline_bs2d", "spline_bs3d", "spline_multiknot3d", "spline_nested2d"]


def _spline(name):
    import pychebyshev_b200 as pcb
    from pychebyshev_b200 import ChebyshevApproximation

    g = G.load(name)
    knots, shape, pieces = G.spline_parts(g, O.diff_matrix)
    dom = [list(map(float, r)) for r in g["domain"]]
    D = len(dom)
    if bool(g["nested"]):
        # nested n_nodes: assemble through the constructor + per-piece values
        per_dim = [[] for _ in range(D)]
        for idx, (t, nodes, w, dm) in zip(np.ndindex(*shape), pieces):
            for d in range(D):
                if len(per_dim[d]) <= idx[d]:
                    per_dim[d].append(len(nodes[d]))
        sp = pcb.ChebyshevSpline(None, D, dom, per_dim, knots, defer_build=True)
        for piece, (t, nodes, w, dm) in zip(sp._pieces, pieces):
            piece.set_original_function_values(t)
        sp._built = True
    else:
        n = [int(v) for v in g["piece_n_nodes"][0]]
        sp = pcb.ChebyshevSpline.from_values([p[0] for p in pieces], D, dom, n, knots)
    for piece, (t, nodes, w, dm) in zip(sp._pieces, pieces):
        assert isinstance(piece, ChebyshevApproximation)
        for d in range(D):
            assert np.array_equal(piece.nodes[d], nodes[d])
            assert np.array_equal(piece.weights[d], w[d])
    return g, sp


@pytest.mark.parametrize("name", SPLINES)
def test_spline_lookup_bit_exact(name):
    g, sp = _spline(name)
    got = sp.find_pieces(g["points"])
    assert got.dtype == np.int32
    assert np.array_equal(got, g["piece"])
    if len(g["lookup_points"]):
        assert np.array_equal(sp.find_pieces(g["lookup_points"]), g["lookup_piece"])


@pytest.mark.parametrize("name", SPLINES)
def test_spline_eval_batch_matches_reference(name):
    g, sp = _spline(name)
    got = sp.eval_batch_multi(g["points"], g["orders"])
    knots, shape, pieces = G.spline_parts(g, O.diff_matrix)
    fac = G.spline_factor(g, knots, pieces)
    for r, o in enumerate(g["orders"]):
        scale_close(got[:, r], g["values"][:, r], f"{name} order={o}", factor=fac)
        one = sp.eval_batch(g["points"], [int(v) for v in o])
        # the single-order call may take another kernel (constant-bank vs. global-memory
        # evaluator): it is held to the reference like the multi-order call, not to bit equality
        scale_close(one, g["values"][:, r], f"{name} order={o} (single)", factor=fac)


def test_spline_single_point_knot_rule():
    g, sp = _spline("spline_bs2d")
    # value at the knot is fine, derivative at the knot raises (spline.py:519-550)
    assert abs(sp.eval([100.0, 0.5], [0, 0])) < 1e-10
    with pytest.raises(ValueError, match="not defined at knot"):
        sp.eval([100.0, 0.5], [1, 0])
    with pytest.raises(ValueError, match="not defined at knot"):
        sp.eval_multi([100.0, 0.5], [[0, 0], [1, 0]])
    # batch: silently the right-hand piece
    out = sp.eval_batch(np.array([[100.0, 0.5]]), [1, 0])
    assert np.isfinite(out).all()


# ------------------------------------------------------------------------------------------
# slider
# ------------------------------------------------------------------------------------------

def test_slider_matches_reference():
    import pychebyshev_b200 as pcb

    g = G.load("slider10d")
    part, pivot_value, slides = G.slider_parts(g, O.diff_matrix)
    dom = [list(map(float, r)) for r in g["domain"]]
    n = [int(v) for v in g["n_nodes"]]
    sl = pcb.ChebyshevSlider.from_slides([s[0] for s in slides], 10, dom, n, part,
                                         list(g["pivot_point"]), pivot_value)
    got = sl.eval_batch_multi(g["points"], g["orders"])
    ref = g["values"]
    for r, o in enumerate(g["orders"]):
        active = {sl._dim_to_slide[d] for d, k in enumerate(o) if k > 0}
        if len(active) > 1:
            assert np.array_equal(got[:, r], np.zeros(len(ref)))  # exactly 0.0
        else:
            # derivative rows: the reference's single-point path interleaves the D^T passes with
            # the contraction; its own paths differ by up to ~5e-12 there
            rel = 1e-12 if not active else 2e-11
            scale_close(got[:, r], ref[:, r], f"slider order={o}", rel=rel)
    v = sl.eval([float(x) for x in g["points"][5]], [0] * 10)
    scale_close(np.array([v]), ref[5:6, 0], "slider.eval")


# ------------------------------------------------------------------------------------------
# constant-bank evaluators (uniform datapath) against the global-memory evaluators
# ------------------------------------------------------------------------------------------

def _both_paths(make, monkeypatch):
    """The same object twice: plans on the constant-bank path and (PCB_NO_BANK) on the generic one."""
    bank = make()
    monkeypatch.setenv("PCB_NO_BANK", "1")
    generic = make()
    generic._plans = {}
    return bank, generic


@pytest.mark.parametrize("name", ["spline_abs1d", "spline_bs2d", "spline_multiknot3d"])
def test_spline_bank_and_generic_evaluators_agree(name, monkeypatch):
    g, _ = _spline(name)
    bank, generic = _both_paths(lambda: _spline(name)[1], monkeypatch)
    dom = np.asarray(g["domain"], dtype=np.float64)
    D = len(dom)
    rng = np.random.default_rng(11)
    knots, shape, pieces = G.spline_parts(g, O.diff_matrix)
    for n in (1, 255, 257, 100_003):  # ragged: no multiple of the 256-query CTA tile
        pts = dom[:, 0] + (dom[:, 1] - dom[:, 0]) * rng.random((n, D))
        if n > 64:
            # node hits (one-hot rows), knots (right-hand piece) and the domain corners
            nodes0 = pieces[0][1]
            pts[3] = [nodes0[d][min(1, len(nodes0[d]) - 1)] for d in range(D)]
            pts[4] = dom[:, 0]
            pts[5] = dom[:, 1]
            for d in range(D):
                if len(knots[d]):
                    pts[6 + d, d] = knots[d][0]
        a = bank.eval_batch(pts, [0] * D)
        b = generic.eval_batch(pts, [0] * D)
        assert np.array_equal(bank.find_pieces(pts), generic.find_pieces(pts))
        tol = 1e-12 * max(1.0, float(np.max(np.abs(b)))) + 1e-14
        assert np.max(np.abs(a - b)) <= tol, (name, n, float(np.max(np.abs(a - b))))
        if n > 64:
            # exact node hit of piece 0: the tensor entry itself, on both paths
            want = pieces[0][0][tuple(min(1, len(nodes0[d]) - 1) for d in range(D))]
            assert a[3] == want and b[3] == want
    # NaN coordinates: NaN out, routed to the last piece (spline.py:677-690), no trap
    pts = dom[:, 0] + (dom[:, 1] - dom[:, 0]) * rng.random((40, D))
    pts[7, 0] = np.nan
    a, b = bank.eval_batch(pts, [0] * D), generic.eval_batch(pts, [0] * D)
    assert np.isnan(a[7]) and np.isnan(b[7])
    keep = np.arange(40) != 7
    assert np.allclose(a[keep], b[keep], rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("width", [1e-3, 1.0, 1e6, 1e20])
def test_spline_bank_path_any_domain_width(width, monkeypatch):
    """The product-form rows are prescaled by powers of two: no overflow for any domain width.
    (Much narrower domains hit the reference's ABSOLUTE 1e-14 node-snap threshold, much wider ones
    underflow the reference's own barycentric weights: neither is an evaluator property.)"""
    import pychebyshev_b200 as pcb
    from pychebyshev_b200 import _grid

    dom = [[-0.25 * width, 0.75 * width], [2.0 * width, 3.0 * width]]
    n = [15, 16]
    knots = [[0.1 * width], []]

    def f(x, y):
        return np.cos(3.0 * x / width) * (y / width) ** 2

    def make():
        sp = pcb.ChebyshevSpline(None, 2, dom, n, knots, defer_build=True)
        for piece in sp._pieces:
            xs, ys = np.meshgrid(piece.nodes[0], piece.nodes[1], indexing="ij")
            piece.set_original_function_values(f(xs, ys))
        sp._built = True
        return sp

    bank, generic = _both_paths(make, monkeypatch)
    rng = np.random.default_rng(5)
    lo = np.array([d[0] for d in dom])
    hi = np.array([d[1] for d in dom])
    pts = lo + (hi - lo) * rng.random((5000, 2))
    a, b = bank.eval_batch(pts, [0, 0]), generic.eval_batch(pts, [0, 0])
    exact = f(pts[:, 0], pts[:, 1])
    assert np.all(np.isfinite(a))
    assert np.max(np.abs(a - b)) <= 1e-12 * np.max(np.abs(b)) + 1e-14
    assert np.max(np.abs(a - exact)) <= 1e-9 * np.max(np.abs(exact))  # interpolation error, both paths
    del _grid


def test_slider_bank_and_generic_evaluators_agree(monkeypatch):
    import pychebyshev_b200 as pcb

    g = G.load("slider10d")
    part, pivot_value, slides = G.slider_parts(g, O.diff_matrix)
    dom = [list(map(float, r)) for r in g["domain"]]
    n = [int(v) for v in g["n_nodes"]]

    def make():
        return pcb.ChebyshevSlider.from_slides([s[0] for s in slides], 10, dom, n, part,
                                               list(g["pivot_point"]), pivot_value)

    bank, generic = _both_paths(make, monkeypatch)
    rng = np.random.default_rng(3)
    lo = np.array([d[0] for d in dom])
    hi = np.array([d[1] for d in dom])
    # 7 rows: more than the 4 the bank kernel keeps in registers (the rest go through global memory)
    orders = [[0] * 10]
    for d in (0, 3, 4, 9):
        o = [0] * 10
        o[d] = 1
        orders.append(o)
    o = [0] * 10
    o[0], o[9] = 1, 1  # cross-slide: exactly zero
    orders.append(o)
    o = [0] * 10
    o[2] = 2
    orders.append(o)
    for rows in (1, 300, 70_001):
        pts = lo + (hi - lo) * rng.random((rows, 10))
        a = bank.eval_batch_multi(pts, orders)
        b = generic.eval_batch_multi(pts, orders)
        assert a.shape == (rows, len(orders))
        assert np.array_equal(a[:, 5], np.zeros(rows)) and np.array_equal(b[:, 5], np.zeros(rows))
        for r in range(len(orders)):
            tol = 2e-11 * max(1.0, float(np.max(np.abs(b[:, r])))) + 1e-14
            assert np.max(np.abs(a[:, r] - b[:, r])) <= tol, (rows, r)


# ------------------------------------------------------------------------------------------
# size-independent properties at scale (BASELINE sizes do not fit a CPU oracle)
# ------------------------------------------------------------------------------------------

def test_tt_large_batch_properties():
    """1e7 queries on device: chunk invariance, permutation equivariance, idempotence."""
    import torch

    g, tt = _tt("tt_bs5d")
    cores, domain, dim_order = G.tt_parts(g)
    n = 10_000_000
    gen = torch.Generator(device="cuda").manual_seed(7)
    lo = torch.tensor([d[0] for d in domain], device="cuda", dtype=torch.float64)
    hi = torch.tensor([d[1] for d in domain], device="cuda", dtype=torch.float64)
    pts = lo + (hi - lo) * torch.rand((n, 5), generator=gen, device="cuda", dtype=torch.float64)
    full = tt.eval_batch(pts)
    again = tt.eval_batch(pts)
    assert torch.equal(full, again)  # idempotent / deterministic
    # a query's value does not depend on where it sits in the batch or on the batch size
    perm = torch.randperm(n, generator=gen, device="cuda")
    assert torch.equal(tt.eval_batch(pts[perm]), full[perm])
    assert torch.equal(tt.eval_batch(pts[123_457:1_234_567]), full[123_457:1_234_567])
    # spot-check a slice against the oracle
    idx = torch.arange(0, n, n // 2000, device="cuda")
    ref = O.tt_eval_batch(cores, domain, dim_order, pts[idx].cpu().numpy())
    scale_close(full[idx].cpu().numpy(), ref, "1e7 spot check")
    # FD rows: value row identical to eval_batch, rows finite
    fd = tt.eval_multi_batch(pts[:1_000_000], g["fd_orders"][:4], algo=1)
    assert torch.isfinite(fd).all()
    scale_close(fd[:, 0].cpu().numpy(), full[:1_000_000].cpu().numpy(), "fd value row", rel=1e-13)


def test_spline_large_batch_properties():
    """3e7 queries on device (C3 2-D spline): determinism, batch-position invariance, routing
    consistency, linearity of the interpolation operator, exactness at the nodes."""
    import torch

    import pychebyshev_b200 as pcb

    g, sp = _spline("spline_bs2d")
    dom = np.asarray(g["domain"], dtype=np.float64)
    n = 30_000_000
    gen = torch.Generator(device="cuda").manual_seed(17)
    lo = torch.tensor(dom[:, 0], device="cuda")
    hi = torch.tensor(dom[:, 1], device="cuda")
    pts = lo + (hi - lo) * torch.rand((n, 2), generator=gen, device="cuda", dtype=torch.float64)
    full = sp.eval_batch(pts, [0, 0])
    assert torch.equal(full, sp.eval_batch(pts, [0, 0]))
    perm = torch.randperm(n, generator=gen, device="cuda")
    assert torch.equal(sp.eval_batch(pts[perm], [0, 0]), full[perm])
    assert torch.equal(sp.eval_batch(pts[1_000_003:2_999_999], [0, 0]), full[1_000_003:2_999_999])
    # routing: the piece index is the number of knots <= S (one knot at 100), whatever the batch
    piece = sp.find_pieces(pts)
    assert torch.equal(piece.to(torch.int64), (pts[:, 0] >= 100.0).to(torch.int64))
    # linearity: spline(2 f - 3 g) = 2 spline(f) - 3 spline(g), g = another tensor on the same grid
    knots, shape, pieces = G.spline_parts(g, O.diff_matrix)
    n_nodes = [int(v) for v in g["piece_n_nodes"][0]]
    rng = np.random.default_rng(4)
    other = [rng.standard_normal(p[0].shape) for p in pieces]
    dom_l = [list(map(float, r)) for r in g["domain"]]
    sp_g = pcb.ChebyshevSpline.from_values(other, 2, dom_l, n_nodes, knots)
    sp_c = pcb.ChebyshevSpline.from_values([2.0 * p[0] - 3.0 * o for p, o in zip(pieces, other)], 2, dom_l,
                                           n_nodes, knots)
    sub = pts[:5_000_000]
    lhs = sp_c.eval_batch(sub, [0, 0])
    rhs = 2.0 * full[:5_000_000] - 3.0 * sp_g.eval_batch(sub, [0, 0])
    scale = float(torch.max(torch.abs(rhs)))
    assert float(torch.max(torch.abs(lhs - rhs))) <= 1e-12 * scale + 1e-14
    # exact at every node of every piece
    for (t, nodes, w, dm) in pieces:
        xs, ys = np.meshgrid(nodes[0], nodes[1], indexing="ij")
        grid = np.column_stack([xs.ravel(), ys.ravel()])
        inside = grid[:, 0] != 100.0  # the knot itself belongs to the right-hand piece
        got = sp.eval_batch(grid, [0, 0])
        assert np.array_equal(got[inside], t.ravel()[inside])
    # spot check against the oracle
    idx = torch.arange(0, n, n // 1500, device="cuda")
    host = pts[idx].cpu().numpy()
    ref = np.empty(len(host))
    which = (host[:, 0] >= 100.0).astype(int)
    for p, (t, nodes, w, dm) in enumerate(pieces):
        m = which == p
        ref[m] = O.full_eval_batch(t, nodes, w, dm, host[m], [0, 0])
    scale_close(full[idx].cpu().numpy(), ref, "3e7 spline spot check")


def test_full_tensor_core_path_properties():
    """C1 (11^5) on the DMMA path, 1e6 queries: determinism, batch-position invariance, linearity in
    the tensor, exactness at grid nodes."""
    import torch

    import pychebyshev_b200 as pcb
    from pychebyshev_b200 import workloads as wl

    g = G.load("full_bs5d")
    nodes = G.split(g["nodes_cat"], [int(v) for v in g["n_nodes"]])
    tensor = wl.grid_values(wl.bs_call_price, nodes)
    rng = np.random.default_rng(9)
    other = rng.standard_normal(tensor.shape)
    mk = lambda t: pcb.ChebyshevApproximation.from_values(t, 5, wl.BS5D_DOMAIN, wl.BS5D_NODES)  # noqa: E731
    a, b, c = mk(tensor), mk(other), mk(0.5 * tensor + 4.0 * other)
    n = 1_000_000
    gen = torch.Generator(device="cuda").manual_seed(3)
    lo = torch.tensor([d[0] for d in wl.BS5D_DOMAIN], device="cuda", dtype=torch.float64)
    hi = torch.tensor([d[1] for d in wl.BS5D_DOMAIN], device="cuda", dtype=torch.float64)
    pts = lo + (hi - lo) * torch.rand((n, 5), generator=gen, device="cuda", dtype=torch.float64)
    zero = [[0, 0, 0, 0, 0]]
    fa = a.eval_batch_multi(pts, zero, algo=2)
    assert torch.equal(fa, a.eval_batch_multi(pts, zero, algo=2))
    perm = torch.randperm(n, generator=gen, device="cuda")
    assert torch.equal(a.eval_batch_multi(pts[perm], zero, algo=2), fa[perm])
    fb = b.eval_batch_multi(pts, zero, algo=2)
    fc = c.eval_batch_multi(pts, zero, algo=2)
    rhs = 0.5 * fa + 4.0 * fb
    assert float(torch.max(torch.abs(fc - rhs))) <= 1e-12 * float(torch.max(torch.abs(rhs))) + 1e-14
    # grid nodes: the tensor entries themselves (one-hot weight rows), also on the tensor-core path
    idx = rng.integers(0, 11, size=(4096, 5))
    grid = np.column_stack([nodes[d][idx[:, d]] for d in range(5)])
    got = a.eval_batch_multi(grid, zero, algo=2)[:, 0]
    assert np.array_equal(got, tensor[tuple(idx.T)])
    # the FMA evaluator agrees
    f1 = a.eval_batch_multi(pts[:20_000], zero, algo=1)
    scale_close(fa[:20_000].cpu().numpy(), f1.cpu().numpy(), "DMMA vs FMA evaluator")


# ------------------------------------------------------------------------------------------
# native .pcb loader (no Python grid arithmetic on the path)
# ------------------------------------------------------------------------------------------

def test_native_pcb_loader_matches_reference(tmp_path):
    """pcb_plan_from_file on the reference's own fixture files (bytes stored in the goldens)
    against the values the reference computed after loading them."""
    import pychebyshev_b200 as pcb

    gold = G.load("pcb_files")
    for key, kind, ndim in (("approx_2d_simple", "approx", 2), ("spline_1d_kink", "spline", 1)):
        path = tmp_path / f"{key}.pcb"
        path.write_bytes(gold[key + "_bytes"].tobytes())
        plan = pcb.load_plan(path)
        assert plan.kind == kind and plan.ndim == ndim
        got = plan.eval(gold[key + "_points"])[:, 0]
        scale_close(got, gold[key + "_values"], f"native loader {key}")
    # a file written by this package (5D), read back natively vs through the Python class
    g = G.load("full_4d")
    n = [int(v) for v in g["n_nodes"]]
    cheb = pcb.ChebyshevApproximation.from_values(g["tensor"], 4, g["domain"].tolist(), n)
    path = tmp_path / "full4d.pcb"
    cheb.save(path)
    native = pcb.load_plan(path).eval(g["points"])[:, 0]
    scale_close(native, g["values"][:, 0], "native loader full_4d",
                factor=G.extrapolation_factor(cheb.domain, cheb.nodes, cheb.weights, g["points"]))


def test_c_host_evaluates_pcb_file_like_the_python_host(tmp_path):
    """examples/pcb_eval.c (no Python, no torch): .pcb file -> pcb_plan_from_file -> pcb_plan_eval."""
    import subprocess

    from test_pcb_format import _build_c_host

    exe = _build_c_host(tmp_path)
    g, sp = _spline("spline_bs2d")
    path = tmp_path / "spline.pcb"
    sp.save(path)
    rng = np.random.default_rng(8)
    dom = np.asarray(g["domain"], dtype=np.float64)
    pts = dom[:, 0] + (dom[:, 1] - dom[:, 0]) * rng.random((10_001, 2))
    (tmp_path / "pts.f64").write_bytes(np.ascontiguousarray(pts).tobytes())
    res = subprocess.run([exe, str(path), str(tmp_path / "pts.f64"), str(tmp_path / "out.f64")],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    assert "ChebyshevSpline, 2 dims, 10001 points" in res.stdout
    got = np.frombuffer((tmp_path / "out.f64").read_bytes(), dtype=np.float64)
    want = sp.eval_batch(pts, [0, 0])
    assert np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want)) + 1e-14


def test_constant_bank_residency_across_plans_streams_and_threads():
    """The constant bank holds one plan image at a time.  Plans that alternate on two streams,
    driven from two host threads, must still see their own data: every result equals the serial one."""
    import threading

    import torch

    g1, tt1 = _tt("tt_bs5d")
    g2, tt2 = _tt("tt_4d")
    g3, tt3 = _tt("tt_rank20_10d")  # per-core path: a new bank image at every launch
    gs, sp = _spline("spline_bs2d")
    cases = []
    for g, obj in ((g1, tt1), (g2, tt2), (g3, tt3)):
        cores, domain, dim_order = G.tt_parts(g)
        D = len(domain)
        udom = np.array([domain[dim_order.index(u)] for u in range(D)])
        rng = np.random.default_rng(D)
        pts = torch.from_numpy(rng.uniform(udom[:, 0], udom[:, 1], size=(150_001, D))).cuda()
        plan = obj._plan()
        cases.append((plan, pts, plan.eval_device(pts).clone()))
        fd = obj._plan().with_orders(np.asarray(g["fd_orders"][:3]), 2)
        cases.append((fd, pts, fd.eval_device(pts).clone()))
    dom = np.asarray(gs["domain"], dtype=np.float64)
    pts = torch.from_numpy(dom[:, 0] + (dom[:, 1] - dom[:, 0]) * np.random.default_rng(1).random((150_001, 2))).cuda()
    plan = sp._plan([[0, 0]])
    cases.append((plan, pts, plan.eval_device(pts).clone()))
    torch.cuda.synchronize()

    errors = []

    def worker(offset):
        try:
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                for it in range(12):
                    plan, pts, want = cases[(it + offset) % len(cases)]
                    got = plan.eval_device(pts)
                    if not torch.equal(got, want):
                        errors.append((offset, it))
            stream.synchronize()
        except Exception as exc:  # noqa: BLE001
            errors.append(repr(exc))

    threads = [threading.Thread(target=worker, args=(k * 3,)) for k in range(3)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


# ------------------------------------------------------------------------------------------
# 2-D grids on the FP64 tensor cores (spline2d_dmma_kernel / slider2d_dmma_kernel)
# ------------------------------------------------------------------------------------------

def test_dmma_2d_spline_matches_bank_evaluator_and_reference(monkeypatch):
    """The tensor-core path for 2-D pieces against the constant-bank evaluator (PCB_NO_DMMA2D=1) on a
    ragged batch with knot hits, node hits, mixed-piece row tiles and 5 outputs (two passes of 4),
    and against the reference-made golden values."""
    import pychebyshev_b200 as pcb

    for name in ("spline_bs2d", "spline_nested2d"):
        g, sp = _spline(name)
        knots, shape, pieces = G.spline_parts(g, O.diff_matrix)
        fac = G.spline_factor(g, knots, pieces)
        got = sp.eval_batch_multi(g["points"], g["orders"])
        for r, o in enumerate(g["orders"]):
            scale_close(got[:, r], g["values"][:, r], f"{name} dmma2d order={o}", factor=fac)
    g, sp = _spline("spline_bs2d")
    rng = np.random.default_rng(99)
    n = 128 * 37 + 51
    pts = np.column_stack([rng.uniform(80.0, 120.0, n), rng.uniform(0.25, 1.0, n)])
    pts[::7, 0] = 100.0                                    # on the knot: right piece
    pts[3::11, 0] = sp._pieces[0].nodes[0][4]              # node hit in dim 0
    pts[5::13, 1] = sp._pieces[1].nodes[1][9]              # node hit in dim 1
    pts[:64, 0] = np.where(np.arange(64) % 2 == 0, 90.0, 110.0)  # alternate pieces inside row tiles
    orders = [[0, 0], [1, 0], [0, 1], [2, 0], [1, 1]]
    a = sp.eval_batch_multi(pts, orders)
    pa = sp.find_pieces(pts)
    monkeypatch.setenv("PCB_NO_DMMA2D", "1")
    sp2 = pcb.ChebyshevSpline.from_values([p.tensor_values for p in sp._pieces], 2, sp.domain,
                                          sp.n_nodes, sp.knots)
    b = sp2.eval_batch_multi(pts, orders)
    assert np.array_equal(pa, sp2.find_pieces(pts))
    for r, o in enumerate(orders):
        scale_close(a[:, r], b[:, r], f"dmma2d vs bank order={o}")


def test_dmma_2d_slider_matches_bank_evaluator(monkeypatch):
    import pychebyshev_b200 as pcb

    g = G.load("slider10d")
    part, pivot_value, slides = G.slider_parts(g, O.diff_matrix)
    dom = [list(map(float, r)) for r in g["domain"]]
    nn = [int(v) for v in g["n_nodes"]]

    def make():
        return pcb.ChebyshevSlider.from_slides([s[0] for s in slides], 10, dom, nn, part,
                                               list(g["pivot_point"]), pivot_value)
    rng = np.random.default_rng(7)
    n = 128 * 9 + 77
    pts = rng.uniform(80.0, 120.0, size=(n, 10))
    pts[::9, 2] = slides[1][1][0][3]                       # node hit in slide 1, dim 0
    orders = [[0] * 10, [1] + [0] * 9, [0] * 9 + [2], [1, 0, 1] + [0] * 7, [0, 1] + [0] * 8,
              [0, 0, 0, 1] + [0] * 6]                      # 6 rows: > SLIDER_ACC
    a = make().eval_batch_multi(pts, orders)
    monkeypatch.setenv("PCB_NO_DMMA2D", "1")
    b = make().eval_batch_multi(pts, orders)
    assert (a[:, 3] == 0.0).all()
    for r, o in enumerate(orders):
        scale_close(a[:, r], b[:, r], f"slider dmma2d vs bank order={o}")


# ------------------------------------------------------------------------------------------
# full tensor: joint-K tensor-core variant (pcb_full_eval algo 3)
# ------------------------------------------------------------------------------------------

@pytest.mark.parametrize("name", ["full_3d", "full_4d", "full_bs5d"])
def test_full_joint_k_tensor_core_variant(name, monkeypatch):
    """The DMMA GEMM over the last two axes jointly (K = n_d * n_e, A fragments in registers) against
    the reference goldens and against the per-row tensor-core variant, ragged batch included."""
    from pychebyshev_b200 import workloads as wl

    monkeypatch.setenv("PCB_FORCE_DMMA2", "1")  # also where the planner would prefer the per-row variant
    if name == "full_bs5d":
        gg = G.load(name)
        nodes = G.split(gg["nodes_cat"], [int(v) for v in gg["n_nodes"]])
        g, cheb = _full(name, wl.grid_values(wl.bs_call_price, nodes))
    else:
        g, cheb = _full(name)
    got = cheb.eval_batch_multi(g["points"], g["orders"], algo=3)
    fac = _full_factor(g, cheb)
    for r in range(got.shape[1]):
        scale_close(got[:, r], g["values"][:, r], f"{name} joint-K order={g['orders'][r]}", factor=fac)
    rng = np.random.default_rng(17)
    dom = np.array(cheb.domain)
    pts = rng.uniform(dom[:, 0], dom[:, 1], size=(128 * 5 + 37, cheb.num_dimensions))
    pts[3] = [cheb.nodes[d][1] for d in range(cheb.num_dimensions)]  # node hits in every dim
    orders = [list(map(int, o)) for o in g["orders"][:3]]
    a = cheb.eval_batch_multi(pts, orders, algo=3)
    b = cheb.eval_batch_multi(pts, orders, algo=2)
    for r in range(len(orders)):
        scale_close(a[:, r], b[:, r], f"{name} joint-K vs per-row order={orders[r]}")


def test_dmma_3d_spline_matches_bank_evaluator_and_reference(monkeypatch):
    """3-D pieces on the tensor cores (joint-K over the last two axes) against the reference goldens
    and the constant-bank evaluator (PCB_NO_DMMA3D=1): ragged batch, knot / node hits, mixed-piece row
    tiles, several outputs (split over launches when the fragment images do not fit together)."""
    import pychebyshev_b200 as pcb

    for name in ("spline_bs3d", "spline_multiknot3d"):
        g, sp = _spline(name)
        knots, shape, pieces = G.spline_parts(g, O.diff_matrix)
        fac = G.spline_factor(g, knots, pieces)
        got = sp.eval_batch_multi(g["points"], g["orders"])
        for r, o in enumerate(g["orders"]):
            scale_close(got[:, r], g["values"][:, r], f"{name} dmma3d order={o}", factor=fac)
    g, sp = _spline("spline_bs3d")
    rng = np.random.default_rng(123)
    n = 256 * 11 + 77
    pts = np.column_stack([rng.uniform(80.0, 120.0, n), rng.uniform(0.25, 1.0, n), rng.uniform(0.01, 0.08, n)])
    pts[::7, 0] = 100.0
    pts[3::11, 0] = sp._pieces[0].nodes[0][4]
    pts[5::13, 2] = sp._pieces[1].nodes[2][9]
    pts[:64, 0] = np.where(np.arange(64) % 2 == 0, 90.0, 110.0)
    orders = [[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 2], [1, 0, 1]]
    a = sp.eval_batch_multi(pts, orders)
    pa = sp.find_pieces(pts)
    monkeypatch.setenv("PCB_NO_DMMA3D", "1")
    sp2 = pcb.ChebyshevSpline.from_values([p.tensor_values for p in sp._pieces], 3, sp.domain,
                                          sp.n_nodes, sp.knots)
    b = sp2.eval_batch_multi(pts, orders)
    assert np.array_equal(pa, sp2.find_pieces(pts))
    for r, o in enumerate(orders):
        scale_close(a[:, r], b[:, r], f"dmma3d vs bank order={o}")
