"""Host-side mirror of the reference API: constructors, validation messages, derivative-id
registry, special-points dispatch, build == from_values, TT construction.  No GPU needed."""

import math

import numpy as np
import pytest

import _golden as G
import pychebyshev_b200 as pcb
from pychebyshev_b200 import _grid, workloads as wl
from pychebyshev_b200.tt import tt_svd


def test_build_equals_from_values_and_reference_tensor():
    g = G.load("full_3d")
    n = [int(v) for v in g["n_nodes"]]
    dom = [list(map(float, r)) for r in g["domain"]]

    def f(x, _):
        return math.sin(x[0]) + math.sin(x[1]) + math.sin(x[2]) + x[0] * x[1] * x[2]

    cheb = pcb.ChebyshevApproximation(f, 3, dom, n)
    cheb.build(verbose=False)
    assert cheb.n_evaluations == int(np.prod(n)) and cheb.is_construction_finished()
    np.testing.assert_allclose(cheb.tensor_values, g["tensor"], rtol=0, atol=1e-15)
    same = pcb.ChebyshevApproximation.from_values(cheb.tensor_values, 3, dom, n)
    for d in range(3):
        assert np.array_equal(same.nodes[d], cheb.nodes[d])
        assert np.array_equal(same.weights[d], cheb.weights[d])
        assert np.array_equal(same.diff_matrices[d], cheb.diff_matrices[d])
    info = pcb.ChebyshevApproximation.nodes(3, dom, n)
    assert info["shape"] == tuple(n) and info["full_grid"].shape == (int(np.prod(n)), 3)
    with pytest.raises(RuntimeError, match="no function assigned"):
        same.build()


def test_derivative_tensor_is_the_reference_recipe():
    g = G.load("full_4d")
    n, nodes, weights, dms = G.full_parts(g)
    cheb = pcb.ChebyshevApproximation.from_values(g["tensor"], 4, g["domain"].tolist(), n)
    for o in g["orders"]:
        t = cheb.derivative_tensor(o)
        assert t.flags.c_contiguous
        assert np.array_equal(t, _grid.differentiate_tensor(g["tensor"], dms, o))
    # rebinding tensor_values invalidates cached derivative tensors / plans
    first = cheb.derivative_tensor([1, 0, 0, 0])
    cheb.tensor_values = cheb.tensor_values * 2.0
    assert np.array_equal(cheb.derivative_tensor([1, 0, 0, 0]), 2.0 * first)


def test_validation_messages():
    with pytest.raises(ValueError, match="max_n must be at least 3"):
        pcb.ChebyshevApproximation(None, 1, [[0, 1]], [3], max_n=2)
    with pytest.raises(ValueError, match="Got neither"):
        pcb.ChebyshevApproximation(lambda x, _: 0.0, 1, [[0, 1]])
    with pytest.raises(ValueError, match="does not match"):
        pcb.ChebyshevApproximation.from_values(np.zeros((2, 2)), 2, [[0, 1], [0, 1]], [2, 3])
    with pytest.raises(ValueError, match="NaN or Inf"):
        pcb.ChebyshevApproximation.from_values(np.array([1.0, np.nan]), 1, [[0, 1]], [2])
    with pytest.raises(ValueError, match="strictly less"):
        pcb.ChebyshevApproximation.from_values(np.zeros(2), 1, [[1, 1]], [2])
    cheb = pcb.ChebyshevApproximation(lambda x, _: float("nan"), 1, [[0, 1]], [3])
    with pytest.raises(ValueError, match="non-finite"):
        cheb.build(verbose=False)
    unbuilt = pcb.ChebyshevApproximation(lambda x, _: 0.0, 1, [[0, 1]], [3])
    with pytest.raises(RuntimeError, match=r"Call build\(\) first"):
        unbuilt.vectorized_eval_batch(np.zeros((1, 1)), [0])


def test_derivative_id_registry():
    cheb = pcb.ChebyshevApproximation.from_values(np.zeros((3, 3)), 2, [[0, 1], [0, 1]], [3, 3])
    a = cheb.get_derivative_id([1, 0])
    b = cheb.get_derivative_id([0, 2])
    assert (a, b) == (0, 1) and cheb.get_derivative_id([1, 0]) == 0
    assert cheb._resolve_derivative_args(None, 1) == [0, 2]
    with pytest.raises(ValueError, match="out of range"):
        cheb.get_derivative_id([3, 0])
    with pytest.raises(ValueError, match="length"):
        cheb.get_derivative_id([1])
    with pytest.raises(ValueError, match="must be int"):
        cheb.get_derivative_id([1.0, 0])
    with pytest.raises(ValueError, match="not both"):
        cheb._resolve_derivative_args([0, 0], 0)
    with pytest.raises(ValueError, match="must provide"):
        cheb._resolve_derivative_args(None, None)
    with pytest.raises(KeyError, match="unknown derivative_id"):
        cheb._resolve_derivative_args(None, 7)


def test_special_points_dispatch_returns_spline():
    f = lambda x, _: max(x[0] - 100.0, 0.0) * math.exp(-0.05 * x[1])  # noqa: E731
    sp = pcb.ChebyshevApproximation(f, 2, wl.SPLINE2D_DOMAIN, n_nodes=[[9, 13], [8]],
                                    special_points=[[100.0], []])
    assert isinstance(sp, pcb.ChebyshevSpline) and not isinstance(sp, pcb.ChebyshevApproximation)
    assert sp._shape == (2, 1) and sp.knots == [[100.0], []]
    sp.build(verbose=False)
    g = G.load("spline_nested2d")
    assert [p.n_nodes for p in sp._pieces] == g["piece_n_nodes"].tolist()
    np.testing.assert_allclose(np.concatenate([p.tensor_values.ravel() for p in sp._pieces]),
                               g["piece_tensors_cat"], rtol=0, atol=1e-13)
    # no special point anywhere -> a plain approximation
    plain = pcb.ChebyshevApproximation(f, 2, wl.SPLINE2D_DOMAIN, [5, 5], special_points=[[], []])
    assert isinstance(plain, pcb.ChebyshevApproximation)
    with pytest.raises(ValueError, match="must have 2 entries"):
        pcb.ChebyshevApproximation(f, 2, wl.SPLINE2D_DOMAIN, [5, 5], special_points=[[100.0]])
    with pytest.raises(ValueError, match="must be nested"):
        pcb.ChebyshevApproximation(f, 2, wl.SPLINE2D_DOMAIN, [5, 5], special_points=[[100.0], []])
    with pytest.raises(ValueError, match="not strictly inside"):
        pcb.ChebyshevApproximation(f, 2, wl.SPLINE2D_DOMAIN, [[5, 5], [5]],
                                   special_points=[[130.0], []])
    with pytest.raises(ValueError, match="one per sub-interval"):
        pcb.ChebyshevApproximation(f, 2, wl.SPLINE2D_DOMAIN, [[5], [5]],
                                   special_points=[[100.0], []])


def test_spline_construction_and_knot_rule():
    with pytest.raises(ValueError, match="not strictly inside"):
        pcb.ChebyshevSpline(None, 1, [[-1, 1]], [5], [[1.0]])
    with pytest.raises(ValueError, match="must be sorted"):
        pcb.ChebyshevSpline(None, 1, [[-1, 1]], [5], [[0.5, 0.2]])
    info = pcb.ChebyshevSpline.nodes(2, wl.SPLINE2D_DOMAIN, [4, 3], wl.SPLINE2D_KNOTS)
    assert info["num_pieces"] == 2 and info["piece_shape"] == (2, 1)
    vals = [wl.grid_values(wl.payoff2d, p["nodes_per_dim"]) for p in info["pieces"]]
    sp = pcb.ChebyshevSpline.from_values(vals, 2, wl.SPLINE2D_DOMAIN, [4, 3], wl.SPLINE2D_KNOTS)
    assert sp._built and sp._pieces[1].domain == [[100.0, 120.0], [0.25, 1.0]]
    with pytest.raises(ValueError, match="Expected 2 piece_values"):
        pcb.ChebyshevSpline.from_values(vals[:1], 2, wl.SPLINE2D_DOMAIN, [4, 3], wl.SPLINE2D_KNOTS)
    # derivative at a knot is rejected before any device work (spline.py:519-550)
    with pytest.raises(ValueError, match="not defined at knot"):
        sp.eval([100.0, 0.5], [1, 0])
    unbuilt = pcb.ChebyshevSpline(lambda x, _: 0.0, 1, [[-1, 1]], [5], [[0.0]])
    with pytest.raises(RuntimeError, match=r"Call build\(\) before eval_batch"):
        unbuilt.eval_batch(np.zeros((1, 1)), [0])


def test_tt_construction_paths():
    # TT-SVD of a separable-ish function reproduces the dense tensor
    dom = [[-1.0, 1.0], [0.0, 2.0], [0.5, 1.5]]
    n = [7, 9, 6]
    nodes = [_grid.cheb_nodes(lo, hi, k) for (lo, hi), k in zip(dom, n)]
    dense = wl.grid_values(lambda x, y, z: np.sin(x + 0.5 * y) * np.exp(-0.3 * z) + x * y, nodes)
    cores = tt_svd(dense, max_rank=8, tol=1e-12)
    back = cores[0]
    for c in cores[1:]:
        back = np.tensordot(back, c, axes=([back.ndim - 1], [0]))
    np.testing.assert_allclose(back.reshape(n), dense, atol=1e-12)
    tt = pcb.ChebyshevTT.from_values(dense, 3, dom, n, max_rank=8, tolerance=1e-12)
    assert tt.tt_ranks[0] == 1 and tt.tt_ranks[-1] == 1 and tt.method == "svd"
    assert tt.compression_ratio > 1.0
    # coefficient cores evaluate the interpolant (checked against the oracle on the host)
    from oracle import np_oracle as O

    pts = wl.uniform_queries(dom, 50, seed=3)
    exact = np.sin(pts[:, 0] + 0.5 * pts[:, 1]) * np.exp(-0.3 * pts[:, 2]) + pts[:, 0] * pts[:, 1]
    approx = O.tt_eval_batch(tt._coeff_cores, tt.domain, tt.dim_order, pts)
    np.testing.assert_allclose(approx, exact, atol=5e-4)
    # values_to_coeffs is the reference transform: compare on the stored reference cores
    g = G.load("tt_4d")
    rcores, rdom, _ = G.tt_parts(g)
    f4 = lambda x, _: (math.sin(x[0] + 0.5 * x[1]) * math.exp(-0.3 * x[2])  # noqa: E731
                       + math.cos(x[3] - x[0]) + x[1] * x[3])
    mine = pcb.ChebyshevTT(f4, 4, rdom, [7, 9, 6, 8], max_rank=8, tolerance=1e-10)
    mine.build(verbose=False, method="svd")
    assert mine.tt_ranks == [int(v) for v in g["ranks"]]
    got = O.tt_eval_batch(mine._coeff_cores, mine.domain, mine.dim_order, g["points"][:200])
    np.testing.assert_allclose(got, g["values"][:200], atol=1e-9)
    with pytest.raises(ValueError, match="boundary TT ranks"):
        pcb.ChebyshevTT.from_cores([np.zeros((2, 3, 1))], [[0, 1]])
    with pytest.raises(RuntimeError, match=r"Call build\(\)"):
        pcb.ChebyshevTT(None, 2, [[0, 1], [0, 1]], [3, 3]).eval_batch(np.zeros((1, 2)))
    with pytest.raises(ValueError, match="not supported"):
        tt.eval_multi_batch(pts, [[3, 0, 0]])


def test_slider_construction():
    with pytest.raises(ValueError, match="Partition must cover"):
        pcb.ChebyshevSlider(None, 3, [[0, 1]] * 3, [4] * 3, [[0], [1]], [0.5] * 3)
    f = lambda x, _: math.sin(x[0]) + x[1] * x[2]  # noqa: E731
    sl = pcb.ChebyshevSlider(f, 3, [[0, 1]] * 3, [5, 4, 4], [[0], [1, 2]], [0.5, 0.5, 0.5])
    sl.build(verbose=False)
    assert len(sl.slides) == 2 and sl.slides[1].n_nodes == [4, 4]
    assert sl.pivot_value == pytest.approx(math.sin(0.5) + 0.25)
    assert sl.get_derivative_id([1, 0, 0]) == 0


def test_tt_cross_build_accuracy():
    """method='cross' (own cross approximation) reaches the accuracy of the reference's TT on
    the 5D Black-Scholes function with a comparable number of evaluations."""
    from oracle import np_oracle as O

    tt = pcb.ChebyshevTT(wl.bs5d_scalar, 5, wl.BS5D_DOMAIN, wl.BS5D_NODES, max_rank=15, max_sweeps=5)
    tt.build(verbose=False, seed=42)
    assert tt.method == "cross" and max(tt.tt_ranks) <= 15
    assert tt.total_build_evals < 0.2 * 11 ** 5
    pts = wl.uniform_queries(wl.BS5D_DOMAIN, 500, 3)
    approx = O.tt_eval_batch(tt._coeff_cores, tt.domain, tt.dim_order, pts)
    exact = wl.bs_call_price(*pts.T)
    assert np.max(np.abs(approx - exact)) < 5e-3  # the reference's own TT: 5.0e-3 on this set
    # deterministic for a fixed seed
    again = pcb.ChebyshevTT(wl.bs5d_scalar, 5, wl.BS5D_DOMAIN, wl.BS5D_NODES, max_rank=15, max_sweeps=5)
    again.build(verbose=False, seed=42)
    assert all(np.array_equal(a, b) for a, b in zip(tt._coeff_cores, again._coeff_cores))


def test_single_point_coordinates_are_read_like_the_reference_reads_them():
    """point[d] for d < D only; a length-1 sequence counts as its element (the reference's own
    tests pass such points: tests/test_from_values.py:246-254)."""
    assert np.array_equal(_grid.point_row([[0.1], [0.5], [0.9]], 1), [[0.1]])
    assert np.array_equal(_grid.point_row((1.0, 2.0, 3.0, 4.0), 2), [[1.0, 2.0]])
    assert np.array_equal(_grid.point_row(np.array([1, 2, 3]), 3), [[1.0, 2.0, 3.0]])
    with pytest.raises(ValueError, match="scalar"):
        _grid.point_row([[0.1, 0.2]], 1)
    with pytest.raises(IndexError):
        _grid.point_row([0.1], 2)
