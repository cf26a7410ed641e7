"""Pin the oracle (oracle/np_oracle.py, oracle/c/oracle.c) and the package's host-side grid
recipes against outputs of the unmodified reference stored in tests/golden/*.npz."""

import numpy as np
import pytest

import _golden as G
from oracle import c_oracle as C
from oracle import np_oracle as O
from pychebyshev_b200 import _grid

FULL = ["full_1d", "full_2d", "full_3d", "full_4d"]
TT = ["tt_bs5d", "tt_4d", "tt_4d_perm", "tt_basket10d", "tt_rank20_10d"]
SPLINES = ["spline_abs1d", "spline_bs2d", "spline_bs3d", "spline_multiknot3d", "spline_nested2d"]


@pytest.mark.parametrize("name", FULL + ["full_bs5d"])
def test_grid_recipes_bit_identical(name):
    g = G.load(name)
    n, nodes, weights, dms = G.full_parts(g)
    for d in range(len(n)):
        lo, hi = g["domain"][d]
        for mod in (O.make_nodes, _grid.cheb_nodes):
            assert np.array_equal(mod(lo, hi, n[d]), nodes[d])
        for mod in (O.barycentric_weights, _grid.bary_weights):
            assert np.array_equal(mod(nodes[d]), weights[d])
        for mod in (O.diff_matrix, _grid.diff_matrix):
            assert np.array_equal(mod(nodes[d], weights[d]), dms[d])


@pytest.mark.parametrize("name", FULL)
def test_numpy_oracle_full_bit_identical(name):
    g = G.load(name)
    n, nodes, weights, dms = G.full_parts(g)
    out = O.full_eval_multi_batch(g["tensor"], nodes, weights, dms, g["points"], g["orders"])
    assert np.array_equal(out, g["values"])
    # the package's derivative passes are the reference's, bit for bit
    for o in g["orders"]:
        assert np.array_equal(_grid.differentiate_tensor(g["tensor"], dms, o),
                              np.ascontiguousarray(O.apply_derivative_passes(g["tensor"], dms, o)))


def test_numpy_oracle_full_bs5d_subset():
    from pychebyshev_b200 import workloads as wl

    g = G.load("full_bs5d")
    n, nodes, weights, dms = G.full_parts(g)
    tensor = wl.grid_values(wl.bs_call_price, nodes)
    import hashlib

    assert hashlib.sha256(tensor.tobytes()).hexdigest() == str(g["tensor_sha256"])
    idx = np.r_[0:40, len(g["points"]) - 50:len(g["points"])]
    out = O.full_eval_multi_batch(tensor, nodes, weights, dms, g["points"][idx], g["orders"][:4])
    assert np.array_equal(out, g["values"][idx][:, :4])


@pytest.mark.parametrize("name", FULL + ["full_bs5d"])
def test_c_oracle_full(name):
    g = G.load(name)
    n, nodes, weights, dms = G.full_parts(g)
    if "tensor" in g:
        tensor = g["tensor"]
    else:
        from pychebyshev_b200 import workloads as wl

        tensor = wl.grid_values(wl.bs_call_price, nodes)
    tens = [_grid.differentiate_tensor(tensor, dms, o) for o in g["orders"]]
    out = C.Full(n, nodes, weights, tens).eval_batch(g["points"], threads=4)
    fac = G.extrapolation_factor(g["domain"], nodes, weights, g["points"])
    for r in range(out.shape[1]):
        G.assert_close_scaled(out[:, r], g["values"][:, r], fac, f"{name} {g['orders'][r]}")


@pytest.mark.parametrize("name", TT)
def test_oracles_tt(name):
    g = G.load(name)
    cores, domain, dim_order = G.tt_parts(g)
    assert np.array_equal(O.tt_eval_batch(cores, domain, dim_order, g["points"]), g["values"])
    ct = C.TT(cores, domain, dim_order)
    G.assert_close_scaled(ct.eval_batch(g["points"], threads=4), g["values"], 1.0, name)
    # finite differences: NumPy oracle bit-identical on a subset, C oracle within the
    # propagated tolerance on everything
    sub = slice(0, 60)
    fd = O.tt_eval_multi_batch(cores, domain, dim_order, g["fd_points"][sub], g["fd_orders"])
    assert np.array_equal(fd, g["fd_values"][sub])
    tol = G.fd_tolerance(g, domain, dim_order)
    cfd = ct.eval_multi_batch(g["fd_points"], g["fd_orders"], threads=4)
    err = np.abs(cfd - g["fd_values"])
    assert (err <= tol[None, :]).all(), float(np.max(err / tol[None, :]))


def test_tt_order3_is_rejected():
    g = G.load("tt_4d")
    cores, domain, dim_order = G.tt_parts(g)
    with pytest.raises(ValueError, match="not supported"):
        O.tt_eval_multi_point(cores, domain, dim_order, [0.0, 1.0, 1.0, -1.0], [[3, 0, 0, 0]])
    with pytest.raises(ValueError, match="not supported"):
        C.TT(cores, domain, dim_order).eval_multi_batch(g["fd_points"][:2], [[0, 3, 0, 0]])


@pytest.mark.parametrize("name", SPLINES)
def test_oracles_spline(name):
    g = G.load(name)
    knots, shape, pieces = G.spline_parts(g, O.diff_matrix)
    assert np.array_equal(O.spline_lookup(knots, shape, g["points"]), g["piece"])
    assert np.array_equal(C.spline_lookup(knots, g["points"]), g["piece"])
    if len(g["lookup_points"]):
        assert np.array_equal(O.spline_lookup(knots, shape, g["lookup_points"]), g["lookup_piece"])
        assert np.array_equal(C.spline_lookup(knots, g["lookup_points"]), g["lookup_piece"])
    for r, o in enumerate(g["orders"]):
        out = O.spline_eval_batch(knots, shape, pieces, g["points"], o)
        assert np.array_equal(out, g["values"][:, r])
    pcs = [([len(a) for a in nodes], nodes, w,
            [_grid.differentiate_tensor(t, dm, o) for o in g["orders"]])
           for (t, nodes, w, dm) in pieces]
    out, piece = C.spline_eval_batch(knots, pcs, g["points"], threads=2)
    assert np.array_equal(piece, g["piece"])
    fac = G.spline_factor(g, knots, pieces)
    for r in range(out.shape[1]):
        G.assert_close_scaled(out[:, r], g["values"][:, r], fac, f"{name} {g['orders'][r]}")


def test_numpy_oracle_slider_bit_identical():
    g = G.load("slider10d")
    part, pivot_value, slides = G.slider_parts(g, O.diff_matrix)
    for r, o in enumerate(g["orders"]):
        out = O.slider_eval_batch(part, pivot_value, slides, g["points"], list(o))
        active = {s for s, grp in enumerate(part) for d in grp if o[d] > 0}
        if len(active) == 1:
            # reference single-point path interleaves D^T passes: rounding-level differences
            G.assert_close_scaled(out, g["values"][:, r], 1.0, f"slider {o}", rel=2e-11)
        else:
            assert np.array_equal(out, g["values"][:, r])
