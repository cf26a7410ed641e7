"""derivative-tensor preparation, slice and extrude on the device.

Acceptance bar: the device derivative tensors agree with the host recipe
(``_grid.differentiate_tensor`` = the reference's ``_apply_derivative_passes``,
``barycentric.py:951-990``) to <= 1e-15 of the tensor's scale on C1 (11^5) and C4 (16^6).  The
kernels are written to be BIT-IDENTICAL (one sequential FMA chain per element, the order OpenBLAS
uses); the tests assert the 1e-15 bar and report bit identity, which is also asserted wherever the
host BLAS is FMA-based (every x86-64 CPU this image runs on).
"""

import numpy as np
import pytest

import _golden as G

pytestmark = pytest.mark.gpu


def _scale_close(got, ref, bar=1e-15):
    scale = float(np.max(np.abs(ref)))
    err = float(np.max(np.abs(got - ref)))
    assert err <= bar * scale, f"max |d| {err:.3e} > {bar:g} x scale {scale:.3e}"
    return bool(np.array_equal(got, ref))


@pytest.mark.parametrize("shape,orders", [
    ((11,) * 5, [[1, 0, 0, 0, 0], [2, 0, 0, 0, 0], [0, 0, 0, 1, 0], [0, 0, 0, 0, 1], [1, 0, 2, 0, 1]]),
    ((7, 9, 6), [[1, 0, 0], [0, 2, 0], [0, 0, 1], [1, 1, 1]]),
    ((15, 15), [[1, 0], [0, 1], [2, 2]]),
    ((5, 8, 3, 4), [[0, 1, 0, 1], [2, 0, 1, 0]]),
])
def test_device_derivative_tensors_bit_identical_to_host_recipe(shape, orders):
    from pychebyshev_b200 import _grid, device_tensor as DT

    D = len(shape)
    rng = np.random.default_rng(D * 100 + shape[0])
    nodes, weights, dms = _grid.grid_arrays([[0.5 * d, 1.0 + 1.5 * d] for d in range(D)], shape)
    T = rng.standard_normal(shape) * 50.0
    for o in orders:
        ref = _grid.differentiate_tensor(T, dms, o)
        got = DT.differentiate(T, dms, o).cpu().numpy()
        assert got.shape == ref.shape
        identical = _scale_close(got, ref)
        assert identical, f"order {o}: within 1e-15 but not bit-identical"


def test_one_dimensional_tensors_differ_from_numpy_only_in_the_summation_order():
    """A 1-D tensor goes through dgemv in NumPy (`vector @ D^T`), whose lane-split partial sums are
    not the sequential chain: agreement to rounding (1e-15 per pass, amplified by the second pass's
    O(n^2) matrix entries), not bit identity -- which is why ChebyshevApproximation keeps 1-D
    derivative tensors on the host recipe."""
    from pychebyshev_b200 import _grid, device_tensor as DT

    nodes, weights, dms = _grid.grid_arrays([[0.0, 1.0]], (33,))
    T = np.random.default_rng(133).standard_normal(33) * 50.0
    _scale_close(DT.differentiate(T, dms, [1]).cpu().numpy(), _grid.differentiate_tensor(T, dms, [1]))
    _scale_close(DT.differentiate(T, dms, [2]).cpu().numpy(), _grid.differentiate_tensor(T, dms, [2]),
                 bar=1e-13)


def test_c1_bs5d_plan_from_values_matches_host_tensor_plan_bitwise():
    """C1: the device-prepared plan and the plan fed with host-made derivative tensors give
    bit-identical prices and Greeks (same tensors -> same kernel -> same bits)."""
    import pychebyshev_b200 as pcb
    from pychebyshev_b200 import _engine, _grid, workloads as wl

    nodes = [_grid.cheb_nodes(lo, hi, n) for (lo, hi), n in zip(wl.BS5D_DOMAIN, wl.BS5D_NODES)]
    cheb = pcb.ChebyshevApproximation.from_values(wl.grid_values(wl.bs_call_price, nodes), 5,
                                                  wl.BS5D_DOMAIN, wl.BS5D_NODES)
    pts = wl.uniform_queries(wl.BS5D_DOMAIN, 5000, 3)
    dev_plan = _engine.FullPlan.from_values(cheb.n_nodes, cheb.nodes, cheb.weights,
                                            cheb.diff_matrices, cheb.tensor_values, wl.BS5D_GREEKS)
    host_plan = _engine.FullPlan(cheb.n_nodes, cheb.nodes, cheb.weights,
                                 [cheb.derivative_tensor(o) for o in wl.BS5D_GREEKS])
    a, b = dev_plan.eval(pts), host_plan.eval(pts)
    assert np.array_equal(a, b)
    # and against the reference-made golden outputs
    g = G.load("full_bs5d")
    got = dev_plan.eval(g["points"])
    cols = [list(map(int, o)) for o in g["orders"]]
    for j, o in enumerate(wl.BS5D_GREEKS):
        if o in cols:
            fac = G.extrapolation_factor(wl.BS5D_DOMAIN, cheb.nodes, cheb.weights, g["points"])
            G.assert_close_scaled(got[:, j], g["values"][:, cols.index(o)], fac, f"C1 {o}")


def test_c4_16p6_device_derivatives_and_plan_creation_time(capsys):
    """C4 (134 MB per tensor): derivative tensors on the device vs the host recipe, and what plan
    creation costs either way."""
    import time

    import pychebyshev_b200 as pcb
    from pychebyshev_b200 import _engine, _grid, device_tensor as DT, workloads as wl

    nodes = [_grid.cheb_nodes(lo, hi, n) for (lo, hi), n in zip(wl.C4_DOMAIN, wl.C4_NODES)]
    cheb = pcb.ChebyshevApproximation.from_values(wl.grid_values(wl.bs6d, nodes), 6, wl.C4_DOMAIN,
                                                  wl.C4_NODES)
    t0 = time.perf_counter()
    host_tensors = [cheb.derivative_tensor(o) for o in wl.C4_GREEKS]
    t_host_deriv = time.perf_counter() - t0
    identical = []
    for o, ref in zip(wl.C4_GREEKS[1:], host_tensors[1:]):
        got = DT.differentiate(cheb.tensor_values, cheb.diff_matrices, o).cpu().numpy()
        identical.append(_scale_close(got, ref))
        del got
    assert all(identical), identical
    import torch

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    p_host = _engine.FullPlan(cheb.n_nodes, cheb.nodes, cheb.weights, host_tensors)
    torch.cuda.synchronize()
    t_host_plan = time.perf_counter() - t0
    t0 = time.perf_counter()
    p_dev = _engine.FullPlan.from_values(cheb.n_nodes, cheb.nodes, cheb.weights, cheb.diff_matrices,
                                         cheb.tensor_values, wl.C4_GREEKS)
    torch.cuda.synchronize()
    t_dev_plan = time.perf_counter() - t0
    pts = wl.uniform_queries(wl.C4_DOMAIN, 1000, 9)
    assert np.array_equal(p_dev.eval(pts), p_host.eval(pts))
    with capsys.disabled():
        print(f"\n[C4 plan creation] host derivative tensors {t_host_deriv:.2f} s + plan from host "
              f"tensors {t_host_plan:.2f} s  vs  plan from the value tensor (device passes) "
              f"{t_dev_plan:.2f} s")


def test_slice_and_extrude_match_the_numpy_recipe():
    from pychebyshev_b200 import _grid, device_tensor as DT

    shape = (7, 9, 6, 5)
    rng = np.random.default_rng(77)
    dom = [[-1.0, 1.0], [0.0, 2.0], [0.5, 1.5], [-2.0, 0.0]]
    nodes, weights, dms = _grid.grid_arrays(dom, shape)
    T = rng.standard_normal(shape)
    t = DT.to_device(T)
    for axis, value in ((0, 0.3), (1, float(nodes[1][4])), (2, 1.49), (3, -1.234)):
        diff = value - nodes[axis]
        k = int(np.argmin(np.abs(diff)))
        if abs(diff[k]) < 1e-14:
            ref = np.take(T, k, axis=axis)
        else:
            w = weights[axis] / diff
            ref = np.tensordot(T, w / np.sum(w), axes=([axis], [0]))
        got = DT.slice_axis(t, axis, nodes[axis], weights[axis], value).cpu().numpy()
        assert got.shape == ref.shape
        _scale_close(got, ref, bar=2e-15)
        if abs(diff[k]) < 1e-14:
            assert np.array_equal(got, ref)
    for axis in (0, 2, 4):
        ref = np.repeat(np.expand_dims(T, axis=axis), 5, axis=axis)
        assert np.array_equal(DT.extrude_axis(t, axis, 5).cpu().numpy(), ref)


def test_slice_extrude_methods_side_by_side_with_the_reference():
    """ChebyshevApproximation.slice / .extrude against the reference's own methods
    (barycentric.py:1977-2154) on the same box."""
    from oracle import reference as R

    try:
        R.load()
    except R.ReferenceUnavailable as exc:
        pytest.skip(str(exc))
    from oracle import ref_objects as RO
    from pychebyshev_b200 import dropin

    def f(x, y, z):
        return np.sin(x + 0.5 * y) * np.exp(-0.3 * z) + x * y

    dom = [[-1.0, 1.0], [0.0, 2.0], [0.5, 1.5]]
    ref = RO.full_from_func(f, dom, [9, 8, 7])
    ours = dropin.adopt(ref, cached=False)
    rs = ref.slice([(1, 0.7)])
    os_ = ours.slice([(1, 0.7)])
    assert os_.n_nodes == list(rs.n_nodes) and os_.domain == [list(b) for b in rs.domain]
    _scale_close(os_.tensor_values, rs.tensor_values, bar=2e-15)
    re = ref.extrude((1, (0.0, 3.0), 5))
    oe = ours.extrude((1, (0.0, 3.0), 5))
    assert np.array_equal(oe.tensor_values, re.tensor_values)
    assert np.array_equal(oe.nodes[1], re.nodes[1])
    pts = np.random.default_rng(5).uniform([-1, 0.0, 0, 0.5], [1, 3.0, 2, 1.5], size=(200, 4))
    G.assert_close_scaled(oe.vectorized_eval_batch(pts, [0, 0, 1, 0]),
                          re.vectorized_eval_batch(pts, [0, 0, 1, 0]), 1.0, "extruded d/dy")
    with pytest.raises(ValueError, match="outside"):
        ours.slice([(0, 5.0)])
    with pytest.raises(ValueError, match="Cannot slice all"):
        ours.slice([(0, 0.0), (1, 1.0), (2, 1.0)])
