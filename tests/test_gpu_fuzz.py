"""Seeded random configurations, the UNMODIFIED reference and the CUDA engine side by side.

The fixed-config tests (test_gpu_vs_reference.py, test_gpu_parity.py) pin BASELINE.json's shapes;
this file walks the space around them -- dimension counts, ragged node counts, ranks, storage
permutations, knot layouts, slide partitions, derivative-order rows, queries on nodes / on knots /
on the domain boundary -- so that every kernel-selection branch of the engine (constant bank or
shared memory, tensor-core or scalar form, per-core chain, host split of stencil rows) meets the
reference's own classes on inputs nobody tuned it for.  Every object is made by the reference's own
constructors (oracle/ref_objects.py); the engine is reached through ``dropin.adopt``.

Tolerances are the north-star ones (tests/_golden.py): 1e-12-class on interpolated values and
analytic derivatives, the propagated finite-difference bound on TT Greeks, bit-exact on routing.
"""

import numpy as np
import pytest

import _golden as G

pytestmark = pytest.mark.gpu

BASE_SEED = 7_2026


def _ref():
    from oracle import reference as R

    try:
        return R.load()
    except R.ReferenceUnavailable as exc:
        pytest.skip(str(exc))


def _domain(rng, D):
    lo = rng.uniform(-3.0, 3.0, D)
    width = rng.choice([0.3, 1.0, 2.5, 40.0], D) * rng.uniform(0.7, 1.3, D)
    return [[float(a), float(a + w)] for a, w in zip(lo, width)]


def _smooth(rng, domain, terms=3):
    """A broadcasting sum of separable smooth terms, frequencies scaled to the domain widths."""
    D = len(domain)
    amp = rng.uniform(0.5, 2.0, terms)
    freq = rng.uniform(0.5, 2.5, (terms, D)) / np.array([hi - lo for lo, hi in domain])
    phase = rng.uniform(0, 2 * np.pi, (terms, D))

    def f(*xs):
        tot = 0.0
        for j in range(terms):
            term = amp[j]
            for d, x in enumerate(xs):
                term = term * np.cos(freq[j, d] * x + phase[j, d])
            tot = tot + term
        return tot

    return f


def _queries(rng, domain, n, nodes=None, knots=None):
    """Uniform draws plus the awkward ones: on grid nodes, a few ulps off them, on the bounds,
    on the knots and one ulp either side."""
    D = len(domain)
    pts = np.stack([rng.uniform(lo, hi, n) for lo, hi in domain], axis=1)
    r = 0
    if nodes is not None:
        for _ in range(6):
            for d in range(D):
                pts[r, d] = nodes[d][rng.integers(len(nodes[d]))]
            r += 1
        for _ in range(4):
            d = int(rng.integers(D))
            x = nodes[d][rng.integers(len(nodes[d]))]
            pts[r, d] = x + rng.choice([-1, 1]) * 4 * np.spacing(abs(x) + 1.0)
            r += 1
    for d in range(D):
        pts[r, d] = domain[d][0]
        pts[r + 1, d] = domain[d][1]
    r += 2
    if knots is not None:
        for d in range(D):
            for k in knots[d]:
                for x in (k, np.nextafter(k, -np.inf), np.nextafter(k, np.inf)):
                    pts[r, d] = x
                    r += 1
    assert r <= n
    return np.ascontiguousarray(pts)


def _order_rows(rng, D, rows, max_active=None, max_order=2):
    out = [[0] * D]
    while len(out) < rows:
        k = int(rng.integers(1, (max_active or D) + 1))
        o = [0] * D
        for d in rng.choice(D, size=min(k, D), replace=False):
            o[int(d)] = int(rng.integers(1, max_order + 1))
        out.append(o)
    return out


# ------------------------------------------------------------------------------------------
# full tensor
# ------------------------------------------------------------------------------------------

def _full_shape(rng, D, cap):
    while True:
        n = [int(v) for v in rng.integers(3, 18, D)]
        if int(np.prod(n)) <= cap:
            return n


@pytest.mark.parametrize("case", range(14))
def test_full_tensor_random_shapes(case):
    _ref()
    from oracle import ref_objects as RO
    from pychebyshev_b200 import dropin

    rng = np.random.default_rng(BASE_SEED + case)
    D = 1 + case % 6
    n_nodes = _full_shape(rng, D, 250_000)
    domain = _domain(rng, D)
    cheb = RO.full_from_func(_smooth(rng, domain), domain, n_nodes)
    pts = _queries(rng, domain, 400, nodes=cheb.nodes)
    orders = _order_rows(rng, D, 5)
    mirror = dropin.adopt(cheb)
    got = mirror.eval_batch_multi(pts, orders)
    for g, o in enumerate(orders):
        ref = cheb.vectorized_eval_batch(pts, list(o))
        G.assert_close_scaled(got[:, g], ref, 1.0, f"full D={D} n={n_nodes} order {o}")
        # one order per call may pick another kernel form (other summation order): same bound
        G.assert_close_scaled(mirror.vectorized_eval_batch(pts, list(o)), ref, 1.0,
                              f"full D={D} n={n_nodes} order {o}, single-order call")
    # the single-point entry points ride the same kernels.  The reference's own single-point path
    # applies D^T to the contracted vectors instead of the tensor (App. B.1 of SURVEY.md), and for
    # a high mixed derivative its two paths drift apart by far more than 1e-12 (5.8e-8 of the
    # value scale for order [2,2,1,1,1] on a 15x12x15x3x6 grid): the bound is the strict one
    # against the hoisted-D oracle and the reference's own spread against its single-point path.
    p = [float(v) for v in pts[17]]
    o = orders[1]
    scale = max(1.0, float(np.abs(got[:, 1]).max()))
    mine = mirror.vectorized_eval(p, o)
    ref_batch = float(cheb.vectorized_eval_batch(pts[17:18], list(o))[0])
    ref_single = float(cheb.vectorized_eval(p, o))
    assert abs(mine - ref_batch) <= G.REL * abs(ref_batch) + G.ABS * scale
    assert abs(mine - ref_single) <= G.REL * abs(ref_single) + G.ABS * scale + 4.0 * abs(ref_single - ref_batch)


# ------------------------------------------------------------------------------------------
# tensor train
# ------------------------------------------------------------------------------------------

def _tt_case(rng, case):
    from pychebyshev_b200 import workloads as wl

    D = int(rng.integers(2, 11))
    n_nodes = [int(v) for v in rng.integers(3, 15, D)]
    rmax = int(rng.choice([2, 5, 8, 11, 14, 20, 26]))
    ranks = [1] + [int(rng.integers(1, rmax + 1)) for _ in range(D - 1)] + [1]
    cores = wl.synthetic_tt_cores(n_nodes, ranks, BASE_SEED + 100 + case)
    domain = _domain(rng, D)
    dim_order = [int(v) for v in rng.permutation(D)] if case % 2 else list(range(D))
    return cores, domain, dim_order


@pytest.mark.parametrize("case", range(16))
def test_tt_random_trains_values(case):
    _ref()
    from oracle import ref_objects as RO
    from pychebyshev_b200 import dropin

    rng = np.random.default_rng(BASE_SEED + 1000 + case)
    cores, sdomain, dim_order = _tt_case(rng, case)
    tt = RO.tt_from_cores(cores, sdomain, dim_order)
    D = len(cores)
    udomain = [sdomain[dim_order.index(u)] for u in range(D)]
    pts = _queries(rng, udomain, 3000)
    ref = tt.eval_batch(pts)
    got = dropin.adopt(tt).eval_batch(pts)
    G.assert_close_scaled(got, ref, 1.0, f"TT values D={D} ranks={tt.tt_ranks} order={dim_order}")


@pytest.mark.parametrize("case", range(12))
def test_tt_random_trains_finite_difference_rows(case):
    _ref()
    from oracle import ref_objects as RO
    from pychebyshev_b200 import dropin

    rng = np.random.default_rng(BASE_SEED + 2000 + case)
    cores, sdomain, dim_order = _tt_case(rng, case)
    tt = RO.tt_from_cores(cores, sdomain, dim_order)
    D = len(cores)
    udomain = [sdomain[dim_order.index(u)] for u in range(D)]
    pts = _queries(rng, udomain, 60)           # includes both domain corners: nudged stencils
    orders = _order_rows(rng, D, 6, max_active=min(D, 3))
    ref = np.array([tt.eval_multi([float(v) for v in p], orders) for p in pts])
    got = dropin.adopt(tt).eval_multi_batch(pts, orders)
    fake = {"fd_orders": np.asarray(orders), "fd_single_values": ref[:, 0]}
    tol = G.fd_tolerance(fake, sdomain, dim_order)
    err = np.abs(got - ref)
    assert (err <= tol[None, :]).all(), (
        f"TT FD D={D} ranks={tt.tt_ranks} orders={orders}: worst err/tol "
        f"{float((err / tol[None, :]).max()):.3g}")


# ------------------------------------------------------------------------------------------
# spline
# ------------------------------------------------------------------------------------------

@pytest.mark.parametrize("case", range(12))
def test_spline_random_knot_layouts(case):
    _ref()
    from oracle import ref_objects as RO
    from pychebyshev_b200 import dropin

    rng = np.random.default_rng(BASE_SEED + 3000 + case)
    D = 1 + case % 3
    domain = _domain(rng, D)
    n_nodes = [int(v) for v in rng.integers(3, 19 if D < 3 else 12, D)]
    knots = []
    for lo, hi in domain:
        k = int(rng.integers(0, 4))
        knots.append(sorted(float(v) for v in rng.uniform(lo + 0.05 * (hi - lo), hi - 0.05 * (hi - lo), k)))
    sp = RO.spline_from_func(_smooth(rng, domain), domain, n_nodes, knots)
    mirror = dropin.adopt(sp)
    pts = _queries(rng, domain, 5000, knots=knots)
    piece = RO.spline_lookup(sp, pts)
    assert np.array_equal(mirror.find_pieces(pts), piece), f"routing D={D} knots={knots}"
    for o in _order_rows(rng, D, 4):
        # eval_batch routes a query ON a knot to the right-hand piece and differentiates there
        # (only the single-point eval refuses, spline.py:519-549): both sides get every query
        G.assert_close_scaled(mirror.eval_batch(pts, o), sp.eval_batch(pts, o), 1.0,
                              f"spline D={D} n={n_nodes} knots={[len(k) for k in knots]} order {o}")
    on_knot = next(((d, k) for d in range(D) for k in knots[d]), None)
    if on_knot is not None:
        d, k = on_knot
        p = [float(v) for v in pts[0]]
        p[d] = k
        o = [0] * D
        o[d] = 1
        with pytest.raises(ValueError):
            sp.eval(p, o)
        with pytest.raises(ValueError):
            mirror.eval(p, o)


# ------------------------------------------------------------------------------------------
# slider
# ------------------------------------------------------------------------------------------

@pytest.mark.parametrize("case", range(10))
def test_slider_random_partitions(case):
    ref_mod = _ref()
    from pychebyshev_b200 import dropin

    rng = np.random.default_rng(BASE_SEED + 4000 + case)
    D = int(rng.integers(2, 9))
    dims = [int(v) for v in rng.permutation(D)]
    partition, i = [], 0
    while i < D:
        k = int(rng.integers(1, 4))
        partition.append(sorted(dims[i:i + k]))
        i += k
    domain = _domain(rng, D)
    n_nodes = [int(v) for v in rng.integers(3, 12, D)]
    pivot = [float(rng.uniform(lo, hi)) for lo, hi in domain]
    f = _smooth(rng, domain, terms=2)

    def func(x, _):
        return float(f(*x))

    sl = ref_mod.ChebyshevSlider(func, D, domain, n_nodes, partition, pivot)
    sl.build(verbose=False)
    mirror = dropin.adopt(sl)
    pts = _queries(rng, domain, 200)
    orders = _order_rows(rng, D, 5, max_active=2)
    ref = np.array([[sl.eval([float(v) for v in p], list(o)) for o in orders] for p in pts])
    got = mirror.eval_batch_multi(pts, orders)
    for g, o in enumerate(orders):
        G.assert_close_scaled(got[:, g], ref[:, g], 1.0, f"slider D={D} partition={partition} order {o}",
                              rel=2e-11 if any(o) else G.REL)


# ------------------------------------------------------------------------------------------
# .pcb files written by the reference, read by the native loader (no Python on the path)
# ------------------------------------------------------------------------------------------

@pytest.mark.parametrize("case", range(8))
def test_pcb_files_written_by_the_reference(case, tmp_path):
    _ref()
    import pychebyshev_b200 as pcb
    from oracle import ref_objects as RO

    rng = np.random.default_rng(BASE_SEED + 5000 + case)
    D = 1 + case % 4
    domain = _domain(rng, D)
    path = tmp_path / "f.pcb"
    if case % 2 == 0:
        n_nodes = _full_shape(rng, D, 60_000)
        obj = RO.full_from_func(_smooth(rng, domain), domain, n_nodes)
        pts = _queries(rng, domain, 600, nodes=obj.nodes)
        kind = "approx"
    else:
        n_nodes = [int(v) for v in rng.integers(3, 12, D)]
        knots = [sorted(float(v) for v in rng.uniform(lo + 0.1 * (hi - lo), hi - 0.1 * (hi - lo),
                                                      int(rng.integers(0, 3))))
                 for lo, hi in domain]
        obj = RO.spline_from_func(_smooth(rng, domain), domain, n_nodes, knots)
        pts = _queries(rng, domain, 600, knots=knots)
        kind = "spline"
    obj.save(str(path), format="binary")
    orders = _order_rows(rng, D, 4)
    plan = pcb.load_plan(path, orders=orders)
    assert plan.kind == kind and plan.G == len(orders)
    got = plan.eval(pts)
    for g, o in enumerate(orders):
        ref = obj.vectorized_eval_batch(pts, list(o)) if kind == "approx" else obj.eval_batch(pts, o)
        # the loader rebuilds nodes, weights and differentiation matrices natively from the header
        # (no arrays of the reference object are shared): same 1e-12-class bound
        G.assert_close_scaled(got[:, g], ref, 1.0, f".pcb {kind} D={D} n={n_nodes} order {o}")


# ------------------------------------------------------------------------------------------
# kernels specialised on the node count of uniform grids (8, 11, 12, 15, 16 nodes in every dimension)
# ------------------------------------------------------------------------------------------

@pytest.mark.parametrize("n", [8, 11, 12, 15, 16, 9])
@pytest.mark.parametrize("kind", ["spline2d", "spline3d", "slider"])
def test_fixed_node_count_variants(kind, n, monkeypatch):
    """Uniform grids take straight-line weight-row code (template argument = node count); 9 has no
    variant and stays on the generic kernel.  Same arithmetic in the same order: the results must be
    BIT-IDENTICAL to the generic kernel's (PCB_NO_NFIX), and within the usual bound of the reference."""
    ref_mod = _ref()
    from oracle import ref_objects as RO
    from pychebyshev_b200 import dropin

    rng = np.random.default_rng(BASE_SEED + 6000 + n)
    if kind == "slider":
        D = 4
        domain = _domain(rng, D)
        f = _smooth(rng, domain, terms=2)
        obj = ref_mod.ChebyshevSlider(lambda x, _: float(f(*x)), D, domain, [n] * D, [[0, 2], [1, 3]],
                                      [float(rng.uniform(lo, hi)) for lo, hi in domain])
        obj.build(verbose=False)
        pts = _queries(rng, domain, 3000)
        orders = [[0] * D, [1, 0, 0, 0], [0, 0, 0, 2], [1, 1, 0, 0]]
        ref = np.array([[obj.eval([float(v) for v in p], list(o)) for o in orders] for p in pts[:150]])
        run = lambda: dropin.adopt(obj, cached=False).eval_batch_multi(pts, orders)  # noqa: E731
    else:
        D = 2 if kind == "spline2d" else 3
        domain = _domain(rng, D)
        knots = [[float(rng.uniform(domain[0][0] + 0.2 * (domain[0][1] - domain[0][0]),
                                    domain[0][1] - 0.2 * (domain[0][1] - domain[0][0])))]] + [[]] * (D - 1)
        obj = RO.spline_from_func(_smooth(rng, domain), domain, [n] * D, knots)
        pts = _queries(rng, domain, 3000, knots=knots)
        orders = [[0] * D, [1] + [0] * (D - 1), [0] * (D - 1) + [1]]
        ref = np.stack([obj.eval_batch(pts[:150], o) for o in orders], axis=1)
        run = lambda: dropin.adopt(obj, cached=False).eval_batch_multi(pts, orders)  # noqa: E731
    fixed = run()
    monkeypatch.setenv("PCB_NO_NFIX", "1")
    generic = run()
    assert np.array_equal(fixed, generic), f"{kind} n={n}: specialised and generic kernels differ"
    for g, o in enumerate(orders):
        G.assert_close_scaled(fixed[:150, g], ref[:, g], 1.0, f"{kind} n={n} order {o}",
                              rel=2e-11 if kind == "slider" and any(o) else G.REL)
