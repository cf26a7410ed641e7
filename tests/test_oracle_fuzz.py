"""The NumPy restatement (oracle/np_oracle.py) against the UNMODIFIED reference, live, on seeded
random configurations: bit-identical outputs on the machine both run on (same NumPy, same BLAS).
tests/test_oracle_golden.py pins the oracle on the committed fixtures; this walks shapes, ranks,
storage permutations, knot layouts and derivative rows that no fixture holds.  CPU only."""

import numpy as np
import pytest

from oracle import np_oracle as O

SEED = 90210


def _ref():
    from oracle import reference as R

    try:
        return R.load()
    except R.ReferenceUnavailable as exc:
        pytest.skip(str(exc))


def _domain(rng, D):
    lo = rng.uniform(-3.0, 3.0, D)
    return [[float(a), float(a + w)] for a, w in zip(lo, rng.choice([0.3, 1.0, 40.0], D))]


def _func(rng, domain):
    D = len(domain)
    freq = rng.uniform(0.5, 2.5, D) / np.array([hi - lo for lo, hi in domain])
    phase = rng.uniform(0, 6.28, D)

    def f(*xs):
        out = 1.0
        for d, x in enumerate(xs):
            out = out * np.cos(freq[d] * x + phase[d])
        return out + 0.25

    return f


def _points(rng, domain, n, nodes=None):
    pts = np.stack([rng.uniform(lo, hi, n) for lo, hi in domain], axis=1)
    if nodes is not None:
        for r in range(4):
            for d in range(len(domain)):
                pts[r, d] = nodes[d][rng.integers(len(nodes[d]))]
        pts[4, 0] = nodes[0][0] + 3e-15
    pts[5] = [lo for lo, _ in domain]
    pts[6] = [hi for _, hi in domain]
    return np.ascontiguousarray(pts)


@pytest.mark.parametrize("case", range(10))
def test_full_tensor_oracle_is_the_reference_bit_for_bit(case):
    _ref()
    from oracle import ref_objects as RO

    rng = np.random.default_rng(SEED + case)
    D = 1 + case % 5
    n = [int(v) for v in rng.integers(3, 13, D)]
    domain = _domain(rng, D)
    cheb = RO.full_from_func(_func(rng, domain), domain, n)
    pts = _points(rng, domain, 60, cheb.nodes)
    for _ in range(3):
        order = [int(v) for v in rng.integers(0, 3, D)]
        want = cheb.vectorized_eval_batch(pts, order)
        got = O.full_eval_batch(cheb.tensor_values, cheb.nodes, cheb.weights, cheb.diff_matrices, pts, order)
        assert np.array_equal(got, want), (D, n, order)
    for d in range(D):
        lo, hi = domain[d]
        assert np.array_equal(O.make_nodes(lo, hi, n[d]), cheb.nodes[d])
        assert np.array_equal(O.barycentric_weights(cheb.nodes[d]), cheb.weights[d])
        assert np.array_equal(O.diff_matrix(cheb.nodes[d], cheb.weights[d]), cheb.diff_matrices[d])


@pytest.mark.parametrize("case", range(10))
def test_tt_oracle_is_the_reference_bit_for_bit(case):
    _ref()
    from oracle import ref_objects as RO
    from pychebyshev_b200 import workloads as wl

    rng = np.random.default_rng(SEED + 100 + case)
    D = int(rng.integers(2, 9))
    n = [int(v) for v in rng.integers(3, 13, D)]
    ranks = [1] + [int(rng.integers(1, 13)) for _ in range(D - 1)] + [1]
    cores = wl.synthetic_tt_cores(n, ranks, SEED + case)
    sdomain = _domain(rng, D)
    dim_order = [int(v) for v in rng.permutation(D)] if case % 2 else list(range(D))
    tt = RO.tt_from_cores(cores, sdomain, dim_order)
    udomain = [sdomain[dim_order.index(u)] for u in range(D)]
    pts = _points(rng, udomain, 40)
    assert np.array_equal(O.tt_eval_batch(cores, sdomain, dim_order, pts), tt.eval_batch(pts))
    orders = [[0] * D]
    for _ in range(4):
        o = [0] * D
        for d in rng.choice(D, size=min(D, int(rng.integers(1, 4))), replace=False):
            o[int(d)] = int(rng.integers(1, 3))
        orders.append(o)
    want = np.array([tt.eval_multi([float(v) for v in p], orders) for p in pts])
    got = O.tt_eval_multi_batch(cores, sdomain, dim_order, pts, np.asarray(orders))
    assert np.array_equal(got, want), (D, ranks, orders)
    with pytest.raises(ValueError):
        tt.eval_multi([float(v) for v in pts[0]], [[3] + [0] * (D - 1)])
    with pytest.raises(ValueError):
        O.tt_eval_multi_batch(cores, sdomain, dim_order, pts[:1], np.asarray([[3] + [0] * (D - 1)]))


@pytest.mark.parametrize("case", range(8))
def test_spline_oracle_is_the_reference_bit_for_bit(case):
    _ref()
    from oracle import ref_objects as RO

    rng = np.random.default_rng(SEED + 200 + case)
    D = 1 + case % 3
    domain = _domain(rng, D)
    n = [int(v) for v in rng.integers(3, 10, D)]
    knots = [sorted(float(v) for v in rng.uniform(lo + 0.1 * (hi - lo), hi - 0.1 * (hi - lo),
                                                  int(rng.integers(0, 3))))
             for lo, hi in domain]
    sp = RO.spline_from_func(_func(rng, domain), domain, n, knots)
    pts = _points(rng, domain, 80)
    r = 7
    for d in range(D):
        for k in knots[d]:
            for x in (k, np.nextafter(k, -np.inf), np.nextafter(k, np.inf)):
                pts[r, d] = x
                r += 1
    pts[r, 0] = np.nan
    shape = list(sp._shape)
    assert np.array_equal(O.spline_lookup(knots, shape, pts), RO.spline_lookup(sp, pts))
    pieces = [(p.tensor_values, p.nodes, p.weights, p.diff_matrices) for p in sp._pieces]
    ok = ~np.isnan(pts[:, 0])
    for _ in range(3):
        order = [int(v) for v in rng.integers(0, 2, D)]
        want = sp.eval_batch(pts[ok], order)
        got = O.spline_eval_batch(knots, shape, pieces, pts[ok], order)
        assert np.array_equal(got, want), (D, n, knots, order)


@pytest.mark.parametrize("case", range(6))
def test_slider_oracle_is_the_reference_bit_for_bit(case):
    ref = _ref()
    rng = np.random.default_rng(SEED + 300 + case)
    D = int(rng.integers(2, 8))
    dims = [int(v) for v in rng.permutation(D)]
    partition, i = [], 0
    while i < D:
        k = int(rng.integers(1, 4))
        partition.append(sorted(dims[i:i + k]))
        i += k
    domain = _domain(rng, D)
    n = [int(v) for v in rng.integers(3, 9, D)]
    pivot = [float(rng.uniform(lo, hi)) for lo, hi in domain]
    f = _func(rng, domain)
    sl = ref.ChebyshevSlider(lambda x, _: float(f(*x)), D, domain, n, partition, pivot)
    sl.build(verbose=False)
    slides = [(s.tensor_values, s.nodes, s.weights, s.diff_matrices) for s in sl.slides]
    pts = _points(rng, domain, 40)
    orders = [[0] * D]
    for _ in range(3):
        o = [0] * D
        for d in rng.choice(D, size=min(D, int(rng.integers(1, 3))), replace=False):
            o[int(d)] = int(rng.integers(1, 3))
        orders.append(o)
    for o in orders:
        want = np.array([sl.eval([float(v) for v in p], list(o)) for p in pts])
        got = O.slider_eval_batch(partition, float(sl.pivot_value), slides, pts, o)
        if not any(o):
            assert np.array_equal(got, want), (D, partition, o)
        else:
            # the reference's single-point path applies D^T to the contracted vectors, the oracle
            # (like its batch path and the kernels) to the tensor: SURVEY App. B.1, <= 5e-12 observed
            scale = max(1.0, float(np.abs(want).max()))
            assert np.all(np.abs(got - want) <= 2e-11 * np.abs(want) + 1e-14 * scale), (D, partition, o)
