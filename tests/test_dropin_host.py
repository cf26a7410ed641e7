"""Host-side logic added in round 2 (no GPU): plan-cache fingerprints, NUMA helpers, adoption of
reference objects, and the row-set splitting of ``ChebyshevTT.eval_multi_batch`` (any row set the
reference's ``eval_multi`` accepts), checked against the UNMODIFIED reference where it is installed
and against the oracle otherwise."""

import numpy as np
import pytest

import _golden as G
from oracle import np_oracle as O
from pychebyshev_b200 import _engine, _numa


def test_fingerprint_sees_rebinding_and_in_place_edits():
    a = np.arange(1000.0)
    t0 = _engine.fingerprint(a)
    assert _engine.fingerprint(a) == t0
    a[17] += 1.0
    assert _engine.fingerprint(a) != t0          # small arrays: full content hash
    b = np.zeros(2_000_000)
    t1 = _engine.fingerprint(b)
    b *= 0.0
    assert _engine.fingerprint(b) == t1
    b += 1.0                                      # whole-array in-place op: sampled hash changes
    assert _engine.fingerprint(b) != t1
    c = b.copy()
    assert _engine.fingerprint(c) != _engine.fingerprint(b)  # other buffer, same content
    v = np.zeros((8, 8))[:, ::2]                  # non-contiguous views work too
    assert _engine.fingerprint(v) == _engine.fingerprint(v)


def test_numa_helpers_are_safe_without_topology():
    assert _numa._parse_cpulist("0-3,8,10-11") == {0, 1, 2, 3, 8, 10, 11}
    assert _numa._parse_cpulist("") == set()
    assert _numa.node_cpus(-1) == set()
    with _numa.bound_to_device(None) as bound:    # no device -> no-op
        assert bound is False
    assert _numa.describe()["nodes"] >= 1


class _FakePlan:
    """Stands in for the device plan: every (<= 16 rows, <= 3 differentiated dims) launch is answered
    by the NumPy oracle, and the limits of one kernel call are enforced like the C ABI does."""

    def __init__(self, cores, domain, dim_order):
        self.cores, self.domain, self.dim_order = cores, domain, dim_order
        self.calls = 0

    def with_orders(self, orders, algo=0):
        orders = np.asarray(orders)
        assert orders.shape[0] <= 16 and ((orders > 0).sum(axis=1) <= 3).all()
        plan = self

        class _View:
            def eval(self, points, out=None):
                plan.calls += 1
                return O.tt_eval_multi_batch(plan.cores, plan.domain, plan.dim_order,
                                             np.asarray(points), [list(map(int, o)) for o in orders])
        return _View()


def _tt4():
    import pychebyshev_b200 as pcb

    g = G.load("tt_4d_perm")
    cores, domain, dim_order = G.tt_parts(g)
    return g, cores, domain, dim_order, pcb.ChebyshevTT.from_cores(cores, domain, dim_order)


def test_tt_row_sets_beyond_one_kernel_call_match_the_reference_recursion():
    g, cores, domain, dim_order, tt = _tt4()
    pts = g["fd_points"][:25]
    rows = [[0, 0, 0, 0], [1, 1, 1, 1], [2, 1, 1, 2], [1, 0, 0, 0], [1, 2, 1, 1], [0, 1, 1, 0]]
    rows += [[(i >> k) & 1 for k in range(4)] for i in range(16)]   # 22 rows in total
    fake = _FakePlan(cores, domain, dim_order)
    got = tt._eval_rows_split(fake, pts, np.asarray(rows, dtype=np.int64), None, 0)
    assert got.shape == (25, len(rows)) and fake.calls > 2
    try:
        from oracle import ref_objects as RO

        ref_tt = RO.tt_from_cores(cores, domain, dim_order)
        ref = np.array([ref_tt.eval_multi(list(map(float, p)), rows) for p in pts])
    except Exception:  # noqa: BLE001  (reference not installed: the oracle restates eval_multi)
        ref = O.tt_eval_multi_batch(cores, domain, dim_order, pts, rows)
    fakeg = {"fd_orders": np.asarray(rows), "fd_single_values": ref[:, 0]}
    tol = G.fd_tolerance(fakeg, domain, dim_order)
    assert (np.abs(got - ref) <= tol[None, :]).all()
    # rows the kernel handles directly are passed through untouched (bit-identical to the oracle)
    direct = O.tt_eval_multi_batch(cores, domain, dim_order, pts, [rows[3], rows[5]])
    assert np.array_equal(got[:, [3, 5]], direct)
    # four differentiated dims: same nesting order as the reference => agreement far inside the
    # propagated tolerance (identical arithmetic up to the inner kernel's rounding)
    assert np.max(np.abs(got[:, 1] - ref[:, 1])) <= 1e-3 * tol[1]


def test_tt_negative_orders_count_as_zero_and_order3_raises():
    g, cores, domain, dim_order, tt = _tt4()
    with pytest.raises(ValueError, match="not supported"):
        tt.eval_multi_batch(g["fd_points"][:2], [[3, 0, 0, 0]])
    with pytest.raises(ValueError, match="entries"):
        tt.eval_multi_batch(g["fd_points"][:2], [[1, 0, 0]])


def test_adopt_shares_arrays_and_follows_edits():
    from oracle import reference as R

    try:
        ref = R.load()
    except R.ReferenceUnavailable as exc:
        pytest.skip(str(exc))
    from oracle import ref_objects as RO
    from pychebyshev_b200 import dropin, workloads as wl

    cheb = RO.full_from_func(lambda x, y: np.sin(x) * y, [[0.0, 1.0], [1.0, 2.0]], [6, 7])
    m = dropin.adopt(cheb)
    assert type(m).__name__ == "ChebyshevApproximation" and m.tensor_values is cheb.tensor_values
    assert m.nodes is cheb.nodes and m.diff_matrices is cheb.diff_matrices
    assert dropin.adopt(cheb) is m                    # cached
    cheb.tensor_values[0, 0] += 1.0                   # in-place edit -> new mirror
    m2 = dropin.adopt(cheb)
    assert m2 is not m and m2.tensor_values is cheb.tensor_values
    cheb.tensor_values = cheb.tensor_values * 2.0     # rebinding -> new mirror
    assert dropin.adopt(cheb) is not m2
    sp = RO.spline2d()
    ms = dropin.adopt(sp)
    assert ms.num_pieces == 2 and ms._pieces[1].tensor_values is sp._pieces[1].tensor_values
    assert ms.knots == sp.knots and ms._shape == tuple(sp._shape)
    tt = RO.tt_from_cores(*G.tt_parts(G.load("tt_4d_perm")))
    mt = dropin.adopt(tt)
    assert mt.dim_order == list(tt._dim_order) and mt.tt_ranks == list(tt.tt_ranks)
    with pytest.raises(TypeError):
        dropin.adopt(object())
    # without a GPU the patched methods must not silently fall back to the CPU
    import torch

    if not torch.cuda.is_available():
        import pychebyshev_b200 as pcb

        with pytest.raises(pcb.BackendUnavailable):
            dropin.install(ref)
        assert not dropin.installed()
    del wl
