"""Same-box, same-run parity: the UNMODIFIED reference and the CUDA engine side by side.

SURVEY.md §8(c): the reference is not bit-reproducible across machines (BLAS / einsum summation
order), so parity is defined against the reference *run on the same box in the same run*.  The
reference package travels to the GPU box under oracle/_ref/site (oracle/ref_install.py); each test
builds a BASELINE.json config with the reference's own constructors, draws FRESH seeded queries
(not in tests/golden), evaluates with the reference's own methods on the host and with the engine
on the B200, and compares under the north-star tolerance (tests/_golden.py).

The engine side goes through ``pychebyshev_b200.dropin``: ``adopt(ref_obj)`` (mirror sharing the
reference object's arrays) and ``install()`` (the reference's methods rebound to the engine).
"""

import numpy as np
import pytest

import _golden as G

pytestmark = pytest.mark.gpu

SEED = 20261018


def _ref():
    from oracle import reference as R

    try:
        return R.load()
    except R.ReferenceUnavailable as exc:
        pytest.skip(str(exc))


def _uniform(domain, n, seed):
    from pychebyshev_b200 import workloads as wl

    return wl.uniform_queries(domain, n, seed)


def scale_close(got, ref, what, factor=1.0):
    G.assert_close_scaled(got, ref, factor, what)


# ------------------------------------------------------------------------------------------
# C1 / C4: full tensor
# ------------------------------------------------------------------------------------------

def test_c1_full_bs5d_price_and_greeks_side_by_side():
    _ref()
    from oracle import ref_objects as RO
    from pychebyshev_b200 import dropin, workloads as wl

    cheb = RO.full_bs5d()
    pts = _uniform(wl.BS5D_DOMAIN, 400, SEED)
    pts[0] = [cheb.nodes[d][3] for d in range(5)]          # all dims on a node
    pts[1, 2] = cheb.nodes[2][7] + 5e-15                    # inside the coincidence window
    ref = np.stack([cheb.vectorized_eval_batch(pts, list(o)) for o in wl.BS5D_GREEKS], axis=1)
    got = dropin.adopt(cheb).eval_batch_multi(pts, wl.BS5D_GREEKS)
    for g, o in enumerate(wl.BS5D_GREEKS):
        scale_close(got[:, g], ref[:, g], f"C1 order {o}")


def test_c4_full_16p6_side_by_side():
    _ref()
    from oracle import ref_objects as RO
    from pychebyshev_b200 import dropin, workloads as wl

    cheb = RO.full_c4()
    pts = _uniform(wl.C4_DOMAIN, 24, SEED + 1)
    orders = [wl.C4_GREEKS[0], wl.C4_GREEKS[1]]
    ref0 = cheb.vectorized_eval_batch(pts, list(orders[0]))       # 18 q/s on the host
    ref1 = cheb.vectorized_eval_batch(pts[:8], list(orders[1]))   # 4.7 q/s
    got = dropin.adopt(cheb).eval_batch_multi(pts, orders)
    scale_close(got[:, 0], ref0, "C4 price")
    scale_close(got[:8, 1], ref1, "C4 delta")


# ------------------------------------------------------------------------------------------
# C2 / C5: tensor train
# ------------------------------------------------------------------------------------------

def test_c2_tt_bs5d_built_here_by_the_reference_side_by_side():
    """TT-Cross build by the reference on THIS box (fresh seed), then values at 1e5 points and
    price+Greeks (eval_multi per point) at 400 points incl. points within 1.5 h of the edges."""
    _ref()
    from oracle import ref_objects as RO
    from pychebyshev_b200 import dropin, workloads as wl

    tt = RO.tt_bs5d_build(seed=SEED % 1000)
    mirror = dropin.adopt(tt)
    pts = _uniform(wl.BS5D_DOMAIN, 100_003, SEED + 2)
    scale_close(mirror.eval_batch(pts), tt.eval_batch(pts), "C2 eval_batch")
    q = pts[:400].copy()
    dom = np.array(wl.BS5D_DOMAIN)
    h = (dom[:, 1] - dom[:, 0]) * 1e-4
    q[0] = dom[:, 0]
    q[1] = dom[:, 1]
    q[2] = dom[:, 0] + 0.7 * h
    q[3] = dom[:, 1] - 1.2 * h
    ref = np.array([tt.eval_multi(list(map(float, p)), wl.BS5D_GREEKS) for p in q])
    got = mirror.eval_multi_batch(q, wl.BS5D_GREEKS)
    fake = {"fd_orders": np.asarray(wl.BS5D_GREEKS), "fd_single_values": ref[:, 0]}
    tol = G.fd_tolerance(fake, tt.domain, list(tt._dim_order))
    assert (np.abs(got - ref) <= tol[None, :]).all(), float(np.max(np.abs(got - ref) / tol))
    scale_close(got[:, 0], ref[:, 0], "C2 eval_multi value row")


@pytest.mark.parametrize("name", ["tt_basket10d", "tt_rank20_10d"])
def test_c5_tt_10d_side_by_side(name):
    _ref()
    from oracle import ref_objects as RO
    from pychebyshev_b200 import dropin

    g = G.load(name)
    cores, domain, dim_order = G.tt_parts(g)
    tt = RO.tt_from_cores(cores, domain, dim_order)
    mirror = dropin.adopt(tt)
    udom = [domain[dim_order.index(u)] for u in range(len(domain))]
    pts = _uniform(udom, 20_011, SEED + 3)
    scale_close(mirror.eval_batch(pts), tt.eval_batch(pts), f"{name} eval_batch")
    orders = [list(map(int, o)) for o in g["fd_orders"][:3]]
    q = pts[:60]
    ref = np.array([tt.eval_multi(list(map(float, p)), orders) for p in q])
    got = mirror.eval_multi_batch(q, orders)
    fake = {"fd_orders": np.asarray(orders), "fd_single_values": ref[:, 0]}
    tol = G.fd_tolerance(fake, domain, dim_order)
    assert (np.abs(got - ref) <= tol[None, :]).all()


def test_c5_slider_side_by_side():
    _ref()
    from oracle import ref_objects as RO
    from pychebyshev_b200 import dropin, workloads as wl

    sl = RO.slider10d()
    mirror = dropin.adopt(sl)
    pts = _uniform(wl.C5_DOMAIN, 500, SEED + 4)
    orders = [[0] * 10, [1] + [0] * 9, [1, 1] + [0] * 8, [1, 0, 1] + [0] * 7, [0] * 9 + [2]]
    ref = np.array([[sl.eval(list(map(float, p)), list(o)) for o in orders] for p in pts])
    got = mirror.eval_batch_multi(pts, orders)
    scale_close(got[:, 0], ref[:, 0], "slider values")
    assert (got[:, 3] == 0.0).all() and (ref[:, 3] == 0.0).all()  # cross-slide partial: exactly 0
    for c in (1, 2, 4):
        # the reference's single-point path interleaves D^T with the contraction (App. B.1)
        G.assert_close_scaled(got[:, c], ref[:, c], 1.0, f"slider order {orders[c]}", rel=2e-11)


# ------------------------------------------------------------------------------------------
# C3: spline
# ------------------------------------------------------------------------------------------

@pytest.mark.parametrize("which", ["spline2d", "spline3d"])
def test_c3_spline_side_by_side(which):
    _ref()
    from oracle import ref_objects as RO
    from pychebyshev_b200 import dropin, workloads as wl

    sp = getattr(RO, which)()
    dom = wl.SPLINE2D_DOMAIN if which == "spline2d" else wl.SPLINE3D_DOMAIN
    D = len(dom)
    pts = _uniform(dom, 20_000, SEED + 5)
    pts[0, 0] = 100.0
    pts[1, 0] = np.nextafter(100.0, -np.inf)
    pts[2, 0] = np.nextafter(100.0, np.inf)
    pts[3, 0] = dom[0][0]
    pts[4, 0] = dom[0][1]
    mirror = dropin.adopt(sp)
    piece = RO.spline_lookup(sp, pts)
    assert np.array_equal(mirror.find_pieces(pts), piece)
    for o in ([0] * D, [1] + [0] * (D - 1), [0, 1] + [0] * (D - 2)):
        scale_close(mirror.eval_batch(pts, o), sp.eval_batch(pts, o), f"{which} order {o}")
    # lookup corner cases the values cannot cover: NaN, +-inf, out of domain
    odd = pts[:8].copy()
    odd[:, 0] = [np.nan, np.inf, -np.inf, 0.0, 1e300, -1e300, 100.0, 99.99999999999999]
    assert np.array_equal(mirror.find_pieces(odd), RO.spline_lookup(sp, odd))


# ------------------------------------------------------------------------------------------
# the reference's own methods, rebound to the engine
# ------------------------------------------------------------------------------------------

def test_install_rebinds_reference_methods_and_uninstall_restores_them():
    ref = _ref()
    from oracle import ref_objects as RO
    from pychebyshev_b200 import _engine, dropin, workloads as wl

    tt = RO.tt_from_cores(*G.tt_parts(G.load("tt_bs5d")))
    sp = RO.spline2d()
    pts5 = _uniform(wl.BS5D_DOMAIN, 2000, SEED + 6)
    pts2 = _uniform(wl.SPLINE2D_DOMAIN, 2000, SEED + 7)
    cpu_tt = tt.eval_batch(pts5)
    cpu_multi = tt.eval_multi(list(pts5[0]), wl.BS5D_GREEKS)
    cpu_sp = sp.eval_batch(pts2, [1, 0])
    orig = ref.ChebyshevTT.eval_batch
    dropin.install(ref)
    try:
        assert ref.ChebyshevTT.eval_batch is not orig and dropin.installed()
        n0 = _engine.launch_count()
        gpu_tt = tt.eval_batch(pts5)
        gpu_multi = tt.eval_multi(list(pts5[0]), wl.BS5D_GREEKS)
        gpu_sp = sp.eval_batch(pts2, [1, 0])
        single = sp.eval([95.0, 0.5], [0, 0])
        assert _engine.launch_count() >= n0 + 4
        assert isinstance(gpu_tt, np.ndarray) and gpu_tt.shape == (2000,)
        assert isinstance(gpu_multi, list) and isinstance(gpu_multi[0], float)
        assert isinstance(single, float)
        scale_close(gpu_tt, cpu_tt, "patched TT eval_batch")
        scale_close(gpu_sp, cpu_sp, "patched spline eval_batch")
        assert abs(gpu_multi[0] - cpu_multi[0]) <= 1e-12 * abs(cpu_multi[0]) + 1e-13
        # error behaviour is the reference's
        with pytest.raises(ValueError, match="not defined at knot"):
            sp.eval([100.0, 0.5], [1, 0])
        with pytest.raises(ValueError, match="not supported"):
            tt.eval_multi(list(pts5[0]), [[3, 0, 0, 0, 0]])
        with pytest.raises(ValueError):
            sp.eval_batch(pts2)  # neither derivative_order nor derivative_id
        # in-place edits of the reference object's arrays invalidate the cached mirror
        tt._coeff_cores[0] *= 2.0
        scale_close(tt.eval_batch(pts5), 2.0 * cpu_tt, "after in-place core edit")
        tt._coeff_cores[0] *= 0.5
    finally:
        dropin.uninstall()
    assert ref.ChebyshevTT.eval_batch is orig and not dropin.installed()
    assert np.array_equal(tt.eval_batch(pts5), cpu_tt)


# ------------------------------------------------------------------------------------------
# N4: TT build-side evaluation through the chain kernel
# ------------------------------------------------------------------------------------------

@pytest.mark.parametrize("name", ["tt_bs5d", "tt_4d_perm"])
def test_n4_to_dense_and_grid_index_sets_side_by_side(name):
    """``to_dense`` (tensor_train.py:1874-1917) and batched evaluation at grid multi-indices (the
    TT-Cross index sets / convergence check, :223-228, :287-297, :321-330) on the device against the
    reference's own ``to_dense`` on this box."""
    _ref()
    from oracle import ref_objects as RO
    from pychebyshev_b200 import dropin

    g = G.load(name)
    cores, domain, dim_order = G.tt_parts(g)
    tt = RO.tt_from_cores(cores, domain, dim_order)
    mirror = dropin.adopt(tt)
    ref = tt.to_dense()
    got = mirror.to_dense()
    assert got.shape == ref.shape
    scale_close(got.ravel(), ref.ravel(), f"{name} to_dense")
    rng = np.random.default_rng(SEED + 8)
    idx = np.column_stack([rng.integers(0, n, size=5000) for n in ref.shape])
    vals = mirror.eval_grid_indices(idx)
    scale_close(vals, ref[tuple(idx.T)], f"{name} grid index sets")
    # a cross-matrix index set of the reference's TT-Cross: left x node x right (C-order rows)
    D = len(ref.shape)
    k = D // 2
    left = np.column_stack([rng.integers(0, ref.shape[d], size=6) for d in range(k)])
    right = np.column_stack([rng.integers(0, ref.shape[d], size=5) for d in range(k + 1, D)])
    rows = [list(a) + [i] + list(b) for a in left for i in range(ref.shape[k]) for b in right]
    C = mirror.eval_grid_indices(np.asarray(rows)).reshape(6 * ref.shape[k], 5)
    want = np.array([ref[tuple(r)] for r in rows]).reshape(C.shape)
    scale_close(C.ravel(), want.ravel(), f"{name} cross matrix")
    with pytest.raises(IndexError):
        mirror.eval_grid_indices(np.full((1, D), 99))
