#!/usr/bin/env python
"""Generate the golden input/output fixtures under tests/golden/ from the REFERENCE.

Run in the build container only (the reference does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 NUMBA_CACHE_DIR=/tmp/nb python tests/golden/make_golden.py

It imports the unmodified reference from /root/reference/src, builds the
interpolants of BASELINE.json's configs (plus small edge-case interpolants),
draws seeded query sets and stores the reference's own outputs:

  * ChebyshevApproximation.vectorized_eval_batch      (barycentric.py:992-1047)
  * ChebyshevTT.eval_batch / eval_multi               (tensor_train.py:2217-2320)
  * ChebyshevSpline.eval_batch + the piece lookup     (spline.py:633-700)
  * ChebyshevSlider.eval                              (slider.py:247-318)
  * .pcb bytes written by the reference               (_binary.py:208-346)

The fixtures pin oracle/ (tests/test_oracle_golden.py) and the CUDA path
(tests/test_gpu_parity.py).  Nothing in the repo reads /root/reference at
test/bench time.
"""

from __future__ import annotations

import hashlib
import io
import math
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, REPO)

import pychebyshev as ref  # noqa: E402  (the reference)
from pychebyshev import _binary as ref_binary  # noqa: E402

import importlib.util  # noqa: E402

_spec = importlib.util.spec_from_file_location(
    "workloads", os.path.join(REPO, "pychebyshev_b200", "workloads.py")
)
wl = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(wl)


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    arrays["reference_version"] = np.array(ref.__version__)
    np.savez_compressed(path, **arrays)
    print(f"  wrote {name}.npz  ({os.path.getsize(path) / 1024:.1f} KiB)")


def sha(a):
    return np.array(hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest())


def cat(arrs):
    return np.concatenate([np.asarray(a, dtype=np.float64).ravel() for a in arrs])


# ------------------------------------------------------------------------------------------
# full tensor
# ------------------------------------------------------------------------------------------

def adversarial_points(cheb, rng, n_uniform):
    """Uniform points + exact nodes, node +- tiny offsets, corners, slightly outside."""
    dom = np.array(cheb.domain, dtype=np.float64)
    D = cheb.num_dimensions
    pts = [rng.uniform(dom[:, 0], dom[:, 1], size=(n_uniform, D))]
    base = rng.uniform(dom[:, 0], dom[:, 1], size=(8 * D, D))
    k = 0
    for d in range(D):
        nodes = cheb.nodes[d]
        mid = nodes[len(nodes) // 2]
        for off in (0.0, 1e-15, -1e-15, 5e-15, 1e-13, -1e-13, 1e-9):
            base[k % len(base), d] = mid + off
            k += 1
        base[k % len(base), d] = nodes[0]
        k += 1
    pts.append(base)
    # all-dims-on-node, corners, and extrapolation just outside the domain
    pts.append(np.array([[cheb.nodes[d][1] for d in range(D)]]))
    pts.append(np.array([[cheb.nodes[d][-1] for d in range(D)]]))
    pts.append(dom[:, 0][None, :].copy())
    pts.append(dom[:, 1][None, :].copy())
    pts.append((dom[:, 1] + 0.01 * (dom[:, 1] - dom[:, 0]))[None, :])
    return np.ascontiguousarray(np.concatenate(pts, axis=0))


def full_case(name, func_vec, domain, n_nodes, orders, n_uniform, seed, store_tensor=True):
    t0 = time.time()
    D = len(n_nodes)
    nodes = [ref.barycentric.ChebyshevApproximation.from_values(
        np.zeros((n,)), 1, [domain[d]], [n]).nodes[0] for d, n in enumerate(n_nodes)]
    tensor = wl.grid_values(func_vec, nodes)
    cheb = ref.ChebyshevApproximation.from_values(tensor, D, domain, n_nodes)
    for d in range(D):
        assert np.array_equal(cheb.nodes[d], nodes[d])
    rng = np.random.default_rng(seed)
    pts = adversarial_points(cheb, rng, n_uniform)
    out = np.stack([cheb.vectorized_eval_batch(pts, list(o)) for o in orders], axis=1)
    arrays = dict(
        domain=np.array(domain, dtype=np.float64),
        n_nodes=np.array(n_nodes, dtype=np.int32),
        nodes_cat=cat(cheb.nodes),
        weights_cat=cat(cheb.weights),
        diff_cat=cat(cheb.diff_matrices),
        orders=np.array(orders, dtype=np.int32),
        points=pts,
        values=out,
        tensor_sha256=sha(tensor),
    )
    if store_tensor:
        arrays["tensor"] = tensor
    save(name, **arrays)
    print(f"  {name}: {len(pts)} pts x {len(orders)} orders in {time.time() - t0:.1f}s")
    return cheb


def make_full():
    print("full tensor")
    full_case("full_1d", lambda x: np.sin(x) + 0.1 * x * x, [[0.0, 3.15]], [20],
              [[0], [1], [2]], 300, 1)
    full_case("full_2d", lambda x, y: np.sin(x) * np.cos(2 * y) + x * y,
              [[-1.0, 1.0], [0.0, 2.0]], [9, 14], [[0, 0], [1, 0], [0, 1], [1, 1], [2, 0], [0, 2]],
              400, 2)
    full_case("full_3d", lambda x, y, z: np.sin(x) + np.sin(y) + np.sin(z) + x * y * z,
              [[-1.0, 1.0], [-1.0, 1.0], [-1.0, 1.0]], [10, 8, 4],
              [[0, 0, 0], [1, 0, 0], [0, 0, 1], [0, 2, 0], [1, 0, 1]], 400, 3)
    full_case("full_4d", lambda a, b, c, d: np.exp(-0.3 * a * b) * np.cos(c - d) + a * d,
              [[0.0, 1.0], [0.5, 2.0], [-1.0, 1.0], [0.0, 3.0]], [7, 5, 6, 12],
              [[0, 0, 0, 0], [1, 0, 0, 0], [0, 0, 0, 2], [0, 1, 1, 0]], 400, 4)
    # C1: 5D Black-Scholes 11^5 (tensor regenerated from workloads.bs_call_price in the tests)
    full_case("full_bs5d", wl.bs_call_price, wl.BS5D_DOMAIN, wl.BS5D_NODES,
              wl.BS5D_GREEKS + [[0, 0, 0, 0, 1], [1, 0, 0, 1, 0]], 800, 5, store_tensor=False)
    # C4: 6D 16^6 (134 MB; tensor regenerated in the tests)
    full_case("full_c4_16p6", wl.bs6d, wl.C4_DOMAIN, wl.C4_NODES, wl.C4_GREEKS, 24, 6,
              store_tensor=False)


# ------------------------------------------------------------------------------------------
# tensor train
# ------------------------------------------------------------------------------------------

def tt_points(tt, rng, n_uniform, user_domain):
    dom = np.array(user_domain, dtype=np.float64)
    D = len(user_domain)
    pts = [rng.uniform(dom[:, 0], dom[:, 1], size=(n_uniform, D))]
    edge = rng.uniform(dom[:, 0], dom[:, 1], size=(6 * D, D))
    k = 0
    for d in range(D):
        a, b = dom[d]
        h = (b - a) * 1e-4
        for v in (a, b, a + 0.5 * h, b - 0.5 * h, a + 1.5 * h, b - 1.4999 * h):
            edge[k, d] = v
            k += 1
    pts.append(edge)
    pts.append(dom[:, 0][None, :].copy())
    pts.append(dom[:, 1][None, :].copy())
    return np.ascontiguousarray(np.concatenate(pts, axis=0))


def tt_case(name, tt, user_domain, n_val, n_fd, fd_orders, seed):
    t0 = time.time()
    rng = np.random.default_rng(seed)
    pts = tt_points(tt, rng, n_val, user_domain)
    vals = tt.eval_batch(pts)
    fd_pts = tt_points(tt, rng, n_fd, user_domain)
    fd_vals = np.array([tt.eval_multi(list(map(float, p)), [list(o) for o in fd_orders])
                        for p in fd_pts])
    single = np.array([tt.eval(list(map(float, p))) for p in fd_pts])
    save(
        name,
        domain=np.array(tt.domain, dtype=np.float64),       # storage frame
        n_nodes=np.array(tt.n_nodes, dtype=np.int32),        # storage frame
        ranks=np.array(tt.tt_ranks, dtype=np.int32),
        dim_order=np.array(tt._dim_order, dtype=np.int32),
        cores_cat=cat(tt._coeff_cores),
        points=pts,
        values=vals,
        fd_points=fd_pts,
        fd_orders=np.array(fd_orders, dtype=np.int32),
        fd_values=fd_vals,
        fd_single_values=single,
    )
    print(f"  {name}: ranks {tt.tt_ranks}, {len(pts)} value pts, {len(fd_pts)} FD pts "
          f"in {time.time() - t0:.1f}s")


def make_tt():
    print("tensor train")
    # C2: the reference fixture tt_bs_5d (tests/conftest.py:127-135)
    def bs5(x, _):
        return wl.bs5d_scalar(x)

    tt = ref.ChebyshevTT(bs5, 5, wl.BS5D_DOMAIN, wl.BS5D_NODES, max_rank=15, max_sweeps=5)
    tt.build(verbose=False, seed=42)
    greeks = wl.BS5D_GREEKS + [[0, 0, 0, 0, 1], [1, 0, 0, 1, 0], [2, 0, 0, 1, 0], [0, 0, 2, 0, 0]]
    tt_case("tt_bs5d", tt, wl.BS5D_DOMAIN, 20000, 1500, greeks, 11)

    # permuted storage order, distinct node counts per dim
    def f4(x, _):
        return math.sin(x[0] + 0.5 * x[1]) * math.exp(-0.3 * x[2]) + math.cos(x[3] - x[0]) + x[1] * x[3]

    dom4 = [[-1.0, 1.0], [0.0, 2.0], [0.5, 1.5], [-2.0, 0.0]]
    tt4 = ref.ChebyshevTT(f4, 4, dom4, [7, 9, 6, 8], max_rank=8, tolerance=1e-10)
    tt4.build(verbose=False, method="svd")
    tt_case("tt_4d", tt4, dom4, 3000, 300,
            [[0, 0, 0, 0], [1, 0, 0, 0], [0, 2, 0, 0], [0, 0, 1, 1], [0, 1, 0, 0]], 12)
    tt4p = tt4.reorder([2, 0, 3, 1])
    tt_case("tt_4d_perm", tt4p, dom4, 3000, 300,
            [[0, 0, 0, 0], [1, 0, 0, 0], [0, 2, 0, 0], [0, 0, 1, 1], [0, 1, 0, 0]], 13)

    # C5: 10D basket, rank <= 10
    tt10 = ref.ChebyshevTT(wl.basket10d_scalar, 10, wl.C5_DOMAIN, wl.C5_NODES, max_rank=10,
                           max_sweeps=4)
    tt10.build(verbose=False, seed=7)
    g10 = [[0] * 10, [1] + [0] * 9, [0] * 9 + [2], [0, 0, 1, 0, 0, 0, 0, 1, 0, 0]]
    tt_case("tt_basket10d", tt10, wl.C5_DOMAIN, 20000, 300, g10, 14)

    # synthetic uniform rank 20 (cores injected the way from_values does, tensor_train.py:2946-2964)
    cores = wl.synthetic_tt_cores(wl.C5_NODES, [1] + [20] * 9 + [1], seed=20)
    obj = ref.ChebyshevTT.__new__(ref.ChebyshevTT)
    obj.function = None
    obj.num_dimensions = 10
    obj.domain = [list(b) for b in wl.C5_DOMAIN]
    obj.n_nodes = list(wl.C5_NODES)
    obj.max_rank = 20
    obj.tolerance = 1e-6
    obj.max_sweeps = 10
    obj.max_derivative_order = 2
    obj.additional_data = None
    obj.descriptor = ""
    obj.method = "svd"
    obj._coeff_cores = cores
    obj._tt_ranks = [1] + [20] * 9 + [1]
    obj._built = True
    obj._build_time = 0.0
    obj._total_build_evals = 0
    obj._cached_error_estimate = None
    obj._dim_order = list(range(10))
    tt_case("tt_rank20_10d", obj, wl.C5_DOMAIN, 5000, 100, g10[:3], 15)


# ------------------------------------------------------------------------------------------
# spline
# ------------------------------------------------------------------------------------------

def spline_lookup_ref(sp, pts):
    """The reference's vectorised routing lines, spline.py:677-690, verbatim semantics."""
    N = pts.shape[0]
    mi = np.zeros((N, sp.num_dimensions), dtype=int)
    for d in range(sp.num_dimensions):
        if len(sp.knots[d]) > 0:
            mi[:, d] = np.searchsorted(sp.knots[d], pts[:, d], side="right")
            np.clip(mi[:, d], 0, sp._shape[d] - 1, out=mi[:, d])
    return np.ravel_multi_index(mi.T, sp._shape).astype(np.int32)


def spline_points(sp, rng, n_uniform, with_nan):
    dom = np.array(sp.domain, dtype=np.float64)
    D = sp.num_dimensions
    pts = [rng.uniform(dom[:, 0], dom[:, 1], size=(n_uniform, D))]
    edge = []
    for d in range(D):
        specials = [dom[d, 0], dom[d, 1], dom[d, 0] - 0.1 * (dom[d, 1] - dom[d, 0]),
                    dom[d, 1] + 0.1 * (dom[d, 1] - dom[d, 0])]
        for kn in sp.knots[d]:
            specials += [kn, np.nextafter(kn, -np.inf), np.nextafter(kn, np.inf),
                         kn - 1e-15 * max(1.0, abs(kn)), kn + 1e-13]
            if kn == 0.0:
                specials += [-0.0, 0.0, 5e-324, -5e-324]
        for v in specials:
            p = rng.uniform(dom[:, 0], dom[:, 1])
            p[d] = v
            edge.append(p)
    pts.append(np.array(edge))
    lookup_only = []
    if with_nan:
        for d in range(D):
            for v in (np.nan, np.inf, -np.inf):
                p = rng.uniform(dom[:, 0], dom[:, 1])
                p[d] = v
                lookup_only.append(p)
    return (np.ascontiguousarray(np.concatenate(pts, axis=0)),
            np.ascontiguousarray(np.array(lookup_only)) if lookup_only else np.zeros((0, D)))


def spline_case(name, sp, orders, n_uniform, seed):
    t0 = time.time()
    rng = np.random.default_rng(seed)
    pts, lookup_pts = spline_points(sp, rng, n_uniform, with_nan=True)
    out = np.stack([sp.eval_batch(pts, list(o)) for o in orders], axis=1)
    piece = spline_lookup_ref(sp, pts)
    nested = bool(sp._n_nodes_nested)
    piece_n = np.array([p.n_nodes for p in sp._pieces], dtype=np.int32)
    arrays = dict(
        domain=np.array(sp.domain, dtype=np.float64),
        nested=np.array(nested),
        piece_n_nodes=piece_n,
        num_knots=np.array([len(k) for k in sp.knots], dtype=np.int32),
        knots_cat=cat([np.asarray(k, dtype=np.float64) for k in sp.knots]) if any(
            len(k) for k in sp.knots) else np.zeros(0),
        shape=np.array(sp._shape, dtype=np.int32),
        piece_tensors_cat=cat([p.tensor_values for p in sp._pieces]),
        piece_nodes_cat=cat([cat(p.nodes) for p in sp._pieces]),
        piece_weights_cat=cat([cat(p.weights) for p in sp._pieces]),
        orders=np.array(orders, dtype=np.int32),
        points=pts,
        values=out,
        piece=piece,
        lookup_points=lookup_pts,
        lookup_piece=spline_lookup_ref(sp, lookup_pts) if len(lookup_pts) else np.zeros(0, np.int32),
    )
    if not nested:
        buf = io.BytesIO()
        ref_binary.write_spline(buf, sp)
        arrays["pcb_bytes"] = np.frombuffer(buf.getvalue(), dtype=np.uint8)
    save(name, **arrays)
    print(f"  {name}: shape {sp._shape}, {len(pts)} pts x {len(orders)} orders in "
          f"{time.time() - t0:.1f}s")


def make_spline():
    print("spline")
    sp1 = ref.ChebyshevSpline(lambda x, _: abs(x[0]), 1, [[-1.0, 1.0]], [15], [[0.0]])
    sp1.build(verbose=False)
    spline_case("spline_abs1d", sp1, [[0], [1], [2]], 500, 21)

    sp2 = ref.ChebyshevSpline.from_values(
        _piece_values(wl.payoff2d, wl.SPLINE2D_DOMAIN, wl.SPLINE2D_NODES, wl.SPLINE2D_KNOTS),
        2, wl.SPLINE2D_DOMAIN, wl.SPLINE2D_NODES, wl.SPLINE2D_KNOTS)
    spline_case("spline_bs2d", sp2, [[0, 0], [1, 0], [0, 1], [2, 0]], 2000, 22)

    sp3 = ref.ChebyshevSpline.from_values(
        _piece_values(wl.payoff3d, wl.SPLINE3D_DOMAIN, wl.SPLINE3D_NODES, wl.SPLINE3D_KNOTS),
        3, wl.SPLINE3D_DOMAIN, wl.SPLINE3D_NODES, wl.SPLINE3D_KNOTS)
    spline_case("spline_bs3d", sp3, [[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], 1500, 23)

    # several knots in several dims: exercises the C-order ravel of piece indices
    def fmk(x, y, z):
        return np.abs(x - 0.3) * np.abs(x + 0.4) + np.abs(z - 1.0) * np.cos(y) + x * y

    dom = [[-1.0, 1.0], [0.0, 2.0], [0.0, 3.0]]
    knots = [[-0.4, 0.3], [], [1.0]]
    spm = ref.ChebyshevSpline.from_values(_piece_values(fmk, dom, [6, 7, 5], knots), 3, dom,
                                          [6, 7, 5], knots)
    spline_case("spline_multiknot3d", spm, [[0, 0, 0], [1, 0, 0], [0, 1, 1]], 1500, 24)

    # nested per-piece node counts through the special_points door (barycentric.py:306-338)
    spn = ref.ChebyshevApproximation(
        lambda x, _: max(x[0] - 100.0, 0.0) * math.exp(-0.05 * x[1]), 2, wl.SPLINE2D_DOMAIN,
        n_nodes=[[9, 13], [8]], special_points=[[100.0], []])
    assert isinstance(spn, ref.ChebyshevSpline)
    spn.build(verbose=False)
    spline_case("spline_nested2d", spn, [[0, 0], [1, 0], [0, 2]], 800, 25)


def _piece_values(func_vec, domain, n_nodes, knots):
    """Per-piece value tensors in the C-order of ChebyshevSpline._pieces."""
    import itertools

    info = ref.ChebyshevSpline.nodes(len(domain), domain, n_nodes, knots)
    out = []
    for piece in info["pieces"]:
        nodes = piece["nodes_per_dim"]
        out.append(wl.grid_values(func_vec, nodes))
    del itertools
    return out


# ------------------------------------------------------------------------------------------
# slider
# ------------------------------------------------------------------------------------------

def make_slider():
    print("slider")
    t0 = time.time()
    sl = ref.ChebyshevSlider(wl.basket10d_scalar, wl.C5_DIM, wl.C5_DOMAIN, wl.C5_NODES,
                             wl.C5_PARTITION, wl.C5_PIVOT)
    sl.build(verbose=False)
    rng = np.random.default_rng(31)
    dom = np.array(wl.C5_DOMAIN)
    pts = rng.uniform(dom[:, 0], dom[:, 1], size=(600, 10))
    pts[0, :] = [s.nodes[0][3] for s in sl.slides for _ in (0, 1)]  # node hits
    pts[1, 0] = sl.slides[0].nodes[0][5]
    orders = [[0] * 10, [1] + [0] * 9, [1, 1] + [0] * 8, [1, 0, 1] + [0] * 7, [0] * 9 + [2]]
    out = np.array([[sl.eval(list(map(float, p)), list(o)) for o in orders] for p in pts])
    save(
        "slider10d",
        domain=np.array(sl.domain, dtype=np.float64),
        n_nodes=np.array(sl.n_nodes, dtype=np.int32),
        partition=np.array(sl.partition, dtype=np.int32),
        pivot_point=np.array(sl.pivot_point, dtype=np.float64),
        pivot_value=np.array(sl.pivot_value),
        slide_tensors_cat=cat([s.tensor_values for s in sl.slides]),
        slide_nodes_cat=cat([cat(s.nodes) for s in sl.slides]),
        slide_weights_cat=cat([cat(s.weights) for s in sl.slides]),
        orders=np.array(orders, dtype=np.int32),
        points=pts,
        values=out,
    )
    print(f"  slider10d: {len(pts)} pts x {len(orders)} orders in {time.time() - t0:.1f}s")


# ------------------------------------------------------------------------------------------
# .pcb bytes
# ------------------------------------------------------------------------------------------

def make_pcb():
    print(".pcb golden bytes")
    out = {}
    for fn in ("approx_2d_simple.pcb", "approx_5d_bs.pcb", "spline_1d_kink.pcb"):
        with open(os.path.join("/root/reference/tests/fixtures", fn), "rb") as f:
            raw = f.read()
        key = fn.replace(".pcb", "")
        out[key + "_sha256"] = np.array(hashlib.sha256(raw).hexdigest())
        if len(raw) < 4096:
            out[key + "_bytes"] = np.frombuffer(raw, dtype=np.uint8)
        # what the reference reads out of it, evaluated at a few points
        if fn.startswith("approx"):
            obj = ref.ChebyshevApproximation.load(os.path.join("/root/reference/tests/fixtures", fn))
        else:
            obj = ref.ChebyshevSpline.load(os.path.join("/root/reference/tests/fixtures", fn))
        dom = np.array(obj.domain, dtype=np.float64)
        rng = np.random.default_rng(41)
        pts = rng.uniform(dom[:, 0], dom[:, 1], size=(64, obj.num_dimensions))
        if fn.startswith("approx"):
            vals = obj.vectorized_eval_batch(pts, [0] * obj.num_dimensions)
            out[key + "_tensor"] = obj.tensor_values
        else:
            vals = obj.eval_batch(pts, [0] * obj.num_dimensions)
        out[key + "_points"] = pts
        out[key + "_values"] = vals
    # a fresh reference-written approximation file, byte for byte
    cheb = ref.ChebyshevApproximation.from_values(
        np.arange(12, dtype=np.float64).reshape(3, 4) * 0.25, 2, [[0.0, 1.0], [-2.0, 2.0]], [3, 4])
    buf = io.BytesIO()
    ref_binary.write_approx(buf, cheb)
    out["approx_3x4_bytes"] = np.frombuffer(buf.getvalue(), dtype=np.uint8)
    out["approx_3x4_tensor"] = cheb.tensor_values
    save("pcb_files", **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["full", "tt", "spline", "slider", "pcb"]
    t0 = time.time()
    for w in which:
        {"full": make_full, "tt": make_tt, "spline": make_spline, "slider": make_slider,
         "pcb": make_pcb}[w]()
    print(f"done in {time.time() - t0:.1f}s")
