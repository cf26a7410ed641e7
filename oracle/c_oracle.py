"""ctypes access to oracle/_build/liboracle.so (the plain-C restatement, oracle/c/oracle.c).

TEST INFRASTRUCTURE ONLY -- same rules as oracle/np_oracle.py.  Used as a fast checker for
parity sets the NumPy oracle is too slow for, and as bench.py's timed CPU baseline ("port").
The C code is single-threaded; ``threads > 1`` shards the query range over Python threads
(ctypes releases the GIL during the call).
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_f64p = C.POINTER(C.c_double)


def build():
    res = subprocess.run(["make", "-C", os.path.join(_HERE, "c")], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"oracle build failed:\n{res.stdout}\n{res.stderr}")
    return LIB_PATH


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_tt_eval_multi_batch.restype = C.c_int
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_f64p)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(_i32p)


def _l(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(_i64p)


def _sharded(n, threads, fn):
    """Run ``fn(lo, hi)`` over contiguous shards of range(n)."""
    threads = max(1, min(int(threads), n)) if n else 1
    if threads == 1:
        fn(0, n)
        return
    cuts = np.linspace(0, n, threads + 1).astype(np.int64)
    with ThreadPoolExecutor(threads) as pool:
        list(pool.map(lambda k: fn(int(cuts[k]), int(cuts[k + 1])), range(threads)))


class TT:
    def __init__(self, cores, domain, dim_order):
        self.D = len(cores)
        self.keep = [_i([c.shape[1] for c in cores]),
                     _i([c.shape[0] for c in cores] + [cores[-1].shape[2]]),
                     _f([d[0] for d in domain]), _f([d[1] for d in domain]), _i(dim_order),
                     _f(np.concatenate([np.asarray(c, dtype=np.float64).ravel() for c in cores]))]

    def _args(self):
        return [self.D] + [k[1] for k in self.keep]

    def eval_batch(self, points, threads=1):
        pts, _ = _f(points)
        out = np.empty(pts.shape[0])
        lib = load()

        def run(lo, hi):
            lib.orc_tt_eval_batch(*self._args(), pts[lo:hi].ctypes.data_as(_f64p),
                                  C.c_int64(hi - lo), out[lo:hi].ctypes.data_as(_f64p))

        _sharded(pts.shape[0], threads, run)
        return out

    def eval_multi_batch(self, points, orders, threads=1):
        pts, _ = _f(points)
        od, od_p = _i(np.asarray(orders).reshape(-1, self.D))
        G = od.shape[0]
        out = np.empty((pts.shape[0], G))
        lib = load()
        bad = []

        def run(lo, hi):
            rc = lib.orc_tt_eval_multi_batch(*self._args(), pts[lo:hi].ctypes.data_as(_f64p),
                                             C.c_int64(hi - lo), G, od_p,
                                             out[lo:hi].ctypes.data_as(_f64p))
            if rc:
                bad.append(rc)

        _sharded(pts.shape[0], threads, run)
        if bad:
            raise ValueError("Derivative order not supported (use 1 or 2)")
        return out


class Full:
    """G pre-differentiated tensors of one grid."""

    def __init__(self, n_nodes, nodes, weights, tensors):
        self.D = len(n_nodes)
        self.G = len(tensors)
        self.keep = [_i(n_nodes), _f(np.concatenate(nodes)), _f(np.concatenate(weights)),
                     _f(np.concatenate([np.ascontiguousarray(t, dtype=np.float64).ravel()
                                        for t in tensors]))]

    def eval_batch(self, points, threads=1):
        pts, _ = _f(points)
        out = np.empty((pts.shape[0], self.G))
        lib = load()
        n_p, nodes_p, w_p, t_p = (k[1] for k in self.keep)

        def run(lo, hi):
            lib.orc_full_eval_batch(self.D, n_p, nodes_p, w_p, self.G, t_p,
                                    pts[lo:hi].ctypes.data_as(_f64p), C.c_int64(hi - lo),
                                    out[lo:hi].ctypes.data_as(_f64p))

        _sharded(pts.shape[0], threads, run)
        return out


def spline_lookup(knots, points):
    pts, pts_p = _f(points)
    _, nk_p = _i([len(k) for k in knots])
    flat = [float(v) for k in knots for v in k] or [0.0]
    _, k_p = _f(flat)
    out = np.empty(pts.shape[0], dtype=np.int32)
    load().orc_spline_lookup(len(knots), nk_p, k_p, pts_p, C.c_int64(pts.shape[0]),
                             out.ctypes.data_as(_i32p))
    return out


def spline_eval_batch(knots, pieces, points, threads=1):
    """``pieces``: C-order list of ``(n_nodes, nodes, weights, [tensor_g...])``."""
    pts, _ = _f(points)
    D = len(knots)
    G = len(pieces[0][3])
    _, nk_p = _i([len(k) for k in knots])
    _, k_p = _f([float(v) for k in knots for v in k] or [0.0])
    _, pn_p = _i([list(p[0]) for p in pieces])
    node_off, tensor_off, no, to = [], [], 0, 0
    for p in pieces:
        node_off.append(no)
        tensor_off.append(to)
        no += int(np.sum(p[0]))
        to += int(np.prod(p[0])) * G
    _, no_p = _l(node_off)
    _, to_p = _l(tensor_off)
    _, nodes_p = _f(np.concatenate([a for p in pieces for a in p[1]]))
    _, w_p = _f(np.concatenate([a for p in pieces for a in p[2]]))
    _, t_p = _f(np.concatenate([np.ascontiguousarray(t, dtype=np.float64).ravel()
                                for p in pieces for t in p[3]]))
    out = np.empty((pts.shape[0], G))
    piece = np.empty(pts.shape[0], dtype=np.int32)
    lib = load()

    def run(lo, hi):
        lib.orc_spline_eval_batch(D, nk_p, k_p, len(pieces), pn_p, no_p, to_p, nodes_p, w_p, G, t_p,
                                  pts[lo:hi].ctypes.data_as(_f64p), C.c_int64(hi - lo),
                                  out[lo:hi].ctypes.data_as(_f64p),
                                  piece[lo:hi].ctypes.data_as(_i32p))

    _sharded(pts.shape[0], threads, run)
    return out, piece
