"""Synthetic workloads of BASELINE.json's configs (tests, golden fixtures, bench).

Everything here is *input generation*: analytic pricing functions evaluated on
Chebyshev grids, domains and seeded query sets.  Nothing here evaluates an
interpolant.  The domains / node counts follow the reference's own fixtures
(reference ``tests/conftest.py:86-172``) and SURVEY.md §8(d).
"""

from __future__ import annotations

import math

import numpy as np

# --- C1 / C2: 5D Black-Scholes call V(S, K, T, sigma, r), dividend q = 0.02 ----------

BS5D_DOMAIN = [[80.0, 120.0], [90.0, 110.0], [0.25, 1.0], [0.15, 0.35], [0.01, 0.08]]
BS5D_NODES = [11, 11, 11, 11, 11]
BS5D_Q = 0.02

#: "price + Greeks" of BASELINE.json config 1: price, delta, gamma, vega
BS5D_GREEKS = [[0, 0, 0, 0, 0], [1, 0, 0, 0, 0], [2, 0, 0, 0, 0], [0, 0, 0, 1, 0]]


def _ndtr(x):
    from scipy.special import ndtr

    return ndtr(x)


def bs_call_price(S, K, T, sigma, r, q=BS5D_Q):
    """Black-Scholes call price, vectorised over NumPy broadcasting."""
    S, K, T, sigma, r = (np.asarray(v, dtype=np.float64) for v in (S, K, T, sigma, r))
    srt = sigma * np.sqrt(T)
    d1 = (np.log(S / K) + (r - q + 0.5 * sigma * sigma) * T) / srt
    d2 = d1 - srt
    return S * np.exp(-q * T) * _ndtr(d1) - K * np.exp(-r * T) * _ndtr(d2)


def bs5d_scalar(x, _=None):
    """Scalar ``f(point, additional_data)`` form used by ``build()``."""
    S, K, T, sigma, r = x
    srt = sigma * math.sqrt(T)
    d1 = (math.log(S / K) + (r - BS5D_Q + 0.5 * sigma * sigma) * T) / srt
    d2 = d1 - srt
    nd1 = 0.5 * math.erfc(-d1 / math.sqrt(2.0))
    nd2 = 0.5 * math.erfc(-d2 / math.sqrt(2.0))
    return S * math.exp(-BS5D_Q * T) * nd1 - K * math.exp(-r * T) * nd2


def grid_values(func_vec, nodes):
    """Evaluate a broadcasting function on the tensor grid of ``nodes`` (C-order).

    Uses sparse open meshes so the 16^6 grid never materialises 6 full arrays.
    """
    mesh = np.meshgrid(*nodes, indexing="ij", sparse=True)
    out = func_vec(*mesh)
    shape = tuple(len(n) for n in nodes)
    return np.ascontiguousarray(np.broadcast_to(out, shape), dtype=np.float64)


# --- C3: spline with a kink at the strike ------------------------------------------------

SPLINE2D_DOMAIN = [[80.0, 120.0], [0.25, 1.0]]
SPLINE2D_NODES = [15, 15]
SPLINE2D_KNOTS = [[100.0], []]


def payoff2d(S, T):
    return np.maximum(np.asarray(S, dtype=np.float64) - 100.0, 0.0) * np.exp(-0.05 * np.asarray(T))


SPLINE3D_DOMAIN = [[80.0, 120.0], [0.25, 1.0], [0.01, 0.08]]
SPLINE3D_NODES = [15, 15, 15]
SPLINE3D_KNOTS = [[100.0], [], []]


def payoff3d(S, T, r):
    return np.maximum(np.asarray(S, dtype=np.float64) - 100.0, 0.0) * np.exp(
        -np.asarray(r) * np.asarray(T)
    )


# --- C4: 6D full tensor 16^6 ----------------------------------------------------------------

C4_DOMAIN = [[80.0, 120.0], [90.0, 110.0], [0.25, 1.0], [0.15, 0.35], [0.01, 0.08], [0.0, 0.04]]
C4_NODES = [16] * 6
C4_GREEKS = [[0] * 6, [1, 0, 0, 0, 0, 0], [2, 0, 0, 0, 0, 0], [0, 0, 0, 1, 0, 0]]


def bs6d(S, K, T, sigma, r, q):
    """6D Black-Scholes call with the dividend yield as sixth coordinate."""
    return bs_call_price(S, K, T, sigma, r, q=np.asarray(q, dtype=np.float64))


# --- C5: 10D basket ---------------------------------------------------------------------------

C5_DIM = 10
C5_DOMAIN = [[80.0, 120.0]] * C5_DIM
C5_NODES = [11] * C5_DIM
C5_PARTITION = [[0, 1], [2, 3], [4, 5], [6, 7], [8, 9]]
C5_PIVOT = [100.0] * C5_DIM


def basket10d_scalar(x, _=None):
    """Smooth basket call: softplus of (mean spot - 100), width 5."""
    m = sum(x) / len(x) - 100.0
    z = m / 5.0
    return 5.0 * (math.log1p(math.exp(-abs(z))) + max(z, 0.0))


def uniform_queries(domain, n, seed):
    """``np.random.default_rng(seed).uniform(lo_d, hi_d, n)`` per dim, column-stacked."""
    rng = np.random.default_rng(seed)
    cols = [rng.uniform(lo, hi, n) for lo, hi in domain]
    return np.ascontiguousarray(np.stack(cols, axis=1))


def synthetic_tt_cores(n_nodes, ranks, seed):
    """Random coefficient cores with controlled ranks."""
    rng = np.random.default_rng(seed)
    cores = []
    for k, n in enumerate(n_nodes):
        r0, r1 = ranks[k], ranks[k + 1]
        g = rng.standard_normal((r0, n, r1)) / math.sqrt(r0 * n)
        # decaying Chebyshev spectrum so the interpolant is a smooth function
        g *= (0.6 ** np.arange(n))[None, :, None]
        cores.append(np.ascontiguousarray(g))
    return cores
