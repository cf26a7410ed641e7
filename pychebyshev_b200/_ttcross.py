"""Tensor-train cross approximation on a Chebyshev grid (build-time, host NumPy).

Alternating one-site cross interpolation: sweep left-to-right and right-to-left over the
dimensions; at site k sample the function on ``(left index set) x (all nodes of dim k) x (right
index set)``, orthogonalise the sampled unfolding with a QR factorisation and pick the next
index set as a (quasi-)maximum-volume row subset of the Q factor.  After the sweeps the train is
rounded (QR sweep + truncated SVDs) to the requested tolerance, which also trims the ranks.
Function values are memoised, so the number of *unique* evaluations is what is reported.

This is the construction step *before* the evaluation hot path; it only has
to deliver value cores of shape ``(r_{k-1}, n_k, r_k)`` on the ascending node grids.
"""

from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import numpy as np
from scipy.linalg import qr as pivoted_qr


def maxvol_rows(a: np.ndarray, tol: float = 1.02, max_iters: int = 200) -> np.ndarray:
    """Indices of ``r`` rows of the tall matrix ``a`` (n x r) whose submatrix has quasi-maximal
    volume: start from a pivoted-QR guess and swap rows while some entry of ``a @ inv(a[rows])``
    exceeds ``tol`` in magnitude."""
    n, r = a.shape
    if n <= r:
        return np.arange(n)
    _, _, piv = pivoted_qr(a.T, mode="economic", pivoting=True)
    rows = np.array(piv[:r])
    coef = np.linalg.solve(a[rows].T, a.T).T  # a @ inv(a[rows])
    for _ in range(max_iters):
        flat = int(np.argmax(np.abs(coef)))
        i, j = divmod(flat, r)
        if abs(coef[i, j]) <= tol:
            break
        rows[j] = i
        # rank-one update of coef after replacing row j of the submatrix by row i
        col = coef[:, j].copy()
        row = coef[i, :].copy()
        row[j] -= 1.0
        coef -= np.outer(col, row) / coef[i, j]
    return rows


def _round_tt(cores: List[np.ndarray], tol: float, max_rank: int) -> List[np.ndarray]:
    """Left-orthogonalise, then truncate right-to-left with relative singular-value cutoff."""
    cores = [c.copy() for c in cores]
    d = len(cores)
    for k in range(d - 1):
        r0, n, r1 = cores[k].shape
        q, rr = np.linalg.qr(cores[k].reshape(r0 * n, r1))
        cores[k] = q.reshape(r0, n, q.shape[1])
        cores[k + 1] = np.tensordot(rr, cores[k + 1], axes=([1], [0]))
    for k in range(d - 1, 0, -1):
        r0, n, r1 = cores[k].shape
        u, s, vt = np.linalg.svd(cores[k].reshape(r0, n * r1), full_matrices=False)
        keep = min(max_rank, len(s))
        if s[0] > 0:
            keep = max(1, min(keep, int(np.sum(s > tol * s[0]))))
        cores[k] = vt[:keep].reshape(keep, n, r1)
        cores[k - 1] = np.tensordot(cores[k - 1], u[:, :keep] * s[:keep], axes=([2], [0]))
    return cores


def tt_cross(func: Callable[[Sequence[float]], float], grids: List[np.ndarray], max_rank: int,
             tol: float, max_sweeps: int, seed=None) -> Tuple[List[np.ndarray], int]:
    """Value cores of a rank-``<= max_rank`` TT approximation of ``func`` on ``grids``.

    Returns ``(cores, unique_function_evaluations)``."""
    rng = np.random.default_rng(seed)
    d = len(grids)
    n = [len(g) for g in grids]
    cache = {}

    def sample(idx_rows: np.ndarray) -> np.ndarray:
        out = np.empty(len(idx_rows))
        for t, row in enumerate(idx_rows):
            key = tuple(int(v) for v in row)
            val = cache.get(key)
            if val is None:
                val = float(func([float(grids[k][key[k]]) for k in range(d)]))
                cache[key] = val
            out[t] = val
        return out

    if d == 1:
        vals = sample(np.arange(n[0])[:, None])
        return [vals.reshape(1, n[0], 1)], len(cache)

    # target ranks: capped by max_rank and by the sizes of the two unfoldings
    ranks = [1] * (d + 1)
    for k in range(1, d):
        left = int(np.prod(n[:k], dtype=np.float64).clip(max=1e9))
        right = int(np.prod(n[k:], dtype=np.float64).clip(max=1e9))
        ranks[k] = int(min(max_rank, left, right))
    for k in range(1, d):  # a bond cannot exceed its neighbour times the mode size
        ranks[k] = min(ranks[k], ranks[k - 1] * n[k - 1])
    for k in range(d - 1, 0, -1):
        ranks[k] = min(ranks[k], ranks[k + 1] * n[k])

    # right index sets J[k]: (ranks[k], d-k) multi-indices over dims k..d-1, random start
    right_sets = [None] * (d + 1)
    right_sets[d] = np.zeros((1, 0), dtype=np.int64)
    for k in range(1, d):
        right_sets[k] = np.stack([rng.integers(0, n[m], size=ranks[k]) for m in range(k, d)], axis=1)
    left_sets = [None] * (d + 1)
    left_sets[0] = np.zeros((1, 0), dtype=np.int64)

    def site_block(k: int) -> np.ndarray:
        """Samples on left_sets[k] x nodes_k x right_sets[k+1] as (rl, n_k, rr)."""
        lset, rset = left_sets[k], right_sets[k + 1]
        rl, rr = len(lset), len(rset)
        idx = np.empty((rl, n[k], rr, d), dtype=np.int64)
        idx[..., :k] = lset[:, None, None, :]
        idx[..., k] = np.arange(n[k])[None, :, None]
        idx[..., k + 1:] = rset[None, None, :, :]
        return sample(idx.reshape(-1, d)).reshape(rl, n[k], rr)

    cores = [None] * d
    prev_norm = None
    for sweep in range(max(1, max_sweeps)):
        # left-to-right: fix left index sets
        for k in range(d - 1):
            block = site_block(k)
            rl, _, rr = block.shape
            q, _ = np.linalg.qr(block.reshape(rl * n[k], rr))
            rows = maxvol_rows(q)
            cores[k] = np.linalg.solve(q[rows].T, q.T).T.reshape(rl, n[k], q.shape[1])
            li, ni = np.divmod(rows, n[k])
            left_sets[k + 1] = np.concatenate([left_sets[k][li], ni[:, None]], axis=1)
        cores[d - 1] = site_block(d - 1)
        # right-to-left: fix right index sets
        for k in range(d - 1, 0, -1):
            block = site_block(k)
            rl, _, rr = block.shape
            q, _ = np.linalg.qr(block.reshape(rl, n[k] * rr).T)
            rows = maxvol_rows(q)
            cores[k] = np.linalg.solve(q[rows].T, q.T).reshape(q.shape[1], n[k], rr)
            ni, ri = np.divmod(rows, rr)
            right_sets[k] = np.concatenate([ni[:, None], right_sets[k + 1][ri]], axis=1)
        cores[0] = site_block(0)
        norm = float(np.sqrt(sum(np.sum(c * c) for c in cores)))
        if prev_norm is not None and abs(norm - prev_norm) <= tol * max(norm, 1e-300):
            break
        prev_norm = norm
    return _round_tt(cores, tol, max_rank), len(cache)
