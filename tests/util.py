"""Seeded synthetic inputs shared by the tests (sparsified-domain data)."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def sample_rows(rng, p, n, m):
    """(m, n) distinct sorted rows per column, uniform without replacement."""
    keys = rng.random((n, p))
    rows = np.argpartition(keys, m - 1, axis=1)[:, :m]
    rows.sort(axis=1)
    return rows.T.copy()


def make_sparsified(p=64, n=500, m=8, K=5, seed=0, kind="mixture", f32=True, ragged=False):
    """Returns (X csc float64 (values fp32-representable if f32), centers p x K, gamma)."""
    rng = np.random.default_rng(seed)
    mu = rng.standard_normal((p, K))
    rows = sample_rows(rng, p, n, m)                       # (m, n)
    cols = np.repeat(np.arange(n), m)
    r = rows.T.reshape(-1)
    scale = p / m
    if kind == "mixture":
        lab = np.arange(n) % K
        vals = (mu[r, lab[cols]] + 0.1 * rng.standard_normal(n * m)) * scale
        centers = mu + 0.05 * rng.standard_normal((p, K))
    elif kind == "unstructured":
        vals = rng.standard_normal(n * m) * scale
        centers = 0.1 * rng.standard_normal((p, K))
    else:
        raise ValueError(kind)
    if ragged:                                             # drop a random subset of entries
        keep = rng.random(n * m) < 0.7
        keep[: 2 * m] = False                              # first two columns become empty
        r, cols, vals = r[keep], cols[keep], vals[keep]
    if f32:
        vals = vals.astype(np.float32).astype(np.float64)
    X = sp.csc_matrix((vals, (r, cols)), shape=(p, n))
    X.sort_indices()
    gamma = m / p
    return X, centers, gamma
