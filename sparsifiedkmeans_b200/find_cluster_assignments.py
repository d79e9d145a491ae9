"""findClusterAssignments: the operator surface of the reference, sparse-X branches on the GPU.

    [assignments, distances, centers] = findClusterAssignments(X, centers, tryBuiltinMex, gamma)
                                                  (private/findClusterAssignments.m:1-33)

X is p x n with points as COLUMNS, `centers` is p x K.  Assignments are 1-based, distances
Euclidean (not squared).  X may be a scipy.sparse matrix (uploaded for the call) or a resident
`engine.Dataset` (what kmeans_sparsified keeps across iterations).
"""
from __future__ import annotations

import numpy as np

from .engine import Context, Dataset, default_context


def findClusterAssignments(X, centers, tryBuiltinMex=None, gamma=None, nargout: int = 2,
                           store: str = "f64", ctx: Context | None = None):
    """Sparse X, dense centres -> private/findClusterAssignments.m:77-80 then :168-171;
    sparse X, sparse centres (scipy sparse `centers`) -> :63-75.  Dense X -> :124-171 (plain
    Euclidean nearest centre, gamma unused there), evaluated on the GPU as sum (x-c)^2 in fp32 with
    fp64 re-evaluation of uncertified columns (skm_second_pass); the reference's closed-source
    pdist2 / BLAS arithmetic is not reproducible bit for bit, ties below ~1e-12 relative are unpinned.

    `store` selects how an uploaded X is held: "f64" (default for this stateless call: every
    bit as the reference) or "f32" (the fast kernel; values are rounded to float first).
    """
    import scipy.sparse as sp
    del tryBuiltinMex                                   # only affects the dense branch (:135-152)
    owns = False
    if isinstance(X, Dataset):
        ds = X
    elif sp.issparse(X):
        ds = Dataset.from_scipy(X, store=store, ctx=ctx or default_context())
        owns = True
    else:
        return _dense(np.asarray(X), centers, nargout, ctx or default_context())
    try:
        sparse_centers = sp.issparse(centers)
        c = np.asarray(centers.todense() if sparse_centers else centers, dtype=np.float64)
        if c.ndim == 1:
            c = c.reshape(-1, 1)
        if c.shape[0] != ds.p:
            raise ValueError("Array of centers not of correct size")             # :55
        if sparse_centers:
            assignments, distances = ds.assign_sparse_centers(c, gamma)
        else:
            assignments, distances = ds.assign(c, gamma)
        if nargout < 3:
            return assignments, distances
        # third output: plain mean of the members, zero for empty clusters (:178-188)
        K = c.shape[1]
        cols = _dense_columns(X, ds)
        out = np.zeros((ds.p, K))
        for ki in range(K):
            ind = np.flatnonzero(assignments == ki + 1)
            if ind.size:
                out[:, ki] = np.asarray(cols[:, ind].mean(axis=1)).ravel()
        return assignments, distances, out
    finally:
        if owns:
            ds.close()


def _dense(X, centers, nargout, ctx):
    """Dense-X branch (:124-171) and the optional mean-centres output (:178-188)."""
    import scipy.sparse as sp
    from .engine import second_pass
    c = np.asarray(centers.todense() if sp.issparse(centers) else centers, dtype=np.float64)
    if c.ndim == 1:
        c = c.reshape(-1, 1)
    if X.ndim == 1:
        X = X.reshape(-1, 1)
    if c.shape[0] != X.shape[0]:
        raise ValueError("Array of centers not of correct size")                 # :55
    res = second_pass(X, centers=c, ctx=ctx)
    a, d = res["assign"], res["dist"]
    if nargout < 3:
        return a, d
    K = c.shape[1]
    out = np.zeros((X.shape[0], K))
    if X.shape[1]:
        m = second_pass(X, assign_in=a, want_assign=False, want_dist=False, ctx=ctx)["centers"]
        out[:, :m.shape[1]] = m
    return a, d, out


def _dense_columns(X, ds):
    import scipy.sparse as sp
    if sp.issparse(X):
        return sp.csc_matrix(X)
    return sp.csc_matrix(np.stack([ds.get_column(j) for j in range(ds.n)], axis=1))
