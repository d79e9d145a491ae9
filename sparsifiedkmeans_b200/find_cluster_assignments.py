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
    sparse X, sparse centres (scipy sparse `centers`) -> :63-75.  Dense X (:124-166) is the
    non-sparsified K-means path and is out of scope for this engine.

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
        raise NotImplementedError(
            "findClusterAssignments: dense X (findClusterAssignments.m:124-166) is outside the "
            "sparsified hot path this engine accelerates")
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


def _dense_columns(X, ds):
    import scipy.sparse as sp
    if sp.issparse(X):
        return sp.csc_matrix(X)
    return sp.csc_matrix(np.stack([ds.get_column(j) for j in range(ds.n)], axis=1))
