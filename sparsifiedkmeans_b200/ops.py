"""Level-1 operators with the reference's MEX names and calling conventions, running on the GPU
in IEEE double in the reference's operation order (bit-identical results).

    dist = SparseMatrixMinusCluster(X, c[, beta])      private/SparseMatrixMinusCluster.c:2-11
    [ip, n2] = SparseMatrixInnerProduct(X, c)          private/SparseMatrixInnerProduct.c:2-9
    n2 = SparseMatrixColumnNormSq(X)                   private/SparseMatrixColumnNormSq.c:2-8
    w = hadamard(x); w = hadamard_pthreads(x)          private/hadamard.c:8-24

X is a scipy.sparse matrix of shape (p, n) (points are columns).  Errors mirror the MEX
gateways' usage errors (raised as ValueError / SkmError instead of mexErrMsgTxt).
"""
from __future__ import annotations

import numpy as np

from ._lib import check
from .engine import Context, _ptr, default_context


def _csc_u64(X):
    import scipy.sparse as sp
    if not sp.issparse(X):
        raise ValueError("Input matrix must be a sparse matrix")        # SparseMatrixMinusCluster.c:62-65
    if np.iscomplexobj(X):
        raise ValueError("Cannot handle complex data yet")               # :66-69
    X = sp.csc_matrix(X, dtype=np.float64)
    X.sort_indices()
    return (X.shape[0], X.shape[1], np.ascontiguousarray(X.indptr, dtype=np.uint64),
            np.ascontiguousarray(X.indices, dtype=np.uint64), np.ascontiguousarray(X.data, dtype=np.float64))


def SparseMatrixMinusCluster(X, c, beta=None, ctx: Context | None = None) -> np.ndarray:
    ctx = ctx or default_context()
    p, n, jc, ir, pr = _csc_u64(X)
    c = np.asarray(c, dtype=np.float64)
    if c.ndim == 1:
        c = c.reshape(-1, 1)
    if c.shape[0] != p:
        raise ValueError("Center vector must be or pxk, but this vector did not have p rows")   # :104-107
    K = c.shape[1]
    cf = np.ascontiguousarray(c.T).reshape(-1)
    out = np.empty(K * n, dtype=np.float64)
    check(ctx._lib.skm_sparse_matrix_minus_cluster(ctx.handle, p, n, K, _ptr(jc), _ptr(ir), _ptr(pr), _ptr(cf),
                                                   int(beta is not None), float(beta or 0.0), _ptr(out)))
    return out.reshape(n, K).T


def SparseMatrixInnerProduct(X, c, ctx: Context | None = None):
    ctx = ctx or default_context()
    p, n, jc, ir, pr = _csc_u64(X)
    c = np.ascontiguousarray(c, dtype=np.float64).reshape(-1)
    if c.shape[0] != p:
        raise ValueError("Center vector must have one entry per row of X")
    ip = np.empty(n, dtype=np.float64)
    n2 = np.empty(n, dtype=np.float64)
    check(ctx._lib.skm_sparse_matrix_inner_product(ctx.handle, p, n, _ptr(jc), _ptr(ir), _ptr(pr), _ptr(c),
                                                   _ptr(ip), _ptr(n2)))
    return ip.reshape(1, n), n2.reshape(1, n)


def SparseMatrixColumnNormSq(X, ctx: Context | None = None) -> np.ndarray:
    ctx = ctx or default_context()
    p, n, jc, _, pr = _csc_u64(X)
    n2 = np.empty(n, dtype=np.float64)
    check(ctx._lib.skm_sparse_matrix_column_normsq(ctx.handle, p, n, _ptr(jc), _ptr(pr), _ptr(n2)))
    return n2.reshape(1, n)


def hadamard(x, ctx: Context | None = None) -> np.ndarray:
    """Unnormalised Sylvester-ordered WHT of each column; rows must be a power of two >= 2."""
    ctx = ctx or default_context()
    import scipy.sparse as sp
    if sp.issparse(x):
        raise ValueError("Input must be a full matrix, not sparse")       # hadamard.c:97-111
    if np.iscomplexobj(x):
        raise ValueError("Input must be real")
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x.reshape(-1, 1)
    m, n = x.shape
    if m < 2 or (m & (m - 1)) != 0:
        raise ValueError("Number of rows must be a power of 2")            # hadamard.c:134-140
    xf = np.ascontiguousarray(x.T).reshape(-1)
    out = np.empty_like(xf)
    check(ctx._lib.skm_hadamard(ctx.handle, m, n, _ptr(xf), _ptr(out)))
    return out.reshape(n, m).T


def hadamard_pthreads(x, ctx: Context | None = None) -> np.ndarray:
    """Same transform as `hadamard` (the reference's pthreads build is bit-identical to its
    serial one; on the GPU the column parallelism is the grid)."""
    return hadamard(x, ctx)
