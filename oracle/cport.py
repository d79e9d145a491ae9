"""ctypes wrapper for oracle/skm_oracle.c -- TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from functools import lru_cache

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libskm_oracle.so")

_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i64 = C.c_int64


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "skm_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    return LIB_PATH


@lru_cache(maxsize=None)
def lib():
    build()
    L = C.CDLL(LIB_PATH)
    L.skmo_masked_dist.argtypes = [_i64, _i64, _i64, _i64p, _i64p, _f64p, _f64p, _f64p]
    L.skmo_masked_dist_beta.argtypes = [_i64, _i64, _i64p, _i64p, _f64p, _f64p, C.c_double, _f64p]
    L.skmo_colmin.argtypes = [_i64, _i64, _f64p, _f64p, _i64p]
    L.skmo_assign.argtypes = [_i64, _i64, _i64, _i64p, _i64p, _f64p, _f64p, _f64p, _i64p, C.c_int]
    L.skmo_inner_product.argtypes = [_i64, _i64p, _i64p, _f64p, _f64p, _f64p, _f64p]
    L.skmo_colnormsq.argtypes = [_i64, _i64p, _f64p, _f64p]
    L.skmo_hadamard.argtypes = [_i64, _i64, _f64p, _f64p]
    L.skmo_centroid_update.argtypes = [_i64, _i64, _i64, _i64p, _i64p, _f64p, _i64p, C.c_double,
                                       C.c_int, _f64p, _f64p, _i64p, _f64p]
    for f in ("skmo_masked_dist", "skmo_masked_dist_beta", "skmo_colmin", "skmo_assign",
              "skmo_inner_product", "skmo_colnormsq", "skmo_hadamard", "skmo_centroid_update"):
        getattr(L, f).restype = None
    return L


def _csc(jc, ir, x):
    return (np.ascontiguousarray(jc, dtype=np.int64), np.ascontiguousarray(ir, dtype=np.int64),
            np.ascontiguousarray(x, dtype=np.float64))


def _colmajor(a):
    """flat float64 buffer holding `a` (2-D) in column-major order."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    return np.ascontiguousarray(a.T).reshape(-1), a.shape


def masked_dist(p, n, jc, ir, x, centers):
    """K x n matrix of masked distances (SparseMatrixMinusCluster.c:169-182)."""
    jc, ir, x = _csc(jc, ir, x)
    c, (pp, K) = _colmajor(centers)
    assert pp == p
    out = np.empty(K * n, dtype=np.float64)
    lib().skmo_masked_dist(p, n, K, jc, ir, x, c, out)
    return out.reshape(n, K).T


def masked_dist_beta(p, n, jc, ir, x, c, beta):
    jc, ir, x = _csc(jc, ir, x)
    c = np.ascontiguousarray(np.asarray(c, dtype=np.float64).reshape(-1))
    out = np.empty(n, dtype=np.float64)
    lib().skmo_masked_dist_beta(p, n, jc, ir, x, c, float(beta), out)
    return out.reshape(1, n)


def colmin(D):
    """[dmin, assign(1-based)] = min(D,[],1) with MATLAB semantics."""
    D = np.asarray(D, dtype=np.float64)
    K, n = D.shape
    flat = np.ascontiguousarray(D.T).reshape(-1)
    dmin = np.empty(n, dtype=np.float64)
    a = np.empty(n, dtype=np.int64)
    lib().skmo_colmin(K, n, flat, dmin, a)
    return dmin, a


def assign(p, n, jc, ir, x, centers, threads=1):
    """Fused masked distance + argmin; returns (assign 1-based int64, dmin float64)."""
    jc, ir, x = _csc(jc, ir, x)
    c, (pp, K) = _colmajor(centers)
    assert pp == p
    dmin = np.empty(n, dtype=np.float64)
    a = np.empty(n, dtype=np.int64)
    lib().skmo_assign(p, n, K, jc, ir, x, c, dmin, a, int(threads))
    return a, dmin


def inner_product(n, jc, ir, x, c):
    jc, ir, x = _csc(jc, ir, x)
    c = np.ascontiguousarray(np.asarray(c, dtype=np.float64).reshape(-1))
    ip = np.empty(n, dtype=np.float64)
    n2 = np.empty(n, dtype=np.float64)
    lib().skmo_inner_product(n, jc, ir, x, c, ip, n2)
    return ip, n2


def colnormsq(n, jc, x):
    jc = np.ascontiguousarray(jc, dtype=np.int64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    n2 = np.empty(n, dtype=np.float64)
    lib().skmo_colnormsq(n, jc, x, n2)
    return n2


def hadamard(X):
    """Unnormalised Sylvester-ordered WHT of each column (hadamard.c:57-92)."""
    flat, (m, n) = _colmajor(X)
    assert m >= 2 and (m & (m - 1)) == 0, "rows must be a power of two >= 2"
    out = np.empty_like(flat)
    lib().skmo_hadamard(m, n, flat, out)
    return out.reshape(n, m).T


def centroid_update(p, n, K, jc, ir, x, assign1, gamma, centers_in, ml_correction=True):
    """ML-corrected centre update (kmeans_sparsified.m:430-453).

    Returns (centers, S, N, counts); columns of empty clusters keep centers_in."""
    jc, ir, x = _csc(jc, ir, x)
    a = np.ascontiguousarray(assign1, dtype=np.int64)
    S = np.empty(p * K, dtype=np.float64)
    N = np.empty(p * K, dtype=np.float64)
    counts = np.empty(K, dtype=np.int64)
    cen, _ = _colmajor(centers_in)
    cen = cen.copy()
    lib().skmo_centroid_update(p, n, K, jc, ir, x, a, float(gamma), int(bool(ml_correction)),
                               S, N, counts, cen)
    return cen.reshape(K, p).T.copy(), S.reshape(K, p).T, N.reshape(K, p).T, counts
