"""Handle wrappers over the level-2 C ABI: Context, Dataset (resident sparsified matrix) and
Lloyd (per-K iteration state).  Everything numeric happens inside libskm_b200.so; this
module only marshals numpy buffers (host) or raw device pointers.
"""
from __future__ import annotations

import ctypes as C
import weakref
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import SKM_F32, SKM_F64, SKM_I32, SKM_I64, SKM_U16, check

_NP_INDEX = {np.dtype(np.int32): SKM_I32, np.dtype(np.int64): SKM_I64, np.dtype(np.uint64): SKM_I64}
_NP_VALUE = {np.dtype(np.float32): SKM_F32, np.dtype(np.float64): SKM_F64}
_NP_ROWS = {**_NP_INDEX, np.dtype(np.uint16): SKM_U16}      # row indices may also be uint16 (p <= 65536)


def _ptr(a) -> int:
    return None if a is None else a.ctypes.data


def _centers(a, rows: int):
    """(rows, K) array -> (contiguous column-major float64 copy (flat), K)."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    if a.ndim != 2 or a.shape[0] != rows:
        raise ValueError("Array of centers not of correct size")   # findClusterAssignments.m:55
    if a.shape[1] < 1:
        raise ValueError("need at least one centre")
    return np.ascontiguousarray(a.T).reshape(-1), int(a.shape[1])


class Context:
    """One per process / GPU (skm_ctx)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self._lib = _lib.load()
        h = C.c_void_p()
        check(self._lib.skm_ctx_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h)))
        self._h = h
        self.device = int(device)

    @property
    def handle(self):
        if self._h is None:
            raise RuntimeError("context destroyed")
        return self._h

    @property
    def stream(self) -> int:
        return int(self._lib.skm_ctx_stream(self.handle) or 0)

    @property
    def launch_count(self) -> int:
        return int(self._lib.skm_ctx_launch_count(self.handle))

    def synchronize(self):
        check(self._lib.skm_ctx_sync(self.handle))

    def tc_chunks(self) -> tuple[int, int]:
        """(chunks of second_pass handled by the tensor-core filter, chunks after which it was switched off)."""
        a, b = C.c_int64(), C.c_int64()
        check(self._lib.skm_ctx_tc_chunks(self.handle, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    TIMING_SLOTS = ("assign", "recheck", "accumulate", "finalize", "prep", "fwht", "kpp", "upload")

    def timing_enable(self, on: bool = True):
        check(self._lib.skm_ctx_timing_enable(self.handle, int(on)))

    def timing_read(self) -> dict:
        """{slot: (milliseconds, launch groups)} since the previous read (synchronises)."""
        ms = np.zeros(8, dtype=np.float64)
        cnt = np.zeros(8, dtype=np.int64)
        check(self._lib.skm_ctx_timing_read(self.handle, _ptr(ms), _ptr(cnt)))
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(self.TIMING_SLOTS)}

    def close(self):
        if self._h is not None:
            self._lib.skm_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx: dict[int, Context] = {}


def default_context(device: int = 0) -> Context:
    ctx = _default_ctx.get(device)
    if ctx is None:
        ctx = _default_ctx[device] = Context(device)
    return ctx


@dataclass
class IterStats:
    dff: float
    sumsq: float
    n_empty: int
    n_rechecked: int
    n_points: int
    has_nan: bool

    @property
    def objective(self) -> float:          # kmeans_sparsified.m:471
        return float(np.sqrt(self.sumsq))


class Dataset:
    """A p x n sparsified matrix resident in HBM (skm_dataset). Points are columns."""

    def __init__(self, ctx: Context, handle):
        self.ctx = ctx
        self._lib = ctx._lib
        self._h = handle
        self._children = weakref.WeakSet()       # Lloyd states bound to this dataset
        info = _lib.DatasetInfo()
        check(self._lib.skm_dataset_get_info(handle, C.byref(info)))
        self.p, self.n, self.nnz = int(info.p), int(info.n), int(info.nnz)
        self.max_col_nnz = int(info.max_col_nnz)
        self.store_dtype = "f32" if info.store_dtype == SKM_F32 else "f64"
        self.device_bytes = int(info.device_bytes)
        self.stream_bytes = int(info.stream_bytes)

    # -- construction -----------------------------------------------------
    @classmethod
    def from_csc(cls, p: int, n: int, jc, ir, val, store: str = "f32", ctx: Context | None = None, K_hint: int = 0):
        """Upload host CSC arrays (any of int32/int64/uint64 indices, float32/float64 values).  K_hint: the number of
        centres the caller will use; large uploads then build that kernel family's entry order while they are in flight."""
        ctx = ctx or default_context()
        jc = np.ascontiguousarray(jc)
        ir = np.ascontiguousarray(ir)
        val = np.ascontiguousarray(val)
        if jc.dtype not in _NP_INDEX:
            jc = jc.astype(np.int64)
        if ir.dtype not in _NP_ROWS:
            ir = ir.astype(np.int64)
        if val.dtype not in _NP_VALUE:
            val = val.astype(np.float64)
        if jc.shape[0] != n + 1:
            raise ValueError("jc must have n+1 entries")
        h = C.c_void_p()
        check(ctx._lib.skm_dataset_create_csc_hint(
            ctx.handle, p, n, _ptr(jc), _NP_INDEX[jc.dtype], _ptr(ir), _NP_ROWS[ir.dtype],
            _ptr(val), _NP_VALUE[val.dtype], SKM_F32 if store == "f32" else SKM_F64, 0, int(K_hint), C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def from_scipy(cls, X, store: str = "f32", ctx: Context | None = None, K_hint: int = 0):
        """X: scipy.sparse matrix of shape (p, n), points as columns."""
        import scipy.sparse as sp
        X = sp.csc_matrix(X)
        X.sort_indices()
        return cls.from_csc(X.shape[0], X.shape[1], X.indptr, X.indices, X.data, store, ctx, K_hint)

    @classmethod
    def from_device_csc(cls, p: int, n: int, jc_ptr: int, jc_type: int, ir_ptr: int, ir_type: int,
                        val_ptr: int, val_type: int, store: str = "f32", ctx: Context | None = None):
        """Adopt CSC arrays that already live on the GPU (copied into the library's layout)."""
        ctx = ctx or default_context()
        h = C.c_void_p()
        check(ctx._lib.skm_dataset_create_csc(
            ctx.handle, p, n, jc_ptr, jc_type, ir_ptr, ir_type, val_ptr, val_type,
            SKM_F32 if store == "f32" else SKM_F64, 1, C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def alloc_csc(cls, p: int, n: int, nnz: int, ctx: Context | None = None):
        """In-place production (skm_dataset_alloc_csc): returns (pending, colptr_ptr, rowidx_ptr, val_ptr); the
        caller's kernels fill the device arrays (int64[n+1], int32[nnz], float32[nnz]) and then call
        `pending.commit()` to get the Dataset."""
        ctx = ctx or default_context()
        h = C.c_void_p()
        check(ctx._lib.skm_dataset_alloc_csc(ctx.handle, int(p), int(n), int(nnz), C.byref(h)))
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(ctx._lib.skm_dataset_csc_ptrs(h, C.byref(a), C.byref(b), C.byref(c)))

        class _Pending:
            def __init__(self):
                self.h = h

            def commit(self_inner):
                if self_inner.h is None:
                    raise RuntimeError("already committed or released")
                hh, self_inner.h = self_inner.h, None
                try:
                    check(ctx._lib.skm_dataset_commit(hh))
                except Exception:
                    ctx._lib.skm_dataset_destroy(hh)
                    raise
                return cls(ctx, hh)

            def release(self_inner):
                if self_inner.h is not None:
                    ctx._lib.skm_dataset_destroy(self_inner.h)
                    self_inner.h = None
        return _Pending(), int(a.value or 0), int(b.value or 0), int(c.value or 0)

    @classmethod
    def from_fwht_sample(cls, p2: int, n: int, m: int, x_ptr: int, signs_ptr: int, rows_ptr: int | None,
                         ctx: Context | None = None, seed: int = 0, col0: int = 0):
        """Fused precondition + row sample on the device (skm_fwht_sample_f32): x is a dense
        p2 x n float32 column-major device matrix, signs float32[p2], rows int32[m*n] (column j
        keeps rows[m*j:m*j+m]); rows_ptr=None draws the rows on the device from (seed, col0 + j).
        The result is resident and ready for Lloyd iterations."""
        ctx = ctx or default_context()
        h = C.c_void_p()
        check(ctx._lib.skm_fwht_sample_f32(ctx.handle, p2, n, m, x_ptr, signs_ptr, rows_ptr, int(seed), int(col0),
                                           C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def from_dense_host(cls, X, signs, m: int, seed: int = 0, col0: int = 0, chunk_cols: int = 0,
                        ctx: Context | None = None):
        """Precondition + sample a dense host matrix X (p x n, points are columns; float32 or
        float64) on the GPU: zero-pad to p2 = len(signs), *(1+2eps), sign flip, FWHT, /sqrt(p2),
        keep m rows per column drawn on the device from (seed, col0 + j), divide by m/p2."""
        ctx = ctx or default_context()
        X = np.asarray(X)
        if X.dtype not in _NP_VALUE:
            X = X.astype(np.float64)
        p, n = X.shape
        Xf = np.ascontiguousarray(X.T).reshape(-1)                 # column-major p x n
        d = np.ascontiguousarray(signs, dtype=np.float64).reshape(-1)
        h = C.c_void_p()
        check(ctx._lib.skm_dataset_from_dense_host(ctx.handle, p, d.shape[0], n, _ptr(Xf), _NP_VALUE[Xf.dtype], _ptr(d),
                                                   int(m), int(seed), int(col0), int(chunk_cols), C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def from_dense_host_dct(cls, X, signs, m: int, seed: int = 0, col0: int = 0, rows=None, chunk_cols: int = 0,
                            ctx: Context | None = None):
        """DCT-sketch twin of from_dense_host (skm_dataset_from_dense_host_dct): X dense (p, n), points
        are columns; keeps m rows per column of dct(D*X*(1+2eps)) / (m/p).  `rows` (m, n) optional
        explicit 0-based rows; otherwise drawn on the device from (seed, col0 + column)."""
        ctx = ctx or default_context()
        X = np.asarray(X)
        if X.dtype not in _NP_VALUE:
            X = X.astype(np.float64)
        p, n = X.shape
        Xf = np.ascontiguousarray(X.T).reshape(-1)
        d = np.ascontiguousarray(signs, dtype=np.float64).reshape(-1)
        if d.shape[0] != p:
            raise ValueError("signs must have p entries")
        r = None
        if rows is not None:
            r = np.ascontiguousarray(np.asarray(rows, dtype=np.int32).T).reshape(-1)      # column j at [j*m, j*m+m)
            if r.shape[0] != m * n:
                raise ValueError("rows must be (m, n)")
        h = C.c_void_p()
        check(ctx._lib.skm_dataset_from_dense_host_dct(ctx.handle, p, n, _ptr(Xf), _NP_VALUE[Xf.dtype], _ptr(d), int(m),
                                                       int(seed), int(col0), _ptr(r), int(chunk_cols), C.byref(h)))
        return cls(ctx, h)

    def to_scipy(self):
        """Download as a scipy CSC matrix (float64 values) -- for tests and small data."""
        import scipy.sparse as sp
        cols = [self.get_column(j) for j in range(self.n)]
        return sp.csc_matrix(np.stack(cols, axis=1)) if cols else sp.csc_matrix((self.p, 0))

    @property
    def handle(self):
        if self._h is None:
            raise RuntimeError("dataset destroyed")
        return self._h

    def close(self):
        if self._h is not None:
            for child in list(self._children):   # a Lloyd state must not outlive its dataset
                child.close()
            self._lib.skm_dataset_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- operators ----------------------------------------------------------
    def layout_check(self, layout: int = -1) -> dict:
        """Verify the streamed (SELL-32) image against the CSC matrix and count the shared-memory
        wavefronts its entry order costs; layout 0/1 first re-orders it for that kernel family."""
        out = np.zeros(4, dtype=np.int64)
        check(self._lib.skm_dataset_layout_check(self.handle, int(layout), _ptr(out)))
        return {"bad_columns": int(out[0]), "steps": int(out[1]), "wavefronts": int(out[2]), "layout": int(out[3])}

    def get_column(self, j: int) -> np.ndarray:
        out = np.empty(self.p, dtype=np.float64)
        check(self._lib.skm_dataset_get_column(self.handle, int(j), _ptr(out)))
        return out

    def assign(self, centers, gamma=None, want_dist: bool = True):
        """findClusterAssignments, dense centres: returns (assign 1-based int32, dist float64)."""
        c, K = _centers(centers, self.p)
        a = np.empty(self.n, dtype=np.int32)
        d = np.empty(self.n, dtype=np.float64) if want_dist else None
        check(self._lib.skm_assign(self.handle, _ptr(c), K, int(gamma is not None),
                                   float(gamma if gamma is not None else 0.0), _ptr(a), _ptr(d)))
        return a, d

    def assign_sparse_centers(self, centers, gamma=None):
        c, K = _centers(centers, self.p)
        a = np.empty(self.n, dtype=np.int32)
        d = np.empty(self.n, dtype=np.float64)
        check(self._lib.skm_assign_sparse_centers(self.handle, _ptr(c), K, int(gamma is not None),
                                                  float(gamma if gamma is not None else 0.0), _ptr(a), _ptr(d)))
        return a, d

    def masked_distances(self, centers) -> np.ndarray:
        """K x n exact masked distances (SparseMatrixMinusCluster on resident data)."""
        c, K = _centers(centers, self.p)
        out = np.empty(K * self.n, dtype=np.float64)
        check(self._lib.skm_masked_distances(self.handle, _ptr(c), K, _ptr(out)))
        return out.reshape(self.n, K).T

    # -- k-means++ ------------------------------------------------------------
    def minmax(self) -> tuple[float, float]:
        """min(X(:)), max(X(:)) over all p*n elements, implicit zeros included (skm_dataset_minmax)."""
        a, b = C.c_double(), C.c_double()
        check(self._lib.skm_dataset_minmax(self.handle, C.byref(a), C.byref(b)))
        return a.value, b.value

    def kpp_update(self, center, gamma=None, first: bool = False, sparse_center: bool = False) -> float:
        """Fold the masked distance to one new centre into the running minimum; returns the local sum of
        D^2.  sparse_center=True is the gamma-less call of Arthur_initialization.m:26, where the centre stays
        sparse and the sum runs over the intersection of supp(x) and supp(centre) only (findClusterAssignments.m:70-74)."""
        c = np.ascontiguousarray(center, dtype=np.float64).reshape(-1)
        if c.shape[0] != self.p:
            raise ValueError("centre must have p entries")
        tot = C.c_double()
        if sparse_center:
            if gamma is not None:
                raise ValueError("the sparse-centre form has no gamma (Arthur_initialization.m:26-29)")
            check(self._lib.skm_kpp_update_sparse(self.handle, _ptr(c), int(first), C.byref(tot)))
            return tot.value
        check(self._lib.skm_kpp_update(self.handle, _ptr(c), int(gamma is not None),
                                       float(gamma if gamma is not None else 0.0), int(first), C.byref(tot)))
        return tot.value

    def kpp_pick(self, target: float) -> int:
        j = C.c_int64()
        check(self._lib.skm_kpp_pick(self.handle, float(target), C.byref(j)))
        return int(j.value)

    def kpp_mindist(self) -> np.ndarray:
        out = np.empty(self.n, dtype=np.float64)
        check(self._lib.skm_kpp_get_mindist(self.handle, _ptr(out)))
        return out


class Lloyd:
    """Iteration state for K centres on one resident shard (skm_lloyd)."""

    def __init__(self, ds: Dataset, K: int, incremental: bool = False, bounded: bool = False):
        self.ds = ds
        self.K = int(K)
        self._lib = ds._lib
        h = C.c_void_p()
        check(self._lib.skm_lloyd_create(ds.handle, self.K, C.byref(h)))
        self._h = h
        ds._children.add(self)
        if incremental:
            self.set_update_mode(True)
        if bounded:
            self.set_assign_mode(True)

    def set_assign_mode(self, bounded: bool):
        """True: bounds carried across calls let most columns keep their centre after one centre evaluation
        (skm_lloyd_set_assign_mode); assignments and distances are unchanged."""
        check(self._lib.skm_lloyd_set_assign_mode(self.handle, 1 if bounded else 0))

    def last_assign_flagged(self) -> int:
        v = C.c_int64(0)
        check(self._lib.skm_lloyd_last_assign(self.handle, C.byref(v)))
        return int(v.value)

    def set_prune(self, mode):
        """None: automatic; True / False: always / never try the partial-distance pruning of the full assignment pass
        (skm_lloyd_set_prune, plans with several launches: K > 16); assignments are the same either way."""
        check(self._lib.skm_lloyd_set_prune(self.handle, -1 if mode is None else (1 if mode else 0)))

    def last_prune(self) -> tuple[int, int]:
        """(columns the last pruned pass could not keep, entry pairs it read per column); (-1, -1): not pruned."""
        a, b = C.c_int64(0), C.c_int64(0)
        check(self._lib.skm_lloyd_last_prune(self.handle, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def set_tc_filter(self, mode):
        """None: automatic; True / False: force the tensor-core plan of the full assignment pass on / off
        (skm_lloyd_set_tc_filter); assignments are the same either way."""
        check(self._lib.skm_lloyd_set_tc_filter(self.handle, -1 if mode is None else (1 if mode else 0)))

    def last_tc(self) -> tuple[int, int]:
        """(columns the winner's exact evaluation could not keep, columns that went to fp64) of the last
        tensor-core pass, valid after finalize(); (-1, -1) if the last pass was not one."""
        a, b = C.c_int64(0), C.c_int64(0)
        check(self._lib.skm_lloyd_last_tc(self.handle, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def debug_tc_scores(self, gamma=None) -> np.ndarray:
        """Test hook: raw filter scores [n, bn] of a tensor-core pass with the current centres."""
        out = np.empty((self.ds.n, 128), dtype=np.float32)
        bn = C.c_int64(0)
        check(self._lib.skm_debug_tc_scores(self.handle, int(gamma is not None), float(gamma if gamma is not None else 0.0),
                                            _ptr(out), C.byref(bn)))
        return out.reshape(-1)[: self.ds.n * bn.value].reshape(self.ds.n, bn.value)

    def set_update_mode(self, incremental: bool):
        """False: per-cluster sums recomputed from all columns every iteration (the reference's way);
        True: only columns whose assignment changed move their entries (skm_lloyd_set_update_mode)."""
        check(self._lib.skm_lloyd_set_update_mode(self.handle, 1 if incremental else 0))

    def last_update(self) -> tuple[str, int]:
        kind, nch = C.c_int(0), C.c_int64(0)
        check(self._lib.skm_lloyd_last_update(self.handle, C.byref(kind), C.byref(nch)))
        return ("full", "incremental", "unchanged")[kind.value], int(nch.value)

    @property
    def handle(self):
        if self._h is None:
            raise RuntimeError("lloyd state destroyed")
        return self._h

    def close(self):
        if self._h is not None:
            self._lib.skm_lloyd_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def kernel_name(self) -> str:
        return self._lib.skm_lloyd_kernel_name(self.handle).decode()

    def set_centers(self, centers):
        c, K = _centers(centers, self.ds.p)
        if K != self.K:
            raise ValueError("centers must be p x K")
        check(self._lib.skm_lloyd_set_centers(self.handle, _ptr(c)))

    def get_centers(self) -> np.ndarray:
        out = np.empty(self.ds.p * self.K, dtype=np.float64)
        check(self._lib.skm_lloyd_get_centers(self.handle, _ptr(out)))
        return out.reshape(self.K, self.ds.p).T.copy()

    def get_centers_old(self) -> np.ndarray:
        out = np.empty(self.ds.p * self.K, dtype=np.float64)
        check(self._lib.skm_lloyd_get_centers_old(self.handle, _ptr(out)))
        return out.reshape(self.K, self.ds.p).T.copy()

    def set_center_column(self, k: int, col):
        c = np.ascontiguousarray(col, dtype=np.float64).reshape(-1)
        check(self._lib.skm_lloyd_set_center_column(self.handle, int(k), _ptr(c)))

    def assign(self, gamma=None):
        check(self._lib.skm_lloyd_assign(self.handle, int(gamma is not None),
                                         float(gamma if gamma is not None else 0.0)))

    def assign_sparse(self, gamma=None):
        """Sparse-centres branch (findClusterAssignments.m:63-75) against the current centres."""
        check(self._lib.skm_lloyd_assign_sparse(self.handle, int(gamma is not None),
                                                float(gamma if gamma is not None else 0.0)))

    def accumulate(self):
        check(self._lib.skm_lloyd_accumulate(self.handle))

    def partials_ptr(self) -> tuple[int, int]:
        n = C.c_int64()
        p = self._lib.skm_lloyd_partials(self.handle, C.byref(n))
        return int(p or 0), int(n.value)

    def partials_tensor(self):
        """The device buffer [S | N | counts | sumsq] as a torch float64 tensor (no copy)."""
        import torch
        ptr, n = self.partials_ptr()

        class _Raw:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False),
                                        "version": 2, "strides": None}
        return torch.as_tensor(_Raw(), device=f"cuda:{self.ds.ctx.device}")

    @staticmethod
    def _stats(s: _lib.IterStats) -> IterStats:
        return IterStats(float(s.dff), float(s.sumsq), int(s.n_empty), int(s.n_rechecked),
                         int(s.n_points), bool(s.has_nan))

    def finalize(self, gamma: float, ml_correction: bool = True) -> IterStats:
        s = _lib.IterStats()
        check(self._lib.skm_lloyd_finalize(self.handle, float(gamma), int(ml_correction), C.byref(s)))
        return self._stats(s)

    def refresh_diff(self) -> IterStats:
        s = _lib.IterStats()
        check(self._lib.skm_lloyd_refresh_diff(self.handle, C.byref(s)))
        return self._stats(s)

    def counts(self) -> np.ndarray:
        out = np.empty(self.K, dtype=np.int64)
        check(self._lib.skm_lloyd_get_counts(self.handle, _ptr(out)))
        return out

    def assignments(self, want_dist: bool = True):
        a = np.empty(self.ds.n, dtype=np.int32)
        d = np.empty(self.ds.n, dtype=np.float64) if want_dist else None
        check(self._lib.skm_lloyd_get_assignments(self.handle, _ptr(a), _ptr(d)))
        return a, d

    def argmax_distance(self) -> tuple[float, int]:
        v = C.c_double()
        j = C.c_int64()
        check(self._lib.skm_lloyd_argmax_distance(self.handle, C.byref(v), C.byref(j)))
        return v.value, int(j.value)

    def step(self, gamma_dist, gamma_update: float, ml_correction: bool = True, reduce=None) -> IterStats:
        """One Lloyd iteration on this shard: K1 assign, K2 accumulate, optional collective
        (`reduce(partials_tensor)`), K3 finalize."""
        self.assign(gamma_dist)
        self.accumulate()
        if reduce is not None:
            reduce(self)
        return self.finalize(gamma_update, ml_correction)


def lloyd_step_host(p: int, n: int, jc, ir, val, centers, gamma_dist, gamma_update: float,
                    ml_correction: bool = True, chunk_cols: int = 0, want_assign: bool = True,
                    want_dist: bool = False, ctx: Context | None = None, reduce=None):
    """One Lloyd iteration with X in HOST memory, streamed over PCIe in column chunks
    (skm_lloyd_step_host).  jc/ir/val are numpy arrays (int32/int64, float32/float64; pinned
    memory overlaps best).  `reduce(tensor)` (optional) receives the partials as a torch float64
    device tensor with the library stream current, for the multi-GPU all-reduce.
    Returns (new_centers, assign 1-based or None, dist or None, IterStats)."""
    ctx = ctx or default_context()
    jc = np.ascontiguousarray(jc)
    ir = np.ascontiguousarray(ir)
    val = np.ascontiguousarray(val)
    if jc.dtype not in _NP_INDEX:
        jc = jc.astype(np.int64)
    if ir.dtype not in _NP_ROWS:
        ir = ir.astype(np.int64)
    if val.dtype not in _NP_VALUE:
        val = val.astype(np.float64)
    c, K = _centers(centers, p)
    out_c = np.empty(p * K, dtype=np.float64)
    a = np.empty(n, dtype=np.int32) if want_assign else None
    d = np.empty(n, dtype=np.float64) if want_dist else None
    st = _lib.IterStats()
    cb = None
    if reduce is not None:
        import torch

        def _cb(ptr, count, stream, _user):
            try:
                class _Raw:
                    __cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f8", "data": (int(ptr), False),
                                                "version": 2, "strides": None}
                t = torch.as_tensor(_Raw(), device=f"cuda:{ctx.device}")
                with torch.cuda.stream(torch.cuda.ExternalStream(int(stream), device=f"cuda:{ctx.device}")):
                    reduce(t)
                return 0
            except Exception:                      # never unwind through the C frame
                import traceback
                traceback.print_exc()
                return 1
        cb = _lib.REDUCE_FN(_cb)
    check(ctx._lib.skm_lloyd_step_host(
        ctx.handle, p, n, _ptr(jc), _NP_INDEX[jc.dtype], _ptr(ir), _NP_ROWS[ir.dtype], _ptr(val),
        _NP_VALUE[val.dtype], _ptr(c), K, int(gamma_dist is not None),
        float(gamma_dist if gamma_dist is not None else 0.0), float(gamma_update), int(ml_correction),
        int(chunk_cols), _ptr(out_c), _ptr(a), _ptr(d), C.byref(st),
        C.cast(cb, C.c_void_p) if cb is not None else None, None))
    return out_c.reshape(K, p).T.copy(), a, d, Lloyd._stats(st)


def second_pass(X, centers=None, assign_in=None, scale: float = 1.0, want_assign: bool = True,
                want_dist: bool = True, chunk_cols: int = 0, ctx: Context | None = None, x_device_ptr: int | None = None,
                shape=None, x_dtype=None):
    """The second pass over the ORIGINAL dense data (skm_second_pass; kmeans_sparsified.m:542-560,
    private/recalculateAssignmentLargeFile.m:85-113).

    X: dense (p, n) array with points as columns -- any array whose transpose is C-contiguous is
    used in place (e.g. a np.memmap of an (n, p) file); float32 or float64.  `centers` (p, K) are
    the centres to re-assign against; `assign_in` (n,) 1-based labels select the columns averaged
    into the new centres.  Returns a dict with the requested pieces of
    centers (p, K), counts (K,), assign (n,) 1-based, dist (n,), n_rechecked."""
    ctx = ctx or default_context()
    if x_device_ptr is not None:
        p, n = shape
        xt = SKM_F32 if np.dtype(x_dtype) == np.float32 else SKM_F64
        xptr, on_dev, keep = int(x_device_ptr), 1, None
    else:
        p, n = X.shape
        XT = X.T                                            # (n, p): a point per row, contiguous
        if XT.dtype not in (np.float32, np.float64) or not XT.flags.c_contiguous:
            XT = np.ascontiguousarray(XT, dtype=np.float64 if XT.dtype != np.float32 else np.float32)
        xt = SKM_F32 if XT.dtype == np.float32 else SKM_F64
        xptr, on_dev, keep = XT.ctypes.data, 0, XT
    want_any_assign = want_assign or want_dist
    c = K = None
    if centers is not None:
        c, K = _centers(centers, p)
    a_in = None
    if assign_in is not None:
        a_in = np.ascontiguousarray(assign_in, dtype=np.int32).reshape(-1)
        if a_in.shape[0] != n:
            raise ValueError("assign_in must have one label per column")
        if K is None:
            K = max(1, int(a_in.max())) if a_in.size else 1
    if K is None:
        raise ValueError("need centers and/or assign_in")
    out = {}
    c_out = np.empty(p * K, dtype=np.float64) if a_in is not None else None
    cnt = np.zeros(K, dtype=np.int64) if a_in is not None else None
    a = np.empty(n, dtype=np.int32) if (want_assign and c is not None) else None
    d = np.empty(n, dtype=np.float64) if (want_dist and c is not None) else None
    nre = C.c_int64(0)
    check(ctx._lib.skm_second_pass(ctx.handle, p, n, C.c_void_p(xptr), xt, on_dev, float(scale),
                                   _ptr(c) if (c is not None and want_any_assign) else None, K, _ptr(a_in),
                                   _ptr(c_out), _ptr(cnt), _ptr(a), _ptr(d), int(chunk_cols), C.byref(nre)))
    del keep
    if c_out is not None:
        out["centers"] = c_out.reshape(K, p).T.copy()
        out["counts"] = cnt
    if a is not None:
        out["assign"] = a
    if d is not None:
        out["dist"] = d
    out["n_rechecked"] = int(nre.value)
    return out


def dct_mix(X, signs=None, inverse: bool = False, ctx: Context | None = None) -> np.ndarray:
    """mix(X) = dct(D*X) or, inverse, unmix(X) = D*idct(X) (kmeans_sparsified.m:256-258,295-296) for a
    dense (p, n) host matrix, fp64, on the GPU (skm_dct_mix)."""
    ctx = ctx or default_context()
    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 1:
        X = X.reshape(-1, 1)
    p, n = X.shape
    d = None if signs is None else np.ascontiguousarray(signs, dtype=np.float64).reshape(-1)
    if d is not None and d.shape[0] != p:
        raise ValueError("signs must have p entries")
    Xf = np.ascontiguousarray(X.T).reshape(-1)
    out = np.empty(p * n, dtype=np.float64)
    check(ctx._lib.skm_dct_mix(ctx.handle, p, n, _ptr(Xf), _ptr(d), int(inverse), _ptr(out)))
    return out.reshape(n, p).T


def sample_rows_general(p: int, n: int, m: int, seed: int, col0: int, rows_ptr: int, ctx: Context | None = None):
    """Row sets of the DCT pipeline's on-device sampler into device int32[m*n] (skm_sample_rows_general)."""
    ctx = ctx or default_context()
    check(ctx._lib.skm_sample_rows_general(ctx.handle, p, n, m, int(seed), int(col0), C.c_void_p(rows_ptr)))


def mix_hadamard(X, signs, compute: str = "f64", ctx: Context | None = None) -> np.ndarray:
    """mix(X) = hadamard(D*[X;0])/sqrt(p2) on the GPU (kmeans_sparsified.m:238-248,286-295).
    X is dense p x n (host), signs has p2 = 2^nextpow2(p) entries; returns p2 x n float64."""
    ctx = ctx or default_context()
    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 1:
        X = X.reshape(-1, 1)
    p, n = X.shape
    d = np.ascontiguousarray(signs, dtype=np.float64).reshape(-1)
    p2 = d.shape[0]
    Xf = np.ascontiguousarray(X.T).reshape(-1)
    out = np.empty(p2 * n, dtype=np.float64)
    check(ctx._lib.skm_mix_hadamard(ctx.handle, p, p2, n, _ptr(Xf), _ptr(d),
                                    SKM_F64 if compute == "f64" else SKM_F32, _ptr(out)))
    return out.reshape(n, p2).T


def fwht_f32_inplace(p2: int, n: int, x_ptr: int, signs_ptr: int | None, ctx: Context | None = None):
    """In-place device FWHT of a dense p2 x n float32 matrix: sign flip, transform, /sqrt(p2)."""
    ctx = ctx or default_context()
    check(ctx._lib.skm_fwht_f32_inplace(ctx.handle, p2, n, x_ptr, signs_ptr))


def sample_rows(p2: int, n: int, m: int, seed: int, col0: int, rows_ptr: int, ctx: Context | None = None):
    """Write the row sets the on-device sampler draws for (seed, col0) into device int32[m*n]."""
    ctx = ctx or default_context()
    check(ctx._lib.skm_sample_rows(ctx.handle, p2, n, m, int(seed), int(col0), rows_ptr))
