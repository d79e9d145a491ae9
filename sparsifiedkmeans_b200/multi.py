"""Several GPUs of one box driven by ONE host process (skm_multi_* of include/skm_b200.h).

The reference is a single MATLAB process; `MultiContext` / `MultiDataset` / `MultiLloyd` are what its MEX
gateway (mex/skm_lloyd_mex.c, 'upload' with a device count) and `kmeans_sparsified(..., Devices=[...])` use to
run the column-sharded Lloyd iteration on all GPUs without torchrun: K1/K2 per shard concurrently, then one
kernel per device that sums every peer's partials through NVLink peer memory in device order and finalises the
centres (the all-reduce fused into K3).  `MultiLloyd` has the same methods as `engine.Lloyd`, with global
column indices, so the host loop of kmeans.py runs unchanged on either.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import SKM_F32, SKM_F64, check
from .engine import _NP_INDEX, _NP_ROWS, _NP_VALUE, Context, Dataset, IterStats, Lloyd, _centers, _ptr


class _BorrowedContext(Context):
    """View of a context owned by a MultiContext (never destroyed from Python)."""

    def __init__(self, lib, handle, device):
        self._lib = lib
        self._h = C.c_void_p(handle)
        self.device = int(device)

    def close(self):
        self._h = None


class MultiContext:
    def __init__(self, devices=None):
        """devices: None / 0 = every visible GPU, an int = the first that many, or a list of device ids."""
        self._lib = _lib.load()
        h = C.c_void_p()
        if devices is None or isinstance(devices, int):
            check(self._lib.skm_multi_create(int(devices or 0), None, C.byref(h)))
        else:
            arr = (C.c_int * len(devices))(*[int(d) for d in devices])
            check(self._lib.skm_multi_create(len(devices), C.cast(arr, C.c_void_p), C.byref(h)))
        self._h = h
        self.ndev = int(self._lib.skm_multi_ndev(h))
        self.contexts = []
        for g in range(self.ndev):
            ch = self._lib.skm_multi_ctx(h, g)
            self.contexts.append(_BorrowedContext(self._lib, ch, self._lib.skm_ctx_device(ch)))
        self.devices = [c.device for c in self.contexts]
        self.peer_access = bool(self._lib.skm_multi_peer_access(h))

    @property
    def handle(self):
        if self._h is None:
            raise RuntimeError("multi context destroyed")
        return self._h

    def close(self):
        if self._h is not None:
            self._lib.skm_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiDataset:
    """A p x n sparsified matrix sharded by contiguous column blocks over the devices of a MultiContext."""

    def __init__(self, mctx: MultiContext, handle):
        self.mctx = mctx
        self.ctx = mctx.contexts[0]
        self._lib = mctx._lib
        self._h = handle
        info = _lib.DatasetInfo()
        check(self._lib.skm_multi_dataset_get_info(handle, C.byref(info)))
        self.p, self.n, self.nnz = int(info.p), int(info.n), int(info.nnz)
        self.max_col_nnz = int(info.max_col_nnz)
        self.store_dtype = "f32" if info.store_dtype == SKM_F32 else "f64"
        self.device_bytes = int(info.device_bytes)
        self.stream_bytes = int(info.stream_bytes)
        self.col0 = []
        for g in range(mctx.ndev):
            c0 = C.c_int64()
            self._lib.skm_multi_dataset_shard(handle, g, C.byref(c0))
            self.col0.append(int(c0.value))
        self._children = []

    @classmethod
    def from_csc(cls, p, n, jc, ir, val, store="f32", mctx: MultiContext | None = None):
        mctx = mctx or MultiContext()
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
        check(mctx._lib.skm_multi_dataset_create_csc(
            mctx.handle, p, n, _ptr(jc), _NP_INDEX[jc.dtype], _ptr(ir), _NP_ROWS[ir.dtype], _ptr(val),
            _NP_VALUE[val.dtype], SKM_F32 if store == "f32" else SKM_F64, C.byref(h)))
        return cls(mctx, h)

    @classmethod
    def from_scipy(cls, X, store="f32", mctx: MultiContext | None = None):
        import scipy.sparse as sp
        X = sp.csc_matrix(X)
        X.sort_indices()
        return cls.from_csc(X.shape[0], X.shape[1], X.indptr, X.indices, X.data, store, mctx)

    @classmethod
    def from_shards(cls, shards, mctx: MultiContext):
        """Adopt per-device `engine.Dataset`s (shards[g] built on mctx.contexts[g], in column order); they are
        owned by the result afterwards."""
        arr = (C.c_void_p * mctx.ndev)(*[s.handle for s in shards])
        h = C.c_void_p()
        check(mctx._lib.skm_multi_dataset_from_shards(mctx.handle, C.cast(arr, C.c_void_p), C.byref(h)))
        for s in shards:
            s._h = None                                      # ownership moved
        return cls(mctx, h)

    @classmethod
    def from_dense_host(cls, X, signs, m, seed=0, mctx: MultiContext | None = None, dct=False, rows=None):
        """Precondition + sample a dense host matrix (p x n, points are columns) block by block, each block on
        its own device (the on-device sampler is a function of (seed, global column), so the result does not
        depend on the number of devices)."""
        from concurrent.futures import ThreadPoolExecutor
        mctx = mctx or MultiContext()
        n = X.shape[1]
        G = mctx.ndev

        def build(g):
            lo, hi = n * g // G, n * (g + 1) // G
            if dct:
                r = None if rows is None else np.asarray(rows)[:, lo:hi]
                return Dataset.from_dense_host_dct(X[:, lo:hi], signs, m, seed=seed, col0=lo, rows=r, ctx=mctx.contexts[g])
            return Dataset.from_dense_host(X[:, lo:hi], signs, m, seed=seed, col0=lo, ctx=mctx.contexts[g])
        with ThreadPoolExecutor(G) as ex:
            shards = list(ex.map(build, range(G)))
        return cls.from_shards(shards, mctx)

    @property
    def handle(self):
        if self._h is None:
            raise RuntimeError("dataset destroyed")
        return self._h

    def close(self):
        if self._h is not None:
            for ch in list(self._children):
                ch.close()
            self._lib.skm_multi_dataset_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def get_column(self, j: int) -> np.ndarray:
        out = np.empty(self.p, dtype=np.float64)
        check(self._lib.skm_multi_dataset_get_column(self.handle, int(j), _ptr(out)))
        return out

    def minmax(self):
        lo, hi = np.inf, -np.inf
        for g in range(self.mctx.ndev):
            a, b = C.c_double(), C.c_double()
            sh = self._lib.skm_multi_dataset_shard(self.handle, g, None)
            check(self._lib.skm_dataset_minmax(sh, C.byref(a), C.byref(b)))
            lo, hi = min(lo, a.value), max(hi, b.value)
        return lo, hi

    def kpp_update(self, center, gamma=None, first=False, sparse_center=False) -> float:
        c = np.ascontiguousarray(center, dtype=np.float64).reshape(-1)
        tot = C.c_double()
        check(self._lib.skm_multi_kpp_update(self.handle, _ptr(c), int(gamma is not None),
                                             float(gamma if gamma is not None else 0.0), int(first), int(sparse_center),
                                             C.byref(tot)))
        return tot.value

    def kpp_pick(self, target: float) -> int:
        j = C.c_int64()
        check(self._lib.skm_multi_kpp_pick(self.handle, float(target), C.byref(j)))
        return int(j.value)


class MultiLloyd:
    """engine.Lloyd over all shards: same methods, global column indices."""

    def __init__(self, ds: MultiDataset, K: int, incremental=False, bounded=False):
        self.ds = ds
        self.K = int(K)
        self._lib = ds._lib
        h = C.c_void_p()
        check(self._lib.skm_multi_lloyd_create(ds.handle, self.K, C.byref(h)))
        self._h = h
        ds._children.append(self)
        if incremental or bounded:
            check(self._lib.skm_multi_lloyd_set_modes(h, int(bool(incremental)), int(bool(bounded))))
        self._sparse = False

    @property
    def handle(self):
        if self._h is None:
            raise RuntimeError("lloyd state destroyed")
        return self._h

    def close(self):
        if self._h is not None:
            self._lib.skm_multi_lloyd_destroy(self._h)
            self._h = None
            if self in self.ds._children:
                self.ds._children.remove(self)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_centers(self, centers):
        c, K = _centers(centers, self.ds.p)
        if K != self.K:
            raise ValueError("centers must be p x K")
        check(self._lib.skm_multi_lloyd_set_centers(self.handle, _ptr(c)))

    def _get(self, fn, *a):
        out = np.empty(self.ds.p * self.K, dtype=np.float64)
        check(fn(self.handle, *a, _ptr(out)))
        return out.reshape(self.K, self.ds.p).T.copy()

    def get_centers(self): return self._get(self._lib.skm_multi_lloyd_get_centers)
    def get_centers_old(self): return self._get(self._lib.skm_multi_lloyd_get_centers_old)
    def get_centers_of(self, g): return self._get(self._lib.skm_multi_lloyd_get_centers_of, int(g))

    def set_center_column(self, k, col):
        c = np.ascontiguousarray(col, dtype=np.float64).reshape(-1)
        check(self._lib.skm_multi_lloyd_set_center_column(self.handle, int(k), _ptr(c)))

    def step(self, gamma_dist, gamma_update, ml_correction=True, sparse_centers=False, reduce=None) -> IterStats:
        s = _lib.IterStats()
        check(self._lib.skm_multi_lloyd_step(self.handle, int(gamma_dist is not None),
                                             float(gamma_dist if gamma_dist is not None else 0.0), float(gamma_update),
                                             int(ml_correction), int(sparse_centers), C.byref(s)))
        return Lloyd._stats(s)

    # the three-call form of engine.Lloyd, so kmeans.py's loop runs unchanged: assign/assign_sparse only record
    # the branch, accumulate is a no-op, finalize runs the whole fused step
    def assign(self, gamma=None):
        self._pending = (gamma, False)

    def assign_sparse(self, gamma=None):
        self._pending = (gamma, True)

    def accumulate(self):
        pass

    def finalize(self, gamma, ml_correction=True) -> IterStats:
        g, sparse = self._pending
        return self.step(g, gamma, ml_correction, sparse)

    def refresh_diff(self) -> IterStats:
        s = _lib.IterStats()
        check(self._lib.skm_multi_lloyd_refresh_diff(self.handle, C.byref(s)))
        return Lloyd._stats(s)

    def counts(self):
        out = np.empty(self.K, dtype=np.int64)
        check(self._lib.skm_multi_lloyd_get_counts(self.handle, _ptr(out)))
        return out

    def assignments(self, want_dist=True):
        a = np.empty(self.ds.n, dtype=np.int32)
        d = np.empty(self.ds.n, dtype=np.float64) if want_dist else None
        check(self._lib.skm_multi_lloyd_get_assignments(self.handle, _ptr(a), _ptr(d)))
        return a, d

    def argmax_distance(self):
        v, j = C.c_double(), C.c_int64()
        check(self._lib.skm_multi_lloyd_argmax_distance(self.handle, C.byref(v), C.byref(j)))
        return v.value, int(j.value)

    @property
    def launch_count(self) -> int:
        v = C.c_int64()
        check(self._lib.skm_multi_lloyd_launch_count(self.handle, C.byref(v)))
        return int(v.value)
