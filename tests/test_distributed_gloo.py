"""World-size-2 gloo test of the sharded Lloyd driver (host logic of the N>1 path): column
shards, one all-reduce of [S|N|counts|sumsq] per iteration, global singleton EmptyAction.
The per-shard engine here is oracle-backed (CPU); on the GPU box the same driver runs over
CudaShardEngine + NCCL.  CPU only."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleShardEngine:
    """Engine protocol of sparsifiedkmeans_b200.distributed over the CPU oracle."""

    def __init__(self, X, K):
        from oracle import host_ref
        self.X = host_ref.as_csc(X)
        self.p, self.n_local = self.X.shape
        self.K = K

    def set_centers(self, C): self.C = np.array(C, dtype=np.float64)
    def get_centers(self): return self.C.copy()

    def assign(self, gamma):
        from oracle import host_ref
        if self.n_local:
            self.a, self.d, _ = host_ref.find_cluster_assignments(self.X, self.C, gamma)
        else:
            self.a, self.d = np.zeros(0, dtype=np.int64), np.zeros(0)

    def accumulate(self):
        from oracle import cport
        p, K = self.p, self.K
        _, S, N, counts = cport.centroid_update(p, self.n_local, K, self.X.indptr, self.X.indices, self.X.data,
                                                self.a, 1.0, np.zeros((p, K)), True)
        flat = np.concatenate([S.T.ravel(), N.T.ravel(), counts.astype(np.float64), [np.sum(self.d ** 2)]])
        self.part = torch.from_numpy(flat)
        return self.part

    def finalize(self, gamma, ml):
        from sparsifiedkmeans_b200.engine import IterStats
        p, K = self.p, self.K
        v = self.part.numpy()
        S, N = v[:p * K].reshape(K, p).T, v[p * K:2 * p * K].reshape(K, p).T
        self._counts = np.rint(v[2 * p * K:2 * p * K + K]).astype(np.int64)
        self.C_old = self.C.copy()
        ok = self._counts > 0
        self.C[:, ok] = gamma * S[:, ok] / (N[:, ok] + 1e-16)
        return IterStats(float(np.linalg.norm(self.C_old - self.C)), float(v[-1]), int(np.sum(~ok)), 0,
                         int(self._counts.sum()), bool(np.isnan(self.C).any()))

    def refresh_diff(self):
        from sparsifiedkmeans_b200.engine import IterStats
        return IterStats(float(np.linalg.norm(self.C_old - self.C)), 0.0, 0, 0, 0, bool(np.isnan(self.C).any()))

    def counts(self): return self._counts
    def argmax_distance(self): j = int(np.argmax(self.d)); return float(self.d[j]), j
    def get_column(self, j): return np.asarray(self.X[:, j].todense()).ravel()
    def set_center_column(self, k, col): self.C[:, k] = col
    def assignments(self): return self.a, self.d

    def kpp_update(self, center, gamma, first):
        from oracle import host_ref
        c = np.asarray(center, dtype=np.float64).reshape(-1, 1)
        _, dd, _ = host_ref.find_cluster_assignments(self.X, c, gamma)
        self.mind = dd if first else np.minimum(self.mind, dd)
        return float(np.sum(self.mind ** 2))

    def kpp_pick(self, target):
        cs = np.cumsum(self.mind ** 2)
        return int(min(np.searchsorted(cs, target, side="right"), self.n_local - 1))


def _worker(rank, world, port, case, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sparsifiedkmeans_b200.distributed import ShardedLloyd, shard_bounds
        from tests.util import make_sparsified
        X, c, gamma = make_sparsified(**case["data"])
        if case.get("kill_cluster"):
            c = c.copy()
            c[:, -1] = 1e6                                # a centre nobody is assigned to
        lo, hi = shard_bounds(X.shape[1], world, rank)
        eng = OracleShardEngine(X[:, lo:hi], c.shape[1])
        drv = ShardedLloyd(eng)
        its, st = drv.run(c, gamma, gamma, max_iter=case["max_iter"], tol=1e-6)
        a, d = eng.assignments()
        out[rank] = dict(its=its, dff=st.dff, sumsq=st.sumsq, centers=eng.get_centers(), a=np.asarray(a), lo=lo, hi=hi)
    finally:
        dist.destroy_process_group()


def _kpp_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sparsifiedkmeans_b200.distributed import shard_bounds, sharded_arthur_initialization
        from tests.util import make_sparsified
        X, _, gamma = make_sparsified(p=64, n=1001, m=8, K=6, seed=71, kind="mixture")
        lo, hi = shard_bounds(X.shape[1], world, rank)
        eng = OracleShardEngine(X[:, lo:hi], 6)
        u = np.random.default_rng(3).random(4000)
        idx, cen = sharded_arthur_initialization(eng, 6, gamma, X.shape[1], lo, first=500, uniforms=iter(u))
        out[rank] = dict(idx=idx, cen=cen)
    finally:
        dist.destroy_process_group()


def test_two_rank_kmeanspp_equals_single_process():
    from oracle import host_ref
    from tests.util import make_sparsified
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_kpp_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    X, _, gamma = make_sparsified(p=64, n=1001, m=8, K=6, seed=71, kind="mixture")
    u = np.random.default_rng(3).random(4000)
    want, cen = host_ref.arthur_initialization(X, 6, gamma, first=500, uniforms=iter(u))
    assert np.array_equal(out[0]["idx"], want) and np.array_equal(out[1]["idx"], want)
    assert np.array_equal(out[0]["cen"], np.asarray(cen.todense()))
    assert np.array_equal(out[0]["cen"], out[1]["cen"])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("case", [
    dict(data=dict(p=64, n=901, m=8, K=5, seed=31, kind="mixture"), max_iter=15),
    dict(data=dict(p=48, n=700, m=6, K=4, seed=32, kind="unstructured"), max_iter=6),
    dict(data=dict(p=64, n=640, m=8, K=5, seed=33, kind="mixture"), max_iter=4, kill_cluster=True),
], ids=["mixture", "unstructured", "empty-cluster-singleton"])
def test_two_rank_lloyd_equals_single_process(case):
    from oracle import host_ref
    from tests.util import make_sparsified
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), case, out), nprocs=world, join=True)
    X, c, gamma = make_sparsified(**case["data"])
    if case.get("kill_cluster"):
        c = c.copy()
        c[:, -1] = 1e6
    ref = host_ref.lloyd(X, c, gamma, max_iter=case["max_iter"], tol=1e-6)
    r0, r1 = out[0], out[1]
    assert r0["its"] == r1["its"] == ref.iterations
    assert np.array_equal(r0["centers"], r1["centers"])                  # ranks never drift apart
    np.testing.assert_allclose(r0["centers"], ref.centers, rtol=1e-9, atol=1e-12)
    a = np.concatenate([r0["a"], r1["a"]])
    assert np.array_equal(a, ref.assignments)                             # bit-exact across the sharding
    np.testing.assert_allclose(np.sqrt(r0["sumsq"]), ref.objective, rtol=1e-12)
