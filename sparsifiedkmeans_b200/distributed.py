"""Column-sharded Lloyd iteration: one process per GPU, one all-reduce per iteration.

The reference is single-process; the path shards naturally because points (columns) are
independent in K1 and K2 needs exactly one exchange: the sum over ranks of
[S (p*K) | N (p*K) | counts (K) | sumsq (1)].  After the all-reduce every
rank holds identical partials, runs K3 identically, and the centres never drift apart.

`ShardedLloyd` is written against a small engine protocol so the host logic can be exercised
on CPU with the gloo backend (tests/test_distributed_gloo.py drives it with an oracle-backed
engine); in production the engine is `CudaShardEngine` (libskm_b200 on this rank's GPU) and the
collective is NCCL over NVLink on the library's own stream.

Engine protocol (all indices local to the shard):
    n_local, p, K
    set_centers(C), get_centers() -> C
    assign(gamma_dist)                       K1 on the local columns
    accumulate() -> partials                 K2; returns the tensor to all-reduce (in place)
    finalize(gamma, ml) -> IterStats         K3 from the (reduced) partials
    refresh_diff() -> IterStats
    counts() -> int64[K]                      global member counts after finalize
    argmax_distance() -> (value, local j)
    get_column(j) -> dense p-vector
    set_center_column(k, col)
    assignments() -> (1-based int32[n_local], float64[n_local])
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous column block [lo, hi) of `rank` (GPU g owns columns [g*n/G, (g+1)*n/G))."""
    return n * rank // world, n * (rank + 1) // world


class CudaShardEngine:
    """libskm_b200 on this rank's GPU."""

    def __init__(self, ds, K: int, incremental: bool = False, bounded: bool = False):
        from .engine import Lloyd
        self.ds = ds
        # both modes are per-shard decisions (each rank keeps its own sums and bounds); the all-reduce buffer
        # is a copy of the kept sums, so the collective is issued by every rank in every iteration either way
        self.L = Lloyd(ds, K, incremental=incremental, bounded=bounded)
        self._modes = (incremental, bounded)
        self.n_local, self.p, self.K = ds.n, ds.p, int(K)
        self._ext = None

    def stream(self):
        """torch view of the library's stream, so the collective is ordered after K2."""
        import torch
        if self._ext is None:
            self._ext = torch.cuda.ExternalStream(self.ds.ctx.stream, device=f"cuda:{self.ds.ctx.device}")
        return self._ext

    def set_centers(self, C): self.L.set_centers(C)
    def get_centers(self): return self.L.get_centers()
    def assign(self, gamma_dist): self.L.assign(gamma_dist)
    def assign_sparse(self, gamma_dist): self.L.assign_sparse(gamma_dist)
    def finalize(self, gamma, ml): return self.L.finalize(gamma, ml)
    def refresh_diff(self): return self.L.refresh_diff()
    def counts(self): return self.L.counts()
    def argmax_distance(self): return self.L.argmax_distance()
    def get_column(self, j): return self.ds.get_column(j)
    def set_center_column(self, k, col): self.L.set_center_column(k, col)
    def assignments(self): return self.L.assignments()

    def accumulate(self):
        self.L.accumulate()
        return self.L.partials_tensor()

    # k-means++ support
    def kpp_update(self, center, gamma, first): return self.ds.kpp_update(center, gamma, first)
    def kpp_pick(self, target): return self.ds.kpp_pick(target)

    def drop_centers(self, keep):
        """EmptyAction='drop' (kmeans_sparsified.m:454-459): continue with the centres `keep` only."""
        from .engine import Lloyd
        cen = self.L.get_centers()[:, keep]
        self.L.close()
        self.K = int(len(keep))
        self.L = Lloyd(self.ds, self.K, incremental=self._modes[0], bounded=self._modes[1])
        self.L.set_centers(cen)

    def close(self):
        self.L.close()


class ShardedLloyd:
    """The Lloyd loop of kmeans_sparsified.m:417-486 over column shards."""

    def __init__(self, engine, group=None):
        self.e = engine
        self.group = group

    # -- collectives ---------------------------------------------------------
    def _dist(self):
        import torch.distributed as dist
        return dist if dist.is_available() and dist.is_initialized() else None

    def _allreduce(self, t):
        dist = self._dist()
        if dist is None or dist.get_world_size(self.group) == 1:
            return
        stream = self.e.stream() if hasattr(self.e, "stream") else None
        if stream is not None:
            import torch
            with torch.cuda.stream(stream):
                dist.all_reduce(t, group=self.group)
        else:
            dist.all_reduce(t, group=self.group)

    def _gather_objects(self, obj):
        dist = self._dist()
        if dist is None or dist.get_world_size(self.group) == 1:
            return [obj]
        out = [None] * dist.get_world_size(self.group)
        dist.all_gather_object(out, obj, group=self.group)
        return out

    def _bcast_object(self, obj, src):
        dist = self._dist()
        if dist is None or dist.get_world_size(self.group) == 1:
            return obj
        box = [obj]
        # `src` is a rank of self.group; broadcast_object_list wants the global rank
        gsrc = dist.get_global_rank(self.group, src) if self.group is not None else src
        dist.broadcast_object_list(box, src=gsrc, group=self.group)
        return box[0]

    # -- one iteration ---------------------------------------------------------
    def step(self, gamma_dist, gamma_update, ml_correction=True, empty_action="singleton", centers_sparse=False):
        """K1, K2, all-reduce, K3, then the reference's EmptyAction (kmeans_sparsified.m:432-445).
        centers_sparse: take the sparse-centres branch of the operator (findClusterAssignments.m:63-75), as the
        reference does while the centres are stored sparse (kmeans_sparsified.m:460-464).
        Returns the IterStats of the iteration (identical on every rank)."""
        e = self.e
        if centers_sparse:
            e.assign_sparse(gamma_dist)
        else:
            e.assign(gamma_dist)
        part = e.accumulate()
        self._allreduce(part)
        st = e.finalize(gamma_update, ml_correction)
        if st.n_empty:
            action = str(empty_action).lower()
            if action == "error":
                raise RuntimeError("One cluster lost all its members")
            if action == "singleton":
                # global first-occurrence arg-max of the distances: ranks own ascending column
                # blocks, so the lowest rank wins ties, then the lowest local index
                v, j = e.argmax_distance() if e.n_local else (-np.inf, -1)
                cands = self._gather_objects((float(v), int(j)))
                owner, best = 0, -np.inf
                for r, (cv, cj) in enumerate(cands):
                    if cj >= 0 and (cv > best or (np.isnan(best) and not np.isnan(cv))):
                        owner, best = r, cv
                dist = self._dist()
                me = dist.get_rank(self.group) if dist is not None else 0
                col = e.get_column(cands[owner][1]) if me == owner else None
                col = self._bcast_object(col, owner)
                for k in np.flatnonzero(e.counts() == 0):
                    e.set_center_column(int(k), col)
                st2 = e.refresh_diff()
                st.dff, st.has_nan = st2.dff, st2.has_nan
            elif action == "drop":
                # kmeans_sparsified.m:454-459: the empty centres are removed and K shrinks.  Every rank sees the same
                # global counts, so every rank drops the same columns.
                if not hasattr(e, "drop_centers"):
                    raise NotImplementedError("EmptyAction='drop' needs an engine with drop_centers(keep)")
                keep = np.flatnonzero(e.counts() > 0)
                e.drop_centers(keep)
            else:
                raise ValueError("invalid EmptyAction choice")
        return st

    def run(self, centers, gamma_dist, gamma_update, max_iter=100, tol=1e-6, ml_correction=True,
            empty_action="singleton", centers_sparse=False):
        """Iterate to convergence; returns (iterations, last IterStats).  centers_sparse=True starts in the
        sparse-centres branch (centres that are columns of X: Start='sample' / k-means++ without denseCenters)
        and leaves it once more than 99% of the centre entries are non-zero (kmeans_sparsified.m:460-464)."""
        self.e.set_centers(centers)
        st, its = None, 0
        for its in range(1, max_iter + 1):
            st = self.step(gamma_dist, gamma_update, ml_correction, empty_action, centers_sparse)
            if centers_sparse:
                cen = self.e.get_centers()
                if np.count_nonzero(cen) / max(cen.size, 1) > 0.99:
                    centers_sparse = False
            if st.dff < tol:
                break
            if st.has_nan:
                raise RuntimeError("Found NaN in centers")
        return its, st


def sharded_arthur_initialization(engine, K: int, gamma, n_total: int, lo: int, first: int, uniforms, group=None):
    """k-means++ (private/Arthur_initialization.m:24-69) over column shards.

    Every rank passes the SAME `first` (global index of the first centre, :35) and the same
    iterable of uniform numbers (one per randsample call, :50,:56).  Per round: each rank folds
    the distance to the newest centre into its running minimum and returns its local sum of D^2;
    the sums are all-gathered; the rank whose cumulative range contains u*total finds the local
    index by a prefix search; it broadcasts the global index and the column.  Returns
    (global indices int64[K], centres p x K) on every rank.  `engine` needs n_local, get_column,
    kpp_update(center, gamma, first) -> local sum, kpp_pick(target) -> local index."""
    import torch.distributed as dist
    on = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    world = dist.get_world_size(group) if on else 1
    me = dist.get_rank(group) if on else 0
    bounds = [None] * world
    if on:
        dist.all_gather_object(bounds, (int(lo), int(lo + engine.n_local)), group=group)
    else:
        bounds[0] = (int(lo), int(lo + engine.n_local))
    p = engine.p
    if on:
        import torch
        # two small tensor collectives per round (NCCL on the GPU, gloo on the CPU): the local D^2 sums
        # (all-gather of one double) and [global index | column] from the rank that owns the pick
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        g_in = torch.zeros(1, dtype=torch.float64, device=dev)
        g_out = torch.zeros(world, dtype=torch.float64, device=dev)
        box = torch.zeros(1 + p, dtype=torch.float64, device=dev)

    def owner_of(g):
        for r, (a, b) in enumerate(bounds):
            if a <= g < b:
                return r
        raise ValueError(f"global column {g} is owned by no rank")

    def publish(src, g_local):
        """rank `src` owns local column g_local: every rank gets (global index, dense column)"""
        if not on:
            return bounds[0][0] + int(g_local), engine.get_column(int(g_local))
        if src == me:
            col = np.asarray(engine.get_column(int(g_local)), dtype=np.float64)
            box[0] = float(bounds[me][0] + int(g_local))
            box[1:] = torch.from_numpy(col).to(dev)
        dist.broadcast(box, src=dist.get_global_rank(group, src) if group is not None else src, group=group)
        h = box.cpu().numpy()
        return int(h[0]), h[1:].copy()

    it = iter(uniforms)
    r0 = owner_of(int(first))
    g0, c0 = publish(r0, int(first) - bounds[r0][0])
    chosen = [g0]
    cols = [c0]
    for _ in range(K - 1):
        local = engine.kpp_update(cols[-1], gamma, first=(len(chosen) == 1)) if engine.n_local else 0.0
        if on:
            g_in[0] = float(local)
            dist.all_gather_into_tensor(g_out, g_in, group=group)
            sums = [float(v) for v in g_out.cpu().numpy()]
        else:
            sums = [float(local)]
        total = 0.0
        prefix = []
        for v in sums:                       # same summation order on every rank
            prefix.append(total)
            total += v

        def pick():
            u = float(next(it))
            if not (total > 0):              # all distances zero: uniform (Arthur_initialization.m:44-48)
                g = min(int(u * n_total), n_total - 1)
                q = owner_of(g)
                return publish(q, g - bounds[q][0])
            target = u * total
            q = world - 1
            for r in range(world):
                if target < prefix[r] + sums[r] and sums[r] > 0:
                    q = r
                    break
            j = engine.kpp_pick(target - prefix[q]) if q == me else 0
            return publish(q, j)

        g, col = pick()
        counter = 1
        while g in chosen and counter < 400:  # :54-61
            g, col = pick()
            counter += 1
        if g in chosen:
            raise RuntimeError("Cannot sample with replacement with this distribution")
        chosen.append(g)
        cols.append(col)
    return np.asarray(chosen, dtype=np.int64), np.stack(cols, axis=1)
