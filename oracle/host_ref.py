"""float64 numpy restatement of the reference's MATLAB host logic on the hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  MATLAB/Octave are not in
this image, so the `.m` files cannot be executed; every function below follows
the cited lines of the reference one statement at a time.  The native kernels
they call are NOT restated here: they go through the reference's own compiled C
(oracle/_ref via refmex) when present, else through the pinned C port (cport).

Random draws are factored out: callers pass the Rademacher signs, the sampled
row sets, the start centres or the uniform numbers k-means++ consumes, because
the reference draws them from MathWorks' closed-source generators
(`randn`, `randperm`, `randsample`), which no reference test pins
("parity unpinned" at those boundaries).

Conventions: X is a scipy.sparse.csc_matrix of shape (p, n) with sorted row
indices (points are columns, as inside kmeans_sparsified.m after :213-218);
assignments are 1-based like MATLAB's.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp

from . import cport, refmex


# ---------------------------------------------------------------------------
# native-kernel dispatch (reference binary when available)
# ---------------------------------------------------------------------------

def _smmc(p, n, jc, ir, x, centers):
    if refmex.ref_available("SparseMatrixMinusCluster"):
        return refmex.SparseMatrixMinusCluster(p, n, jc, ir, x, centers)
    return cport.masked_dist(p, n, jc, ir, x, centers)


def _hadamard(A):
    if refmex.ref_available("hadamard"):
        return refmex.hadamard(A)
    return cport.hadamard(A)


def matlab_round(v: float) -> int:
    """MATLAB round(): half away from zero (Python's round is half-to-even)."""
    return int(math.floor(abs(v) + 0.5) * (1 if v >= 0 else -1))


def nextpow2_size(p: int) -> int:
    """2^nextpow2(p) (kmeans_sparsified.m:242)."""
    p2 = 1
    while p2 < p:
        p2 <<= 1
    return p2


def as_csc(X) -> sp.csc_matrix:
    X = sp.csc_matrix(X, dtype=np.float64)
    X.sum_duplicates()
    X.eliminate_zeros()      # MATLAB sparse matrices never hold explicit zeros
    X.sort_indices()
    return X


# ---------------------------------------------------------------------------
# findClusterAssignments.m, sparse-X branches
# ---------------------------------------------------------------------------

def find_cluster_assignments(X: sp.csc_matrix, centers, gamma=None, centers_sparse=None):
    """[assignments, distances] = findClusterAssignments(X, centers, [], gamma).

    Follows /root/reference/private/findClusterAssignments.m:53-83 and :168-171.
    `centers` is a dense (p, k) array; `centers_sparse` says whether MATLAB would
    hold it as a sparse matrix (then exact zeros are structural zeros and the
    per-centre row-subset path :63-75 is taken).  If `centers` is a scipy sparse
    matrix, centers_sparse defaults to True.
    Returns (assign 1-based int64 (n,), dist float64 (n,), D float64 (k, n)).
    """
    if sp.issparse(centers):
        if centers_sparse is None:
            centers_sparse = True
        centers = np.asarray(centers.todense(), dtype=np.float64)
    centers = np.asarray(centers, dtype=np.float64)
    if centers.ndim == 1:
        centers = centers.reshape(-1, 1)
    p, n = X.shape
    pp, k = centers.shape
    if p != pp:
        raise ValueError("Array of centers not of correct size")       # :55
    if centers_sparse:
        D = np.zeros((k, n))                                           # :57
        Xr = X.tocsr()
        for ki in range(k):                                            # :64 / :71
            ind = np.flatnonzero(centers[:, ki])                       # find(centers(:,ki))
            sub = Xr[ind, :].tocsc()
            sub.sort_indices()
            cfull = centers[ind, ki]
            if gamma is not None:
                gamma_center = ind.size / p                            # :67 nnz/size(centers,1)
                vals = sub.data / gamma_center
                cvec = cfull / gamma                                   # :68
            else:
                vals = sub.data
                cvec = cfull
            if ind.size == 0:
                D[ki, :] = 0.0    # empty row subset: every masked sum is empty -> sqrt(0)
            else:
                D[ki, :] = _smmc(ind.size, n, sub.indptr, sub.indices, vals,
                                 cvec.reshape(-1, 1))[0]
    else:
        c = centers / gamma if gamma is not None else centers          # :78 division, not reciprocal
        D = _smmc(p, n, X.indptr, X.indices, X.data, c)
    dist, assign = cport.colmin(D)                                     # :169 first-occurrence min
    return assign, dist, D


def find_cluster_assignments_dense(X: np.ndarray, centers: np.ndarray):
    """Dense-X branch of findClusterAssignments, the non-pdist2 arm
    (/root/reference/private/findClusterAssignments.m:154-163, :168-175):
    distances(k,:) = nrm2 - 2*(X'*c_k)' + norm(c_k)^2, min over k (first occurrence), then
    sqrt(max(0, .)) of the minimum only.  float64; the BLAS summation order of X'*c_k is not
    pinned by the reference (parity unpinned at the 1e-15 level).
    Returns (assign 1-based (n,), dist (n,), D2 (k, n) squared distances)."""
    X = np.asarray(X, dtype=np.float64)
    centers = np.asarray(centers, dtype=np.float64)
    p, n = X.shape
    k = centers.shape[1]
    nrm2 = np.sum(X * X, axis=0)                                         # :155
    D2 = np.empty((k, n))
    for ki in range(k):
        D2[ki, :] = nrm2 - 2.0 * (X.T @ centers[:, ki]) + np.linalg.norm(centers[:, ki]) ** 2   # :157
    assign = np.argmin(D2, axis=0) + 1                                   # :169 (first occurrence)
    dist = np.sqrt(np.maximum(0.0, D2[assign - 1, np.arange(n)]))        # :173
    return assign, dist, D2


def second_pass(XFull: np.ndarray, best_centers: np.ndarray, best_assign1: np.ndarray, K: int):
    """The in-core two-pass block of kmeans_sparsified.m:542-560: XFull is the ORIGINAL data after
    X*(1+2*eps) (:292, :310), best_centers the unmixed centres of the sparsified run (:523).
    Returns (centers_twoPass (p,K), assignments_twoPass, distances_twoPass)."""
    p, n = XFull.shape
    c2 = np.zeros((p, K))                                                # :545
    for ki in range(K):
        ind = np.flatnonzero(best_assign1 == ki + 1)                     # :548
        if ind.size:
            c2[:, ki] = np.mean(XFull[:, ind], axis=1)                   # :550
    a2, d2, _ = find_cluster_assignments_dense(XFull, best_centers)      # :558
    return c2, a2, d2


# ---------------------------------------------------------------------------
# kmeans_sparsified.m: preconditioning + sampling (explicit randomness)
# ---------------------------------------------------------------------------

def mix_hadamard(X: np.ndarray, d: np.ndarray) -> np.ndarray:
    """mix(X) = H(DD*upsample(X)), H = hadamard(.)/sqrt(p2).

    Follows kmeans_sparsified.m:238-248 (upsample, H) and :286-295 (DD, mix).
    X is dense (p, n); d is the +-1 vector of length p2 = 2^nextpow2(p)."""
    X = np.asarray(X, dtype=np.float64)
    p, n = X.shape
    p2 = d.shape[0]
    if p < p2:
        X = np.vstack([X, np.zeros((p2 - p, n))])                      # :243
    return _hadamard(d.reshape(-1, 1) * X) / math.sqrt(p2)             # :248 division


def unmix_hadamard(C: np.ndarray, d: np.ndarray, p: int) -> np.ndarray:
    """unmix(X) = downsample(DD*Ht(X)), Ht = H (kmeans_sparsified.m:244,:254,:296)."""
    p2 = d.shape[0]
    Y = _hadamard(np.asarray(C, dtype=np.float64)) / math.sqrt(p2)
    return (d.reshape(-1, 1) * Y)[:p, :]


def sparsity_params(p: int, p2: int, sparsity_level: float):
    """small_p and the redefined SparsityLevel (kmeans_sparsified.m:325-331).

    Note the reference divides by the ORIGINAL p, not p2 (both branches)."""
    small_p = max(1, matlab_round(sparsity_level * p2))
    return small_p, small_p / p


def sample_fixed_entries(Xmixed: np.ndarray, rows: np.ndarray) -> sp.csc_matrix:
    """Y = randsample_fixedNumberEntries(X, small_p) with the row sets given.

    rows is (small_p, n): for each column the 0-based rows kept (distinct).
    Value kept = X(i,j) / (small_p/p2) (randsample_fixedNumberEntries.m:30-31,
    :62); `sparse()` sorts rows within a column and drops exact zeros."""
    p2, n = Xmixed.shape
    small_p = rows.shape[0]
    level = small_p / p2
    cols = np.repeat(np.arange(n), small_p)
    r = rows.T.reshape(-1)
    vals = Xmixed[r, cols] / level
    Y = sp.csc_matrix((vals, (r, cols)), shape=(p2, n))
    return as_csc(Y)


# ---------------------------------------------------------------------------
# kmeans_sparsified.m: Lloyd loop
# ---------------------------------------------------------------------------

@dataclass
class LloydResult:
    assignments: np.ndarray           # 1-based, from the LAST findClusters call (pre-update centres)
    distances: np.ndarray
    centers: np.ndarray               # post-update centres (p2, K)
    centers_sparse: bool
    iterations: int
    stopping_diff: float
    objective: float
    history: list = field(default_factory=list)   # per-iteration (dff, obj)
    dropped: list = field(default_factory=list)


class EmptyClusterError(RuntimeError):
    pass


def centroid_update_ml(X: sp.csc_matrix, assign1: np.ndarray, K: int, gamma: float):
    """Per-cluster S, N and gamma*S./(N+1e-16) (kmeans_sparsified.m:447-448).

    Column sums are evaluated sequentially in ascending column order (MATLAB's
    own order inside sparse `sum` is not documented; agreement is to rounding)."""
    p, n = X.shape
    cen, S, N, counts = cport.centroid_update(p, n, K, X.indptr, X.indices, X.data, assign1,
                                              gamma, np.zeros((p, K)), True)
    return cen, S, N, counts


def lloyd(X: sp.csc_matrix, centers, gamma, max_iter=100, tol=1e-6,
          empty_action="singleton", ml_correction=True, centers_sparse=False,
          unbiased_distance=True) -> LloydResult:
    """The replicate body of kmeans_sparsified.m:417-486 for Sparsify=true.

    `gamma` is the redefined SparsityLevel (:326-329); it is passed to
    findClusterAssignments only when unbiased_distance (:369-373) and always used
    in the ML-corrected update (:448)."""
    X = as_csc(X)
    p, n = X.shape
    centers = np.array(centers, dtype=np.float64, copy=True)
    K = centers.shape[1]
    g_dist = gamma if unbiased_distance else None
    history, dropped_all = [], []
    dff = obj = float("nan")
    assign = dist = None
    its = 0
    for its in range(1, max_iter + 1):
        assign, dist, _ = find_cluster_assignments(X, centers, g_dist, centers_sparse)  # :420
        if np.any(dist < 0):
            raise RuntimeError("Found negative distance estimates, something went wrong")
        centers_old = centers.copy()                                   # :428
        drop = []
        iMax = None
        if ml_correction:
            newc, _, _, counts = centroid_update_ml(X, assign, K, gamma)
        else:
            newc, _, _, counts = cport.centroid_update(p, n, K, X.indptr, X.indices, X.data,
                                                       assign, gamma, np.zeros((p, K)), False)
        for ki in range(K):                                            # :430
            if counts[ki] == 0:                                        # :432
                ea = empty_action.lower()
                if ea == "singleton":
                    if iMax is None:
                        iMax = int(np.argmax(dist))                    # :435 first max
                    centers[:, ki] = np.asarray(X[:, iMax].todense()).ravel()   # :436
                elif ea == "error":
                    raise EmptyClusterError("One cluster lost all its members")
                elif ea == "drop":
                    drop.append(ki)
                else:
                    raise ValueError("invalid EmptyAction choice")
            else:
                centers[:, ki] = newc[:, ki]                           # :448 / :450
        if drop:                                                       # :454-459
            keep = [k for k in range(K) if k not in drop]
            centers = centers[:, keep]
            centers_old = centers_old[:, keep]
            assign = np.zeros(0, dtype=np.int64)                       # :457 assignments = []
            K = centers.shape[1]
            dropped_all.extend(drop)
        if centers_sparse and np.count_nonzero(centers) / centers.size > 0.99:   # :460-464
            centers_sparse = False
        dff = float(np.linalg.norm(centers_old - centers, "fro"))      # :470
        obj = float(math.sqrt(np.sum(dist ** 2)))                      # :471
        history.append((dff, obj))
        if dff < tol:                                                  # :476
            break
        if np.any(np.isnan(centers)):                                  # :480
            raise RuntimeError("Found NaN in centers")
    return LloydResult(assign, dist, centers, centers_sparse, its, dff, obj, history, dropped_all)


# ---------------------------------------------------------------------------
# Arthur_initialization.m (k-means++), random draws supplied by the caller
# ---------------------------------------------------------------------------

def weighted_pick(weights: np.ndarray, u: float) -> int:
    """0-based index drawn from weights/sum(weights) by inverting the CDF at u in [0,1).

    Stands in for MathWorks' randsample(n,1,true,w) (Arthur_initialization.m:50):
    closed source, so this is OUR contract: the first i with cumsum(w)[i] > u*sum(w)."""
    cs = np.cumsum(weights, dtype=np.float64)
    tot = cs[-1]
    i = int(np.searchsorted(cs, u * tot, side="right"))
    return min(i, weights.shape[0] - 1)


def arthur_initialization(X: sp.csc_matrix, K: int, gamma, first: int, uniforms):
    """centers = Arthur_initialization(X, K, gamma) (Arthur_initialization.m:24-69).

    `first` is the 0-based index randi would return (:35); `uniforms` is an
    iterable of numbers in [0,1) consumed one per randsample call (:50,:56).
    Returns (chosen indices 0-based, centers as a csc matrix of the chosen columns).
    Each round recomputes the distance to ALL chosen centres (:39) with
    findClusterAssignments(X, full(centers), [], gamma) (:31)."""
    X = as_csc(X)
    p, n = X.shape
    if K < 1:
        raise ValueError("K must be >= 1")
    it = iter(uniforms)
    chosen = [int(first)]
    for _ in range(K - 1):
        cen = np.asarray(X[:, chosen].todense(), dtype=np.float64)
        if gamma is None:
            # findDist = findClusterAssignments(X, ref) with ref sparse (:29)
            _, dist, _ = find_cluster_assignments(X, cen, None, centers_sparse=True)
        else:
            _, dist, _ = find_cluster_assignments(X, cen, gamma, centers_sparse=False)
        if np.any(dist < 0):
            raise RuntimeError("distance has negative components! Debug please")
        w = dist ** 2 if np.linalg.norm(dist) > 0 else np.ones(n)
        i = weighted_pick(w, next(it))
        counter = 1
        while i in chosen and counter < 400:                           # :54-61
            i = weighted_pick(w, next(it))
            counter += 1
        if i in chosen:
            raise RuntimeError("Cannot sample with replacement with this distribution")
        chosen.append(i)
    return np.array(chosen, dtype=np.int64), X[:, chosen]
