"""kmeans_sparsified: host-side mirror of the reference's entry point for the sparsified path.

    [IDX, C, SUMD, D, OUTPUT] = kmeans_sparsified(X, K, 'Name', value, ...)   (kmeans_sparsified.m:1)

Same option names, defaults and error behaviour as the reference's inputParser
(kmeans_sparsified.m:130-156).  The preconditioning, the per-iteration assignment, the
per-cluster sums and the centre update run on the GPU through libskm_b200; this file only
sequences them the way kmeans_sparsified.m:213-607 does.

Scope: the sparsified path (`Sparsify=True`) with the Hadamard sketch, the DCT
sketch (p not a power of two, :226-231) or 'none', from memory or from a file on disk (`DataFile`, a
MATLAB -v7.3 .mat read by matfile73.py, or a memory-mapped .npy).  The dense path
(Sparsify=False) is outside the hot path and raises NotImplementedError.  The two-pass outputs (`nargout` 6..9:
centers_twoPass, assignments_twoPass, distances_twoPass, SUMD_twoPass; kmeans_sparsified.m:525-571)
come from one streamed pass over the original data (skm_second_pass).

Randomness: MATLAB's generators are closed source, so the random draws (Rademacher signs, row
samples, k-means++ picks) come from `numpy.random.default_rng(Seed)`; every draw can also be
supplied explicitly (`Signs`, `SampleRows`, `StartIndices`) so a run can be reproduced
bit-for-bit against the oracle.
"""
from __future__ import annotations

import math
import time
import warnings

import numpy as np

from . import ops
from .engine import Context, Dataset, Lloyd, default_context

__version__ = 2.1          # the reference's version number (kmeans_sparsified.m:121)

_DEFAULTS = dict(
    Replicates=1, Start="Arthur", MaxIter=100, Display=False, PrintEvery=10, Tol=1e-6,
    Sparsify=False, SparsityLevel=0.01, SketchType="auto", EmptyAction="singleton",
    ColumnSamples=False, MLcorrection=True, DataFile=None, MB_limit=500, DataFileVerbose=False,
    SparsityIgnoreUpsampling=False, FORCE_BUG=False, tryBuiltinMex=True, unbiasedDistance=True,
    unbiasedInitialization=True, denseCenters=False,
)
_EXTRA = dict(Seed=None, Signs=None, SampleRows=None, StartIndices=None, Store="f32", Device=0, Devices=None,
              MixDtype="f64", nargout=5, Context=None, Pipeline="auto", IncrementalUpdate=True, BoundedAssign=True)


class SkmWarning(UserWarning):
    """warning('id','msg') of the reference; `identifier` carries the id (e.g. kmeans_sparsified:dropCluster,
    kmeans_sparsified.m:433)."""

    def __init__(self, msg, identifier=""):
        super().__init__(msg)
        self.identifier = identifier


class KMeansError(RuntimeError):
    """error('id','msg') of the reference; `identifier` carries the id."""

    def __init__(self, msg, identifier=""):
        super().__init__(msg)
        self.identifier = identifier


def matlab_round(v: float) -> int:
    """MATLAB round(): half away from zero."""
    return int(math.floor(abs(v) + 0.5) * (1 if v >= 0 else -1))


def _nextpow2_size(p: int) -> int:
    p2 = 1
    while p2 < p:
        p2 <<= 1
    return p2


def randsample_block(rng, n: int, k: int, n_rep: int) -> np.ndarray:
    """(k, n_rep) 0-based distinct row indices per column, uniform without replacement
    (contract of private/randsample_block.m:44-84; the order inside a column is irrelevant
    because sparse() sorts, randsample_fixedNumberEntries.m:62)."""
    out = np.empty((k, n_rep), dtype=np.int32)
    step = max(1, (64 << 20) // max(n, 1))
    for c0 in range(0, n_rep, step):
        c = min(step, n_rep - c0)
        keys = rng.random((c, n), dtype=np.float32)
        idx = np.argpartition(keys, k - 1, axis=1)[:, :k] if k < n else np.tile(np.arange(n), (c, 1))
        idx.sort(axis=1)
        out[:, c0:c0 + c] = idx.T
    return out


def randsample_fixedNumberEntries(Xmixed: np.ndarray, small_p: int, rows: np.ndarray):
    """Y = randsample_fixedNumberEntries(X, small_p) with the row sets `rows` (small_p, n):
    keeps X(rows(:,j), j) / (small_p/p2); exact zeros are dropped as sparse() does
    (private/randsample_fixedNumberEntries.m:30-31,62)."""
    import scipy.sparse as sp
    p2, n = Xmixed.shape
    level = small_p / p2
    cols = np.repeat(np.arange(n), small_p)
    r = np.asarray(rows).T.reshape(-1)
    vals = Xmixed[r, cols] / level
    Y = sp.csc_matrix((vals, (r, cols)), shape=(p2, n))
    Y.sum_duplicates()
    Y.eliminate_zeros()
    Y.sort_indices()
    return Y


def Arthur_initialization(ds: Dataset, K: int, gamma, rng=None, first=None, uniforms=None):
    """k-means++ on the resident sparsified matrix (private/Arthur_initialization.m:24-69).

    Each round folds the masked distance to the newest centre into a running minimum on the
    GPU (identical values to the reference's recomputation against all chosen centres), then
    samples the next index with probability proportional to the squared distance.  Returns the
    chosen 0-based column indices."""
    n = ds.n
    rng = rng or np.random.default_rng()
    it = iter(uniforms) if uniforms is not None else None

    def draw():
        return float(next(it)) if it is not None else float(rng.random())

    chosen = [int(first) if first is not None else int(rng.integers(n))]           # :35
    for _ in range(K - 1):
        # with gamma the reference densifies the centres (full(ref), :28); without, they stay sparse and
        # findClusterAssignments takes its sparse-centres branch (:26, findClusterAssignments.m:70-74)
        tot = ds.kpp_update(ds.get_column(chosen[-1]), gamma, first=(len(chosen) == 1), sparse_center=gamma is None)
        if not (tot > 0):                                                          # :44-48 all-zero distances
            pick = lambda: min(int(draw() * n), n - 1)                              # noqa: E731
        else:
            pick = lambda: ds.kpp_pick(draw() * tot)                                # noqa: E731
        i = pick()
        counter = 1
        while i in chosen and counter < 400:                                       # :54-61
            i = pick()
            counter += 1
        if i in chosen:
            raise KMeansError("Cannot sample with replacement with this distribution")
        chosen.append(i)
    return np.asarray(chosen, dtype=np.int64)


def _open_data_file(path: str):
    """DataFile mode (kmeans_sparsified.m:180-207, private/sampleAndMixFromLargeFile.m:60-107): the matrix
    stays on disk and is read in column chunks.  The reference reads a MATLAB -v7.3 .mat through `matfile`
    (HDF5): `.mat` paths go through matfile73.py, a reader of the HDF5 subset MATLAB writes for a dense matrix
    (contiguous datasets are memory-mapped in place, chunked / deflated ones are decoded into a memory-mapped
    temporary file).  A NumPy .npy file (2-D, float32 or float64), memory-mapped, is accepted as well.  Either way
    the precondition+sample pipeline and the second pass stream the matrix chunk by chunk without loading it."""
    import os
    cand = [path, path + ".mat", path + ".npy"]                                       # :187-189: '.mat' is appended
    found = next((c for c in cand if os.path.isfile(c)), None)
    if found is None:
        raise KMeansError("Cannot find specified data file to load")                  # :191
    if not found.endswith(".npy"):
        from . import matfile73
        try:
            A, _ = matfile73.open_matrix(found)
        except matfile73.MatFileError as e:
            raise KMeansError(str(e)) from e
        if A.ndim != 2 or A.shape[0] < 1 or A.shape[1] < 1:
            raise KMeansError("Error reading file; returned bad size for matrix")     # :204
        if A.dtype not in (np.float32, np.float64):
            A = np.asarray(A, dtype=np.float64)                                       # integer classes: as MATLAB's double()
        return A
    A = np.load(found, mmap_mode="r")
    if A.ndim != 2 or A.shape[0] < 1 or A.shape[1] < 1:
        raise KMeansError("Error reading file; returned bad size for matrix")         # :204
    if A.dtype not in (np.float32, np.float64):
        raise KMeansError("the data file must hold float32 or float64 values")
    return A


def kmeans_sparsified(X=None, K=None, **opts):
    """See the module docstring.  Returns (IDX, C, SUMD, D, OUTPUT)."""
    if X is None and K is None:                                                    # :120-125
        print(f"Sparsified K-Means, Version {__version__:.1f}, June 1 2016")
        return __version__
    t0 = time.perf_counter()
    unknown = set(opts) - set(_DEFAULTS) - set(_EXTRA)
    if unknown:
        raise KMeansError(f"'{sorted(unknown)[0]}' is not a recognized parameter", "MATLAB:InputParser:UnmatchedParameter")
    o = {**_DEFAULTS, **_EXTRA, **opts}
    if not (0 < o["SparsityLevel"] <= 1):
        raise KMeansError("The value of 'SparsityLevel' is invalid", "MATLAB:InputParser:ArgumentFailedValidation")
    if str(o["EmptyAction"]).lower() not in ("singleton", "error", "drop"):
        raise KMeansError("The value of 'EmptyAction' is invalid", "MATLAB:InputParser:ArgumentFailedValidation")
    load_from_disk = isinstance(X, str) or bool(o["DataFile"])
    if load_from_disk:                                                                # :180-207
        if not o["Sparsify"]:
            raise KMeansError('No reason to turn on "LoadFromDisk" option if not sampling')
        if not isinstance(X, str) and X is not None:
            warnings.warn('Loading data from disk, ignoring "X" input. Are you sure code is OK?')
        X = _open_data_file(X if isinstance(X, str) else o["DataFile"])
    if not o["Sparsify"]:
        raise NotImplementedError("Sparsify=false (dense K-means through pdist2) is outside the sparsified hot path")
    import scipy.sparse as sp
    rng = np.random.default_rng(o["Seed"])
    # Devices=[0,1,...] (or an int count): the columns are sharded over several GPUs of this box and driven from
    # this one process (skm_multi_*, multi.py); everything below runs unchanged on the sharded handles
    mctx = None
    if o["Devices"] is not None and not (isinstance(o["Devices"], (list, tuple)) and len(o["Devices"]) == 1):
        from .multi import MultiContext, MultiDataset, MultiLloyd
        mctx = o["Devices"] if isinstance(o["Devices"], MultiContext) else MultiContext(o["Devices"])
    elif o["Devices"] is not None:
        o["Device"] = int(o["Devices"][0])
    ctx: Context = o["Context"] or (mctx.contexts[0] if mctx is not None else default_context(int(o["Device"])))
    own_mctx = mctx is not None and not isinstance(o["Devices"], type(mctx))
    OUTPUT_devices = mctx.devices if mctx is not None else [ctx.device]
    OUTPUT = {"LoadFromDisk": load_from_disk, "Options": {k: o[k] for k in _DEFAULTS}, "Sparsify": True}

    if np.iscomplexobj(X):
        raise KMeansError("Code and distance computations require real data")        # :311-313
    if isinstance(X, np.ndarray) and X.dtype in (np.float32, np.float64):
        Xd = X                            # used in place (a memory-mapped file is streamed chunk by chunk)
    else:
        Xd = np.asarray(X.todense() if sp.issparse(X) else X, dtype=np.float64)
    if not o["ColumnSamples"]:
        Xd = Xd.T                                                                     # :213-215
    p, n = Xd.shape
    K = int(K)
    if n < K:
        raise KMeansError("X must have more samples than the number of clusters.", "kmeans_sparsified:badDimensions")

    # ---- sketch selection (:224-296) ----
    sketch = o["SketchType"]
    if isinstance(sketch, str) and sketch.lower() == "auto":
        sketch = "Hadamard" if p == _nextpow2_size(p) else "DCT"
        OUTPUT["SketchType"] = sketch
    H_user = Ht_user = None
    if isinstance(sketch, (list, tuple)):                                             # :263-270 cell of two handles
        if len(sketch) == 2 and callable(sketch[0]) and callable(sketch[1]):
            H_user, Ht_user = sketch
            sketch = "function handles"
        else:
            raise KMeansError("If SketchType is a cell, then both entries should be function handles for forward "
                              "and adjoint transform")
    if not isinstance(sketch, str):
        raise KMeansError('bad type for "SketchType"')
    sk = sketch.lower()
    p2 = p
    if H_user is not None:
        sk = "handles"
    elif sk == "hadamard":
        p2 = _nextpow2_size(p)
        OUTPUT["SlowHadamard"] = False
    elif sk == "dct":
        if p > 32768:
            raise NotImplementedError("the DCT sketch is applied as a dense p x p product; p <= 32768")
    elif sk not in ("nothing", "none"):
        raise KMeansError('bad type for "SketchType"')
    if sk in ("hadamard", "dct", "handles"):
        d = o["Signs"]
        if d is None:
            d = np.ones(p2) if o["FORCE_BUG"] else np.sign(rng.standard_normal(p2))    # :283-287
            d[d == 0] = 1.0
        d = np.asarray(d, dtype=np.float64).reshape(-1)
        if d.shape[0] != p2:
            raise ValueError("Signs must have 2^nextpow2(p) entries")
    else:
        d = None
    scale_eps = 1 + 2 * np.finfo(np.float64).eps                                      # :281 / :292 (applied below)

    def mix(A):                                                                       # :295
        if d is None:
            return np.asarray(A, dtype=np.float64)
        if sk == "handles":                                                           # user transform on the host
            return np.asarray(H_user(d.reshape(-1, 1) * np.asarray(A, dtype=np.float64)), dtype=np.float64)
        if sk == "dct":
            from .engine import dct_mix
            return dct_mix(A, d, False, ctx)                                          # H = dct, :257
        from .engine import mix_hadamard
        return mix_hadamard(A, d, o["MixDtype"], ctx)

    def unmix(C):                                                                     # :296
        if d is None:
            return C
        if sk == "handles":
            return d.reshape(-1, 1) * np.asarray(Ht_user(C), dtype=np.float64)
        if sk == "dct":
            from .engine import dct_mix
            return dct_mix(C, d, True, ctx)                                           # Ht = idct, :258
        Y = ops.hadamard(C, ctx) / math.sqrt(p2)
        return (d.reshape(-1, 1) * Y)[:p, :]

    small_p = max(1, matlab_round(o["SparsityLevel"] * p2))                           # :325-331
    gamma = small_p / p
    display = str(o["Display"]).lower() if o["Display"] else "off"
    pipeline = str(o["Pipeline"]).lower()
    if pipeline == "auto":
        # all-GPU precondition + sample when nothing pins the random rows and the sizes allow it
        ok = (o["SampleRows"] is None and 32 <= p2 <= 32768) if sk == "hadamard" else (sk == "dct")
        if sk == "handles":
            ok = False                                   # a user transform runs on the host
        pipeline = "device" if (d is not None and ok and o["Store"] == "f32") else "host"
    if pipeline == "device":
        if d is None or sk == "handles" or (sk == "hadamard" and o["SampleRows"] is not None):
            raise ValueError("Pipeline='device' needs the Hadamard sketch with on-device row sampling, or the DCT sketch")
        t1 = time.perf_counter()
        seed = int(rng.integers(0, 2 ** 63 - 1))
        if mctx is not None:
            ds = MultiDataset.from_dense_host(Xd, d, small_p, seed=seed, mctx=mctx, dct=(sk == "dct"), rows=o["SampleRows"])
        elif sk == "dct":
            ds = Dataset.from_dense_host_dct(Xd, d, small_p, seed=seed, rows=o["SampleRows"], ctx=ctx)
        else:
            ds = Dataset.from_dense_host(Xd, d, small_p, seed=seed, ctx=ctx)      # applies *(1+2eps) itself
        OUTPUT["TimeToSketch"] = time.perf_counter() - t1
        OUTPUT["TimeToSample"] = 0.0                                                  # fused into the sketch
        Xs = None
    else:
        t1 = time.perf_counter()
        Xm = mix(Xd * scale_eps)
        OUTPUT["TimeToSketch"] = time.perf_counter() - t1
        t1 = time.perf_counter()
        rows = o["SampleRows"]
        if rows is None:
            rows = randsample_block(rng, p2, small_p, n)
        Xs = randsample_fixedNumberEntries(Xm, small_p, np.asarray(rows))             # :334
        del Xm
        OUTPUT["TimeToSample"] = time.perf_counter() - t1
        ds = (MultiDataset.from_scipy(Xs, store=o["Store"], mctx=mctx) if mctx is not None
              else Dataset.from_scipy(Xs, store=o["Store"], ctx=ctx, K_hint=K))
    OUTPUT["Pipeline"] = pipeline
    OUTPUT["Devices"] = OUTPUT_devices
    make_lloyd = (lambda Kx, inc, bnd: MultiLloyd(ds, Kx, incremental=inc, bounded=bnd)) if mctx is not None else \
                 (lambda Kx, inc, bnd: Lloyd(ds, Kx, incremental=inc, bounded=bnd))
    if display in ("iter", "final"):
        print(f"Randomly mixing of type {sketch}")
        print("Randomly taking %.1f%% of the data; actual dataset is %.1f%% sparse"
              % (100 * gamma, 100 * ds.nnz / (ds.p * max(ds.n, 1))))

    def columns(ind):
        """dense p2 x len(ind) matrix of the sparsified columns `ind` (X(:,ind) in the reference)"""
        if Xs is not None:
            return np.asarray(Xs[:, ind].todense())
        return np.stack([ds.get_column(int(j)) for j in ind], axis=1)

    ml = bool(o["MLcorrection"])
    g_dist = gamma if o["unbiasedDistance"] else None                                 # :369-373
    start = o["Start"]
    R = int(o["Replicates"])
    OUTPUT.update(iterations=np.zeros(R, dtype=np.int64), stoppingDiff=np.zeros(R), objectives=np.zeros(R),
                  replicateTimes=np.zeros(R), replicateTimesJustInitialization=np.zeros(R))
    best = dict(obj=math.inf)
    distances = None
    L = None
    try:
        for trial in range(R):
            t1 = time.perf_counter()
            centers_sparse = False
            if isinstance(start, str):
                s = start.lower()
                if s == "sample":
                    ind = rng.choice(n, size=K, replace=False) if o["StartIndices"] is None else np.asarray(o["StartIndices"])
                    centers = columns(ind)
                    centers_sparse = True
                elif s == "uniform":
                    mn, mx = (float(Xs.min()), float(Xs.max())) if Xs is not None else ds.minmax()   # :388-389
                    centers = (mx - mn) * rng.random((p2, K)) - mn                     # :390 (sign as in the reference)
                elif s in ("arthur", "++", "kmeans++", "k-means++", "k-means-++"):
                    if o["StartIndices"] is not None:
                        ind = np.asarray(o["StartIndices"], dtype=np.int64)
                    else:
                        ind = Arthur_initialization(ds, K, gamma if o["unbiasedInitialization"] else None, rng)
                    centers = columns(ind)
                    centers_sparse = True
                else:
                    raise KMeansError('cannot handle other types of "Start" values')
            else:
                st = np.asarray(start, dtype=np.float64)
                if not o["ColumnSamples"]:
                    st = st.T
                centers = mix(st)                                                     # :401-405
                if R > 1:
                    warnings.warn(SkmWarning("initialization is specified, so running more than 1 replicate is not helpful",
                                             "kmeans_sparsified:deterministicCenters"))
            if o["denseCenters"]:
                centers_sparse = False                                                # :412-414
            OUTPUT["replicateTimesJustInitialization"][trial] = time.perf_counter() - t1

            Kt = K
            # IncrementalUpdate: the per-cluster sums only depend on the assignments, so from the second
            # iteration on only the columns that changed cluster move their entries (same centres up to fp64
            # rounding; False recomputes them from all columns every iteration like kmeans_sparsified.m:430-453)
            # BoundedAssign: bounds carried across iterations let a column keep its centre after one centre
            # evaluation once the centres move little (same assignments; fp32 datasets only)
            # (both modes add a host synchronisation per iteration, which only pays off on matrices of some size)
            big = ds.nnz >= 2_000_000
            bounded = bool(o["BoundedAssign"]) and ds.store_dtype == "f32" and big
            incremental = bool(o["IncrementalUpdate"]) and big
            L = make_lloyd(Kt, incremental, bounded)
            L.set_centers(centers)
            its = 0
            dff = obj = math.nan
            assignments = None
            dropped_last = False
            for its in range(1, int(o["MaxIter"]) + 1):
                dropped_last = False
                if centers_sparse:
                    L.assign_sparse(g_dist)                                           # findClusterAssignments.m:63-75
                else:
                    L.assign(g_dist)                                                  # :420
                L.accumulate()
                stt = L.finalize(gamma, ml)                                           # :447-450
                if stt.n_empty:
                    action = str(o["EmptyAction"]).lower()
                    warnings.warn(SkmWarning("cluster has lost all its members", "kmeans_sparsified:dropCluster"))   # :433
                    counts = L.counts()
                    empty = np.flatnonzero(counts == 0)
                    if action == "singleton":                                         # :434-437
                        _, imax = L.argmax_distance()
                        col = ds.get_column(imax)
                        for ki in empty:
                            L.set_center_column(int(ki), col)
                        stt2 = L.refresh_diff()
                        stt.dff, stt.has_nan = stt2.dff, stt2.has_nan
                    elif action == "error":
                        raise KMeansError("One cluster lost all its members")
                    else:                                                             # drop, :454-459
                        keep = np.flatnonzero(counts > 0)
                        cen = L.get_centers()[:, keep]
                        old = L.get_centers_old()[:, keep]
                        _, distances = L.assignments()
                        L.close()
                        Kt = keep.size
                        L = make_lloyd(Kt, incremental, bounded)
                        L.set_centers(cen)
                        assignments = np.zeros(0, dtype=np.int32)                     # :457 assignments = []
                        dropped_last = True
                        stt.dff = float(np.linalg.norm(old - cen, "fro"))
                        stt.has_nan = bool(np.isnan(cen).any())
                if centers_sparse:                                                    # :460-464
                    cen = L.get_centers()
                    if np.count_nonzero(cen) / cen.size > 0.99:
                        centers_sparse = False
                dff, obj = stt.dff, stt.objective                                     # :470-471
                if display == "iter" and its % int(o["PrintEvery"]) == 0:
                    print("Iter: %3d; change in cluster centers: %.2e; objective: %.2e" % (its, dff, obj))
                if dff < o["Tol"]:                                                    # :476-478
                    break
                if stt.has_nan:
                    raise KMeansError("Found NaN in centers")                         # :480-484
            if not dropped_last:
                assignments, distances = L.assignments()
            OUTPUT["replicateTimes"][trial] = time.perf_counter() - t1
            OUTPUT["stoppingDiff"][trial] = dff
            OUTPUT["objectives"][trial] = obj
            OUTPUT["iterations"][trial] = its
            is_best = obj < best["obj"]
            if is_best:                                                               # :493-503
                best = dict(obj=obj, assignments=assignments, distances=distances, centers=L.get_centers(), K=Kt)
            if display == "iter" or (display == "final" and is_best):
                tail = " (this is the best trial so far)" if is_best else " (best so far was %.2e)" % best["obj"]
                print("Trial %3d of %3d total, objective %.2e%s" % (trial + 1, R, obj, tail))
            L.close()
    finally:
        if L is not None:
            L.close()
        ds.close()
    if "assignments" not in best:
        raise KMeansError("no replicate produced a finite objective")
    OUTPUT["TimeInitialization"] = float(np.sum(OUTPUT["replicateTimesJustInitialization"]))
    OUTPUT["TimeAlgo_wo_initialization"] = float(np.sum(OUTPUT["replicateTimes"])) - OUTPUT["TimeInitialization"]

    # SUMD uses the LAST trial's distances with the best assignments, as the reference does (:514-518)
    Kb = best["K"]
    SUMD = np.zeros(Kb)
    if best["assignments"].size:
        for ki in range(Kb):
            SUMD[ki] = np.sum(distances[best["assignments"] == ki + 1] ** 2)
    OUTPUT["TimeOverall_OnePass"] = time.perf_counter() - t0
    C = unmix(best["centers"])                                                        # :523
    IDX, D = best["assignments"], best["distances"]
    two = None
    nargout = int(o["nargout"])
    if nargout > 5:                                                                   # :525-571, in-core arm :542-560
        # one streamed pass over the original data X*(1+2eps) (= XFull, :310): per-cluster means of the
        # best assignments (:545-551) and, for nargout > 6, dense re-assignment against the unmixed
        # centres (:558).  The reference only builds the assignments when nargout > 6.
        from .engine import second_pass
        if IDX.size != n:
            raise KMeansError("two-pass outputs need assignments for every sample (EmptyAction='drop' removed them)")
        t1 = time.perf_counter()
        res = second_pass(Xd, centers=C if nargout > 6 else None, assign_in=IDX, scale=scale_eps,
                          want_assign=nargout > 6, want_dist=nargout > 6, ctx=ctx)
        OUTPUT["TimeSecondPass_Overall"] = time.perf_counter() - t1
        # the reference times its two in-core passes separately (:551, :560); here both happen in ONE streamed
        # pass over the data, so the same wall time is reported under both names
        OUTPUT["TimeSecondPass_Centers"] = OUTPUT["TimeSecondPass_Overall"]
        if nargout > 6:
            OUTPUT["TimeSecondPass_Assignments"] = OUTPUT["TimeSecondPass_Overall"]
        C2 = res["centers"]
        if C2.shape[1] < Kb:                                                          # labels never reached Kb
            C2 = np.concatenate([C2, np.zeros((p, Kb - C2.shape[1]))], axis=1)
        two = [C2]
        if nargout > 6:
            IDX2, D2 = res["assign"], res["dist"]
            OUTPUT["SecondPassRechecked"] = res["n_rechecked"]
            two += [IDX2, D2]
            if nargout >= 9:                                                          # :562-570
                SUMD2 = np.zeros(Kb)
                for ki in range(Kb):
                    # the reference sums the ONE-pass `distances` here (not distances_twoPass), :567
                    SUMD2[ki] = np.sum(distances[IDX2 == ki + 1] ** 2)
                two.append(SUMD2)
                OUTPUT["TimeSecondPass_SUMD"] = 0.0
    if not o["ColumnSamples"]:                                                        # :586-605
        C = C.T
        if two is not None:
            two[0] = two[0].T
    OUTPUT["TimeOverall"] = time.perf_counter() - t0
    if own_mctx:
        mctx.close()
    if two is None:
        return IDX, C, SUMD, D, OUTPUT
    return (IDX, C, SUMD, D, OUTPUT, *two)
