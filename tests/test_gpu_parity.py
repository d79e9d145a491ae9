"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs.  Bar: bit-exact for assignments and for every fp64 operator; centroids within
1e-6 relative (the tolerance BASELINE.json's north_star states)."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import cport, host_ref, refmex
from tests.util import make_sparsified

pytestmark = pytest.mark.gpu

CENTROID_RTOL = 1e-6


def _ref_dist(X, c):
    p, n = X.shape
    if refmex.ref_available():
        return refmex.SparseMatrixMinusCluster(p, n, X.indptr, X.indices, X.data, c)
    return cport.masked_dist(p, n, X.indptr, X.indices, X.data, c)


# ---------------------------------------------------------------- level 1 ---
@pytest.mark.parametrize("K", [1, 2, 3, 4, 7, 10, 33, 64, 100])
def test_sparse_matrix_minus_cluster_bit_exact(ctx, K):
    from sparsifiedkmeans_b200 import SparseMatrixMinusCluster
    X, c, _ = make_sparsified(p=96, n=400, m=9, K=K, seed=K, kind="unstructured", f32=False, ragged=True)
    got = SparseMatrixMinusCluster(X, c)
    want = _ref_dist(X, c)
    assert got.shape == (K, X.shape[1])
    assert np.array_equal(got, want)


def test_sparse_matrix_minus_cluster_beta(ctx):
    from sparsifiedkmeans_b200 import SparseMatrixMinusCluster
    X, c, _ = make_sparsified(p=50, n=300, m=5, K=1, seed=3, kind="unstructured", f32=False)
    got = SparseMatrixMinusCluster(X, c, beta=0.37)
    p, n = X.shape
    want = cport.masked_dist_beta(p, n, X.indptr, X.indices, X.data, c[:, 0], 0.37)
    if refmex.ref_available():
        assert np.array_equal(want, refmex.SparseMatrixMinusCluster(p, n, X.indptr, X.indices, X.data, c, beta=0.37))
    assert np.array_equal(got, want)
    with pytest.raises(Exception):
        SparseMatrixMinusCluster(X, np.zeros((p, 2)), beta=0.5)


def test_inner_product_and_norms_bit_exact(ctx):
    from sparsifiedkmeans_b200 import SparseMatrixColumnNormSq, SparseMatrixInnerProduct
    X, c, _ = make_sparsified(p=80, n=257, m=11, K=1, seed=5, kind="unstructured", f32=False, ragged=True)
    ip, n2 = SparseMatrixInnerProduct(X, c[:, 0])
    wip, wn2 = cport.inner_product(X.shape[1], X.indptr, X.indices, X.data, c[:, 0])
    assert np.array_equal(ip.ravel(), wip) and np.array_equal(n2.ravel(), wn2)
    assert np.array_equal(SparseMatrixColumnNormSq(X).ravel(), wn2)


@pytest.mark.parametrize("m,n", [(2, 5), (8, 3), (256, 17), (4096, 4), (32768, 2), (65536, 1)])
def test_hadamard_bit_exact(ctx, m, n):
    from sparsifiedkmeans_b200 import hadamard, hadamard_pthreads
    rng = np.random.default_rng(m + n)
    x = rng.standard_normal((m, n))
    want = refmex.hadamard(x) if refmex.ref_available("hadamard") else cport.hadamard(x)
    assert np.array_equal(hadamard(x), want)
    assert np.array_equal(hadamard_pthreads(x), want)


def test_hadamard_rejects_bad_sizes(ctx):
    from sparsifiedkmeans_b200 import hadamard
    with pytest.raises(ValueError):
        hadamard(np.zeros((12, 3)))
    with pytest.raises(ValueError):
        hadamard(sp.csc_matrix(np.eye(4)))


# ------------------------------------------------------------ operator -----
@pytest.mark.parametrize("store", ["f64", "f32"])
@pytest.mark.parametrize("kind", ["mixture", "unstructured"])
@pytest.mark.parametrize("K,p,m", [(1, 32, 4), (2, 40, 6), (5, 64, 8), (10, 784, 78), (16, 128, 7),
                                   (33, 100, 10), (64, 1024, 51), (100, 256, 13), (130, 64, 5),
                                   (6, 64, 8), (9, 200, 20), (13, 512, 26), (14, 96, 12),   # LDS.64 dual-table kernel
                                   (16, 32768, 40), (5, 20000, 25)])       # tables too large for shared memory
def test_assign_matches_reference(ctx, store, kind, K, p, m):
    from sparsifiedkmeans_b200 import Dataset
    X, c, gamma = make_sparsified(p=p, n=3000, m=m, K=K, seed=K * 7 + p, kind=kind, f32=True, ragged=(K % 2 == 1))
    ds = Dataset.from_scipy(X, store=store, ctx=ctx)
    a, d = ds.assign(c, gamma)
    wa, wd, _ = host_ref.find_cluster_assignments(X, c, gamma)
    assert np.array_equal(a, wa), f"{np.count_nonzero(a != wa)} assignments differ"
    if store == "f64":
        assert np.array_equal(d, wd)
    else:
        np.testing.assert_allclose(d, wd, rtol=2e-5, atol=1e-30)
    ds.close()


@pytest.mark.parametrize("ragged", [False, True])
@pytest.mark.parametrize("p,m,n", [(784, 78, 5000), (1024, 51, 3001), (64, 8, 700), (40, 40, 300)])
def test_streamed_image_layouts(ctx, p, m, n, ragged):
    """Both entry orders of the SELL image hold exactly the CSC entries; the dual-table order
    (bipartite edge colouring per half-warp) is conflict-free: one wavefront per 8-byte gather."""
    from sparsifiedkmeans_b200 import Dataset
    X, c, gamma = make_sparsified(p=p, n=n, m=m, K=10, seed=p + m, kind="mixture", ragged=ragged)
    ds = Dataset.from_scipy(X, store="f32", ctx=ctx)
    r0 = ds.layout_check(0)
    assert r0["bad_columns"] == 0 and r0["layout"] == 0 and r0["steps"] > 0
    r1 = ds.layout_check(1)
    assert r1["bad_columns"] == 0 and r1["layout"] == 1
    if p >= 512:                                             # random rows, enough entries per class to balance the two copies
        assert r1["wavefronts"] == r1["steps"], r1          # conflict-free
    else:                                                    # too few / too regular rows to balance: still a valid order
        assert r1["wavefronts"] / r1["steps"] <= 1.25 * r0["wavefronts"] / r0["steps"]
    r2 = ds.layout_check(2)                                  # quarter-warp variant for the 16-byte kernels
    assert r2["bad_columns"] == 0 and r2["layout"] == 2 and r2["steps"] == r0["steps"]
    if p >= 512:
        assert r2["wavefronts"] == r2["steps"], r2
    assert r2["wavefronts"] <= 1.25 * r0["wavefronts"]
    # the kernels agree whatever order the image was left in
    wa, _, _ = host_ref.find_cluster_assignments(X, c, gamma)
    a1, _ = ds.assign(c, gamma)                              # K = 10 -> dual-table kernel
    a2, _ = ds.assign(c[:, :4], gamma)                       # K = 4  -> 16-byte kernel, re-orders the image (layout 0 or 2)
    a3, _ = ds.assign(c, gamma)
    assert np.array_equal(a1, wa) and np.array_equal(a3, wa)
    assert np.array_equal(a2, host_ref.find_cluster_assignments(X, c[:, :4], gamma)[0])
    assert ds.layout_check()["layout"] == 1
    ds.close()


@pytest.mark.parametrize("K", [2, 10, 16, 40])
def test_assign_long_columns_and_tiny_shards(ctx, K):
    """Columns longer than the dual-table scheduler's byte-sized state (> 254 entries) fall back to the
    single-table order; shards smaller than one 32-column slice and single-column shards still work."""
    from sparsifiedkmeans_b200 import Dataset
    X, c, gamma = make_sparsified(p=1024, n=200, m=300, K=K, seed=K, kind="mixture")
    ds = Dataset.from_scipy(X, store="f32", ctx=ctx)
    assert ds.max_col_nnz == 300
    a, d = ds.assign(c, gamma)
    wa, wd, _ = host_ref.find_cluster_assignments(X, c, gamma)
    assert np.array_equal(a, wa)
    assert ds.layout_check()["layout"] == 0 and ds.layout_check()["bad_columns"] == 0
    ds.close()
    for n in (1, 5, 33):
        X, c, gamma = make_sparsified(p=96, n=n, m=12, K=K, seed=K + n, kind="unstructured")
        ds = Dataset.from_scipy(X, store="f32", ctx=ctx)
        a, _ = ds.assign(c, gamma)
        assert np.array_equal(a, host_ref.find_cluster_assignments(X, c, gamma)[0])
        assert ds.layout_check()["bad_columns"] == 0
        ds.close()


def test_assign_near_ties_are_resolved_exactly(ctx):
    """Duplicate centres and 1-ulp perturbations: the fp32 kernel must hand these to the fp64 path."""
    from sparsifiedkmeans_b200 import Dataset
    X, c, gamma = make_sparsified(p=64, n=4000, m=8, K=6, seed=11, kind="unstructured")
    c[:, 3] = c[:, 1]                                   # exact tie -> lower index wins
    c[:, 5] = np.nextafter(c[:, 2], np.inf)             # 1 ulp away
    ds = Dataset.from_scipy(X, store="f32", ctx=ctx)
    a, _ = ds.assign(c, gamma)
    wa, _, _ = host_ref.find_cluster_assignments(X, c, gamma)
    assert np.array_equal(a, wa)
    assert not np.any(a == 4)                            # centre 3 (1-based 4) never beats its twin
    ds.close()


def test_assign_nan_and_inf_centres(ctx):
    from sparsifiedkmeans_b200 import Dataset
    X, c, gamma = make_sparsified(p=48, n=1000, m=6, K=4, seed=13, kind="mixture")
    c[5, 2] = np.nan
    c[7, 3] = np.inf
    for store in ("f64", "f32"):
        ds = Dataset.from_scipy(X, store=store, ctx=ctx)
        a, d = ds.assign(c, None)
        wa, wd, _ = host_ref.find_cluster_assignments(X, c, None)
        assert np.array_equal(a, wa)
        if store == "f64":
            assert np.array_equal(d, wd, equal_nan=True)
        ds.close()


def test_assign_empty_and_tiny(ctx):
    from sparsifiedkmeans_b200 import Dataset
    X = sp.csc_matrix((16, 5))                           # all columns empty -> distance 0, centre 1
    ds = Dataset.from_scipy(X, store="f32", ctx=ctx)
    a, d = ds.assign(np.ones((16, 3)), 0.5)
    assert np.array_equal(a, np.ones(5, dtype=np.int32)) and np.array_equal(d, np.zeros(5))
    ds.close()
    ds = Dataset.from_scipy(sp.csc_matrix((16, 0)), store="f32", ctx=ctx)
    a, d = ds.assign(np.ones((16, 3)), 0.5)
    assert a.shape == (0,) and d.shape == (0,)
    ds.close()


@pytest.mark.parametrize("gamma", [None, 0.125])
def test_assign_sparse_centres_branch(ctx, gamma):
    from sparsifiedkmeans_b200 import Dataset, findClusterAssignments
    X, _, _ = make_sparsified(p=64, n=800, m=8, K=4, seed=17, kind="mixture")
    cen = X[:, [3, 50, 200, 601]]                        # k-means++ style: centres are sampled columns
    ds = Dataset.from_scipy(X, store="f32", ctx=ctx)
    a, d = findClusterAssignments(ds, cen, None, gamma)
    wa, wd, _ = host_ref.find_cluster_assignments(X, cen, gamma)
    assert np.array_equal(a, wa) and np.array_equal(d, wd)
    ds.close()


def test_find_cluster_assignments_errors(ctx):
    from sparsifiedkmeans_b200 import findClusterAssignments
    X, c, gamma = make_sparsified(p=32, n=50, m=4, K=3, seed=1)
    with pytest.raises(ValueError, match="not of correct size"):
        findClusterAssignments(X, c[:-1], None, gamma)
    with pytest.raises(ValueError, match="not of correct size"):
        findClusterAssignments(np.asarray(X.todense()), c[:-1], None, gamma)


def test_upload_rejects_bad_csc(ctx):
    from sparsifiedkmeans_b200 import Dataset
    from sparsifiedkmeans_b200._lib import SkmError
    with pytest.raises(SkmError):
        Dataset.from_csc(4, 2, [0, 2, 3], [0, 9, 1], [1.0, 2.0, 3.0], ctx=ctx)     # row 9 >= p
    with pytest.raises(SkmError):
        Dataset.from_csc(4, 2, [0, 3, 2], [0, 1, 2], [1.0, 2.0, 3.0], ctx=ctx)     # decreasing jc


# ------------------------------------------------------------ Lloyd --------
@pytest.mark.parametrize("store", ["f64", "f32"])
@pytest.mark.parametrize("K,p,m", [(5, 64, 8), (10, 784, 78), (64, 1024, 51), (70, 128, 9)])
def test_centroid_update_matches_reference(ctx, store, K, p, m):
    from sparsifiedkmeans_b200 import Dataset, Lloyd
    X, c, gamma = make_sparsified(p=p, n=5000, m=m, K=K, seed=K, kind="mixture")
    ds = Dataset.from_scipy(X, store=store, ctx=ctx)
    L = Lloyd(ds, K)
    L.set_centers(c)
    st = L.step(gamma, gamma, True)
    a, d = L.assignments()
    wa, wd, _ = host_ref.find_cluster_assignments(X, c, gamma)
    assert np.array_equal(a, wa)
    want, S, N, counts = cport.centroid_update(p, X.shape[1], K, X.indptr, X.indices, X.data, wa, gamma, c, True)
    got = L.get_centers()
    assert np.array_equal(L.counts(), counts)
    scale = np.max(np.abs(want))
    assert np.max(np.abs(got - want)) <= CENTROID_RTOL * scale
    np.testing.assert_allclose(got, want, rtol=CENTROID_RTOL, atol=CENTROID_RTOL * scale * 1e-3)
    assert st.n_points == X.shape[1] and st.n_empty == int(np.sum(counts == 0))
    np.testing.assert_allclose(st.dff, np.linalg.norm(c - want, "fro"), rtol=1e-9)
    np.testing.assert_allclose(st.sumsq, np.sum(wd ** 2), rtol=1e-5 if store == "f32" else 1e-12)
    L.close(); ds.close()


def test_lloyd_trajectory_matches_reference(ctx):
    """Whole loop (kmeans_sparsified.m:417-486) from the same start: identical assignments,
    iteration count, and centres within tolerance."""
    from sparsifiedkmeans_b200 import Dataset, Lloyd
    X, c, gamma = make_sparsified(p=128, n=4000, m=10, K=6, seed=21, kind="mixture")
    ref = host_ref.lloyd(X, c, gamma, max_iter=30, tol=1e-6)
    ds = Dataset.from_scipy(X, store="f32", ctx=ctx)
    L = Lloyd(ds, 6)
    L.set_centers(c)
    its = 0
    for its in range(1, 31):
        st = L.step(gamma, gamma, True)
        assert st.n_empty == 0
        if st.dff < 1e-6:
            break
    a, d = L.assignments()
    assert its == ref.iterations
    assert np.array_equal(a, ref.assignments)
    np.testing.assert_allclose(L.get_centers(), ref.centers, rtol=1e-6, atol=1e-9)
    L.close(); ds.close()


def test_argmax_distance_and_get_column(ctx):
    from sparsifiedkmeans_b200 import Dataset, Lloyd
    X, c, gamma = make_sparsified(p=64, n=2000, m=8, K=5, seed=23, kind="unstructured")
    ds = Dataset.from_scipy(X, store="f64", ctx=ctx)
    L = Lloyd(ds, 5)
    L.set_centers(c)
    L.assign(gamma)
    _, wd, _ = host_ref.find_cluster_assignments(X, c, gamma)
    v, j = L.argmax_distance()
    assert j == int(np.argmax(wd)) and v == wd[j]
    assert np.array_equal(ds.get_column(j), np.asarray(X[:, j].todense()).ravel())
    L.close(); ds.close()


# ------------------------------------------------------------ k-means++ ----
def test_kpp_running_min_matches_reference(ctx):
    from sparsifiedkmeans_b200 import Dataset
    X, _, gamma = make_sparsified(p=64, n=1500, m=8, K=4, seed=29, kind="mixture")
    ds = Dataset.from_scipy(X, store="f32", ctx=ctx)
    chosen = [7, 400, 1200]
    for i, j in enumerate(chosen):
        tot = ds.kpp_update(np.asarray(X[:, j].todense()).ravel(), gamma, first=(i == 0))
        cen = np.asarray(X[:, chosen[: i + 1]].todense())
        _, wd, _ = host_ref.find_cluster_assignments(X, cen, gamma, centers_sparse=False)
        assert np.array_equal(ds.kpp_mindist(), wd)
        np.testing.assert_allclose(tot, np.sum(wd ** 2), rtol=1e-12)
    w = wd ** 2
    for u in (0.0, 0.1, 0.5, 0.999999):
        assert ds.kpp_pick(u * tot) == host_ref.weighted_pick(w, u)
    ds.close()


# ------------------------------------------------------------ streamed -----
@pytest.mark.parametrize("K,p,m,chunk", [(5, 64, 8, 0), (10, 784, 78, 1024), (70, 128, 9, 512)])
@pytest.mark.parametrize("dtypes", [(np.int32, np.int32, np.float32), (np.int64, np.int64, np.float64),
                                    (np.int64, np.uint16, np.float32)])
def test_streamed_host_iteration_matches_reference(ctx, K, p, m, chunk, dtypes):
    """skm_lloyd_step_host (X in host memory, chunks over PCIe) == resident path == oracle."""
    from sparsifiedkmeans_b200 import lloyd_step_host
    X, c, gamma = make_sparsified(p=p, n=5000, m=m, K=K, seed=K + 3, kind="mixture", ragged=(K == 5))
    n = X.shape[1]
    jt, it, vt = dtypes
    newc, a, d, st = lloyd_step_host(p, n, X.indptr.astype(jt), X.indices.astype(it), X.data.astype(vt), c, gamma, gamma,
                                     True, chunk_cols=chunk, want_dist=True, ctx=ctx)
    wa, wd, _ = host_ref.find_cluster_assignments(X, c, gamma)
    assert np.array_equal(a, wa)
    np.testing.assert_allclose(d, wd, rtol=2e-5, atol=1e-30)
    want, _, _, counts = cport.centroid_update(p, n, K, X.indptr, X.indices, X.data, wa, gamma, c, True)
    scale = np.max(np.abs(want))
    assert np.max(np.abs(newc - want)) <= CENTROID_RTOL * scale
    assert st.n_points == n and st.n_empty == int(np.sum(counts == 0))
    np.testing.assert_allclose(st.dff, np.linalg.norm(c - want, "fro"), rtol=1e-9)


def test_streamed_rejects_bad_rows(ctx):
    from sparsifiedkmeans_b200 import lloyd_step_host
    from sparsifiedkmeans_b200._lib import SkmError
    with pytest.raises(SkmError):
        lloyd_step_host(4, 2, np.array([0, 2, 3]), np.array([0, 9, 1]), np.array([1.0, 2.0, 3.0]), np.ones((4, 2)),
                        None, 1.0, ctx=ctx)


def test_upload_accepts_uint16_rows(ctx):
    from sparsifiedkmeans_b200 import Dataset
    X, c, gamma = make_sparsified(p=300, n=700, m=20, K=4, seed=77)
    ds = Dataset.from_csc(300, 700, X.indptr.astype(np.int32), X.indices.astype(np.uint16), X.data.astype(np.float32), ctx=ctx)
    a, _ = ds.assign(c, gamma)
    wa, _, _ = host_ref.find_cluster_assignments(X, c, gamma)
    assert np.array_equal(a, wa)
    ds.close()


@pytest.mark.parametrize("kind", ["mixture", "unstructured"])
def test_kpp_filtered_rounds_match_reference(ctx, kind):
    """Rounds after the first run an fp32 filter in front of the exact pass (n >= 65536): the running minimum stays
    bit-identical to the reference's recomputation, with and without the filter."""
    import os
    from sparsifiedkmeans_b200 import Dataset
    X, _, gamma = make_sparsified(p=128, n=70000, m=12, K=6, seed=31, kind=kind, ragged=True)
    ds = Dataset.from_scipy(X, store="f32", ctx=ctx)
    chosen = [11, 40000, 123, 69999, 5, 31000]
    for i, j in enumerate(chosen):
        tot = ds.kpp_update(np.asarray(X[:, j].todense()).ravel(), gamma, first=(i == 0))
        cen = np.asarray(X[:, chosen[: i + 1]].todense())
        _, wd, _ = host_ref.find_cluster_assignments(X, cen, gamma, centers_sparse=False)
        assert np.array_equal(ds.kpp_mindist(), wd), f"round {i}"
        np.testing.assert_allclose(tot, np.sum(wd ** 2), rtol=1e-12)
    got = ds.kpp_mindist().copy()
    ds.close()
    os.environ["SKM_NO_KPP_FILTER"] = "1"
    try:
        ds = Dataset.from_scipy(X, store="f32", ctx=ctx)
        for i, j in enumerate(chosen):
            ds.kpp_update(np.asarray(X[:, j].todense()).ravel(), gamma, first=(i == 0))
        assert np.array_equal(ds.kpp_mindist(), got)
        ds.close()
    finally:
        del os.environ["SKM_NO_KPP_FILTER"]


def test_objective_is_bit_reproducible(ctx):
    """The sum of squared distances is reduced in a fixed order (ordered_block_sum in update.cu), not with one fp64 atomic
    per block: replicates that reach the same clustering are told apart by it (kmeans_sparsified.m:489-497), so two
    identical passes must give the same bits, in the recompute and in the incremental update."""
    from sparsifiedkmeans_b200 import Dataset, Lloyd
    X, c, gamma = make_sparsified(p=128, n=150000, m=12, K=7, seed=77, kind="unstructured")
    ds = Dataset.from_scipy(X, store="f32", ctx=ctx)
    for incremental in (False, True):
        seen = set()
        # every row of this matrix is one work unit of K2 (< 16 K entries), so the recomputed centres are reproducible too;
        # the incremental +/- updates are fp64 atomics in arbitrary order, so only its first two objectives are
        for rep in range(6):
            L = Lloyd(ds, 7, incremental=incremental)
            L.set_centers(c)
            vals = []
            for _ in range(2 if incremental else 3):
                st = L.step(gamma, gamma, True)
                vals.append(st.sumsq)
            seen.add(tuple(np.float64(v).tobytes() for v in vals))
            L.close()
        assert len(seen) == 1, f"incremental={incremental}: {len(seen)} different objective bit patterns over 6 identical runs"
    ds.close()
