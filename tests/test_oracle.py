"""The oracle pinned: oracle/skm_oracle.c (the C restatement) and oracle/_ref (the reference's
own C compiled unmodified) against the committed golden vectors and against each other, plus
the properties the reference's header comments state.  CPU only."""
import glob
import os

import numpy as np
import pytest
import scipy.linalg
import scipy.sparse as sp

from oracle import cport, host_ref, refmex
from tests.util import make_sparsified

GOLD = os.path.join(os.path.dirname(__file__), "golden")
needs_ref = pytest.mark.skipif(not refmex.ref_available(), reason="oracle/_ref not built")


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "smmc_K*.npz"))))
def test_masked_distance_golden(path):
    g = np.load(path)
    p, n = int(g["p"]), int(g["n"])
    got = cport.masked_dist(p, n, g["jc"], g["ir"], g["x"], g["centers"])
    assert np.array_equal(got, g["dist"])
    if refmex.ref_available():
        assert np.array_equal(refmex.SparseMatrixMinusCluster(p, n, g["jc"], g["ir"], g["x"], g["centers"]), g["dist"])
    # fused assign == distances followed by MATLAB min
    a, d = cport.assign(p, n, g["jc"], g["ir"], g["x"], g["centers"], threads=3)
    d2, a2 = cport.colmin(g["dist"])
    assert np.array_equal(a, a2) and np.array_equal(d, d2)


def test_beta_inner_norm_golden():
    g = np.load(os.path.join(GOLD, "beta_inner_norm.npz"))
    p, n = int(g["p"]), int(g["n"])
    assert np.array_equal(cport.masked_dist_beta(p, n, g["jc"], g["ir"], g["x"], g["c"], float(g["beta"])), g["dist_beta"])
    ip, n2 = cport.inner_product(n, g["jc"], g["ir"], g["x"], g["c"])
    assert np.array_equal(ip, g["inner"].ravel()) and np.array_equal(n2, g["normsq"].ravel())
    assert np.array_equal(cport.colnormsq(n, g["jc"], g["x"]), g["normsq_only"].ravel())


@pytest.mark.parametrize("m", [2, 16, 512])
def test_hadamard_golden(m):
    g = np.load(os.path.join(GOLD, f"hadamard_m{m}.npz"))
    assert np.array_equal(cport.hadamard(g["x"]), g["w"])
    assert np.array_equal(g["w"], g["w_pthreads"])           # serial == pthreads, bit for bit
    # hadamard.c:8-11,18-24: symmetric, H*H = m*I, equals the Sylvester matrix
    H = scipy.linalg.hadamard(m)
    np.testing.assert_allclose(g["w"], H @ g["x"], rtol=0, atol=1e-10 * m)
    np.testing.assert_allclose(cport.hadamard(g["w"]) / m, g["x"], rtol=0, atol=1e-12 * m)


@needs_ref
@pytest.mark.parametrize("K", [1, 2, 3, 5, 17])
def test_port_matches_compiled_reference(K):
    X, c, gamma = make_sparsified(p=70, n=333, m=6, K=K, seed=K, kind="unstructured", f32=False, ragged=True)
    p, n = X.shape
    want = refmex.SparseMatrixMinusCluster(p, n, X.indptr, X.indices, X.data, c / gamma)
    assert np.array_equal(cport.masked_dist(p, n, X.indptr, X.indices, X.data, c / gamma), want)
    # header spec: dist(i) = norm(X(ind,i) - c(ind)), ind = find(X(:,i))  (SparseMatrixMinusCluster.c:2-11)
    Xd = np.asarray(X.todense())
    for j in (0, 5, 100, n - 1):
        ind = np.flatnonzero(Xd[:, j])
        for k in range(K):
            np.testing.assert_allclose(want[k, j], np.linalg.norm(Xd[ind, j] - (c / gamma)[ind, k]), rtol=1e-13, atol=1e-300)


@needs_ref
def test_reference_argument_errors():
    X, c, _ = make_sparsified(p=20, n=10, m=3, K=2, seed=0, f32=False)
    with pytest.raises(refmex.MexError):
        refmex.SparseMatrixMinusCluster(20, 10, X.indptr, X.indices, X.data, c[:-1])       # wrong rows
    with pytest.raises(refmex.MexError):
        refmex.SparseMatrixMinusCluster(20, 10, X.indptr, X.indices, X.data, c, beta=0.5)  # beta needs K == 1
    with pytest.raises(refmex.MexError):
        refmex.hadamard(np.zeros((12, 2)))                                                  # not a power of two


def test_colmin_matlab_semantics():
    D = np.array([[2.0, np.nan, np.nan, 1.0], [2.0, 3.0, np.nan, 1.0], [1.0, 0.5, np.nan, 5.0]])
    d, a = cport.colmin(D)
    assert list(a) == [3, 3, 1, 1]                     # first occurrence; NaN skipped; all-NaN -> index 1
    assert d[0] == 1.0 and d[1] == 0.5 and np.isnan(d[2]) and d[3] == 1.0


def test_centroid_update_formula():
    X, c, gamma = make_sparsified(p=40, n=300, m=5, K=4, seed=9, f32=False)
    a, _, _ = host_ref.find_cluster_assignments(X, c, gamma)
    a[a == 4] = 1                                                         # leave cluster 4 empty
    new, S, N, counts = cport.centroid_update(40, 300, 4, X.indptr, X.indices, X.data, a, gamma, c, True)
    ones = sp.csc_matrix((np.ones_like(X.data), X.indices, X.indptr), shape=X.shape)
    for k in range(3):
        ind = np.flatnonzero(a == k + 1)
        s = np.asarray(X[:, ind].sum(axis=1)).ravel()
        nn = np.asarray(ones[:, ind].sum(axis=1)).ravel()
        np.testing.assert_allclose(new[:, k], gamma * s / (nn + 1e-16), rtol=1e-12, atol=1e-300)   # kmeans_sparsified.m:448
    assert counts[3] == 0 and np.array_equal(new[:, 3], c[:, 3])


def test_host_ref_sparse_centres_and_lloyd():
    X, c, gamma = make_sparsified(p=64, n=600, m=8, K=4, seed=3, kind="mixture")
    cen = X[:, [1, 2, 3, 4]]
    a, d, D = host_ref.find_cluster_assignments(X, cen, gamma)
    assert D.shape == (4, 600) and np.all(d >= 0)
    # a point is at distance 0 from itself in the sparse-centres metric only if gamma_center == gamma ... it is
    # at least the minimiser among centres sharing its full support
    assert a[1] == 1 or D[0, 1] >= d[1]
    res = host_ref.lloyd(X, c, gamma, max_iter=20)
    assert res.iterations >= 1 and res.stopping_diff < 1e-6
    lab = np.arange(600) % 4
    # planted partition recovered up to relabelling (example_sparseKMeans.m:12-22)
    for k in range(4):
        assert len(set(res.assignments[lab == k])) == 1


def test_arthur_initialization_contract():
    X, _, gamma = make_sparsified(p=64, n=400, m=8, K=5, seed=4, kind="mixture")
    rng = np.random.default_rng(0)
    idx, cen = host_ref.arthur_initialization(X, 5, gamma, first=10, uniforms=iter(rng.random(4000)))
    assert idx[0] == 10 and len(set(idx.tolist())) == 5 and cen.shape == (64, 5)
