"""GPU parity of the second pass over the original dense data:
skm_second_pass against the oracle's restatement of kmeans_sparsified.m:542-560 and of the dense
branch of private/findClusterAssignments.m:154-175.

Bars: per-cluster means within 1e-6 relative; assignments equal to the fp64 oracle wherever the
oracle's own top-2 gap exceeds 1e-9 relative (the reference's BLAS / pdist2 summation order is
closed source, so closer ties are unpinned -- DESIGN.md); distances within 2e-5 relative (fp32)."""
import numpy as np
import pytest

from oracle import host_ref

pytestmark = pytest.mark.gpu


def _mixture(p, n, K, seed, sigma=0.3, dtype=np.float64):
    rng = np.random.default_rng(seed)
    mu = rng.standard_normal((p, K))
    lab = rng.integers(0, K, n)
    X = (mu[:, lab] + sigma * rng.standard_normal((p, n))).astype(dtype)
    centers = mu + 0.05 * rng.standard_normal((p, K))
    return X, centers, lab + 1


def _check_assign(X, c, a, d, n_tie_ok=0):
    wa, wd, D2 = host_ref.find_cluster_assignments_dense(np.asarray(X, dtype=np.float64), c)
    srt = np.sort(D2, axis=0)
    gap = (srt[1] - srt[0]) / np.maximum(srt[1], 1e-300) if D2.shape[0] > 1 else np.ones(X.shape[1])
    clear = gap > 1e-9
    assert np.array_equal(a[clear], wa[clear]), f"{np.count_nonzero(a[clear] != wa[clear])} assignments differ"
    assert np.count_nonzero(~clear) <= n_tie_ok
    np.testing.assert_allclose(d, wd, rtol=2e-5, atol=1e-6 * float(np.sqrt(np.max(srt[-1]))))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("p,n,K", [(50, 1000, 5), (784, 3000, 10), (33, 517, 1), (128, 2000, 16),
                                   (100, 999, 33), (1024, 1500, 64), (7, 300, 3), (4000, 200, 20)])
def test_second_pass_matches_oracle(ctx, p, n, K, dtype):
    from sparsifiedkmeans_b200 import second_pass
    X, c, lab = _mixture(p, n, K, seed=p + n + K, dtype=dtype)
    scale = 1 + 2 * np.finfo(np.float64).eps
    res = second_pass(X, centers=c, assign_in=lab, scale=scale, chunk_cols=700, ctx=ctx)
    XF = np.asarray(X, dtype=np.float64) * scale
    wc, wa, wd = host_ref.second_pass(XF, c, lab, K)
    err = np.max(np.abs(res["centers"] - wc)) / np.max(np.abs(wc))
    assert err <= 1e-6, err
    assert np.array_equal(res["counts"], np.bincount(lab - 1, minlength=K))
    _check_assign(XF, c, res["assign"], res["dist"])


def test_second_pass_unstructured_and_ties(ctx):
    """No cluster structure (small gaps everywhere), duplicated centres and a 1-ulp twin: the
    guard must hand the close calls to the fp64 kernel; exact ties go to the lower index."""
    from sparsifiedkmeans_b200 import second_pass
    rng = np.random.default_rng(5)
    p, n, K = 64, 4000, 8
    X = rng.standard_normal((p, n))
    c = 0.1 * rng.standard_normal((p, K))
    c[:, 5] = c[:, 2]
    c[:, 7] = np.nextafter(c[:, 1], np.inf)
    res = second_pass(X, centers=c, ctx=ctx)
    assert not np.any(res["assign"] == 6)
    wa, wd, D2 = host_ref.find_cluster_assignments_dense(X, c)
    # direct fp64 evaluation decides the 1-ulp twins; the expanded formula of the oracle cannot
    direct = np.stack([np.sqrt(((X - c[:, [k]]) ** 2).sum(axis=0)) for k in range(K)])
    best = direct.min(axis=0)
    chosen = direct[res["assign"] - 1, np.arange(n)]
    assert np.all(chosen <= best * (1 + 1e-12))
    assert res["n_rechecked"] > 0


def test_second_pass_empty_cluster_unassigned_and_halves(ctx):
    from sparsifiedkmeans_b200 import second_pass
    X, c, lab = _mixture(40, 600, 4, seed=3)
    lab = lab.copy()
    lab[lab == 3] = 1                      # cluster 3 empty -> zero column (kmeans_sparsified.m:545)
    lab[:5] = 0                            # unassigned
    res = second_pass(X, assign_in=lab, want_assign=False, want_dist=False, ctx=ctx)
    assert "assign" not in res and res["centers"].shape == (40, 4)
    assert np.all(res["centers"][:, 2] == 0) and res["counts"][2] == 0
    keep = lab > 0
    wc, _, _ = host_ref.second_pass(X[:, keep], c, lab[keep], 4)
    np.testing.assert_allclose(res["centers"], wc, rtol=1e-6, atol=1e-9)
    only = second_pass(X, centers=c, ctx=ctx)
    assert "centers" not in only and only["assign"].shape == (600,)
    with pytest.raises(Exception):
        second_pass(X, assign_in=np.full(600, 9), centers=c, ctx=ctx)   # label outside 0..K


def test_dense_operator_and_kmeans_two_pass_outputs(ctx):
    from sparsifiedkmeans_b200 import findClusterAssignments, kmeans_sparsified
    X, c, lab = _mixture(64, 1500, 4, seed=9, sigma=0.2)
    a, d, m = findClusterAssignments(X, c, None, None, nargout=3, ctx=ctx)
    _check_assign(X, c, a, d)
    wm = np.stack([X[:, a == k + 1].mean(axis=1) for k in range(4)], axis=1)
    np.testing.assert_allclose(m, wm, rtol=1e-6, atol=1e-9)
    # rows are points for the driver (ColumnSamples=false); 9 outputs like the reference
    out = kmeans_sparsified(X.T.copy(), 4, Sparsify=True, SparsityLevel=0.25, Seed=4, nargout=9, Replicates=2, Context=ctx)
    IDX, C, SUMD, D, OUTPUT, C2, IDX2, D2, SUMD2 = out
    assert C2.shape == C.shape == (4, 64) and IDX2.shape == IDX.shape and SUMD2.shape == (4,)
    scale = 1 + 2 * np.finfo(np.float64).eps
    wc2, wa2, wd2 = host_ref.second_pass(X * scale, C.T, IDX, 4)
    np.testing.assert_allclose(C2.T, wc2, rtol=1e-6, atol=1e-9)
    assert np.mean(IDX2 == wa2) > 0.999
    # planted clusters are recovered by both passes (up to a permutation): two-pass means sit on the truth
    truth = np.stack([X[:, lab == k + 1].mean(axis=1) for k in range(4)], axis=1)
    dists = np.linalg.norm(C2.T[:, :, None] - truth[:, None, :], axis=0)
    assert np.all(dists.min(axis=1) < 0.05 * np.linalg.norm(truth, axis=0).mean())
    six = kmeans_sparsified(X.T.copy(), 4, Sparsify=True, SparsityLevel=0.25, Seed=4, nargout=6, Context=ctx)
    assert len(six) == 6


@pytest.mark.parametrize("p,n,K,sigma", [(256, 50_000, 64, 0.3), (1024, 20_000, 64, 0.5), (96, 30_000, 17, 0.2),
                                         (784, 8_000, 100, 0.3), (2048, 3_000, 256, 0.3), (64, 40_000, 16, 1.0)])
def test_tensor_core_filter_keeps_the_reference_winners(ctx, p, n, K, sigma):
    """K >= 16: the distance product runs on the tensor cores (tcgen05, tf32) as a filter and the candidates are
    evaluated exactly (csrc/tcgemm.cu).  Several 128-point tiles per CTA (persistent loop, double-buffered TMEM
    accumulator), p not a multiple of the 32-float k-block, K not a multiple of 16, several chunks."""
    from sparsifiedkmeans_b200 import second_pass
    X, c, lab = _mixture(p, n, K, seed=p + K, sigma=sigma, dtype=np.float32)
    kept0, dropped0 = ctx.tc_chunks()
    res = second_pass(X, centers=c, assign_in=None, scale=1.0, chunk_cols=max(n // 3, 1), ctx=ctx)
    kept1, dropped1 = ctx.tc_chunks()
    assert (kept1 - kept0) + (dropped1 - dropped0) >= 1, "the tensor-core path did not run"
    _check_assign(np.asarray(X, dtype=np.float64), c, res["assign"], res["dist"])
    if sigma <= 0.5:
        assert kept1 - kept0 >= 3 and res["n_rechecked"] <= n // 8     # clustered data: the filter certifies


def test_tensor_core_filter_falls_back_on_unstructured_data(ctx):
    from sparsifiedkmeans_b200 import second_pass
    rng = np.random.default_rng(11)
    p, n, K = 128, 20_000, 32
    X = rng.standard_normal((p, n)).astype(np.float32)
    c = 0.1 * rng.standard_normal((p, K))
    c[:, 9] = c[:, 3]                                            # exact duplicate: ties go to the lower index
    res = second_pass(X, centers=c, scale=1.0, chunk_cols=5000, ctx=ctx)
    _check_assign(np.asarray(X, dtype=np.float64), c, res["assign"], res["dist"], n_tie_ok=n)
    assert not np.any(res["assign"] == 10)
