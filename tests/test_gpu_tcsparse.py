"""Tensor-core plan of K1 (csrc/tcsparse.cu): the fp16 filter's scores stay inside the bound the kernel relies on,
and the assignments of the whole pass (filter -> exact winner -> best three -> fp64) are the oracle's."""
import numpy as np
import pytest

from oracle import host_ref
from tests.util import make_sparsified

pytestmark = pytest.mark.gpu


def _scores_f64(X, c, gamma):
    """s[j, k] = sum_{r in supp(x_j)} (c'_rk^2 - 2 x_jr c'_rk), c' = c / gamma."""
    cs = c / gamma if gamma is not None else c
    M = X.copy()
    M.data = np.ones_like(M.data)
    return (M.T @ (cs * cs) - 2.0 * (X.T @ cs))


@pytest.mark.parametrize("kind", ["mixture", "unstructured"])
@pytest.mark.parametrize("K,p,m,n,ragged", [(64, 1024, 51, 3000, False), (40, 200, 20, 1000, True), (100, 256, 13, 2077, True),
                                            (20, 100, 10, 700, False), (2, 64, 8, 300, True), (128, 4096, 64, 520, False)])
def test_filter_scores_within_bound(ctx, kind, K, p, m, n, ragged):
    from sparsifiedkmeans_b200 import Dataset, Lloyd
    X, c, gamma = make_sparsified(p=p, n=n, m=m, K=K, seed=K + p, kind=kind, f32=True, ragged=ragged)
    ds = Dataset.from_scipy(X, store="f32", ctx=ctx)
    L = Lloyd(ds, K)
    L.set_tc_filter(True)
    L.set_centers(c)
    got = L.debug_tc_scores(gamma)[:, :K].astype(np.float64)
    want = _scores_f64(X, c, gamma)
    cs = c / gamma
    A = abs(X)
    M = X.copy(); M.data = np.ones_like(M.data)
    # the rounding model of skm_launch_tcs_filter: 1.25 * 2^-11 * (4 sum|x c| + sum c^2) + tiny absolute part
    mag = 4.0 * (A.T @ np.abs(cs)) + (M.T @ (cs * cs))
    mx = max(np.abs(X.data).max(), np.abs(cs).max())
    bound = 1.25 / 2048 * mag + X.getnnz(axis=0).max() * 260.0 / 2 ** 25 * (2.0 * mx / 64.0) ** 2
    err = np.abs(got - want)
    assert np.all(err <= bound + 1e-30), f"max err/bound {np.max(err / (bound + 1e-300))}"
    # and it is a useful filter: typical error well below the bound
    assert np.median(err / (bound + 1e-300)) < 0.2
    a, _ = L.assignments()
    wa, _, _ = host_ref.find_cluster_assignments(X, c, gamma)
    assert np.array_equal(a, wa)
    L.close(); ds.close()


@pytest.mark.parametrize("kind", ["mixture", "unstructured"])
@pytest.mark.parametrize("K,p,m", [(64, 1024, 51), (24, 784, 78), (33, 100, 10), (100, 256, 13), (128, 512, 26), (3, 64, 8)])
def test_tc_assign_matches_reference(ctx, kind, K, p, m):
    from sparsifiedkmeans_b200 import Dataset, Lloyd
    X, c, gamma = make_sparsified(p=p, n=5000, m=m, K=K, seed=K * 7 + p, kind=kind, f32=True, ragged=(K % 2 == 1))
    ds = Dataset.from_scipy(X, store="f32", ctx=ctx)
    wa, wd, _ = host_ref.find_cluster_assignments(X, c, gamma)
    L = Lloyd(ds, K)
    L.set_tc_filter(True)
    L.set_centers(c)
    L.assign(gamma)
    a, d = L.assignments()
    assert np.array_equal(a, wa), f"{np.count_nonzero(a != wa)} assignments differ"
    np.testing.assert_allclose(d, wd, rtol=2e-5, atol=1e-30)
    L.accumulate()
    st = L.finalize(gamma)
    kept, f64 = L.last_tc()
    assert kept >= 0 and 0 <= f64 <= kept
    if kind == "mixture":
        assert kept <= X.shape[1] // 50, "well separated clusters: the winner's exact evaluation keeps (nearly) every column"
    # same centres as the gather kernels
    L0 = Lloyd(ds, K)
    L0.set_tc_filter(False)
    L0.set_centers(c)
    L0.assign(gamma); L0.accumulate(); L0.finalize(gamma)
    a0, _ = L0.assignments()
    assert np.array_equal(a0, a)
    np.testing.assert_allclose(L.get_centers(), L0.get_centers(), rtol=1e-12, atol=1e-300)
    L.close(); L0.close(); ds.close()


def test_tc_near_ties_and_bad_values(ctx):
    """Duplicate centres (exact ties -> first index), a NaN centre, an empty column, huge values."""
    from sparsifiedkmeans_b200 import Dataset, Lloyd
    X, c, gamma = make_sparsified(p=256, n=2000, m=16, K=32, seed=5, kind="mixture", f32=True, ragged=True)
    c[:, 7] = c[:, 3]                       # exact tie: the reference takes the first index
    c[:, 20] = c[:, 3] * (1 + 1e-9)         # closer than any filter can see
    ds = Dataset.from_scipy(X, store="f32", ctx=ctx)
    L = Lloyd(ds, 32)
    L.set_tc_filter(True)
    L.set_centers(c)
    L.assign(gamma)
    a, d = L.assignments()
    wa, wd, _ = host_ref.find_cluster_assignments(X, c, gamma)
    assert np.array_equal(a, wa)
    c2 = c.copy(); c2[5, 11] = np.nan
    L.set_centers(c2)
    L.assign(gamma)
    a, _ = L.assignments()
    wa, _, _ = host_ref.find_cluster_assignments(X, c2, gamma)
    assert np.array_equal(a, wa)
    L.close(); ds.close()
    Xb = X.copy(); Xb.data = Xb.data * 1e20
    dsb = Dataset.from_scipy(Xb.astype(np.float64), store="f32", ctx=ctx)
    Lb = Lloyd(dsb, 32); Lb.set_tc_filter(True); Lb.set_centers(c * 1e20)
    Lb.assign(gamma)
    a, _ = Lb.assignments()
    Xr = Xb.copy(); Xr.data = Xr.data.astype(np.float32).astype(np.float64)
    wa, _, _ = host_ref.find_cluster_assignments(Xr, c * 1e20, gamma)
    assert np.array_equal(a, wa)
    Lb.close(); dsb.close()


def test_tc_lloyd_trajectory_identical(ctx):
    """Ten Lloyd iterations from a poor start: the tensor-core plan and the gather kernels agree every iteration,
    with and without the bounded / incremental modes on top."""
    from sparsifiedkmeans_b200 import Dataset, Lloyd
    X, c, gamma = make_sparsified(p=512, n=20000, m=26, K=48, seed=11, kind="mixture", f32=True)
    rng = np.random.default_rng(3)
    start = c[:, rng.integers(0, 48, 48)] + 0.3 * rng.standard_normal(c.shape)      # duplicates + noise
    ds = Dataset.from_scipy(X, store="f32", ctx=ctx)
    runs = []
    for tc, modes in [(False, False), (True, False), (True, True)]:
        L = Lloyd(ds, 48, incremental=modes, bounded=modes)
        L.set_tc_filter(tc)
        L.set_centers(start)
        hist = []
        for _ in range(10):
            L.step(gamma, gamma)
            a, _ = L.assignments(want_dist=False)
            hist.append(a.copy())
        runs.append((hist, L.get_centers()))
        L.close()
    for hist, cen in runs[1:]:
        for it, (h0, h1) in enumerate(zip(runs[0][0], hist)):
            assert np.array_equal(h0, h1), f"iteration {it}: {np.count_nonzero(h0 != h1)} differ"
        np.testing.assert_allclose(cen, runs[0][1], rtol=1e-9, atol=1e-12)
    ds.close()
