"""Partial-distance pruning of the full assignment pass (skm_lloyd_set_prune): a prefix pass over ~15 % of every column's
entries + the exact evaluation of its winner + the fallback give the oracle's assignments on separated AND on structureless
data; the automatic mode backs off when nothing can be pruned."""
import numpy as np
import pytest

from oracle import host_ref
from tests.util import make_sparsified

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("prefix", ["half", "f32"])
@pytest.mark.parametrize("kind", ["mixture", "unstructured"])
@pytest.mark.parametrize("K,p,m,ragged", [(64, 1024, 51, False), (33, 256, 26, True), (100, 512, 40, False), (17, 784, 78, True),
                                          (130, 300, 40, True), (40, 2048, 40, False), (40, 4096, 40, False)])
def test_pruned_pass_matches_reference(ctx, monkeypatch, prefix, kind, K, p, m, ragged):
    """Both prefix kernels: the one-launch half-precision table (prefix16.cu, default) and the fp32 K1 kernels.  p = 2048:
    the 64-centre half table does not fit, two launches of 32 centres; p = 4096: no half table fits, fp32 prefix."""
    from sparsifiedkmeans_b200 import Dataset, Lloyd
    if prefix == "f32":
        monkeypatch.setenv("SKM_PRUNE_F32", "1")
    else:
        monkeypatch.delenv("SKM_PRUNE_F32", raising=False)
    X, c, gamma = make_sparsified(p=p, n=6000, m=m, K=K, seed=K * 3 + p, kind=kind, f32=True, ragged=ragged)
    ds = Dataset.from_scipy(X, store="f32", ctx=ctx)
    wa, wd, _ = host_ref.find_cluster_assignments(X, c, gamma)
    L = Lloyd(ds, K)
    L.set_prune(True)
    L.set_centers(c)
    L.assign(gamma)
    a, d = L.assignments()
    assert np.array_equal(a, wa), f"{np.count_nonzero(a != wa)} assignments differ"
    np.testing.assert_allclose(d, wd, rtol=2e-5, atol=1e-30)
    not_kept, pairs = L.last_prune()
    assert pairs >= 2 and 0 <= not_kept <= X.shape[1]
    limit = X.shape[1] // (4 if K <= 128 else 16)          # with / without the fp32 list kernel behind the pruned pass (K <= 128)
    assert ("k_prefix16" in L.kernel_name) == (prefix == "half" and p < 4096 and not_kept <= limit), L.kernel_name
    if p == 2048 and prefix == "half" and not_kept <= limit:
        assert "k_prefix16<32> x2" in L.kernel_name, L.kernel_name
    if kind == "mixture" and not ragged:
        assert not_kept <= X.shape[1] // 20, "separated clusters: the prefix pass bounds (nearly) every other centre away"
    # the sums that follow are the same as without pruning
    L.accumulate(); L.finalize(gamma)
    L0 = Lloyd(ds, K)
    L0.set_prune(False)
    L0.set_centers(c)
    L0.assign(gamma); L0.accumulate(); L0.finalize(gamma)
    assert L0.last_prune() == (-1, -1)
    a0, _ = L0.assignments()
    assert np.array_equal(a0, a)
    np.testing.assert_array_equal(L.get_centers(), L0.get_centers())
    L.close(); L0.close(); ds.close()


def test_prune_trajectory_and_backoff(ctx):
    """Automatic mode from a poor start, with and without the stateful modes: identical assignments every iteration;
    on structureless data the pruned pass is tried, fails, and sits out the following calls."""
    from sparsifiedkmeans_b200 import Dataset, Lloyd
    X, c, gamma = make_sparsified(p=512, n=20000, m=26, K=48, seed=11, kind="mixture", f32=True)
    rng = np.random.default_rng(3)
    start = c[:, rng.integers(0, 48, 48)] + 0.3 * rng.standard_normal(c.shape)
    ds = Dataset.from_scipy(X, store="f32", ctx=ctx)
    runs = []
    for prune, modes in [(False, False), (None, False), (True, True)]:
        L = Lloyd(ds, 48, incremental=modes, bounded=modes)
        L.set_prune(prune)
        L.set_centers(start)
        hist = []
        for _ in range(10):
            L.step(gamma, gamma)
            hist.append(L.assignments(want_dist=False)[0].copy())
        runs.append((hist, L.get_centers()))
        L.close()
    for hist, cen in runs[1:]:
        for it, (h0, h1) in enumerate(zip(runs[0][0], hist)):
            assert np.array_equal(h0, h1), f"iteration {it}: {np.count_nonzero(h0 != h1)} differ"
        np.testing.assert_allclose(cen, runs[0][1], rtol=1e-9, atol=1e-12)
    ds.close()
    Xu, cu, gu = make_sparsified(p=512, n=20000, m=26, K=48, seed=12, kind="unstructured", f32=True)
    dsu = Dataset.from_scipy(Xu, store="f32", ctx=ctx)
    L = Lloyd(dsu, 48)
    L.set_centers(cu)
    tried = []
    for _ in range(6):
        L.assign(gu)
        tried.append(L.last_prune()[0] >= 0)
    wa, _, _ = host_ref.find_cluster_assignments(Xu, cu, gu)
    assert np.array_equal(L.assignments(want_dist=False)[0], wa)
    assert tried[0] and not all(tried), tried             # tried once, then backed off
    L.close(); dsu.close()


def test_modes_enabled_after_a_pruned_pass(ctx):
    """The pruned pass allocates some of the bounded mode's buffers; switching the modes on afterwards must still work
    (regression: set_assign_mode used to skip its own allocations when the bound array already existed)."""
    from sparsifiedkmeans_b200 import Dataset, Lloyd
    X, c, gamma = make_sparsified(p=512, n=12000, m=26, K=40, seed=21, kind="mixture", f32=True)
    ds = Dataset.from_scipy(X, store="f32", ctx=ctx)
    L = Lloyd(ds, 40)
    L.set_centers(c + 0.2 * np.random.default_rng(1).standard_normal(c.shape))
    L.step(gamma, gamma)
    assert L.last_prune()[0] >= 0
    L.set_update_mode(True)
    L.set_assign_mode(True)
    L0 = Lloyd(ds, 40)
    L0.set_prune(False)
    L0.set_centers(L.get_centers())
    for _ in range(6):
        L.step(gamma, gamma)
        L0.step(gamma, gamma)
        assert np.array_equal(L.assignments(want_dist=False)[0], L0.assignments(want_dist=False)[0])
    np.testing.assert_allclose(L.get_centers(), L0.get_centers(), rtol=1e-9, atol=1e-12)
    L.close(); L0.close(); ds.close()


@pytest.mark.parametrize("scale", [1e-30, 1e-12, 1e-6, 1.0, 3e4, 1e12, 1e25])
def test_half_table_prefix_is_scale_free(ctx, monkeypatch, scale):
    """The half-precision table stores s c' with s a power of two chosen from max |c'|: data and centres far outside the
    fp16 range prune exactly like data of order one (same columns kept), and the assignments stay the oracle's."""
    from sparsifiedkmeans_b200 import Dataset, Lloyd
    monkeypatch.delenv("SKM_PRUNE_F32", raising=False)
    X, c, gamma = make_sparsified(p=512, n=8000, m=40, K=48, seed=5, kind="mixture", f32=True)
    X = X.copy()
    X.data = (X.data.astype(np.float32) * np.float32(scale)).astype(X.data.dtype)
    c = c * float(np.float32(scale))
    ds = Dataset.from_scipy(X, store="f32", ctx=ctx)
    wa, wd, _ = host_ref.find_cluster_assignments(X, c, gamma)
    L = Lloyd(ds, 48)
    L.set_prune(True)
    L.set_centers(c)
    L.assign(gamma)
    a, d = L.assignments()
    assert np.array_equal(a, wa)
    np.testing.assert_allclose(d, wd, rtol=2e-5, atol=0)
    if 1e-13 < scale < 1e13:                 # beyond that the fp32 SQUARES leave the fp32 range: everything goes to fp64
        assert L.last_prune()[0] <= X.shape[1] // 20, L.last_prune()
    L.close(); ds.close()


def test_half_table_prefix_nan_and_huge_centres(ctx, monkeypatch):
    """A NaN centre entry and a centre 1e30 away: nothing is kept on a wrong bound, the oracle's MATLAB-min semantics hold."""
    from sparsifiedkmeans_b200 import Dataset, Lloyd
    monkeypatch.delenv("SKM_PRUNE_F32", raising=False)
    X, c, gamma = make_sparsified(p=256, n=4000, m=26, K=20, seed=9, kind="mixture", f32=True)
    for variant in ("nan", "huge"):
        c2 = c.copy()
        if variant == "nan":
            c2[7, 3] = np.nan
        else:
            c2[:, 5] = 1e30
        ds = Dataset.from_scipy(X, store="f32", ctx=ctx)
        wa, wd, _ = host_ref.find_cluster_assignments(X, c2, gamma)
        L = Lloyd(ds, 20)
        L.set_prune(True)
        L.set_centers(c2)
        L.assign(gamma)
        a, d = L.assignments()
        assert np.array_equal(a, wa), variant
        L.close(); ds.close()


@pytest.mark.parametrize("kind", ["mixture", "unstructured"])
@pytest.mark.parametrize("K,p,m,ragged", [(64, 1024, 51, False), (33, 256, 26, True), (100, 512, 40, True), (20, 784, 78, False)])
def test_list_pass_reevaluates_what_was_not_kept(ctx, monkeypatch, kind, K, p, m, ragged):
    """Columns a pruned or bounded pass cannot keep are re-evaluated by the fp32 list kernel (k_assign_cols: a warp per
    column, K1's arithmetic and guard) with fp64 behind it; SKM_LIST_MIN=0 sends even short lists that way.  Centres moved off the
    planted ones so that a good share of the columns is NOT kept; assignments, distances and bounds as without it."""
    from sparsifiedkmeans_b200 import Dataset, Lloyd
    X, c, gamma = make_sparsified(p=p, n=9000, m=m, K=K, seed=K + p, kind=kind, f32=True, ragged=ragged)
    rng = np.random.default_rng(K)
    c1 = c + 0.6 * rng.standard_normal(c.shape) * (np.abs(c).mean() if kind == "mixture" else 1.0)
    ds = Dataset.from_scipy(X, store="f32", ctx=ctx)
    for lm in ("0", None):
        if lm is None:
            monkeypatch.delenv("SKM_LIST_MIN", raising=False)
        else:
            monkeypatch.setenv("SKM_LIST_MIN", lm)
        L = Lloyd(ds, K, bounded=True)
        L.set_prune(True)
        L.set_centers(c1)
        hist = []
        for it in range(4):                                   # pass 0: pruned + list; later passes: bounded + list
            wa, wd, _ = host_ref.find_cluster_assignments(X, L.get_centers(), gamma)
            L.assign(gamma)
            a, d = L.assignments()
            assert np.array_equal(a, wa), f"list_min={lm} pass {it}: {np.count_nonzero(a != wa)} assignments differ"
            np.testing.assert_allclose(d, wd, rtol=2e-5, atol=1e-30)
            hist.append(L.last_prune())
            L.accumulate(); L.finalize(gamma)
        L.close()
    ds.close()


def test_pruned_pass_with_duplicate_centres(ctx, monkeypatch):
    """Two identical centres tie exactly for half of the columns: the pruned pass can never keep those (its bound on the twin
    equals the candidate's own distance), the list kernel cannot certify them, and the fp64 kernel applies MATLAB's
    first-index rule -- the reference's assignments, also with the twin in a different 64-centre chunk."""
    from sparsifiedkmeans_b200 import Dataset, Lloyd
    monkeypatch.delenv("SKM_PRUNE_F32", raising=False)
    for K, twin in ((40, (3, 17)), (100, (5, 90))):
        X, c, gamma = make_sparsified(p=256, n=5000, m=26, K=K, seed=K, kind="mixture", f32=True)
        c = c.copy()
        c[:, twin[1]] = c[:, twin[0]]
        ds = Dataset.from_scipy(X, store="f32", ctx=ctx)
        wa, wd, _ = host_ref.find_cluster_assignments(X, c, gamma)
        assert np.count_nonzero(wa == twin[0] + 1) > 0 and np.count_nonzero(wa == twin[1] + 1) == 0   # labels are 1-based; the first index wins every tie
        L = Lloyd(ds, K)
        L.set_prune(True)
        L.set_centers(c)
        L.assign(gamma)
        a, d = L.assignments()
        assert np.array_equal(a, wa), f"K={K}: {np.count_nonzero(a != wa)} assignments differ"
        np.testing.assert_allclose(d, wd, rtol=2e-5, atol=1e-30)
        L.close(); ds.close()
