"""Several GPUs behind ONE host process (skm_multi_* / multi.py / the MEX gateway's 'upload' with a device count):
the CUDA path on N devices against the oracle.  The N = 1 cases always run; the N >= 2 cases run when the box shows
that many devices (`gpurun --gpus 2`), which is SURVEY.md section 4 test plan (iv): the same assignments and
centres (<= 1e-6) whatever the number of GPUs."""
import numpy as np
import pytest

from oracle import cport, host_ref, refmex
from tests.util import make_sparsified

pytestmark = pytest.mark.gpu


def _ndev():
    import torch
    return torch.cuda.device_count()


def _device_counts():
    return [g for g in (1, 2, 4, 8) if g <= max(_ndev(), 1)]


@pytest.fixture(scope="module", params=[1, 2, 4, 8])
def mctx(request):
    if request.param > _ndev():
        pytest.skip(f"needs {request.param} GPUs")
    from sparsifiedkmeans_b200 import MultiContext
    m = MultiContext(request.param)
    yield m
    m.close()


@pytest.mark.parametrize("kind,K,p,m_", [("mixture", 5, 64, 8), ("unstructured", 7, 96, 9), ("mixture", 64, 128, 12)])
def test_multi_lloyd_iterations_match_oracle(mctx, kind, K, p, m_):
    from sparsifiedkmeans_b200 import MultiDataset, MultiLloyd
    X, c, gamma = make_sparsified(p=p, n=4003, m=m_, K=K, seed=11 + K, kind=kind)
    ds = MultiDataset.from_scipy(X, store="f32", mctx=mctx)
    assert ds.n == X.shape[1] and ds.col0[0] == 0
    L = MultiLloyd(ds, K)
    L.set_centers(c)
    cen = c.copy()
    for it in range(4):
        st = L.step(gamma, gamma, True)
        wa, wd, _ = host_ref.find_cluster_assignments(X, cen, gamma)
        a, d = L.assignments()
        assert np.array_equal(a, wa), f"iteration {it}"                       # bit-exact on every shard
        np.testing.assert_allclose(d, wd, rtol=2e-5)
        want, _, _, counts = cport.centroid_update(p, X.shape[1], K, X.indptr, X.indices, X.data, wa, gamma, cen, True)
        got = L.get_centers()
        assert np.max(np.abs(got - want)) <= 1e-6 * np.max(np.abs(want))
        assert np.array_equal(L.counts(), counts)
        np.testing.assert_allclose(st.sumsq, np.sum(wd ** 2), rtol=1e-5)
        for g in range(mctx.ndev):                                             # identical bits on every device
            assert np.array_equal(L.get_centers_of(g), got)
        cen = got
    v, j = L.argmax_distance()
    _, dd = L.assignments()
    assert j == int(np.argmax(dd)) and v == dd[j]
    np.testing.assert_array_equal(ds.get_column(j), np.asarray(X[:, j].todense()).ravel())
    L.close(); ds.close()


def test_multi_result_does_not_depend_on_device_count(mctx):
    """Same run on 1 device (plain handles) and on the group: identical assignments, centres <= 1e-9 apart
    (the sums run in a different order)."""
    from sparsifiedkmeans_b200 import Dataset, Lloyd, MultiDataset, MultiLloyd
    X, c, gamma = make_sparsified(p=64, n=9000, m=8, K=6, seed=3, kind="mixture")
    ds1 = Dataset.from_scipy(X, store="f32", ctx=mctx.contexts[0])
    L1 = Lloyd(ds1, 6); L1.set_centers(c)
    dsm = MultiDataset.from_scipy(X, store="f32", mctx=mctx)
    Lm = MultiLloyd(dsm, 6); Lm.set_centers(c)
    for _ in range(6):
        s1 = L1.step(gamma, gamma, True)
        sm = Lm.step(gamma, gamma, True)
        assert np.array_equal(L1.assignments()[0], Lm.assignments()[0])
        np.testing.assert_allclose(Lm.get_centers(), L1.get_centers(), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(sm.dff, s1.dff, rtol=1e-6, atol=1e-12)
    L1.close(); ds1.close(); Lm.close(); dsm.close()


def test_multi_modes_keep_assignments(mctx):
    """bounded assignment + incremental update per shard: same assignments as the default iteration."""
    from sparsifiedkmeans_b200 import MultiDataset, MultiLloyd
    X, c, gamma = make_sparsified(p=64, n=8000, m=8, K=6, seed=5, kind="mixture")
    ds = MultiDataset.from_scipy(X, store="f32", mctx=mctx)
    A = MultiLloyd(ds, 6); A.set_centers(c + 0.3)
    B = MultiLloyd(ds, 6, incremental=True, bounded=True); B.set_centers(c + 0.3)
    for _ in range(8):
        A.step(gamma, gamma, True); B.step(gamma, gamma, True)
        assert np.array_equal(A.assignments()[0], B.assignments()[0])
        np.testing.assert_allclose(B.get_centers(), A.get_centers(), rtol=1e-9, atol=1e-11)
    A.close(); B.close(); ds.close()


def test_multi_kmeanspp_equals_oracle(mctx):
    from sparsifiedkmeans_b200 import Arthur_initialization, MultiDataset
    X, _, gamma = make_sparsified(p=64, n=3001, m=8, K=6, seed=71, kind="mixture")
    ds = MultiDataset.from_scipy(X, store="f64", mctx=mctx)
    u = np.random.default_rng(3).random(4000)
    for g in (gamma, None):                                 # dense-centre and sparse-centre (no gamma) forms
        idx = Arthur_initialization(ds, 6, g, first=700, uniforms=iter(u))
        want = host_ref.arthur_initialization(X, 6, g, 700, iter(u))
        assert np.array_equal(idx, np.asarray(want[0] if isinstance(want, tuple) else want))
    ds.close()


def test_kmeans_sparsified_devices_option(mctx):
    """kmeans_sparsified(..., Devices=group): same clustering as the single-device call with the same seed."""
    from sparsifiedkmeans_b200 import kmeans_sparsified
    rng = np.random.default_rng(0)
    K, p, n = 4, 64, 3000
    mu = 3.0 * rng.standard_normal((K, p))
    lab = rng.integers(0, K, n)
    Xr = mu[lab] + 0.3 * rng.standard_normal((n, p))
    common = dict(Sparsify=True, SparsityLevel=0.2, SketchType="Hadamard", Seed=7, Replicates=2)
    I1, C1, S1, D1, O1 = kmeans_sparsified(Xr, K, Context=mctx.contexts[0], **common)
    I2, C2, S2, D2, O2 = kmeans_sparsified(Xr, K, Devices=mctx, **common)
    assert O2["Devices"] == mctx.devices
    assert np.array_equal(I1, I2)
    np.testing.assert_allclose(C2, C1, rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(S2, S1, rtol=1e-6)


def test_lloyd_gateway_multi_gpu(mctx):
    """mex/skm_lloyd_mex.c 'upload' with a device count: the gateway a MATLAB session would call."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["make", "-C", os.path.join(root, "mex")], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    lib = refmex._load(os.path.join(root, "mex", "_build", "libmex_skm_lloyd_mex.so"))
    X, c, gamma = make_sparsified(p=64, n=2000, m=8, K=5, seed=43, kind="mixture")
    p, n = X.shape
    keep = []
    h = refmex.call_mex(lib, [refmex.mx_string("upload", keep), refmex.mx_sparse(p, n, X.indptr, X.indices, X.data, keep),
                              refmex.mx_dense(np.array([[5.0]]), keep), refmex.mx_dense(np.array([[float(mctx.ndev)]]), keep)], 1)[0]
    hm = refmex.mx_dense(h, keep)
    g = refmex.mx_dense(np.array([[gamma]]), keep)
    one = refmex.mx_dense(np.array([[1.0]]), keep)
    a, d, cen, dff, sumsq, counts = refmex.call_mex(lib, [refmex.mx_string("iterate", keep), hm, refmex.mx_dense(c, keep), g, g, one], 6)
    wa, wd, _ = host_ref.find_cluster_assignments(X, c, gamma)
    assert np.array_equal(a.ravel().astype(np.int64), wa)
    want, _, _, wc = cport.centroid_update(p, n, 5, X.indptr, X.indices, X.data, wa, gamma, c, True)
    np.testing.assert_allclose(cen, want, rtol=1e-6, atol=1e-9)
    assert np.array_equal(counts.ravel().astype(np.int64), wc)
    a2, d2 = refmex.call_mex(lib, [refmex.mx_string("assign", keep), hm, refmex.mx_dense(c, keep), refmex.mx_empty()], 2)
    wa2, _, _ = host_ref.find_cluster_assignments(X, c, None)
    assert np.array_equal(a2.ravel().astype(np.int64), wa2)
    refmex.call_mex(lib, [refmex.mx_string("free", keep), hm], 0)


def test_incremental_update_structural_zeros(ctx):
    """A cell (row, cluster) whose contributors all leave while the cluster stays non-empty must come out as the
    reference's exact 0, not a rounding residue divided by 1e-16 (advisor finding, round 1): a tiny cluster loses
    the columns that touched some rows under update_mode = 1, compared with the full recompute."""
    from sparsifiedkmeans_b200 import Dataset, Lloyd
    rng = np.random.default_rng(9)
    X, c, gamma = make_sparsified(p=48, n=600, m=6, K=3, seed=21, kind="mixture")
    ds = Dataset.from_scipy(X, store="f32", ctx=ctx)
    A = Lloyd(ds, 3); B = Lloyd(ds, 3, incremental=True)
    cen = c.copy()
    cen[:, 2] += 4.0                                        # cluster 2 starts far away: few members
    A.set_centers(cen); B.set_centers(cen)
    for it in range(12):
        A.step(gamma, gamma, True); B.step(gamma, gamma, True)
        ca, cb = A.get_centers(), B.get_centers()
        assert np.array_equal(ca == 0.0, cb == 0.0), f"iteration {it}: structural zeros differ"
        np.testing.assert_allclose(cb, ca, rtol=1e-9, atol=1e-12)
        # pull a different random subset of columns towards cluster 2 so that members come and go
        pert = ca.copy()
        pert[:, 2] = ca[:, rng.integers(0, 2)] + 0.5 * rng.standard_normal(48)
        A.set_centers(pert); B.set_centers(pert)
    A.close(); B.close(); ds.close()
