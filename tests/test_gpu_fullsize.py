"""Parity at BASELINE.json's full single-GPU size (configs[1]: n=1e7, p=784, K=10, 78 entries per
point) through size-independent properties, plus an oracle check on column slices copied back to
the host (SURVEY.md section 8d 'parity check at scale')."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu

N, P, K, M = 10_000_000, 784, 10, 78


@pytest.fixture(scope="module")
def big(ctx):
    import torch
    import bench
    from sparsifiedkmeans_b200 import Dataset, Lloyd
    from sparsifiedkmeans_b200._lib import SKM_F32, SKM_I32, SKM_I64
    dev = torch.device("cuda:0")
    colptr, rowidx, val, mu, start = bench.gen_shard_device(dev, N, P, M, K, col0=0, ctx=ctx)
    torch.cuda.synchronize()
    # host copies of three slices (head, middle, tail) for the oracle
    slices = {}
    for name, j0 in (("head", 0), ("middle", N // 2 - 3), ("tail", N - 100_000)):
        j1 = j0 + 100_000
        r = rowidx[j0 * M:j1 * M].cpu().numpy().astype(np.int64)
        v = val[j0 * M:j1 * M].cpu().numpy().astype(np.float64)
        jc = np.arange(100_001, dtype=np.int64) * M
        slices[name] = (j0, sp.csc_matrix((v, r, jc), shape=(P, 100_000)))
    total_val = float(val.double().sum().item())
    ds = Dataset.from_device_csc(P, N, colptr.data_ptr(), SKM_I64, rowidx.data_ptr(), SKM_I32, val.data_ptr(), SKM_F32,
                                 store="f32", ctx=ctx)
    del colptr, rowidx, val
    torch.cuda.empty_cache()
    L = Lloyd(ds, K)
    yield dict(ds=ds, L=L, start=start, slices=slices, total_val=total_val, gamma=M / P)
    L.close(); ds.close()


def test_fullsize_assignments_match_oracle_on_slices(big):
    from oracle import host_ref
    L, start, gamma = big["L"], big["start"], big["gamma"]
    L.set_centers(start)
    L.assign(gamma)
    a, d = L.assignments()
    assert a.shape == (N,) and a.min() >= 1 and a.max() <= K
    for name, (j0, Xs) in big["slices"].items():
        wa, wd, _ = host_ref.find_cluster_assignments(Xs, start, gamma)
        assert np.array_equal(a[j0:j0 + 100_000], wa), name
        np.testing.assert_allclose(d[j0:j0 + 100_000], wd, rtol=2e-5)
    big["a"], big["d"] = a, d


def test_fullsize_conservation_laws(big):
    """Checksums that hold for any correct K2/K3 whatever n: members sum to n, support counts sum
    to nnz, per-row sums of S equal the per-row sums of X, centres = gamma*S/(N+1e-16)."""
    import torch
    L, gamma = big["L"], big["gamma"]
    L.set_centers(big["start"])
    st = L.step(gamma, gamma, True)
    assert st.n_points == N and st.n_empty == 0 and not st.has_nan
    counts = L.counts()
    assert counts.sum() == N
    a = big.get("a")
    if a is not None:
        assert np.array_equal(counts, np.bincount(a - 1, minlength=K))
    part = L.partials_tensor().cpu().numpy()
    S = part[:P * K].reshape(K, P).T
    Nn = part[P * K:2 * P * K].reshape(K, P).T
    assert Nn.sum() == N * M                                   # every stored entry counted exactly once
    assert np.all(Nn == np.rint(Nn))
    np.testing.assert_allclose(S.sum(), big["total_val"], rtol=1e-9)
    C = L.get_centers()
    np.testing.assert_allclose(C, gamma * S / (Nn + 1e-16), rtol=1e-14, atol=0)
    np.testing.assert_allclose(st.sumsq, float(np.sum(big["d"] ** 2)) if "d" in big else st.sumsq, rtol=1e-6)


def test_fullsize_idempotence_and_fixed_point(big):
    """Re-assigning against the same centres reproduces the assignments; iterating to the fixed point
    and stepping once more changes nothing (dff == 0)."""
    L, gamma = big["L"], big["gamma"]
    L.set_centers(big["start"])
    for _ in range(20):
        st = L.step(gamma, gamma, True)
        if st.dff == 0.0:
            break
    assert st.dff < 1e-9
    a1, _ = L.assignments(want_dist=False)
    c1 = L.get_centers()
    L.assign(gamma)
    a2, _ = L.assignments(want_dist=False)
    assert np.array_equal(a1, a2)
    st2 = L.step(gamma, gamma, True)
    np.testing.assert_allclose(L.get_centers(), c1, rtol=1e-12, atol=0)
    assert st2.dff < 1e-9
    lab = (np.arange(N) % K)
    # planted partition: each planted label maps to one cluster
    for k in range(K):
        vals = a2[k::K]
        assert vals.min() == vals.max()


def test_fullsize_streamed_equals_resident(big):
    """skm_lloyd_step_host over a 2e6-column slab of the same data == resident path on that slab."""
    from oracle import cport
    from sparsifiedkmeans_b200 import lloyd_step_host
    j0, Xs = big["slices"]["head"]
    start, gamma = big["start"], big["gamma"]
    newc, a, _, st = lloyd_step_host(P, Xs.shape[1], Xs.indptr, Xs.indices.astype(np.uint16), Xs.data.astype(np.float32),
                                     start, gamma, gamma, True, chunk_cols=20_000, ctx=big["ds"].ctx)
    from oracle import host_ref
    wa, _, _ = host_ref.find_cluster_assignments(Xs, start, gamma)
    assert np.array_equal(a, wa)
    want, _, _, _ = cport.centroid_update(P, Xs.shape[1], K, Xs.indptr, Xs.indices, Xs.data, wa, gamma, start, True)
    assert np.max(np.abs(newc - want)) <= 1e-6 * np.max(np.abs(want))


def test_fullsize_config3_shard_k64(ctx):
    """BASELINE.json configs[2]'s per-GPU shard (n=1.25e7, p=1024, K=64, 51 entries per point): the K=64
    plan (four launches of 16 centres on the dual table, merged through best2) against the oracle on
    slices, the conservation laws of one full iteration, and a conflict-free image."""
    import torch
    import bench
    from oracle import host_ref
    from sparsifiedkmeans_b200 import Dataset, Lloyd
    from sparsifiedkmeans_b200._lib import SKM_F32, SKM_I32, SKM_I64
    n, p, k, m = 12_500_000, 1024, 64, 51
    dev = torch.device("cuda:0")
    colptr, rowidx, val, mu, start = bench.gen_shard_device(dev, n, p, m, k, col0=0)
    torch.cuda.synchronize()
    slices = {}
    w = 30_000
    for name, j0 in (("head", 0), ("middle", n // 2 - 7), ("tail", n - w)):
        r = rowidx[j0 * m:(j0 + w) * m].cpu().numpy().astype(np.int64)
        v = val[j0 * m:(j0 + w) * m].cpu().numpy().astype(np.float64)
        slices[name] = (j0, sp.csc_matrix((v, r, np.arange(w + 1, dtype=np.int64) * m), shape=(p, w)))
    ds = Dataset.from_device_csc(p, n, colptr.data_ptr(), SKM_I64, rowidx.data_ptr(), SKM_I32, val.data_ptr(), SKM_F32,
                                 store="f32", ctx=ctx)
    del colptr, rowidx, val
    torch.cuda.empty_cache()
    L = Lloyd(ds, k)
    try:
        gamma = m / p
        L.set_centers(start)
        st = L.step(gamma, gamma, True)
        a, d = L.assignments()
        assert st.n_points == n and a.min() >= 1 and a.max() <= k
        for name, (j0, Xs) in slices.items():
            wa, wd, _ = host_ref.find_cluster_assignments(Xs, start, gamma)
            assert np.array_equal(a[j0:j0 + w], wa), name
            np.testing.assert_allclose(d[j0:j0 + w], wd, rtol=2e-5)
        assert np.array_equal(np.bincount(a - 1, minlength=k), L.counts())
        np.testing.assert_allclose(st.sumsq, float(np.sum(d.astype(np.float64) ** 2)), rtol=1e-6)
        chk = ds.layout_check()
        assert chk["bad_columns"] == 0
        if "dual table" in L.kernel_name:
            assert chk["layout"] == 2 and chk["wavefronts"] <= 1.0001 * chk["steps"]
    finally:
        L.close(); ds.close()
