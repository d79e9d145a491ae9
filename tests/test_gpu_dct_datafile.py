"""GPU tests of the DCT sketch (kmeans_sparsified.m:226-231,256-258) and of
the DataFile mode (rank 4; kmeans_sparsified.m:180-207, private/sampleAndMixFromLargeFile.m).

MATLAB's dct/idct are the orthonormal DCT-II / DCT-III along columns = scipy.fft.dct(type=2,
norm='ortho') / idct, which is the oracle here (the reference holds no vectors for them)."""
import numpy as np
import pytest
import scipy.fft

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("p,n", [(50, 7), (784, 33), (100, 1), (513, 20)])
def test_dct_mix_matches_scipy(ctx, p, n):
    from sparsifiedkmeans_b200 import dct_mix
    rng = np.random.default_rng(p)
    X = rng.standard_normal((p, n))
    d = np.sign(rng.standard_normal(p))
    Y = dct_mix(X, d, False, ctx)
    want = scipy.fft.dct(d[:, None] * X, type=2, norm="ortho", axis=0)
    np.testing.assert_allclose(Y, want, rtol=0, atol=1e-12 * np.abs(want).max() * p)
    back = dct_mix(Y, d, True, ctx)                      # unmix(mix(X)) = X
    np.testing.assert_allclose(back, X, rtol=0, atol=1e-11)
    np.testing.assert_allclose(dct_mix(X, None, True, ctx), scipy.fft.idct(X, type=2, norm="ortho", axis=0), atol=1e-11)


def test_dct_pipeline_with_explicit_rows_matches_oracle(ctx):
    from sparsifiedkmeans_b200 import Dataset
    rng = np.random.default_rng(3)
    p, n, m = 200, 1500, 20
    X = rng.standard_normal((p, n))
    d = np.sign(rng.standard_normal(p))
    rows = np.stack([np.sort(rng.choice(p, m, replace=False)) for _ in range(n)], axis=1)
    ds = Dataset.from_dense_host_dct(X, d, m, rows=rows, chunk_cols=400, ctx=ctx)
    got = ds.to_scipy().toarray()
    eps2 = 1 + 2 * np.finfo(np.float64).eps
    Y = scipy.fft.dct(d[:, None] * (X * eps2), type=2, norm="ortho", axis=0)
    want = np.zeros_like(Y)
    cols = np.repeat(np.arange(n), m)
    r = rows.T.reshape(-1)
    want[r, cols] = Y[r, cols] / (m / p)                                     # randsample_fixedNumberEntries.m:30-31
    assert np.array_equal(got != 0, want != 0)
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-5 * np.abs(want).max())   # fp32 product
    ds.close()
    with pytest.raises(Exception):
        bad = rows.copy(); bad[1, 0] = bad[0, 0]
        Dataset.from_dense_host_dct(X, d, m, rows=bad, ctx=ctx)


def test_general_row_sampler_contract(ctx):
    """Exactly m distinct ascending rows per column, uniform marginals, a pure function of
    (seed, global column) -- the contract of private/randsample_block.m:44-84 for any p."""
    import torch
    from sparsifiedkmeans_b200 import sample_rows_general
    p, n, m = 784, 20000, 78
    buf = torch.empty(n * m, dtype=torch.int32, device="cuda:0")
    sample_rows_general(p, n, m, 11, 0, buf.data_ptr(), ctx)
    ctx.synchronize()
    R = buf.cpu().numpy().reshape(n, m)
    assert R.min() >= 0 and R.max() < p
    assert np.all(np.diff(R, axis=1) > 0)                                    # ascending, distinct
    freq = np.bincount(R.reshape(-1), minlength=p) / (n * m / p)
    assert abs(freq.mean() - 1) < 1e-12 and freq.std() < 4 * np.sqrt((1 - m / p) / (n * m / p))
    part = torch.empty(5000 * m, dtype=torch.int32, device="cuda:0")
    sample_rows_general(p, 5000, m, 11, 7000, part.data_ptr(), ctx)          # a shard starting at column 7000
    ctx.synchronize()
    assert np.array_equal(part.cpu().numpy().reshape(5000, m), R[7000:12000])
    other = torch.empty(100 * m, dtype=torch.int32, device="cuda:0")
    sample_rows_general(p, 100, m, 12, 0, other.data_ptr(), ctx)
    ctx.synchronize()
    assert not np.array_equal(other.cpu().numpy().reshape(100, m), R[:100])


def _mixture(p, n, K, seed, sigma=0.1):
    rng = np.random.default_rng(seed)
    mu = rng.standard_normal((p, K))
    lab = rng.integers(0, K, n)
    return (mu[:, lab] + sigma * rng.standard_normal((p, n))).T.copy(), lab      # rows are points


def _accuracy(idx, lab, K):
    from itertools import permutations
    return max(np.mean(np.asarray(perm)[idx - 1] == lab) for perm in permutations(range(K)))


def test_kmeans_with_auto_dct_sketch(ctx):
    """p = 50 is not a power of two, so SketchType 'auto' picks the DCT (kmeans_sparsified.m:226-231)."""
    from sparsifiedkmeans_b200 import kmeans_sparsified
    X, lab = _mixture(50, 3000, 4, seed=2)
    for pipe in ("device", "host"):
        IDX, C, SUMD, D, OUT = kmeans_sparsified(X, 4, Sparsify=True, SparsityLevel=0.3, Seed=5, Replicates=3,
                                                 Pipeline=pipe, Context=ctx)
        assert OUT["SketchType"] == "DCT" and OUT["Pipeline"] == pipe
        assert _accuracy(IDX, lab, 4) > 0.97
        truth = np.stack([X[lab == k].mean(axis=0) for k in range(4)])
        dd = np.linalg.norm(C[:, None, :] - truth[None, :, :], axis=2)           # unmixed centres sit on the truth
        assert np.all(dd.min(axis=1) < 0.35 * np.linalg.norm(truth, axis=1).mean())


def test_kmeans_from_data_file(ctx, tmp_path):
    """DataFile mode: the matrix is memory-mapped and streamed; same answer as the in-core call with
    the same seed, including the two-pass outputs."""
    from sparsifiedkmeans_b200 import kmeans_sparsified
    X, lab = _mixture(64, 4000, 3, seed=8)
    path = str(tmp_path / "points.npy")
    np.save(path, X)
    kw = dict(Sparsify=True, SparsityLevel=0.25, Seed=3, Replicates=2, nargout=9, Context=ctx)
    mem = kmeans_sparsified(X, 3, **kw)
    dsk = kmeans_sparsified(path, 3, **kw)
    via_opt = kmeans_sparsified(None, 3, DataFile=path[:-4], **kw)               # extension is optional, :187-189
    assert dsk[4]["LoadFromDisk"] and not mem[4]["LoadFromDisk"]
    for got in (dsk, via_opt):
        assert np.array_equal(got[0], mem[0])
        np.testing.assert_allclose(got[1], mem[1], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(got[5], mem[5], rtol=1e-9, atol=1e-12)
        assert np.array_equal(got[6], mem[6])
    assert _accuracy(dsk[0], lab, 3) > 0.98
    np.save(str(tmp_path / "f32.npy"), X.astype(np.float32))
    f32 = kmeans_sparsified(str(tmp_path / "f32.npy"), 3, **kw)
    assert _accuracy(f32[0], lab, 3) > 0.98


@pytest.mark.parametrize("layout", ["contiguous", "chunked_deflate"])
def test_kmeans_from_v73_mat_file(ctx, tmp_path, layout):
    """DataFile pointing at a MATLAB -v7.3 .mat (the reference's container, private/sampleAndMixFromLargeFile.m:60-66):
    same answer as the in-core call, in both orientations (ColumnSamples)."""
    from sparsifiedkmeans_b200 import kmeans_sparsified, matfile73
    X, lab = _mixture(64, 4000, 3, seed=8)
    kwf = dict(chunks=(512, 48), compress=2) if layout == "chunked_deflate" else {}
    matfile73.write_matrix(str(tmp_path / "rows.mat"), X, name="X", **kwf)                 # n x p, rows are samples
    matfile73.write_matrix(str(tmp_path / "cols.mat"), np.ascontiguousarray(X.T), name="data", **kwf)
    kw = dict(Sparsify=True, SparsityLevel=0.25, Seed=3, Replicates=2, nargout=9, Context=ctx)
    mem = kmeans_sparsified(X, 3, **kw)
    a = kmeans_sparsified(str(tmp_path / "rows.mat"), 3, **kw)
    b = kmeans_sparsified(None, 3, DataFile=str(tmp_path / "cols"), ColumnSamples=True, **kw)   # '.mat' appended (:187-189)
    assert a[4]["LoadFromDisk"] and b[4]["LoadFromDisk"]
    for got in (a, b):
        assert np.array_equal(got[0], mem[0])
        assert np.array_equal(got[6], mem[6])
    np.testing.assert_allclose(a[1], mem[1], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(b[1], mem[1].T, rtol=1e-9, atol=1e-12)
    assert _accuracy(a[0], lab, 3) > 0.98
