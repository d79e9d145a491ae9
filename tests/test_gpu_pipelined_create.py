"""Pipelined dataset creation (skm_dataset_create_csc_hint): the chunked upload with per-chunk conversion, validation,
row-major counting and entry-order build gives the same images and results as the one-shot path."""
import os

import numpy as np
import pytest

from oracle import host_ref
from tests.util import make_sparsified

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows_dtype,val_dtype", [(np.int32, np.float32), (np.uint16, np.float32), (np.int64, np.float64)])
@pytest.mark.parametrize("K", [10, 16, 0])
def test_pipelined_equals_one_shot(ctx, rows_dtype, val_dtype, K):
    from sparsifiedkmeans_b200 import Dataset, Lloyd
    p, m, n = 784, 78, 90000                     # 4.9e6 entries: above the pipelining threshold, ragged chunks
    X, c, gamma = make_sparsified(p=p, n=n, m=m, K=max(K, 10), seed=3, kind="mixture", f32=True, ragged=True)
    jc = X.indptr.astype(np.int64)
    ir = X.indices.astype(rows_dtype)
    val = X.data.astype(val_dtype)
    Kc = max(K, 10)
    os.environ["SKM_PIPELINE_CHUNK"] = "700000"                  # seven chunks instead of one
    try:
        ds = Dataset.from_csc(p, n, jc, ir, val, store="f32", ctx=ctx, K_hint=K)
    finally:
        del os.environ["SKM_PIPELINE_CHUNK"]
    for layout in ([1] if K == 10 else [2] if K == 16 else [0]):
        r = ds.layout_check(layout)
        assert r["bad_columns"] == 0, r
    L = Lloyd(ds, Kc)
    L.set_centers(c[:, :Kc])
    st = L.step(gamma, gamma)
    a, d = L.assignments()
    wa, wd, _ = host_ref.find_cluster_assignments(X, c[:, :Kc], gamma)
    assert np.array_equal(a, wa)
    np.testing.assert_allclose(d, wd, rtol=2e-5, atol=1e-30)
    got = L.get_centers()
    os.environ["SKM_NO_PIPELINE"] = "1"
    try:
        ds0 = Dataset.from_csc(p, n, jc, ir, val, store="f32", ctx=ctx)
    finally:
        del os.environ["SKM_NO_PIPELINE"]
    L0 = Lloyd(ds0, Kc)
    L0.set_centers(c[:, :Kc])
    L0.step(gamma, gamma)
    np.testing.assert_array_equal(got, L0.get_centers())
    Xb = ds.to_scipy()
    assert (abs(Xb - X.astype(np.float32).astype(np.float64)) > 0).nnz == 0
    L.close(); L0.close(); ds.close(); ds0.close()


def test_pipelined_rejects_bad_rows(ctx):
    from sparsifiedkmeans_b200 import Dataset
    p, m, n = 256, 64, 80000
    X, c, gamma = make_sparsified(p=p, n=n, m=m, K=5, seed=1, kind="mixture", f32=True)
    ir = X.indices.astype(np.int32).copy()
    ir[-7] = p + 3
    with pytest.raises(Exception, match="row index"):
        Dataset.from_csc(p, n, X.indptr.astype(np.int64), ir, X.data.astype(np.float32), store="f32", ctx=ctx, K_hint=5)
    jc = X.indptr.astype(np.int64).copy()
    jc[1000] = jc[1001] + 5
    with pytest.raises(Exception, match="column pointers"):
        Dataset.from_csc(p, n, jc, X.indices.astype(np.int32), X.data.astype(np.float32), store="f32", ctx=ctx)
