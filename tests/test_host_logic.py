"""Host-side logic of the kmeans_sparsified mirror that needs no GPU.  CPU only."""
import numpy as np
import pytest

from oracle import host_ref
from sparsifiedkmeans_b200 import kmeans as km
from sparsifiedkmeans_b200.distributed import shard_bounds


def test_matlab_round_is_half_away_from_zero():
    assert [km.matlab_round(v) for v in (0.5, 1.5, 2.5, -0.5, -2.5, 2.4999)] == [1, 2, 3, -1, -3, 2]
    assert km.matlab_round(0.05 * 1024) == 51 and km.matlab_round(0.1 * 784) == 78
    assert km.matlab_round(0.05 * 50) == 3                                            # Python's round() gives 2


def test_randsample_block_contract():
    rng = np.random.default_rng(0)
    rows = km.randsample_block(rng, 64, 5, 2000)        # randsample_block.m: k distinct indices per column
    assert rows.shape == (5, 2000) and rows.min() >= 0 and rows.max() < 64
    assert np.all(np.diff(rows, axis=0) > 0)            # distinct and sorted
    freq = np.bincount(rows.ravel(), minlength=64) / (5 * 2000)
    assert abs(freq.max() - 1 / 64) < 0.006 and abs(freq.min() - 1 / 64) < 0.006       # uniform


def test_randsample_fixed_entries_matches_oracle():
    rng = np.random.default_rng(1)
    Xm = rng.standard_normal((32, 50))
    Xm[3, 7] = 0.0                                       # exact zero must be dropped by sparse()
    rows = km.randsample_block(rng, 32, 4, 50)
    rows[0, 7] = 3 if 3 not in rows[1:, 7] else rows[0, 7]
    Y = km.randsample_fixedNumberEntries(Xm, 4, rows)
    W = host_ref.sample_fixed_entries(Xm, rows)
    assert (Y != W).nnz == 0
    assert np.all(np.diff(Y.indptr) <= 4)
    j = 0
    np.testing.assert_array_equal(Y[:, j].data, Xm[rows[:, j], j] / (4 / 32))          # randsample_fixedNumberEntries.m:30-31


def test_option_validation_mirrors_input_parser():
    X = np.zeros((10, 4))
    with pytest.raises(km.KMeansError):
        km.kmeans_sparsified(X, 2, Sparsify=True, NoSuchOption=1)
    with pytest.raises(km.KMeansError):
        km.kmeans_sparsified(X, 2, Sparsify=True, SparsityLevel=0.0)
    with pytest.raises(km.KMeansError):
        km.kmeans_sparsified(X, 2, Sparsify=True, EmptyAction="explode")
    with pytest.raises(NotImplementedError):
        km.kmeans_sparsified(X, 2)                       # dense path is out of scope
    with pytest.raises(km.KMeansError, match="Cannot find specified data file"):
        km.kmeans_sparsified("no_such_data_file", 2, Sparsify=True)                   # kmeans_sparsified.m:187-192
    with pytest.raises(km.KMeansError, match="No reason"):
        km.kmeans_sparsified("no_such_data_file", 2)                                  # :197-199
    assert km.kmeans_sparsified() == 2.1


def test_shard_bounds_cover_all_columns():
    for n in (0, 1, 7, 1000, 10**8):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(n, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_oracle_dense_branch_and_second_pass_restatement():
    """The oracle's restatement of the dense branch (findClusterAssignments.m:154-175, expand-quadratic arm) and of the
    in-core two-pass block (kmeans_sparsified.m:542-560) against a direct evaluation of their definitions."""
    from oracle import host_ref
    rng = np.random.default_rng(4)
    p, n, K = 30, 400, 5
    X = rng.standard_normal((p, n))
    c = rng.standard_normal((p, K))
    a, d, D2 = host_ref.find_cluster_assignments_dense(X, c)
    direct = np.stack([np.sqrt(((X - c[:, [k]]) ** 2).sum(axis=0)) for k in range(K)])
    assert np.array_equal(a, direct.argmin(axis=0) + 1)
    np.testing.assert_allclose(d, direct.min(axis=0), rtol=1e-10)
    np.testing.assert_allclose(D2, direct ** 2, rtol=1e-9, atol=1e-9)
    c[:, 3] = c[:, 1]                                      # a duplicated centre never wins over its twin (first occurrence)
    a2, _, _ = host_ref.find_cluster_assignments_dense(X, c)
    assert not np.any(a2 == 4)
    lab = rng.integers(1, K + 1, n)
    lab[lab == 2] = 1                                      # an empty cluster keeps its zero column (:545)
    c2, a3, d3 = host_ref.second_pass(X, c, lab, K)
    assert np.all(c2[:, 1] == 0)
    for k in (0, 2, 3, 4):
        np.testing.assert_allclose(c2[:, k], X[:, lab == k + 1].mean(axis=1), rtol=1e-12)
    assert np.array_equal(a3, a2)


def test_data_file_container_checks(tmp_path):
    """DataFile mode (kmeans_sparsified.m:180-207): what the file opener accepts, before any GPU work."""
    f64 = tmp_path / "ok.npy"
    np.save(f64, np.zeros((6, 4)))
    A = km._open_data_file(str(f64))
    assert isinstance(A, np.memmap) and A.shape == (6, 4)
    assert km._open_data_file(str(tmp_path / "ok")).shape == (6, 4)           # extension optional (:187-189)
    np.save(tmp_path / "vec.npy", np.zeros(5))
    with pytest.raises(km.KMeansError, match="bad size"):
        km._open_data_file(str(tmp_path / "vec.npy"))
    np.save(tmp_path / "ints.npy", np.zeros((3, 3), dtype=np.int32))
    with pytest.raises(km.KMeansError, match="float32 or float64"):
        km._open_data_file(str(tmp_path / "ints.npy"))
    with pytest.raises(km.KMeansError, match="Cannot find"):
        km._open_data_file(str(tmp_path / "missing"))
    (tmp_path / "data.mat").write_bytes(b"MATLAB 5.0 MAT-file, Platform: GLNXA64" + b" " * 200)
    with pytest.raises(km.KMeansError, match="not -v7.3"):
        km._open_data_file(str(tmp_path / "data.mat"))
    from sparsifiedkmeans_b200 import matfile73
    X = np.random.default_rng(0).standard_normal((6, 40))
    matfile73.write_matrix(str(tmp_path / "v73.mat"), X, chunks=(16, 6), compress=3)
    A = km._open_data_file(str(tmp_path / "v73"))                             # '.mat' appended (:187-189)
    assert A.shape == (6, 40) and np.array_equal(np.asarray(A), X)
    assert A.T.flags["C_CONTIGUOUS"], "columns are contiguous: the chunked upload streams them without a copy"


def test_driver_accepts_the_iteration_mode_options():
    X = np.zeros((10, 4))
    for opt in ("IncrementalUpdate", "BoundedAssign"):
        with pytest.raises(NotImplementedError):           # recognised option; stops at Sparsify=false as before
            km.kmeans_sparsified(X, 2, **{opt: False})
