"""The MEX shims in mex/ (the reference-facing gateway code): compiled against the stub mex.h
and driven through ctypes exactly like the reference's own MEX files in oracle/_ref.
CPU part: every argument error the reference raises is raised by the shim with the same
identifier/message, before any CUDA call.  GPU part: same outputs, bit for bit."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import refmex
from tests.util import make_sparsified

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MEXDIR = os.path.join(ROOT, "mex")


@pytest.fixture(scope="module")
def shims():
    from sparsifiedkmeans_b200 import build
    build.build_library()
    subprocess.check_call(["make", "-C", MEXDIR], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return {n: refmex._load(os.path.join(MEXDIR, "_build", f"libmex_{n}.so"))
            for n in ("SparseMatrixMinusCluster", "SparseMatrixInnerProduct", "SparseMatrixColumnNormSq",
                      "hadamard", "skm_lloyd_mex", "skm_second_pass_mex")}


def _bad_calls():
    X, c, _ = make_sparsified(p=20, n=10, m=3, K=2, seed=0, f32=False)
    keep = []
    Xs = refmex.mx_sparse(20, 10, X.indptr, X.indices, X.data, keep)
    cd = refmex.mx_dense(c, keep)
    yield "SparseMatrixMinusCluster", [Xs], 1, keep                                   # one input
    yield "SparseMatrixMinusCluster", [Xs, cd], 2, keep                               # two outputs
    yield "SparseMatrixMinusCluster", [refmex.mx_dense(np.zeros((20, 10)), keep), cd], 1, keep   # X not sparse
    yield "SparseMatrixMinusCluster", [Xs, refmex.mx_dense(c[:-1], keep)], 1, keep    # wrong rows
    yield "SparseMatrixMinusCluster", [Xs, cd, refmex.mx_dense(np.array([[0.5]]), keep)], 1, keep  # beta with K=2
    yield "SparseMatrixInnerProduct", [Xs], 1, keep
    yield "SparseMatrixInnerProduct", [refmex.mx_dense(np.zeros((20, 10)), keep), cd], 1, keep
    yield "SparseMatrixColumnNormSq", [refmex.mx_dense(np.zeros((20, 10)), keep)], 1, keep
    yield "hadamard", [], 1, keep
    yield "hadamard", [refmex.mx_dense(np.zeros((12, 3)), keep)], 1, keep             # not a power of two
    yield "hadamard", [refmex.mx_dense(np.zeros((1, 3)), keep)], 1, keep              # length 1
    yield "hadamard", [refmex.mx_dense(np.zeros((8, 3)), keep, is_complex=True)], 1, keep
    yield "hadamard", [refmex.mx_dense(np.zeros((8, 3)), keep, classid=refmex.mxSINGLE_CLASS)], 1, keep


@pytest.mark.parametrize("idx", range(13))
def test_shim_argument_errors_match_reference(shims, idx):
    name, args, nlhs, keep = list(_bad_calls())[idx]
    with pytest.raises(refmex.MexError) as ours:
        refmex.call_mex(shims[name], args, nlhs)
    if refmex.ref_available(name):
        with pytest.raises(refmex.MexError) as ref:
            refmex.call_mex(refmex.load_ref(name), args, nlhs)
        assert ours.value.identifier == ref.value.identifier
        assert str(ours.value) == str(ref.value)


def test_lloyd_gateway_rejects_bad_commands(shims):
    keep = []
    with pytest.raises(refmex.MexError):
        refmex.call_mex(shims["skm_lloyd_mex"], [refmex.mx_dense(np.zeros((1, 1)), keep)], 1)   # not a string


@pytest.mark.gpu
def test_shims_match_reference_outputs(shims):
    X, c, gamma = make_sparsified(p=96, n=300, m=9, K=7, seed=41, kind="unstructured", f32=False, ragged=True)
    p, n = X.shape
    keep = []
    Xs = refmex.mx_sparse(p, n, X.indptr, X.indices, X.data, keep)
    got = refmex.call_mex(shims["SparseMatrixMinusCluster"], [Xs, refmex.mx_dense(c, keep)], 1)[0]
    from oracle import cport
    want = refmex.SparseMatrixMinusCluster(p, n, X.indptr, X.indices, X.data, c) if refmex.ref_available() \
        else cport.masked_dist(p, n, X.indptr, X.indices, X.data, c)
    assert np.array_equal(got, want)
    ip, n2 = refmex.call_mex(shims["SparseMatrixInnerProduct"], [Xs, refmex.mx_dense(c[:, 0], keep)], 2)
    wip, wn2 = cport.inner_product(n, X.indptr, X.indices, X.data, c[:, 0])
    assert np.array_equal(ip.ravel(), wip) and np.array_equal(n2.ravel(), wn2)
    assert np.array_equal(refmex.call_mex(shims["SparseMatrixColumnNormSq"], [Xs], 1)[0].ravel(), wn2)
    x = np.random.default_rng(0).standard_normal((256, 5))
    assert np.array_equal(refmex.call_mex(shims["hadamard"], [refmex.mx_dense(x, keep)], 1)[0], cport.hadamard(x))


@pytest.mark.gpu
def test_lloyd_gateway_iterates_like_the_reference_loop(shims):
    from oracle import host_ref
    X, c, gamma = make_sparsified(p=64, n=2000, m=8, K=5, seed=43, kind="mixture")
    p, n = X.shape
    lib = shims["skm_lloyd_mex"]
    keep = []
    h = refmex.call_mex(lib, [refmex.mx_string("upload", keep), refmex.mx_sparse(p, n, X.indptr, X.indices, X.data, keep),
                              refmex.mx_dense(np.array([[5.0]]), keep)], 1)[0]
    hm = refmex.mx_dense(h, keep)
    g = refmex.mx_dense(np.array([[gamma]]), keep)
    one = refmex.mx_dense(np.array([[1.0]]), keep)
    outs = refmex.call_mex(lib, [refmex.mx_string("iterate", keep), hm, refmex.mx_dense(c, keep), g, g, one], 6)
    a, d, cen, dff, sumsq, counts = outs
    wa, wd, _ = host_ref.find_cluster_assignments(X, c, gamma)
    assert np.array_equal(a.ravel().astype(np.int64), wa)
    want, _, _, wc = __import__("oracle").cport.centroid_update(p, n, 5, X.indptr, X.indices, X.data, wa, gamma, c, True)
    np.testing.assert_allclose(cen, want, rtol=1e-6, atol=1e-9)
    assert np.array_equal(counts.ravel().astype(np.int64), wc)
    np.testing.assert_allclose(dff[0, 0], np.linalg.norm(c - want), rtol=1e-9)
    a2, d2 = refmex.call_mex(lib, [refmex.mx_string("assign", keep), hm, refmex.mx_dense(c, keep), refmex.mx_empty()], 2)
    wa2, wd2, _ = host_ref.find_cluster_assignments(X, c, None)
    assert np.array_equal(a2.ravel().astype(np.int64), wa2)
    refmex.call_mex(lib, [refmex.mx_string("free", keep), hm], 0)


def test_second_pass_gateway_argument_errors(shims):
    keep = []
    X = refmex.mx_dense(np.zeros((6, 4)), keep)
    c = refmex.mx_dense(np.zeros((6, 2)), keep)
    a = refmex.mx_dense(np.ones((1, 4)), keep)
    lib = shims["skm_second_pass_mex"]
    for args, nlhs in (([X, c], 1), ([X, c, a], 4), ([X, refmex.mx_dense(np.zeros((5, 2)), keep), a], 1),
                       ([X, c, refmex.mx_dense(np.ones((1, 3)), keep)], 1)):
        with pytest.raises(refmex.MexError):
            refmex.call_mex(lib, args, nlhs)


@pytest.mark.gpu
def test_second_pass_gateway_matches_two_pass_block(shims):
    """[centers2, a2, d2] = skm_second_pass_mex(XFull, bestCenters, bestAssignments) against the oracle's
    restatement of kmeans_sparsified.m:542-560."""
    from oracle import host_ref
    rng = np.random.default_rng(2)
    p, n, K = 40, 900, 4
    mu = rng.standard_normal((p, K))
    lab = rng.integers(0, K, n)
    X = mu[:, lab] + 0.2 * rng.standard_normal((p, n))
    cen = mu + 0.05 * rng.standard_normal((p, K))
    keep = []
    c2, a2, d2 = refmex.call_mex(shims["skm_second_pass_mex"],
                                 [refmex.mx_dense(X, keep), refmex.mx_dense(cen, keep),
                                  refmex.mx_dense((lab + 1.0).reshape(1, -1), keep)], 3)
    wc, wa, wd = host_ref.second_pass(X, cen, lab + 1, K)
    np.testing.assert_allclose(c2, wc, rtol=1e-6, atol=1e-9)
    assert np.array_equal(a2.ravel().astype(np.int64), wa)
    np.testing.assert_allclose(d2.ravel(), wd, rtol=2e-5)
