"""MATLAB -v7.3 (HDF5) reader of the DataFile mode (sparsifiedkmeans_b200/matfile73.py), CPU only.

Pinned on a file MATLAB itself wrote (scipy's fixture testhdf5_7.4_GLNX86.mat, compared with scipy.io.loadmat of the
same variable saved as -v7); chunked / deflate / shuffle storage and deep chunk B-trees through the module's own writer."""
import os

import numpy as np
import pytest

from sparsifiedkmeans_b200 import matfile73 as m


def _scipy_fixture(name):
    import scipy.io
    return os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data", name)


def test_reads_a_file_written_by_matlab():
    import scipy.io
    path = _scipy_fixture("testhdf5_7.4_GLNX86.mat")
    if not os.path.isfile(path):
        pytest.skip("scipy's -v7.3 fixture is not installed")
    A, name = m.open_matrix(path)
    want = scipy.io.loadmat(_scipy_fixture("testdouble_7.4_GLNX86.mat"))["testdouble"]
    assert name == "testdouble" and A.dtype == np.float64
    assert A.shape == want.shape == (1, 9)                       # HDF5 holds (9, 1): MATLAB's dimensions reversed
    assert np.array_equal(A, want)
    r = m.H5Reader(path)
    (n, addr), = r.variables()
    obj = r.object(addr)
    assert obj.attrs["MATLAB_class"] == "double" and obj.shape == (9, 1) and obj.layout[0] == "contiguous"
    r.close()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kw", [dict(), dict(chunks=(7, 5)), dict(chunks=(16, 33), compress=4),
                                dict(chunks=(3, 4), compress=1, shuffle=True), dict(compress=6), dict(chunks=(64, 64))])
def test_round_trip_layouts(tmp_path, dtype, kw):
    X = np.random.default_rng(1).standard_normal((33, 50)).astype(dtype)
    path = str(tmp_path / "a.mat")
    m.write_matrix(path, X, name="Xvar", **kw)
    with open(path, "rb") as f:
        head = f.read(520)
    assert head.startswith(b"MATLAB 7.3 MAT-file") and head[512:520] == m.SIGNATURE
    Y, name = m.open_matrix(path)
    assert name == "Xvar" and Y.dtype == dtype and Y.shape == X.shape
    assert np.array_equal(np.asarray(Y), X)
    assert Y.T.flags["C_CONTIGUOUS"]                              # a column of X is contiguous, as in MATLAB
    if not kw:
        assert isinstance(Y.base, np.memmap) or isinstance(Y, np.memmap), "contiguous data is mapped in place"


def test_deep_chunk_tree_and_out_of_core_decode(tmp_path):
    """> 64 and > 4096 chunks: two- and three-level version-1 B-trees; a small in-memory limit sends the decoded matrix
    to a memory-mapped temporary file."""
    rng = np.random.default_rng(2)
    X = rng.standard_normal((3, 9000))
    path = str(tmp_path / "deep.mat")
    m.write_matrix(path, X, chunks=(2, 3), compress=1)           # 4500 chunks
    r = m.H5Reader(path)
    (_, addr), = r.variables()
    Y = r.read_matrix(addr, tmpdir=str(tmp_path), in_memory_limit=1024)
    assert np.array_equal(np.asarray(Y), X)
    assert isinstance(Y.base, np.memmap) or isinstance(Y, np.memmap)
    assert sum(1 for _ in r._chunks(r.object(addr).layout[1], 3)) == 4500
    r.close()


def test_errors(tmp_path):
    p5 = tmp_path / "v5.mat"
    p5.write_bytes(b"MATLAB 5.0 MAT-file, Platform: GLNXA64, Created on: x" + b" " * 100)
    with pytest.raises(m.MatFileError, match="not -v7.3"):
        m.open_matrix(str(p5))
    junk = tmp_path / "junk.mat"
    junk.write_bytes(os.urandom(4096))
    with pytest.raises(m.MatFileError, match="no HDF5 signature"):
        m.open_matrix(str(junk))
    with pytest.raises(m.MatFileError, match="2-D"):
        m.write_matrix(str(tmp_path / "v.mat"), np.zeros(5))
    # a truncated file: addresses beyond the end are reported, not read
    good = tmp_path / "good.mat"
    m.write_matrix(str(good), np.arange(600.0).reshape(20, 30))
    data = good.read_bytes()
    (tmp_path / "cut.mat").write_bytes(data[:len(data) - 3000])
    with pytest.raises((m.MatFileError, ValueError)):
        Y, _ = m.open_matrix(str(tmp_path / "cut.mat"))
        np.asarray(Y).sum()


def test_integer_classes_and_big_files_are_mapped_not_read(tmp_path):
    X = (np.arange(48).reshape(6, 8) % 7).astype(np.int16)
    m.write_matrix(str(tmp_path / "i16.mat"), X)
    Y, _ = m.open_matrix(str(tmp_path / "i16.mat"))
    assert Y.dtype == np.int16 and np.array_equal(Y, X)
    r = m.H5Reader(str(tmp_path / "i16.mat"))
    assert r.object(r.variables()[0][1]).attrs["MATLAB_class"] == "int16"
    r.close()


@pytest.mark.parametrize("message_versions", ["new", "v1layout"])
@pytest.mark.parametrize("kw", [dict(), dict(chunks=(7, 5)), dict(chunks=(16, 33), compress=4, shuffle=True)])
def test_other_message_versions(tmp_path, message_versions, kw):
    """The later versions of the header messages (dataspace 2, filter pipeline 2, attribute 3) and the HDF5 1.4 layout
    message (version 1): files from other HDF5 writers that keep version-1 object headers and B-trees."""
    X = np.random.default_rng(5).standard_normal((33, 50))
    path = str(tmp_path / "a.mat")
    m.write_matrix(path, X, message_versions=message_versions, **kw)
    Y, name = m.open_matrix(path)
    assert name == "X" and np.array_equal(np.asarray(Y), X)
    r = m.H5Reader(path)
    obj = r.object(r.variables()[0][1])
    assert obj.attrs["MATLAB_class"] == "double" and obj.shape == (50, 33)
    assert [f[0] for f in obj.filters] == ([2, 1] if kw.get("shuffle") else [])
    r.close()


def test_object_header_continuation(tmp_path):
    """The MATLAB_class attribute in a continuation block of the object header (message 0x0010)."""
    X = np.arange(12.0).reshape(3, 4)
    path = str(tmp_path / "c.mat")
    m.write_matrix(path, X, continuation=True, chunks=(2, 2), compress=1)
    Y, _ = m.open_matrix(path)
    assert np.array_equal(np.asarray(Y), X)
    r = m.H5Reader(path)
    msgs = r._messages(r.variables()[0][1])
    assert [t for t, _ in msgs].count(0x10) == 1 and msgs[-1][0] == 0x0C
    assert r.object(r.variables()[0][1]).attrs["MATLAB_class"] == "double"
    r.close()
