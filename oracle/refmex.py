"""ctypes driver for the reference's own MEX files -- TEST INFRASTRUCTURE ONLY.

The five C files under /root/reference/private are compiled *unmodified* into
oracle/_ref/libref_<name>.so by oracle/Makefile against the stub mex.h in
oracle/mexstub/.  This module builds stub ``mxArray`` structs around numpy
arrays and calls each library's ``mexFunction`` gateway, exactly the way MATLAB
would (SparseMatrixMinusCluster.c:44-45, hadamard.c:115-116, ...).

Nothing here reads /root/reference at run time; only the prebuilt .so files are
needed (they travel to the GPU box with the repo snapshot).
"""
from __future__ import annotations

import ctypes as C
import os
from functools import lru_cache

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")

mxCHAR_CLASS = 4
mxDOUBLE_CLASS = 6
mxSINGLE_CLASS = 7


class MxArray(C.Structure):
    _fields_ = [
        ("m", C.c_size_t), ("n", C.c_size_t),
        ("pr", C.c_void_p), ("ir", C.c_void_p), ("jc", C.c_void_p),
        ("is_sparse", C.c_int), ("is_complex", C.c_int),
        ("classid", C.c_int), ("owns", C.c_int),
        ("nzmax", C.c_size_t),
    ]


class MexError(RuntimeError):
    def __init__(self, msg, ident=""):
        super().__init__(msg)
        self.identifier = ident


def ref_available(name: str = "SparseMatrixMinusCluster") -> bool:
    return os.path.exists(os.path.join(REF_DIR, f"libref_{name}.so"))


@lru_cache(maxsize=None)
def _load(path: str):
    lib = C.CDLL(path)
    lib.mexstub_call.restype = C.c_int
    lib.mexstub_call.argtypes = [C.c_int, C.POINTER(C.POINTER(MxArray)), C.c_int,
                                 C.POINTER(C.POINTER(MxArray))]
    lib.mexstub_last_error.restype = C.c_char_p
    lib.mexstub_last_error_id.restype = C.c_char_p
    lib.mxDestroyArray.argtypes = [C.POINTER(MxArray)]
    lib.mxDestroyArray.restype = None
    return lib


def load_ref(name: str):
    path = os.path.join(REF_DIR, f"libref_{name}.so")
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"{path} missing: run `make -C oracle ref` where /root/reference exists")
    return _load(path)


def mx_dense(a: np.ndarray, keep: list, classid: int = mxDOUBLE_CLASS,
             is_complex: bool = False) -> MxArray:
    """Wrap a numpy array (made Fortran-ordered) as a full MATLAB matrix."""
    dt = np.float64 if classid == mxDOUBLE_CLASS else np.float32
    a = np.asarray(a, dtype=dt)
    if a.ndim == 0:
        a = a.reshape(1, 1)
    elif a.ndim == 1:
        a = a.reshape(-1, 1)
    a = np.asfortranarray(a)
    keep.append(a)
    return MxArray(a.shape[0], a.shape[1], a.ctypes.data, None, None, 0,
                   int(is_complex), classid, 0, 0)


def mx_string(text: str, keep: list) -> MxArray:
    """A MATLAB char row vector (the stub stores one byte per character)."""
    b = np.frombuffer(text.encode(), dtype=np.uint8).copy()
    keep.append(b)
    return MxArray(1, b.shape[0], b.ctypes.data, None, None, 0, 0, mxCHAR_CLASS, 0, 0)


def mx_empty() -> MxArray:
    return MxArray(0, 0, None, None, None, 0, 0, mxDOUBLE_CLASS, 0, 0)


def mx_sparse(p: int, n: int, jc: np.ndarray, ir: np.ndarray, x: np.ndarray,
              keep: list) -> MxArray:
    """Wrap CSC arrays as a MATLAB sparse double matrix (jc/ir as mwIndex=uint64)."""
    jc = np.ascontiguousarray(jc, dtype=np.uint64)
    ir = np.ascontiguousarray(ir, dtype=np.uint64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    assert jc.shape[0] == n + 1
    keep.extend([jc, ir, x])
    return MxArray(p, n, x.ctypes.data, ir.ctypes.data, jc.ctypes.data, 1, 0,
                   mxDOUBLE_CLASS, 0, x.shape[0])


def call_mex(lib, prhs: list, nlhs: int):
    """Call lib's mexFunction with the given inputs; returns a list of numpy outputs."""
    nrhs = len(prhs)
    in_arr = (C.POINTER(MxArray) * max(nrhs, 1))()
    for i, a in enumerate(prhs):
        in_arr[i] = C.pointer(a)
    nout = max(nlhs, 1)
    out_arr = (C.POINTER(MxArray) * nout)()
    rc = lib.mexstub_call(nlhs, out_arr, nrhs, in_arr)
    if rc != 0:
        raise MexError(lib.mexstub_last_error().decode(),
                       lib.mexstub_last_error_id().decode())
    outs = []
    for i in range(nout):
        if not out_arr[i]:
            continue
        mx = out_arr[i].contents
        cnt = mx.m * mx.n
        dt = np.float64 if mx.classid == mxDOUBLE_CLASS else (
            np.float32 if mx.classid == mxSINGLE_CLASS else np.int32)
        if cnt:
            buf = np.ctypeslib.as_array(C.cast(mx.pr, C.POINTER(np.ctypeslib.as_ctypes_type(dt))),
                                        shape=(cnt,)).copy()
        else:
            buf = np.zeros(0, dtype=dt)
        outs.append(buf.reshape((mx.m, mx.n), order="F"))
        lib.mxDestroyArray(out_arr[i])
    return outs


# ---------------------------------------------------------------------------
# convenience wrappers, one per reference MEX file
# ---------------------------------------------------------------------------

def SparseMatrixMinusCluster(p, n, jc, ir, x, centers, beta=None):
    """dist = SparseMatrixMinusCluster(X, c[, beta]); K x n (SparseMatrixMinusCluster.c:2-11)."""
    lib = load_ref("SparseMatrixMinusCluster")
    keep = []
    args = [mx_sparse(p, n, jc, ir, x, keep), mx_dense(centers, keep)]
    if beta is not None:
        args.append(mx_dense(np.array([[beta]]), keep))
    return call_mex(lib, args, 1)[0]


def SparseMatrixInnerProduct(p, n, jc, ir, x, c):
    """[innerProd, normX2] = SparseMatrixInnerProduct(X, c) (SparseMatrixInnerProduct.c:2-9)."""
    lib = load_ref("SparseMatrixInnerProduct")
    keep = []
    outs = call_mex(lib, [mx_sparse(p, n, jc, ir, x, keep), mx_dense(c, keep)], 2)
    return outs[0], outs[1]


def SparseMatrixColumnNormSq(p, n, jc, ir, x):
    """normX2 = SparseMatrixColumnNormSq(X) (SparseMatrixColumnNormSq.c:2-8)."""
    lib = load_ref("SparseMatrixColumnNormSq")
    keep = []
    return call_mex(lib, [mx_sparse(p, n, jc, ir, x, keep)], 1)[0]


def hadamard(x):
    """w = hadamard(x), serial (hadamard.c:115-152)."""
    lib = load_ref("hadamard")
    keep = []
    return call_mex(lib, [mx_dense(x, keep)], 1)[0]


def pthreads_variants():
    out = []
    if os.path.isdir(REF_DIR):
        for f in os.listdir(REF_DIR):
            if f.startswith("libref_hadamard_pthreads_") and f.endswith(".so"):
                out.append(int(f[len("libref_hadamard_pthreads_"):-3]))
    return sorted(out)


def hadamard_pthreads(x, nthreads=4):
    """w = hadamard_pthreads(x) built with -DNTHREADS=nthreads (hadamard_pthreads.c:227-264)."""
    lib = load_ref(f"hadamard_pthreads_{nthreads}")
    keep = []
    return call_mex(lib, [mx_dense(x, keep)], 1)[0]
