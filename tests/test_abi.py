"""The C-ABI library: builds, loads, exports every symbol include/skm_b200.h declares, and fails
loudly without a GPU (no compute calls here).  CPU only."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "skm_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(skm_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_loads():
    from sparsifiedkmeans_b200 import _lib, build
    path = build.build_library()
    assert os.path.exists(path)
    lib = _lib.load()
    assert lib.skm_abi_version() == 1


def test_every_declared_symbol_is_exported_and_bound():
    from sparsifiedkmeans_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in skm_b200.h but not exported: {missing}"
    unbound = [n for n in names if n not in _lib.SIGNATURES]
    assert not unbound, f"declared in skm_b200.h but not bound in _lib.py: {unbound}"
    extra = [n for n in _lib.SIGNATURES if n not in names]
    assert not extra, f"bound in _lib.py but not declared in skm_b200.h: {extra}"


def test_header_cites_the_reference_interfaces():
    src = open(HEADER).read()
    for cite in ("SparseMatrixMinusCluster.c", "SparseMatrixInnerProduct.c", "SparseMatrixColumnNormSq.c",
                 "hadamard.c", "hadamard_pthreads.c", "findClusterAssignments.m", "kmeans_sparsified.m",
                 "Arthur_initialization.m", "randsample_fixedNumberEntries.m"):
        assert cite in src


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from sparsifiedkmeans_b200 import Context
    from sparsifiedkmeans_b200._lib import SkmError
    with pytest.raises(SkmError, match="no CPU fallback"):
        Context(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "sparsifiedkmeans_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "skm_oracle" not in text and "libref_" not in text, f
