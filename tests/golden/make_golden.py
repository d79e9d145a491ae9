#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE's own C files (compiled unmodified
into oracle/_ref by oracle/Makefile) on seeded inputs.  Run in the build container, where
/root/reference exists:

    make -C oracle ref && python tests/golden/make_golden.py

The reference repository ships no test vectors; these fixtures pin the
oracle (tests/test_oracle.py) and travel to the GPU box, where /root/reference does not exist.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refmex          # noqa: E402
from tests.util import make_sparsified   # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    assert refmex.ref_available(), "build oracle/_ref first (make -C oracle ref)"
    # masked distances: K = 1,2,3 (unrolled paths) and general K; ragged + empty columns
    for K in (1, 2, 3, 4, 10, 64):
        X, c, gamma = make_sparsified(p=96, n=200, m=9, K=K, seed=100 + K, kind="unstructured", f32=False, ragged=True)
        cs = c / gamma
        D = refmex.SparseMatrixMinusCluster(96, 200, X.indptr, X.indices, X.data, cs)
        np.savez_compressed(os.path.join(HERE, f"smmc_K{K}.npz"), p=96, n=200, jc=X.indptr.astype(np.int64),
                            ir=X.indices.astype(np.int64), x=X.data, centers=cs, dist=D)
    X, c, _ = make_sparsified(p=64, n=150, m=7, K=1, seed=7, kind="unstructured", f32=False)
    D = refmex.SparseMatrixMinusCluster(64, 150, X.indptr, X.indices, X.data, c, beta=0.37)
    ip, n2 = refmex.SparseMatrixInnerProduct(64, 150, X.indptr, X.indices, X.data, c[:, 0])
    n2b = refmex.SparseMatrixColumnNormSq(64, 150, X.indptr, X.indices, X.data)
    np.savez_compressed(os.path.join(HERE, "beta_inner_norm.npz"), p=64, n=150, jc=X.indptr.astype(np.int64),
                        ir=X.indices.astype(np.int64), x=X.data, c=c[:, 0], beta=0.37, dist_beta=D,
                        inner=ip, normsq=n2, normsq_only=n2b)
    rng = np.random.default_rng(5)
    for m in (2, 16, 512):
        x = rng.standard_normal((m, 6))
        np.savez_compressed(os.path.join(HERE, f"hadamard_m{m}.npz"), x=x, w=refmex.hadamard(x),
                            w_pthreads=refmex.hadamard_pthreads(x, 4))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
