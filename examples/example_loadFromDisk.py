#!/usr/bin/env python
"""Python twin of the reference's example_loadFromDisk.m: the data live in a MATLAB -v7.3 .mat file and are streamed
from disk (`DataFile`), never loaded as a whole (example_loadFromDisk.m:26-29,60-72).

    python examples/example_loadFromDisk.py [--p 500 --n 5000]

The file is written here with sparsifiedkmeans_b200.matfile73.write_matrix (chunked + deflate, the way MATLAB's
`save(..., '-v7.3')` stores a large matrix); a file saved by MATLAB itself is read the same way.  p = 500 is not a power of
two, so SketchType 'auto' takes the DCT (kmeans_sparsified.m:226-231), exactly as the reference example does.
"""
import argparse
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--p", type=int, default=500)
    ap.add_argument("--n", type=int, default=5000)
    ap.add_argument("--k", type=int, default=5)
    ap.add_argument("--gamma", type=float, default=0.05)
    ap.add_argument("--replicates", type=int, default=20)
    a = ap.parse_args()
    from sparsifiedkmeans_b200 import kmeans_sparsified, matfile73
    rng = np.random.default_rng(234)
    centers_true = rng.standard_normal((a.p, a.k))
    lab = np.repeat(np.arange(a.k), a.n // a.k)[:a.n]
    X = centers_true[:, lab] + 0.1 * rng.standard_normal((a.p, lab.size))         # p x n: columns are points (:15-20)

    myfile = os.path.join(tempfile.gettempdir(), "sampleInput.mat")               # :28-29
    matfile73.write_matrix(myfile, X, name="X", chunks=(min(512, X.shape[1]), a.p), compress=3)
    print(f"wrote {myfile}: {os.path.getsize(myfile) / 1e6:.1f} MB (-v7.3: HDF5, chunked + deflate)")

    common = dict(ColumnSamples=True, Display="off", Replicates=a.replicates, Sparsify=True, SparsityLevel=a.gamma, Seed=1)
    kmeans_sparsified(X[:, :200], a.k, **{**common, "Replicates": 1})            # warm-up (context, kernels)
    t0 = time.perf_counter()
    idx, C, sumd, D, out = kmeans_sparsified(X, a.k, **common)                    # in core (:50-58)
    t_mem = time.perf_counter() - t0
    print(f"our sparse version:\t\t\tobjective {np.linalg.norm(D):.3e}, time {t_mem:.2e}")

    t0 = time.perf_counter()
    idx2, C2, sumd2, D2, out2 = kmeans_sparsified(None, a.k, DataFile=myfile, MB_limit=1, **common)   # from disk (:65-69)
    t_disk = time.perf_counter() - t0
    print(f"our sparse version (from disk):\tobjective {np.linalg.norm(D2):.3e}, time {t_disk:.2e}")
    assert out2["LoadFromDisk"]
    same = np.array_equal(idx, idx2)
    found = np.asarray(idx2) - 1
    acc = sum(np.bincount(lab[found == c], minlength=a.k).max() for c in range(a.k) if np.any(found == c)) / lab.size
    print(f"assignments identical to the in-core run: {same}; points in the majority cluster of their label: {acc:.3f}")
    os.remove(myfile)


if __name__ == "__main__":
    main()
