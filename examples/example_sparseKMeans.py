#!/usr/bin/env python
"""Python twin of the reference's example_sparseKMeans.m (BASELINE.json configs[0]): a 5-component
Gaussian mixture (sigma = 0.1, centres ~ N(0, I)), clustered with sparsified K-means at
gamma = 0.05, 20 replicates (example_sparseKMeans.m:12-22,60-68).

    python examples/example_sparseKMeans.py [--p 512 --n 5000] [--cpu]

Prints accuracy against the planted labels and wall-clock time on the GPU; with --cpu it also times
the CPU oracle (the reference's compiled C kernel + the numpy restatement of its MATLAB loop) on the
same sparsified data, for orientation (figs/example.png quotes 0.79 s for the MATLAB original on
unstated 2015 hardware).
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--p", type=int, default=512)
    ap.add_argument("--n", type=int, default=5000)
    ap.add_argument("--k", type=int, default=5)
    ap.add_argument("--gamma", type=float, default=0.05)
    ap.add_argument("--replicates", type=int, default=20)
    ap.add_argument("--cpu", action="store_true")
    a = ap.parse_args()
    from sparsifiedkmeans_b200 import kmeans_sparsified
    rng = np.random.default_rng(0)
    mu = rng.standard_normal((a.k, a.p))
    lab = rng.integers(a.k, size=a.n)
    X = mu[lab] + 0.1 * rng.standard_normal((a.n, a.p))                   # rows are points, as in the reference
    kmeans_sparsified(X[:200], a.k, Sparsify=True, SparsityLevel=a.gamma, Seed=1)     # warm-up (context, kernels)
    t0 = time.perf_counter()
    IDX, C, SUMD, D, OUT = kmeans_sparsified(X, a.k, Sparsify=True, SparsityLevel=a.gamma, Replicates=a.replicates, Seed=1)
    t = time.perf_counter() - t0
    # accuracy up to relabelling (majority vote per cluster)
    acc = sum(np.bincount(lab[IDX == c + 1], minlength=a.k).max() for c in range(a.k) if np.any(IDX == c + 1)) / a.n
    print(f"GPU  kmeans_sparsified: {t * 1e3:8.1f} ms  accuracy {acc:.4f}  iterations/replicate {OUT['iterations'].mean():.1f}  "
          f"pipeline={OUT['Pipeline']} sketch {OUT['TimeToSketch'] * 1e3:.1f} ms  "
          f"replicates {OUT['replicateTimes'].sum() * 1e3:.1f} ms (init {OUT['TimeInitialization'] * 1e3:.1f} ms)")
    if a.cpu:
        from oracle import host_ref
        from sparsifiedkmeans_b200.kmeans import matlab_round, randsample_block
        p2 = host_ref.nextpow2_size(a.p)
        d = np.sign(rng.standard_normal(p2)); d[d == 0] = 1
        m = max(1, matlab_round(a.gamma * p2))
        t0 = time.perf_counter()
        Xm = host_ref.mix_hadamard(X.T * (1 + 2 * np.finfo(float).eps), d)
        Xs = host_ref.sample_fixed_entries(Xm, randsample_block(rng, p2, m, a.n))
        best = None
        for _ in range(a.replicates):
            first = int(rng.integers(a.n))
            idx, cen = host_ref.arthur_initialization(Xs, a.k, m / a.p, first, iter(rng.random(4000)))
            res = host_ref.lloyd(Xs, np.asarray(cen.todense()), m / a.p, centers_sparse=True)
            if best is None or res.objective < best.objective:
                best = res
        tc = time.perf_counter() - t0
        accc = sum(np.bincount(lab[best.assignments == c + 1], minlength=a.k).max() for c in range(a.k)
                   if np.any(best.assignments == c + 1)) / a.n
        print(f"CPU  oracle (reference C kernels + numpy host loop): {tc * 1e3:8.1f} ms  accuracy {accc:.4f}")


if __name__ == "__main__":
    main()
