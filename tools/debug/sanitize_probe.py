#!/usr/bin/env python
"""Small end-to-end exercise of every kernel family, meant to run under compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/debug/sanitize_probe.py
    compute-sanitizer --tool racecheck python tools/debug/sanitize_probe.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    from sparsifiedkmeans_b200 import Dataset, Lloyd, default_context, dct_mix, lloyd_step_host, second_pass
    from tests.util import make_sparsified
    ctx = default_context(0)
    for (p, n, m, K) in ((96, 700, 12, 10), (128, 900, 9, 16), (256, 600, 13, 40), (64, 500, 8, 3)):
        X, c, gamma = make_sparsified(p=p, n=n, m=m, K=K, seed=p, kind="unstructured", ragged=(K == 16))
        ds = Dataset.from_scipy(X, store="f32", ctx=ctx)
        for layout in (0, 1, 2):
            r = ds.layout_check(layout)
            assert r["bad_columns"] == 0
        L = Lloyd(ds, K, incremental=True, bounded=True)
        L.set_centers(c)
        for _ in range(6):
            L.step(gamma, gamma, True)
        L.assign_sparse(gamma)
        L.close()
        ds.kpp_update(ds.get_column(0), gamma, first=True)
        ds.kpp_pick(0.5)
        lloyd_step_host(p, n, X.indptr, X.indices, X.data, c, gamma, gamma, True, ctx=ctx)
        ds.close()
        ds64 = Dataset.from_scipy(X, store="f64", ctx=ctx)
        ds64.assign(c, gamma)
        ds64.close()
    rng = np.random.default_rng(0)
    for (p, n, m) in ((100, 300, 13), (2048, 40, 100)):
        p2 = 1 << (p - 1).bit_length()
        Xd = rng.standard_normal((p, n))
        Dataset.from_dense_host(Xd, np.sign(rng.standard_normal(p2)), m, seed=3, ctx=ctx).close()
    Xd = rng.standard_normal((50, 400))
    Dataset.from_dense_host_dct(Xd, np.sign(rng.standard_normal(50)), 9, seed=4, ctx=ctx).close()
    dct_mix(Xd, None, False, ctx)
    cen = rng.standard_normal((50, 7))
    second_pass(Xd, centers=cen, assign_in=rng.integers(1, 8, 400), ctx=ctx)
    print("sanitize probe done")


if __name__ == "__main__":
    main()
