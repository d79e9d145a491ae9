import numpy as np, sys
sys.path.insert(0, "/root/repo")
from sparsifiedkmeans_b200 import Dataset, default_context
ctx = default_context(0)
rng = np.random.default_rng(0)
for (p, n, m) in ((100, 1500, 13), (512, 3000, 26), (4096, 300, 205)):
    p2 = 1 << (p - 1).bit_length()
    X = rng.standard_normal((p, n))
    d = np.sign(rng.standard_normal(p2))
    outs = []
    for rep in range(3):
        ds = Dataset.from_dense_host(X, d, m, seed=123, ctx=ctx)
        S = ds.to_scipy(); S.sort_indices()
        outs.append((S.indptr.copy(), S.indices.copy(), S.data.copy()))
        ds.close()
    same = all(np.array_equal(outs[0][i], o[i]) for o in outs[1:] for i in range(3))
    print(p, n, m, "deterministic:", same, "nnz", outs[0][1].shape, [o[1].shape for o in outs])
    if not same:
        a, b = outs[0], outs[1]
        if a[1].shape == b[1].shape:
            print("  rows differ at", int(np.count_nonzero(a[1] != b[1])), "vals differ at", int(np.count_nonzero(a[2] != b[2])))
