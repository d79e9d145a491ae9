"""Where the time of a dataset upload goes (SKM_TRACE=1 prints the stages)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from sparsifiedkmeans_b200 import Context, Dataset, Lloyd
ctx = Context(0); dev = torch.device("cuda:0")
n, p, K, m = 10_000_000, 784, 10, 78
colptr, rowidx, val, mu, start = bench.gen_shard_device(dev, n, p, m, K, 0, ctx=ctx)
hj = torch.empty(n + 1, dtype=torch.int64, pin_memory=True); hj.copy_(colptr)
hr = torch.empty(n * m, dtype=torch.uint16, pin_memory=True); hr.copy_(rowidx.to(torch.uint16))
hv = torch.empty(n * m, dtype=torch.float32, pin_memory=True); hv.copy_(val)
del colptr, rowidx, val; torch.cuda.empty_cache(); torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter()
    ds = Dataset.from_csc(p, n, hj.numpy(), hr.numpy(), hv.numpy(), store="f32", ctx=ctx)
    t1 = time.perf_counter()
    L = Lloyd(ds, K); L.set_centers(start); L.step(m / p, m / p, True)
    t2 = time.perf_counter()
    L.step(m / p, m / p, True)
    t3 = time.perf_counter()
    L.close(); ds.close()
    t4 = time.perf_counter()
    print(f"rep {rep}: from_csc {t1-t0:.3f}s first step {t2-t1:.3f}s second step {t3-t2:.4f}s close {t4-t3:.3f}s", file=sys.stderr)
