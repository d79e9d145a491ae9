#!/usr/bin/env python
"""Times one k-means++ round (kpp_update) at n=5e6, p=784, m=78 for the value of SKM_KPP_CTAS in the environment."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench
from sparsifiedkmeans_b200 import Context, Dataset
from sparsifiedkmeans_b200._lib import SKM_F32, SKM_I32, SKM_I64
dev = torch.device("cuda:0"); ctx = Context(0)
n = 5_000_000
colptr, rowidx, val, mu, start = bench.gen_shard_device(dev, n, 784, 78, 10, col0=0)
torch.cuda.synchronize()
ds = Dataset.from_device_csc(784, n, colptr.data_ptr(), SKM_I64, rowidx.data_ptr(), SKM_I32, val.data_ptr(), SKM_F32, store="f32", ctx=ctx)
c = ds.get_column(0)
ds.kpp_update(c, 78 / 784, first=True)
t0 = time.perf_counter()
for _ in range(5):
    tot = ds.kpp_update(c, 78 / 784, first=False)
ms = (time.perf_counter() - t0) / 5 * 1e3
print(json.dumps({"SKM_KPP_CTAS": os.environ.get("SKM_KPP_CTAS"), "ms_per_round": ms, "GBps": n * (78 * 8 + 16) / 1e9 / ms * 1e3, "sum": tot}))
