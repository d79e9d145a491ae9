import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sparsifiedkmeans_b200 import Context, Dataset
p2, n, m = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda:0")
ctx = Context(0)
x = torch.randn(n, p2, device=dev)
signs = torch.ones(p2, device=dev)
keys = torch.rand(min(n, 20000), p2, device=dev)
rows1 = keys.topk(m, dim=1, largest=False).indices.to(torch.int32)
rows = rows1.repeat((n + rows1.shape[0] - 1) // rows1.shape[0], 1)[:n].contiguous()
torch.cuda.synchronize()
ds = Dataset.from_fwht_sample(p2, n, m, x.data_ptr(), signs.data_ptr(), rows.data_ptr(), ctx=ctx)
print("ok", ds.nnz, ds.max_col_nnz)
