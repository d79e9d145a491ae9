"""K2 (k_accumulate_csr) at a given K: ms per pass for the bin width in SKM_K2_BW."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from sparsifiedkmeans_b200 import Context, Lloyd
K = int(sys.argv[1]); p = int(sys.argv[2]); m = int(sys.argv[3]); n = int(sys.argv[4]); gamma = m / p
ctx = Context(0); dev = torch.device("cuda:0")
ds, views, mu, start = bench.gen_dataset(ctx, dev, n, p, m, K, 0, kind="mixture")
L = Lloyd(ds, K); L.set_centers(start)
for _ in range(3): L.step(gamma, gamma, True)
ctx.timing_enable(True); ctx.timing_read()
for _ in range(5): L.step(gamma, gamma, True)
t = ctx.timing_read()
print(json.dumps({"K": K, "p": p, "m": m, "n": n, "bw": os.environ.get("SKM_K2_BW"), "accumulate_ms": round(t["accumulate"][0] / 5, 4),
                  "assign_ms": round(t["assign"][0] / 5, 4), "kernel": L.kernel_name}))
