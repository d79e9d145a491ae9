"""Pruned full pass at the config-3 shape (n columns): a few steps, for ncu / timing."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from sparsifiedkmeans_b200 import Context, Lloyd
cfg = bench.CONFIGS["config3"]; p, K, m = cfg["p"], cfg["K"], cfg["m"]; gamma = m / p
n = int(sys.argv[1]) if len(sys.argv) > 1 else cfg["n"]
ctx = Context(0); dev = torch.device("cuda:0")
ds, views, mu, start = bench.gen_dataset(ctx, dev, n, p, m, K, 0, kind="mixture")
L = Lloyd(ds, K); L.set_centers(start)
for _ in range(3): L.step(gamma, gamma, True)
ctx.timing_enable(True); ctx.timing_read()
for _ in range(5): st = L.step(gamma, gamma, True)
t = ctx.timing_read()
print(json.dumps({"n": n, "kernel": L.kernel_name, "assign_ms": t["assign"][0] / 5, "recheck_ms": t["recheck"][0] / 5, "accumulate_ms": t["accumulate"][0] / 5, "last_prune": L.last_prune()}))
