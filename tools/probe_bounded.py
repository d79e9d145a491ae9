"""Bounded assignment pass at the config-3 shape at its fixed point: ms per pass (tuning knob SKM_BOUNDED_KSM)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from sparsifiedkmeans_b200 import Context, Lloyd

cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "config3"]
n, p, K, m = cfg["n"], cfg["p"], cfg["K"], cfg["m"]
gamma = m / p
ctx = Context(0)
dev = torch.device("cuda:0")
ds, views, mu, start = bench.gen_dataset(ctx, dev, n, p, m, K, 0)
L = Lloyd(ds, K, incremental=True, bounded=True)
L.set_centers(start)
for _ in range(6):
    L.step(gamma, gamma, True)
ctx.timing_enable(True); ctx.timing_read()
for _ in range(10):
    st = L.step(gamma, gamma, True)
t = ctx.timing_read()
print(json.dumps({"ksm_env": os.environ.get("SKM_BOUNDED_KSM"), "assign_ms": t["assign"][0] / 10, "recheck_ms": t["recheck"][0] / 10,
                  "accumulate_ms": t["accumulate"][0] / 10, "prep_ms": t["prep"][0] / 10, "flagged": L.last_assign_flagged(),
                  "GBps": n * (m * 8 + 8) / (t["assign"][0] / 10 * 1e-3) / 1e9}))
