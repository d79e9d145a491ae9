"""Bounded + incremental iterations from a k-means++-like start at the config-3 shape: what the re-evaluation of the columns
that left their bound costs (for ncu: -k regex:"k_assign_list|k_exact_assign|k_assign_cols")."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from sparsifiedkmeans_b200 import Context, Lloyd
cfg = bench.CONFIGS["config3"]; p, K, m = cfg["p"], cfg["K"], cfg["m"]; gamma = m / p
n = int(sys.argv[1]) if len(sys.argv) > 1 else cfg["n"]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 16
ctx = Context(0); dev = torch.device("cuda:0")
ds, views, mu, start = bench.gen_dataset(ctx, dev, n, p, m, K, 0, kind="mixture")
rng = np.random.default_rng(5)
c0 = start[:, rng.integers(0, K, K)] + 0.05 * rng.standard_normal(start.shape)      # several centres inside one planted cluster
L = Lloyd(ds, K, incremental=True, bounded=True); L.set_centers(c0)
ctx.timing_enable(True); ctx.timing_read()
out = []
for it in range(steps):
    st = L.step(gamma, gamma, True)
    t = ctx.timing_read()
    out.append((round(t["assign"][0], 3), round(t["recheck"][0], 3), L.last_assign_flagged(), L.last_prune()[0], st.n_rechecked))
print(json.dumps({"n": n, "list_min": os.environ.get("SKM_LIST_MIN"), "assign_ms,recheck_ms,bounded_flagged,prune_not_kept,fp64_columns": out}))
