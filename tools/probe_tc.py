"""Full assignment pass at the config-3 shape: gather kernels vs the tensor-core plan (ms per pass, same assignments)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from sparsifiedkmeans_b200 import Context, Lloyd

cfg = dict(bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "config3"])
n = int(sys.argv[2]) if len(sys.argv) > 2 else cfg["n"]
kind = sys.argv[3] if len(sys.argv) > 3 else "mixture"
p, K, m = cfg["p"], cfg["K"], cfg["m"]
if len(sys.argv) > 4:
    K = int(sys.argv[4])
gamma = m / p
ctx = Context(0)
dev = torch.device("cuda:0")
ds, views, mu, start = bench.gen_dataset(ctx, dev, n, p, m, K, 0, kind=kind)
out = {"n": n, "p": p, "K": K, "m": m, "kind": kind}
ref = None
for label, tc in (("gather", False), ("tc", True)):
    L = Lloyd(ds, K)
    L.set_tc_filter(tc)
    L.set_centers(start)
    t0 = time.perf_counter()
    st = L.step(gamma, gamma, True)           # includes the one-off image build
    ctx.synchronize()
    first = time.perf_counter() - t0
    for _ in range(3):
        L.step(gamma, gamma, True)
    ctx.timing_enable(True); ctx.timing_read()
    steps = 10
    for _ in range(steps):
        st = L.step(gamma, gamma, True)
    t = ctx.timing_read()
    ctx.timing_enable(False)
    a, _ = L.assignments(want_dist=False)
    if ref is None:
        ref = a
    out[label] = {"kernel": L.kernel_name, "first_step_s": first, "assign_ms": t["assign"][0] / steps, "recheck_ms": t["recheck"][0] / steps,
                  "accumulate_ms": t["accumulate"][0] / steps, "prep_ms": t["prep"][0] / steps, "last_tc": L.last_tc(),
                  "rechecked": st.n_rechecked, "same_as_gather": bool(np.array_equal(a, ref)),
                  "alg_GBps": n * (m * 8 + 8) / (t["assign"][0] / steps * 1e-3) / 1e9}
    L.close()
print(json.dumps(out))
