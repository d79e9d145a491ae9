#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_prune.py -q -x > gpurun_out/r3e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3e_pytest.log
tail -12 gpurun_out/r3e_pytest.log
python - <<'PY' > gpurun_out/r3e_probe.json 2> gpurun_out/r3e_probe.err
import json, os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch, bench
from sparsifiedkmeans_b200 import Context, Lloyd
cfg = bench.CONFIGS["config3"]; n, p, K, m = cfg["n"], cfg["p"], cfg["K"], cfg["m"]; gamma = m / p
ctx = Context(0); dev = torch.device("cuda:0")
out = {}
for kind in ("mixture", "unstructured"):
    ds, views, mu, start = bench.gen_dataset(ctx, dev, n, p, m, K, 0, kind=kind)
    ref = None
    for label, pr in (("off", False), ("auto", None)):
        L = Lloyd(ds, K); L.set_prune(pr); L.set_centers(start)
        for _ in range(4): L.step(gamma, gamma, True)
        ctx.timing_enable(True); ctx.timing_read()
        steps = 9
        for _ in range(steps): st = L.step(gamma, gamma, True)
        t = ctx.timing_read(); ctx.timing_enable(False)
        a, _ = L.assignments(want_dist=False)
        if ref is None: ref = a
        out[kind + "_" + label] = {"kernel": L.kernel_name, "assign_ms": t["assign"][0] / steps, "recheck_ms": t["recheck"][0] / steps,
                                   "last_prune": L.last_prune(), "rechecked": st.n_rechecked, "same": bool(np.array_equal(a, ref))}
        L.close()
    ds.close(); del views
    torch.cuda.empty_cache()
print(json.dumps(out))
PY
tail -3 gpurun_out/r3e_probe.err; cat gpurun_out/r3e_probe.json
