#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_prune.py -q > gpurun_out/r3f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3f_pytest.log
tail -5 gpurun_out/r3f_pytest.log
for PAIRS in 2 3 4 6 8; do
SKM_PRUNE_PAIRS=$PAIRS python - <<'PY'
import json, os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch, bench
from sparsifiedkmeans_b200 import Context, Lloyd
cfg = bench.CONFIGS["config3"]; n, p, K, m = cfg["n"], cfg["p"], cfg["K"], cfg["m"]; gamma = m / p
ctx = Context(0); dev = torch.device("cuda:0")
ds, views, mu, start = bench.gen_dataset(ctx, dev, n, p, m, K, 0, kind="mixture")
L = Lloyd(ds, K); L.set_prune(True); L.set_centers(start)
for _ in range(4): L.step(gamma, gamma, True)
ctx.timing_enable(True); ctx.timing_read()
for _ in range(9): st = L.step(gamma, gamma, True)
t = ctx.timing_read()
print(json.dumps({"pairs": os.environ["SKM_PRUNE_PAIRS"], "assign_ms": t["assign"][0] / 9, "recheck_ms": t["recheck"][0] / 9, "last_prune": L.last_prune()}))
PY
done
