#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -30 gpurun_out/r2b_pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/r2b_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2b_bench.json'))
def short(o, depth=0):
    if isinstance(o, dict):
        return {k: short(v, depth+1) for k,v in o.items()}
    if isinstance(o, list) and len(o)>12: return o[:12]+['...']
    return o
print(json.dumps(short(d), indent=1)[:12000])
PY
