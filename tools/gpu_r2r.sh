#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2r_pytest.log
tail -6 gpurun_out/r2r_pytest.log
SKM_TRACE=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-extra --no-cpu > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; echo "bench rc=$?"
grep -v "skm trace" gpurun_out/r2r_bench.err | tail -3
grep "skm trace" gpurun_out/r2r_bench.err | tail -14
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2r_bench.json') if l.startswith('{')][-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'])
print('e2e', d['e2e']['value'], d['e2e'].get('whole_job_variant'))
PY
