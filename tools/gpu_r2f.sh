#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_second_pass.py -x -q 2>&1 | tail -40 | tee gpurun_out/r2f_pytest_sp.log
for k in 47 46 39 38 31 24; do SKM_BOUNDED_KSM=$k timeout 120 python tools/probe_bounded.py config3; done 2>&1 | tee gpurun_out/r2f_bounded_ksm.txt
SKM_TRACE=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-extra > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench_trace.txt
grep -c . gpurun_out/r2f_bench_trace.txt; tail -40 gpurun_out/r2f_bench_trace.txt
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2f_bench.json') if l.startswith('{')][-1])
print(d['e2e'].get('whole_job_variant'))
PY
