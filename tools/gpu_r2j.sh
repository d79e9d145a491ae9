#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
tail -6 gpurun_out/r2j_pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2j_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2j_bench.json') if l.startswith('{')][-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'])
print('e2e', d['e2e']['value'], d['e2e'].get('whole_job_variant'))
print('cpu', d['cpu_baseline'])
for k in ('config3','unstructured','strong_scaling'):
    v=d.get(k,{})
    print(k, {x:v.get(x) for x in ('ms_per_step','value','parity','rechecked_last_step','error','k1_ms_per_pass')}, v.get('roofline',{}).get('frac'), v.get('bounded_assign',{}).get('ms_per_step'), v.get('bounded_assign',{}).get('assign_ms'))
for k in ('trajectory_config2','trajectory_config3','trajectory_unstructured'):
    v=d.get(k,{})
    print(k, v.get('speedup'), v.get('error'), v.get('default',{}).get('iterations'), v.get('default',{}).get('total_ms'), v.get('bounded_incremental',{}).get('total_ms'), v.get('iterations_with_identical_assignments'))
PY
