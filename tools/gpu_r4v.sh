#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r4v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r4v_pytest.log
tail -5 gpurun_out/r4v_pytest.log
timeout 1200 python bench.py > gpurun_out/r4v_bench.json 2> gpurun_out/r4v_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r4v_bench.err | cut -c1-300
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r4v_bench.json') if l.startswith('{')][-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'parity', d.get('parity',{}).get('assignment_mismatches'))
print('whole_job', d.get('whole_job',{}).get('seconds'))
for k in ('config3','unstructured','strong_scaling'):
    v=d.get(k,{})
    print(k, {x:v.get(x) for x in ('ms_per_step','value','rechecked_last_step','error','k1_ms_per_pass')}, (v.get('parity') or {}).get('assignment_mismatches'), v.get('roofline',{}).get('frac'), v.get('roofline',{}).get('kernel'))
for k in ('trajectory_config2','trajectory_config3','trajectory_unstructured'):
    v=d.get(k,{})
    print(k, v.get('speedup'), v.get('error'), v.get('default',{}).get('iterations'), v.get('default',{}).get('total_ms'), v.get('bounded_incremental',{}).get('total_ms'), v.get('iterations_with_identical_assignments'))
print(d['config3'].get('bounded_assign',{}).get('ms_per_step'))
PY
